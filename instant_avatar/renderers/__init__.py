"""alias of instantavatar_b200.renderers"""
