"""alias of instantavatar_b200.deformers"""
