"""alias of instantavatar_b200.models (reference: instant_avatar/models/__init__.py)"""
