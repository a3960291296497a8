"""alias of instantavatar_b200.models.structures"""
