"""alias of the Fast-SNARF state holder"""
