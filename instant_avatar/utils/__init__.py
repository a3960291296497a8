"""alias of the loss mirrors"""
