"""`instant_avatar.models.structures.utils.Rays` -> instantavatar_b200 mirror"""
from instantavatar_b200.models.dnerf import Rays  # noqa: F401
