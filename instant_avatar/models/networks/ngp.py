"""`instant_avatar.models.networks.ngp.NeRFNGPNet` (confs/network/ngp.yaml) -> instantavatar_b200 mirror"""
from instantavatar_b200.models.networks.ngp import EPS, NeRFNGPNet  # noqa: F401
