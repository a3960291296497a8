"""`instant_avatar.models.structures.density_grid` -> instantavatar_b200 mirror"""
from instantavatar_b200.models.structures.density_grid import DensityGrid, denormalize, max_connected_component  # noqa: F401
