"""`instant_avatar.renderers.raymarcher_acc.Raymarcher` (confs/renderer/raymarcher_acc.yaml) -> instantavatar_b200 mirror"""
from instantavatar_b200.renderers.raymarcher_acc import BoundModel, Raymarcher  # noqa: F401
