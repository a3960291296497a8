"""`instant_avatar.models.DNeRF` (confs/SNARF_NGP*.yaml: `model._target_`) -> instantavatar_b200.models.dnerf"""
from instantavatar_b200.models.dnerf import DNeRFModel, Rays  # noqa: F401
