"""`instant_avatar.utils.loss.{NeRFLoss,NGPLoss}` (confs/SNARF_NGP*.yaml: `model.opt.loss._target_`) -> mirrors"""
from instantavatar_b200.utils_loss import NeRFLoss, NGPLoss  # noqa: F401
