"""`instant_avatar.deformers.smpl_deformer.SMPLDeformer` (confs/deformer/smpl.yaml) -> instantavatar_b200 mirror"""
from instantavatar_b200.deformers.smpl_deformer import SMPLDeformer  # noqa: F401
