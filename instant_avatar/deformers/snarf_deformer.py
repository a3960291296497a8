"""`instant_avatar.deformers.snarf_deformer.SNARFDeformer` (confs/deformer/fast_snarf*.yaml) -> instantavatar_b200 mirror"""
from instantavatar_b200.deformers.snarf_deformer import (SNARFDeformer, get_bbox_from_smpl,  # noqa: F401
                                                         get_predefined_rest_pose)
