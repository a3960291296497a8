"""`instant_avatar.deformers.fast_snarf.deformer_torch.ForwardDeformer` -> instantavatar_b200 mirror"""
from instantavatar_b200.deformers.snarf_deformer import ForwardDeformer  # noqa: F401
