"""`instant_avatar.models.structures.body_model_param.SMPLParamEmbedding` -> instantavatar_b200 mirror"""
from instantavatar_b200.models.structures.body_model_param import SMPLParamEmbedding  # noqa: F401
