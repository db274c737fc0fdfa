from .ms_deform_attn_func import MSDeformAttnFunction, MSDeformAttnFusedFunction  # noqa: F401  (transoar/models/ops/functions/__init__.py:9)
