from .ms_deform_attn_func import MSDeformAttnFunction, MSDeformAttnFusedFunction, MSDeformAttnMergedFunction  # noqa: F401  (transoar/models/ops/functions/__init__.py:9)
