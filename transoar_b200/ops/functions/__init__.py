from .ms_deform_attn_func import MSDeformAttnFunction  # noqa: F401  (transoar/models/ops/functions/__init__.py:9)
