from .ms_deform_attn import MSDeformAttn  # noqa: F401  (transoar/models/ops/modules/__init__.py)
