"""transoar_b200 -- B200-native (sm_100a) drop-in for the 3D multi-scale deformable attention path of
bwittmann/transoar (``transoar/models/ops``): same autograd surface, same module, same ``use_cuda`` flag.

    from transoar_b200.ops.functions import MSDeformAttnFunction      # transoar.models.ops.functions
    from transoar_b200.ops.modules import MSDeformAttn                # transoar.models.ops.modules
    import transoar_b200.MultiScaleDeformableAttention as MSDA        # the extension module the reference imports

The compute lives in ``libmsda3d.so`` (C ABI in ``include/msda3d.h``, CUDA in ``transoar_b200/csrc``).  There is no
CPU fallback: importing works anywhere, calling the op without the library or without a CUDA tensor raises.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"


def install_into_reference():
    """Make the UNMODIFIED reference models run on these kernels.

    The reference's ``use_cuda=True`` path is broken as shipped: ``import MultiScaleDeformableAttention as MSDA`` is
    commented out (transoar/models/ops/functions/ms_deform_attn_func.py:18), so ``MSDeformAttnFunction.apply`` raises
    NameError.  This registers our module under that name and binds ``MSDA`` inside the reference's function module,
    which is all a maintainer has to do (INTEGRATION.md shows the two-line permanent patch).
    """
    import importlib
    import sys

    from . import MultiScaleDeformableAttention as msda
    sys.modules.setdefault("MultiScaleDeformableAttention", msda)
    ref = importlib.import_module("transoar.models.ops.functions.ms_deform_attn_func")
    ref.MSDA = msda
    return ref
