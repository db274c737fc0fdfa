"""ctypes binding of libmsda3d.so (C ABI: include/msda3d.h).  Fails loudly: no fallback of any kind."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# MSDA3D_LIB lets profiling sessions load an experimental build of the same library; never a different implementation.
LIB_PATH = os.environ.get("MSDA3D_LIB") or os.path.join(_HERE, "libmsda3d.so")
_lib = None

F32, F64, BF16, F16 = 0, 1, 2, 3

_vp, _ci = ctypes.c_void_p, ctypes.c_int
_DIMS = [_ci] * 7

_SIGNATURES = {
    "msda3d_abi_version": (_ci, []),
    "msda3d_error_string": (ctypes.c_char_p, [_ci]),
    "msda3d_launch_count": (ctypes.c_ulonglong, []),
    "msda3d_set_tuning": (_ci, [ctypes.c_char_p, _ci]),
    "msda3d_forward": (_ci, [_vp, _ci] + [_vp] * 5 + _DIMS + [_vp]),
    "msda3d_backward": (_ci, [_vp, _ci] + [_vp] * 6 + _DIMS + [_vp] * 3),
    "msda3d_fused_supported": (_ci, [_ci, _ci, _ci]),
    "msda3d_forward_fused": (_ci, [_vp] * 5 + [_ci] + [_vp] * 2 + _DIMS + [_vp]),
    "msda3d_backward_fused": (_ci, [_vp] * 6 + [_ci] + [_vp] * 2 + _DIMS + [_vp] * 3),
    "msda3d_forward_fused_ld": (_ci, [_vp] * 5 + [_ci] + [_vp] * 2 + [ctypes.c_longlong] + _DIMS + [_vp]),
    "msda3d_backward_fused_ld": (_ci, [_vp] * 6 + [_ci] + [_vp] * 2 + [ctypes.c_longlong] + _DIMS + [_vp] * 3),
    "msda3d_forward_host": (_ci, [_ci, _ci] + [_vp] * 5 + _DIMS + [_vp]),
    "msda3d_backward_host": (_ci, [_ci, _ci] + [_vp] * 6 + _DIMS + [_vp] * 3),
    "msda3d_forward_backward_host": (_ci, [_ci, _ci] + [_vp] * 6 + _DIMS + [_vp] * 4),
    "msda3d_host_release": (None, []),
    "msda3d_debug_indices": (_ci, [_vp, _ci, _vp, _vp] + [_ci] * 5 + [_vp, _vp]),
    # include/instnorm.h
    "instnorm_workspace_floats": (ctypes.c_longlong, [_ci, _ci, _ci, ctypes.c_longlong]),
    "instnorm_relu_forward": (_ci, [_vp, _ci, _vp, _vp, _vp, _ci, _ci, ctypes.c_longlong, ctypes.c_float, _vp, _vp, _vp, _vp]),
    "instnorm_relu_backward": (_ci, [_vp, _ci] + [_vp] * 6 + [_ci, _ci, ctypes.c_longlong] + [_vp] * 4),
    "instnorm_ndhwc_workspace_floats": (ctypes.c_longlong, [_ci, _ci, ctypes.c_longlong]),
    "instnorm_relu_forward_ndhwc": (_ci, [_vp] * 4 + [_ci, _ci, ctypes.c_longlong, ctypes.c_float] + [_vp] * 4),
    "instnorm_relu_backward_ndhwc": (_ci, [_vp] * 7 + [_ci, _ci, ctypes.c_longlong] + [_vp] * 4),
    "instnorm_relu_forward_ndhwc_bf16": (_ci, [_vp] * 4 + [_ci, _ci, ctypes.c_longlong, ctypes.c_float] + [_vp] * 4),
    "instnorm_relu_backward_ndhwc_bf16": (_ci, [_vp] * 7 + [_ci, _ci, ctypes.c_longlong] + [_vp] * 4),
    # include/roi_attn.h
    "roi_attn_workspace_floats": (ctypes.c_longlong, [_ci] * 5),
    "roi_attn_forward": (_ci, [_vp] * 5 + [_ci] * 8 + [_vp, _vp, _vp, ctypes.c_longlong]),
    "roi_attn_backward": (_ci, [_vp] * 5 + [_ci] * 8 + [_vp] * 6),
    "roi_attn_forward_tf32": (_ci, [_vp] * 5 + [_ci] * 8 + [_vp, _vp, _vp, ctypes.c_longlong]),
    "roi_attn_backward_tf32": (_ci, [_vp] * 5 + [_ci] * 8 + [_vp] * 6),
    # include/win_attn.h
    "win_attn_supported": (_ci, [_ci, _ci]),
    "win_attn_forward": (_ci, [_vp] * 4 + [_ci] * 5 + [ctypes.c_float, _vp, _vp]),
    "win_attn_backward": (_ci, [_vp] * 8 + [_ci] * 5 + [ctypes.c_float, _vp, _vp]),
    # include/conv3d_tc.h
    "conv3d_tc_supported": (_ci, [_ci, _ci]),
    "hash_rng_set_epoch": (None, [_vp]),
    "conv3d_tc_wgrad_workspace_floats": (ctypes.c_longlong, []),
    "conv3d_tc_k3_wgrad": (_ci, [_vp, _vp, _vp] + [_ci] * 6 + [_vp, _vp]),
    "conv3d_tc_debug_mode": (None, [_ci]),
    "conv3d_tc_debug_mn_probe": (_ci, [_vp] * 4 + [_ci]),
    "conv3d_tc_k3_forward": (_ci, [_vp, _vp, _vp] + [_ci] * 6 + [_vp]),
    # include/conv3d_gen.h
    "conv3d_gen_supported": (_ci, [_ci, _ci, _ci]),
    "conv3d_gen_forward": (_ci, [_vp] * 4 + [_ci] * 7 + [_vp]),
    "conv3d_gen_dgrad": (_ci, [_vp] * 3 + [_ci] * 7 + [_vp]),
    "conv3d_gen_dgrad_s2_folded": (_ci, [_vp] * 3 + [_ci] * 6 + [_vp]),
    "conv3d_gen_wgrad": (_ci, [_vp] * 3 + [_ci] * 7 + [_vp]),
    "conv3d_gen_set_path": (None, [_ci]),
    "conv3d_gen_debug_mma_rate": (_ci, [_vp, _ci, _ci, _ci, _vp]),
    "conv3d_gen_debug_k_probe": (_ci, [_vp] * 4 + [_ci] * 3),
    # include/criterion.h
    "criterion_fused_supported": (_ci, [_ci, _ci]),
    "criterion_fused": (_ci, [_vp] * 7 + [_ci] * 4 + [ctypes.c_float] * 3 + [_vp] * 4),
    # include/fused_ln.h
    "fused_ln_workspace_floats": (ctypes.c_longlong, [_ci]),
    "fused_ln_forward": (_ci, [_vp] * 5 + [ctypes.c_longlong, _ci, ctypes.c_float, ctypes.c_float, ctypes.c_ulonglong] + [_vp] * 4),
    "fused_ln_backward": (_ci, [_vp] * 6 + [ctypes.c_longlong, _ci, ctypes.c_float, ctypes.c_ulonglong] + [_vp] * 5),
    "fused_ln_forward_bf16b": (_ci, [_vp] * 5 + [ctypes.c_longlong, _ci, ctypes.c_float, ctypes.c_float, ctypes.c_ulonglong] + [_vp] * 4),
    "fused_ln_backward_bf16b": (_ci, [_vp] * 6 + [ctypes.c_longlong, _ci, ctypes.c_float, ctypes.c_ulonglong] + [_vp] * 5),
    # include/stem_conv.h
    "stem_conv3d_workspace_floats": (ctypes.c_longlong, [_ci]),
    "stem_conv3d_forward": (_ci, [_vp, _vp, _vp] + [_ci] * 5 + [_vp]),
    "stem_conv3d_wgrad": (_ci, [_vp, _vp, _vp] + [_ci] * 5 + [_vp, _vp]),
    # include/tc_gemm.h
    "tc_colsum_workspace_floats": (ctypes.c_longlong, [_ci]),
    "tc_colsum": (_ci, [_vp, _vp, ctypes.c_longlong, _ci, ctypes.c_longlong, _vp, _vp]),
    "tc_gemm_debug_profile": (None, [_vp]),
    "tc_gemm_tf32_ex": (_ci, [_vp, _vp, _ci, ctypes.c_longlong, _vp, _ci, ctypes.c_longlong, _vp, ctypes.c_longlong, _vp] + [_ci] * 6
                        + [_vp, ctypes.c_float, ctypes.c_float, ctypes.c_ulonglong]),
    "tc_gemm_tf32": (_ci, [_vp, _vp, _ci, ctypes.c_longlong, _vp, _ci, ctypes.c_longlong, _vp, ctypes.c_longlong, _vp] + [_ci] * 6),
    "tc_gemm_bf16": (_ci, [_vp, _vp, _ci, ctypes.c_longlong, _vp, _ci, ctypes.c_longlong, _vp, _ci, ctypes.c_longlong, _vp] + [_ci] * 6
                     + [_vp, ctypes.c_float, ctypes.c_float, ctypes.c_ulonglong]),
}


def build(verbose: bool = False) -> str:
    """Compile csrc/ -> libmsda3d.so for sm_100a with nvcc (no GPU needed)."""
    res = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libmsda3d.so failed:\n" + (res.stdout or "") + (res.stderr or ""))
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """The loaded library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C transoar_b200/csrc`).  transoar_b200 has no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.msda3d_abi_version() != 1:
            raise RuntimeError("libmsda3d.so ABI version mismatch; rebuild it")
        # experiment switches of include/msda3d.h (msda3d_set_tuning) from the environment, e.g. TRANSOAR_B200_TUNING="duo=0,rot=0"
        for item in filter(None, os.environ.get("TRANSOAR_B200_TUNING", "").split(",")):
            key, _, val = item.partition("=")
            if handle.msda3d_set_tuning(key.strip().encode(), int(val)) != 0:
                raise RuntimeError(f"TRANSOAR_B200_TUNING: unknown switch or value {item!r}")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib().msda3d_error_string(rc).decode()} (code {rc})")


def exported_symbols():
    return list(_SIGNATURES)
