"""The C-ABI library loads on a CPU-only host and exports exactly what include/msda3d.h declares.
Argument validation that returns before touching the device is exercised too (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from transoar_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for header in sorted(os.listdir(os.path.join(ROOT, "include"))):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b((?:msda3d|roi_attn|win_attn|instnorm|tc_gemm|tc_colsum|stem_conv3d|conv3d_tc|conv3d_gen|criterion|fused_ln|hash_rng)(?:_[a-z0-9_]+)?)\s*\(", text))
    return sorted(names)


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.lib()
    declared = _declared()
    assert len(declared) >= 12 and "roi_attn_forward" in declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/msda3d.h but not exported"
    assert sorted(_lib.exported_symbols()) == declared, "python binding table out of sync with the header"


def test_abi_version_and_error_strings():
    lib = _lib.lib()
    assert lib.msda3d_abi_version() == 1
    assert lib.msda3d_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4):
        assert lib.msda3d_error_string(code).startswith(b"msda3d:")
    assert b"invalid" in lib.msda3d_error_string(1).lower()           # cudaErrorInvalidValue


def test_null_and_bad_dimension_arguments_are_rejected_before_any_launch():
    lib = _lib.lib()
    buf = (ctypes.c_double * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    before = lib.msda3d_launch_count()
    dims = (1, 8, 1, 1, 1, 1, 1)
    assert lib.msda3d_forward(None, 0, None, p, p, p, p, *dims, p) == -1
    assert lib.msda3d_forward(None, 0, p, p, p, p, p, 0, 8, 1, 1, 1, 1, 1, p) == -1
    assert lib.msda3d_forward(None, 9, p, p, p, p, p, *dims, p) == -1
    assert lib.msda3d_backward(None, 0, p, p, p, p, p, p, *dims, p, p, None) == -1
    assert lib.msda3d_debug_indices(None, 0, p, p, 1, 1, 1, 0, 1, p, p) == -1
    odd = ctypes.c_void_p(p.value + 2)
    assert lib.msda3d_forward(None, 0, odd, p, p, p, p, *dims, p) == -3
    assert lib.msda3d_forward_host(99, 0, p, p, p, p, p, *dims, p) == -4
    assert lib.tc_gemm_tf32(None, None, 0, 4, p, 0, 4, p, 4, None, 4, 4, 4, 0, 0, 1) == -1         # null operand
    assert lib.tc_gemm_tf32(None, p, 0, 6, p, 0, 4, p, 4, None, 4, 4, 4, 0, 0, 1) == -3            # lda not a multiple of 4 (TMA stride)
    assert lib.tc_gemm_tf32(None, p, 0, 4, p, 0, 4, p, 4, None, 4, 4, 4, 0, 0, 3) == -1            # split-K without accumulate
    assert lib.msda3d_launch_count() == before


def test_header_cites_the_reference_interfaces_it_replaces():
    text = open(os.path.join(ROOT, "include", "msda3d.h")).read()
    for needle in ("ms_deform_im2col_cuda.cuh:1094-1125", "cuh:1127-1507", "ms_deform_attn_cuda.cu", "vision.cpp:13-16"):
        assert needle in text


def test_missing_library_fails_loudly_and_nothing_falls_back():
    """The product path must fail when the CUDA extension is missing: a fresh interpreter whose MSDA3D_LIB points at a file that does not
    exist raises from the first native call -- the loader names the build command and there is no CPU / PyTorch route to fall into."""
    import subprocess
    import sys
    code = (
        "import torch\n"
        "from transoar_b200 import _lib\n"
        "from transoar_b200.ops.functions import MSDeformAttnFunction\n"
        "try:\n"
        "    _lib.lib()\n"
        "except RuntimeError as e:\n"
        "    assert 'is missing' in str(e) and 'no CPU or PyTorch fallback' in str(e), e\n"
        "else:\n"
        "    raise SystemExit('loader returned a library that does not exist')\n"
        "import transoar_b200.instnorm as I\n"
        "try:\n"
        "    I.instance_norm_relu(torch.zeros(1, 2, 2, 2, 2), torch.ones(2), torch.zeros(2))\n"
        "except RuntimeError as e:\n"
        "    print('LOUD', type(e).__name__)\n"
        "else:\n"
        "    raise SystemExit('instance_norm_relu computed something without the library')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MSDA3D_LIB="/nonexistent/libmsda3d.so", PYTHONPATH=root)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=root, timeout=300)
    assert res.returncode == 0 and "LOUD" in res.stdout, res.stdout + res.stderr


@pytest.mark.parametrize("zero", range(7))
def test_empty_problems_are_rejected_not_launched(zero):
    """any of N, S, M, C, L, Lq, P equal to 0 is MSDA3D_EINVAL for forward and backward (the reference launches
    zero blocks and prints the launch error, cuh:1119-1123)."""
    lib = _lib.lib()
    buf = (ctypes.c_double * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    dims = [1, 8, 1, 4, 1, 1, 1]
    dims[zero] = 0
    before = lib.msda3d_launch_count()
    assert lib.msda3d_forward(None, 0, p, p, p, p, p, *dims, p) == -1
    assert lib.msda3d_backward(None, 0, p, p, p, p, p, p, *dims, p, p, p) == -1
    assert lib.msda3d_launch_count() == before
