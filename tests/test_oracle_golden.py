"""The CPU oracle against the golden fixtures produced by the REFERENCE's own Python path
(tests/golden/make_golden.py imports ms_deform_attn_core_pytorch from /root/reference).  Tolerances are the
reference's own (transoar/models/ops/test.py:69-97): fp64 torch.allclose defaults, fp32 rtol 1e-2 / atol 1e-3 --
and much tighter bounds on top, since both sides compute the same trilinear blend."""
import numpy as np
import pytest
import torch

from oracle import msda3d_oracle as O

CASES = ["small", "tiny", "border", "heads6"]


def _inputs(z, dt):
    return (z["value"].astype(dt), z["shapes"], z["starts"], z["loc"].astype(dt), z["aw"].astype(dt), z["grad_out"].astype(dt))


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("contract", [True, False])
def test_forward_fp64_matches_reference_python_path(golden_cases, case, contract):
    z = golden_cases[case]
    v, sh, st, loc, aw, _ = _inputs(z, np.float64)
    out = O.forward(v, sh, st, loc, aw, contract)
    assert np.allclose(out, z["out_f64"], rtol=1e-5, atol=1e-8)          # ops/test.py:77
    assert np.abs(out - z["out_f64"]).max() < 1e-15


@pytest.mark.parametrize("case", CASES)
def test_forward_fp32_matches_reference_python_path(golden_cases, case):
    z = golden_cases[case]
    v, sh, st, loc, aw, _ = _inputs(z, np.float32)
    out = O.forward(v, sh, st, loc, aw)
    assert np.allclose(out, z["out_f32"], rtol=1e-2, atol=1e-3)          # ops/test.py:93
    assert np.abs(out - z["out_f32"]).max() < 1e-7                       # values are O(1e-2)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("tag,dt,tol", [("f64", np.float64, 1e-13), ("f32", np.float32, 2e-5)])
def test_backward_matches_reference_autograd(golden_cases, case, tag, dt, tol):
    z = golden_cases[case]
    v, sh, st, loc, aw, go = _inputs(z, dt)
    gv, gl, ga = O.backward(go, v, sh, st, loc, aw)
    for got, key in ((gv, "grad_value"), (gl, "grad_loc"), (ga, "grad_aw")):
        want = z[f"{key}_{tag}"]
        scale = max(1.0, float(np.abs(want).max()))
        assert np.abs(got - want).max() <= tol * scale, key


@pytest.mark.parametrize("case", CASES)
def test_gridsample_restatement_is_the_reference_python_path(golden_cases, case):
    """oracle.gridsample_path restates func.py:41-65; same ATen ops, so fp64 results agree to round-off."""
    z = golden_cases[case]
    v, loc, aw = (torch.from_numpy(z[k]).requires_grad_(True) for k in ("value", "loc", "aw"))
    out = O.gridsample_path(v, [tuple(r) for r in z["shapes"].tolist()], loc, aw)
    assert torch.allclose(out, torch.from_numpy(z["out_f64"]), rtol=1e-12, atol=1e-15)
    out.backward(torch.from_numpy(z["grad_out"]))
    assert torch.allclose(v.grad, torch.from_numpy(z["grad_value_f64"]), rtol=1e-10, atol=1e-14)
    assert torch.allclose(loc.grad, torch.from_numpy(z["grad_loc_f64"]), rtol=1e-10, atol=1e-14)
    assert torch.allclose(aw.grad, torch.from_numpy(z["grad_aw_f64"]), rtol=1e-10, atol=1e-14)


def test_zero_padding_and_range_test_edges():
    """Hand-built known answers on a 2x2x2 volume (cuh:60-107 corner guards, cuh:428 range test)."""
    shapes = np.array([[2, 2, 2]], dtype=np.int64)
    starts = np.array([0], dtype=np.int64)
    value = np.arange(8, dtype=np.float64).reshape(1, 8, 1, 1) + 1.0      # voxel (d,h,w) -> 1 + 4d + 2h + w
    aw = np.ones((1, 1, 1, 1, 1), dtype=np.float64)

    def f(x, y, z):
        loc = np.array([x, y, z], dtype=np.float64).reshape(1, 1, 1, 1, 1, 3)
        return float(O.forward(value, shapes, starts, loc, aw)[0, 0, 0])

    assert f(0.25, 0.25, 0.25) == 1.0          # exactly voxel centre (0,0,0): pixel coord 0
    assert f(0.75, 0.75, 0.75) == 8.0          # centre of voxel (1,1,1)
    assert f(0.5, 0.5, 0.5) == 4.5             # middle of the volume: mean of all eight
    assert f(0.0, 0.25, 0.25) == 0.5           # pixel w = -0.5: half of voxel (0,0,0), other half is zero padding
    assert f(-0.25, 0.25, 0.25) == 0.0         # pixel w = -1.0 fails `w_im > -1`
    assert f(1.0, 0.75, 0.75) == 4.0           # pixel w = 1.5: half of voxel (1,1,1)
    assert f(1.25, 0.75, 0.75) == 0.0          # pixel w = 2.0 fails `w_im < W`
    assert f(float("nan"), 0.5, 0.5) == 0.0    # NaN fails every comparison -> sample skipped


def test_contract_mode_only_moves_knife_edge_samples():
    """fma(loc,size,-0.5) vs round(loc*size)-0.5 differ by at most 1 ulp of the pixel coordinate; floor() may flip only there."""
    rng = np.random.default_rng(7)
    shapes = np.array([[40, 40, 64], [5, 5, 8]], dtype=np.int64)
    loc = rng.random((1, 4096, 2, 2, 4, 3), dtype=np.float32)
    ia, fa = O.indices(shapes, loc, contract=True)
    ib, fb = O.indices(shapes, loc, contract=False)
    same = (ia == ib).all(-1)
    assert same.mean() > 0.999
    # where the integer parts agree the fractions agree to an ulp of the coordinate (< 64 * 2^-23)
    assert np.abs(fa[same] - fb[same]).max() < 1e-5
