"""Parity tests proper: the sm_100a kernels, called through the C ABI (ctypes -> libmsda3d.so), against
  (1) the CPU oracle (oracle/msda3d_oracle.c, pinned to the reference's Python path by tests/golden),
  (2) the golden fixtures themselves,
  (3) the reference's OWN CUDA kernels (oracle/_ref/libmsda3d_refcuda.so, compiled from /root/reference by oracle/Makefile),
and, at BASELINE.json's full sizes, size-independent properties (linearity, adjoint identities, partition of unity).

Tolerances (north_star): fp32 1e-4, bf16/fp16 1e-2, sampling-index arithmetic bit-exact.  The fp32 forward is in fact
bit-identical to both the oracle and the compiled reference, which is asserted.
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import msda3d_oracle as O
from transoar_b200 import MultiScaleDeformableAttention as MSDA
from transoar_b200 import _lib, synth
from transoar_b200.ops.functions import MSDeformAttnFunction
from transoar_b200.ops.modules import MSDeformAttn

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _np(t):
    return t.detach().cpu().numpy()


def _geom(name, **kw):
    g = synth.GEOMETRIES[name]
    return synth.Geometry(**{**g.__dict__, **kw}) if kw else g


def _run_fwd(x, step=64):
    return MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], step)


def _run_bwd(x, step=64):
    return MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], x["grad_out"], step)


def _oracle(x, dt=np.float32):
    v, loc, aw, go = (_np(x[k].float() if dt == np.float32 else x[k]).astype(dt) for k in ("value", "loc", "aw", "grad_out"))
    sh, st = _np(x["shapes"]), _np(x["starts"])
    out = O.forward(v, sh, st, loc, aw)
    gv, gl, ga = O.backward(go, v, sh, st, loc, aw)
    return out, gv, gl, ga


def _relerr(got, want):
    return float(np.abs(got - want).max() / max(float(np.abs(want).max()), 1e-30))


# geometry name -> overrides; covers every vector-kernel group width, tails (L*P not a multiple of G), and the generic path
SHAPES = {
    "small": ("test_small", {}),                                              # ops/test.py "Small": C=4 -> G=1
    "tiny": ("test_tiny", {}),                                                # C=1 -> generic kernel
    "medium": ("test_medium", {"queries": 600}),                              # M=16, C=16 -> G=4
    "c64_l4": ("visceral_refine", {"shapes": ((6, 6, 8), (3, 3, 4), (2, 2, 2), (1, 1, 1)), "queries": 0}),   # C=64 -> G=16, LP=16
    "c64_l3": ("amos_refine", {"shapes": ((4, 4, 2), (2, 2, 1), (1, 1, 1)), "queries": 77}),                 # LP=12 < G: tail lanes
    "c32_p5": ("config1", {"shapes": ((5, 4, 3),), "points": 5, "queries": 50}),                             # G=8, LP=5
    "c128": ("config1", {"shapes": ((3, 3, 3), (2, 2, 2)), "channels": 128, "heads": 2, "points": 9, "queries": 31}),  # G=32, LP=18 > 16
    "c256": ("config1", {"shapes": ((3, 3, 3),), "channels": 256, "heads": 1, "points": 2, "queries": 9}),   # G=32, NV=2
    "c8": ("config1", {"shapes": ((3, 2, 2), (1, 1, 1)), "channels": 8, "heads": 6, "queries": 13}),         # G=2
    "c24_generic": ("config1", {"shapes": ((3, 3, 3),), "channels": 24, "heads": 3, "queries": 20}),          # 24/4 = 6 lanes: generic
    "c65_generic": ("config1", {"shapes": ((2, 3, 2),), "channels": 65, "heads": 2, "queries": 5}),
    "l17_generic": ("config1", {"shapes": tuple((1, 1, 2) for _ in range(17)), "channels": 16, "heads": 2, "points": 1, "queries": 6}),
}


def _inputs(key, batch=2, dist="A", seed=3, dtype=torch.float32, widen=False):
    base, kw = SHAPES[key]
    x = synth.make_inputs(_geom(base, **kw), batch, dist, seed=seed, device=DEV, dtype=dtype)
    if widen:   # push ~25 % of the samples outside [0,1] to hit the padding / range-test branches
        x["loc"] = (x["loc"] * 1.5 - 0.25).contiguous()
    return x


@pytest.mark.parametrize("key", list(SHAPES))
@pytest.mark.parametrize("widen", [False, True])
def test_forward_fp32_is_bit_identical_to_oracle(key, widen):
    x = _inputs(key, widen=widen)
    out = _np(_run_fwd(x))
    want = _oracle(x)[0]
    assert np.abs(out - want).max() <= 1e-4 * max(1.0, np.abs(want).max())
    assert np.array_equal(out, want), f"max diff {np.abs(out - want).max():.3e}"


@pytest.mark.parametrize("key", list(SHAPES))
@pytest.mark.parametrize("widen", [False, True])
def test_backward_fp32_matches_oracle(key, widen):
    x = _inputs(key, widen=widen)
    gv, gl, ga = (_np(t) for t in _run_bwd(x))
    _, wv, wl, wa = _oracle(x)
    assert _relerr(gv, wv) < 1e-4 and _relerr(gl, wl) < 1e-4 and _relerr(ga, wa) < 1e-4
    # samples that fail the range test get exactly zero gradient (cuh:618-621)
    assert np.array_equal(gl == 0, wl == 0) or _relerr(gl, wl) < 1e-6


@pytest.mark.parametrize("key", ["small", "c64_l4", "c65_generic", "c128"])
def test_fp64_matches_oracle(key):
    x = _inputs(key, dtype=torch.float64, widen=True)
    out = _np(_run_fwd(x))
    gv, gl, ga = (_np(t) for t in _run_bwd(x))
    wo, wv, wl, wa = _oracle(x, np.float64)
    assert np.array_equal(out, wo)
    assert _relerr(gv, wv) < 1e-12 and _relerr(gl, wl) < 1e-12 and _relerr(ga, wa) < 1e-12


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_sampling_index_arithmetic_is_bit_exact(dtype):
    """idx = (in_range, d_low, h_low, w_low) and the three fractions, bit for bit, incl. knife-edge and out-of-range samples."""
    g = _geom("visceral_refine", queries=4096)
    x = synth.make_inputs(g, 1, "B", seed=11, device=DEV, dtype=dtype)
    loc = x["loc"]
    # add exact voxel-boundary locations: k / size and (k + 0.5) / size land on integer / half-integer pixel coordinates
    k = torch.arange(loc[..., 0].numel(), device=DEV).reshape(loc.shape[:-1]) % 41
    loc[:, ::7, :, :, :, 0] = (k[:, ::7].to(dtype) / 64)
    loc[:, 1::7, :, :, :, 1] = ((k[:, 1::7].to(dtype) + 0.5) / 40)
    loc[:, 2::7, :, :, :, 2] = -0.5 / 40
    N, Lq, M, L, P, _ = loc.shape
    idx = torch.empty(N, Lq, M, L, P, 4, dtype=torch.int32, device=DEV)
    frac = torch.empty(N, Lq, M, L, P, 3, dtype=dtype, device=DEV)
    rc = _lib.lib().msda3d_debug_indices(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
                                         _lib.F64 if dtype == torch.float64 else _lib.F32,
                                         ctypes.c_void_p(x["shapes"].data_ptr()), ctypes.c_void_p(loc.data_ptr()),
                                         N, M, L, Lq, P, ctypes.c_void_p(idx.data_ptr()), ctypes.c_void_p(frac.data_ptr()))
    _lib.check(rc, "msda3d_debug_indices")
    widx, wfrac = O.indices(_np(x["shapes"]), _np(loc))
    assert np.array_equal(_np(idx)[..., 0], widx[..., 0])
    inr = widx[..., 0] == 1
    assert np.array_equal(_np(idx)[inr], widx[inr])
    assert np.array_equal(_np(frac)[inr].view(np.uint64 if dtype == torch.float64 else np.uint32),
                          wfrac[inr].view(np.uint64 if dtype == torch.float64 else np.uint32))
    assert 0.5 < inr.mean() < 1.0


@pytest.mark.parametrize("case", ["small", "tiny", "border", "heads6"])
def test_against_golden_fixtures_from_reference_python_path(golden_cases, case):
    """Same checks as the reference's own test (ops/test.py:69-97) with its tolerances, on the committed fixtures."""
    z = golden_cases[case]
    for tag, dt, tol in (("f64", torch.float64, dict(rtol=1e-5, atol=1e-8)), ("f32", torch.float32, dict(rtol=1e-2, atol=1e-3))):
        x = {k: torch.from_numpy(z[k]).to(DEV) for k in ("shapes", "starts")}
        for k in ("value", "loc", "aw", "grad_out"):
            x[k] = torch.from_numpy(z[k]).to(dt).to(DEV).contiguous()
        out = _run_fwd(x, 2).cpu()
        assert torch.allclose(out, torch.from_numpy(z[f"out_{tag}"]), **tol)
        gv, gl, ga = (t.cpu() for t in _run_bwd(x, 2))
        assert torch.allclose(gv, torch.from_numpy(z[f"grad_value_{tag}"]), **tol)
        assert torch.allclose(gl, torch.from_numpy(z[f"grad_loc_{tag}"]), **tol)
        assert torch.allclose(ga, torch.from_numpy(z[f"grad_aw_{tag}"]), **tol)
        if tag == "f32":   # far tighter than the reference asks
            assert _relerr(_np(out), z["out_f32"]) < 1e-5 and _relerr(_np(gl), z["grad_loc_f32"]) < 1e-4


@pytest.mark.skipif(not O.refcuda_available(), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("key", ["small", "medium", "c64_l4", "c64_l3", "c128", "c65_generic"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_against_the_references_own_compiled_cuda_op(key, dtype):
    x = _inputs(key, dtype=dtype, widen=True)
    out = _run_fwd(x)
    ref = O.refcuda_forward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"])
    assert torch.equal(out, ref), f"forward differs from the compiled reference: {(out - ref).abs().max():.3e}"
    gv, gl, ga = _run_bwd(x)
    rv, rl, ra = O.refcuda_backward(x["grad_out"], x["value"], x["shapes"], x["starts"], x["loc"], x["aw"])
    tol = 1e-4 if dtype == torch.float32 else 1e-12
    assert _relerr(_np(gv), _np(rv)) < tol and _relerr(_np(gl), _np(rl)) < tol and _relerr(_np(ga), _np(ra)) < tol


@pytest.mark.parametrize("channels", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 32, 64, 65, 66, 67, 68, 69, 70, 71, 128])
def test_gradcheck_fp64_like_reference_test(channels):
    """ops/test.py:100-123: torch.autograd.gradcheck on MSDeformAttnFunction.apply in fp64, Small preset, per channel count."""
    g = _geom("test_small", channels=channels)
    x = synth.make_inputs(g, 2, "A", seed=channels, device=DEV, dtype=torch.float64)
    v, loc, aw = (x[k].clone().requires_grad_(True) for k in ("value", "loc", "aw"))
    assert torch.autograd.gradcheck(MSDeformAttnFunction.apply, (v, x["shapes"], x["starts"], loc, aw, 2))


@pytest.mark.parametrize("channels", [256, 1024, 1025, 2048, 2049])
def test_wide_channel_counts_fp64(channels):
    """The remaining channel counts of ops/test.py:122 (its >1024-channel dispatch branches); gradcheck would need
    millions of forward calls, so the analytic gradient is compared with the oracle instead."""
    g = _geom("test_small", channels=channels)
    x = synth.make_inputs(g, 2, "A", seed=channels, device=DEV, dtype=torch.float64)
    wo, wv, wl, wa = _oracle(x, np.float64)
    assert np.array_equal(_np(_run_fwd(x, 2)), wo)
    gv, gl, ga = (_np(t) for t in _run_bwd(x, 2))
    assert _relerr(gv, wv) < 1e-12 and _relerr(gl, wl) < 1e-12 and _relerr(ga, wa) < 1e-12


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("key", ["c64_l4", "c128", "c8", "c24_generic", "medium"])
def test_16bit_value_with_fp32_locations(dtype, key):
    """16-bit value / grad_output with fp32 loc / aw.  Oracle runs on the same rounded inputs in fp32."""
    x = _inputs(key, dtype=dtype, widen=True)
    assert x["loc"].dtype == torch.float32
    out = _run_fwd(x)
    assert out.dtype == dtype
    gv, gl, ga = _run_bwd(x)
    assert gv.dtype == dtype and gl.dtype == torch.float32 and ga.dtype == torch.float32
    wo, wv, wl, wa = _oracle(x)
    assert _relerr(_np(out.float()), wo) < 1e-2
    assert _relerr(_np(gv.float()), wv) < 1e-2 and _relerr(_np(gl), wl) < 1e-4 and _relerr(_np(ga), wa) < 1e-4


def test_autograd_function_surface_and_streams():
    x = _inputs("c64_l3")
    v, loc, aw = (x[k].clone().requires_grad_(True) for k in ("value", "loc", "aw"))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        out = MSDeformAttnFunction.apply(v, x["shapes"], x["starts"], loc, aw, 64)
        out.backward(x["grad_out"])
    side.synchronize()
    wo, wv, wl, wa = _oracle(x)
    assert out.shape == (2, 77, 6 * 64)
    assert np.array_equal(_np(out), wo)
    assert _relerr(_np(v.grad), wv) < 1e-4 and _relerr(_np(loc.grad), wl) < 1e-4 and _relerr(_np(aw.grad), wa) < 1e-4


def test_error_behaviour_on_device():
    x = _inputs("small", batch=3)
    with pytest.raises(RuntimeError, match="must divide im2col_step"):           # ms_deform_attn_cuda.cu:52
        _run_fwd(x, step=2)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        MSDA.ms_deform_attn_forward(x["value"], x["shapes"].cpu(), x["starts"], x["loc"], x["aw"], 64)
    with pytest.raises(RuntimeError, match="sampling_loc tensor has to be contiguous"):
        MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["starts"], x["loc"].transpose(1, 2), x["aw"], 64)


def test_host_buffer_entry_points_match_device_path():
    x = _inputs("c64_l3")
    want_out = _np(_run_fwd(x))
    want = [_np(t) for t in _run_bwd(x)]
    h = {k: x[k].cpu().pin_memory() for k in x}
    out = torch.empty(want_out.shape).pin_memory()
    gv, gl, ga = torch.empty(want[0].shape).pin_memory(), torch.empty(want[1].shape).pin_memory(), torch.empty(want[2].shape).pin_memory()
    N, S, M, C = x["value"].shape
    _, Lq, _, L, P, _ = x["loc"].shape
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = _lib.lib().msda3d_forward_backward_host(0, _lib.F32, p(h["grad_out"]), p(h["value"]), p(h["shapes"]), p(h["starts"]),
                                                 p(h["loc"]), p(h["aw"]), N, S, M, C, L, Lq, P, p(out), p(gv), p(gl), p(ga))
    _lib.check(rc, "msda3d_forward_backward_host")
    assert np.array_equal(out.numpy(), want_out)
    assert _relerr(gv.numpy(), want[0]) < 1e-5 and np.array_equal(gl.numpy(), want[1]) and np.array_equal(ga.numpy(), want[2])
    out.zero_()
    rc = _lib.lib().msda3d_forward_host(0, _lib.F32, p(h["value"]), p(h["shapes"]), p(h["starts"]), p(h["loc"]), p(h["aw"]),
                                        N, S, M, C, L, Lq, P, p(out))
    _lib.check(rc, "msda3d_forward_host")
    assert np.array_equal(out.numpy(), want_out)
    _lib.lib().msda3d_host_release()


def test_module_forward_backward_against_gridsample_composition():
    """MSDeformAttn (module mirror) vs the same projections + the restated grid_sample route, fp32 on the GPU."""
    torch.manual_seed(0)
    shapes_py = [(6, 5, 4), (3, 3, 2), (2, 1, 1)]
    m = MSDeformAttn(d_model=96, n_levels=3, n_heads=6, n_points=4).to(DEV)
    with torch.no_grad():   # non-trivial offsets / attention logits
        m.sampling_offsets.weight.normal_(0, 0.05)
        m.attention_weights.weight.normal_(0, 0.2)
    shapes, starts = synth.level_tensors(shapes_py, DEV)
    S = int(shapes.prod(1).sum())
    src = torch.randn(2, S, 96, device=DEV, requires_grad=True)
    ref_pts = synth.reference_points(shapes_py, DEV)[None, :, None, :].expand(1, S, 3, 3).contiguous()
    out = m(src, ref_pts, src, shapes, starts)
    g = torch.randn_like(out)
    out.backward(g)
    got = [out.detach().clone(), src.grad.clone()] + [p.grad.clone() for p in m.parameters()]
    src.grad = None
    m.zero_grad()
    # composition with the oracle's grid_sample restatement
    N, M, L, P, C = 2, 6, 3, 4, 16
    value = m.value_proj(src).view(N, S, M, C)
    off = m.sampling_offsets(src).view(N, S, M, L, P, 3)
    aw = torch.softmax(m.attention_weights(src).view(N, S, M, L * P), -1).view(N, S, M, L, P)
    loc = ref_pts[:, :, None, :, None, :] + off / shapes.flip(-1)[None, None, None, :, None, :]
    want_out = m.output_proj(O.gridsample_path(value, shapes_py, loc, aw))
    want_out.backward(g)
    want = [want_out.detach(), src.grad] + [p.grad for p in m.parameters()]
    for a, b in zip(got, want):
        assert _relerr(_np(a), _np(b)) < 1e-4


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties + the compiled reference
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def visceral():
    return {d: synth.make_inputs(synth.GEOMETRIES["visceral_refine"], 1, d, seed=1234, device=DEV) for d in ("A", "B")}


@pytest.mark.parametrize("dist", ["A", "B"])
def test_full_size_linearity_and_adjoint_identities(visceral, dist):
    x = visceral[dist]
    out = _run_fwd(x)
    # linear in value: f(2.5 v) = 2.5 f(v)  (power-of-two-free scale -> tolerance, not equality)
    y = dict(x, value=(x["value"] * 2.5).contiguous())
    assert _relerr(_np(_run_fwd(y)), 2.5 * _np(out)) < 1e-6
    # adjoints: <gOut, f> = <gV, v> (f linear in v) = <gAw, aw> (f linear in aw)
    gv, gl, ga = _run_bwd(x)
    lhs = float((x["grad_out"].double() * out.double()).sum())
    assert abs(float((gv.double() * x["value"].double()).sum()) - lhs) < 1e-5 * abs(lhs) + 1e-9
    assert abs(float((ga.double() * x["aw"].double()).sum()) - lhs) < 1e-5 * abs(lhs) + 1e-9
    # partition of unity: a constant volume sampled strictly inside returns the constant times the in-range weight mass
    ones = dict(x, value=torch.ones_like(x["value"]))
    px = x["loc"] * x["shapes"].flip(-1)[None, None, None, :, None, :].float() - 0.5
    inside = ((px >= 0) & (px <= (x["shapes"].flip(-1)[None, None, None, :, None, :] - 1))).all(-1)
    fully = inside.all(-1).all(-1)                       # units whose every sample is interior
    o1 = _run_fwd(ones).view(1, -1, 6, 64)
    mass = x["aw"].sum((-1, -2))
    assert fully.float().mean() > 0.005
    assert torch.allclose(o1[fully], mass[fully][:, None].expand(-1, 64), atol=2e-6)


@pytest.mark.skipif(not O.refcuda_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("dist", ["A", "B"])
def test_full_size_against_compiled_reference(visceral, dist):
    x = visceral[dist]
    out = _run_fwd(x)
    ref = O.refcuda_forward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"])
    assert torch.equal(out, ref)
    gv, gl, ga = _run_bwd(x)
    rv, rl, ra = O.refcuda_backward(x["grad_out"], x["value"], x["shapes"], x["starts"], x["loc"], x["aw"])
    assert _relerr(_np(gv), _np(rv)) < 1e-4 and _relerr(_np(gl), _np(rl)) < 1e-4 and _relerr(_np(ga), _np(ra)) < 1e-4


def test_batch_elements_are_independent_and_one_launch_covers_the_batch():
    g = synth.GEOMETRIES["amos_refine"]
    x = synth.make_inputs(g, 4, "B", seed=9, device=DEV)
    before = _lib.lib().msda3d_launch_count()
    out = _run_fwd(x, step=2)                             # the reference would loop twice (ms_deform_attn_cuda.cu:56)
    assert _lib.lib().msda3d_launch_count() == before + 1
    for b in range(4):
        xb = {k: (v[b:b + 1].contiguous() if k in ("value", "loc", "aw", "grad_out") else v) for k, v in x.items()}
        assert torch.equal(_run_fwd(xb), out[b:b + 1])


# ---------------------------------------------------------------------------------------------------------------
# Module-level rows: mirrors loaded with the REFERENCE modules' weights vs the reference's outputs
# (fixtures from tests/golden/make_golden_blocks.py, reference run on CPU through its use_cuda=False route)
# ---------------------------------------------------------------------------------------------------------------
def _load(name):
    import os
    from conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, name))


def test_msdeformattn_module_against_reference_module_fixture():
    z = _load("module_msdeformattn.npz")
    m = MSDeformAttn(d_model=48, n_levels=2, n_heads=6, n_points=2).to(DEV).eval()
    m.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")})
    t = lambda k: torch.from_numpy(z[k]).to(DEV)
    query, src = t("query").requires_grad_(True), t("src").requires_grad_(True)
    out = m(query, t("ref"), src, t("shapes"), t("starts"))
    out.backward(t("g"))
    assert _relerr(_np(out), z["out"]) < 1e-4
    assert _relerr(_np(query.grad), z["grad_query"]) < 1e-4 and _relerr(_np(src.grad), z["grad_src"]) < 1e-4
    for k, p in m.named_parameters():
        assert _relerr(_np(p.grad), z["pg." + k]) < 1e-4, k


def test_refine_block_against_reference_block_fixture():
    from transoar_b200.position_encoding import PositionEmbeddingSine3D
    from transoar_b200.refine import DecoderDefAttnBlock
    z = _load("block_defattn.npz")
    blk = DecoderDefAttnBlock(d_model=48, nhead=6, num_layers=2, dim_feedforward=64, dropout=0.1,
                              feature_levels=["P2", "P3", "P4"], n_points=2).to(DEV).eval()
    blk.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")})
    fmaps = [torch.from_numpy(z[f"fmap{i}"]).to(DEV).requires_grad_(True) for i in range(3)]
    pe = PositionEmbeddingSine3D(channels=48)
    outs = blk(fmaps, [pe(f) for f in fmaps])
    sum((o * torch.from_numpy(z[f"g{i}"]).to(DEV)).sum() for i, o in enumerate(outs)).backward()
    for i in range(3):
        assert outs[i].shape == fmaps[i].shape
        assert _relerr(_np(outs[i]), z[f"out{i}"]) < 1e-4
        assert _relerr(_np(fmaps[i].grad), z[f"grad_fmap{i}"]) < 2e-4
    for k, p in blk.named_parameters():
        assert _relerr(_np(p.grad), z["pg." + k]) < 2e-4, k


@pytest.mark.parametrize("geom", [((8, 8, 16), (4, 4, 8), (2, 2, 4), (1, 1, 2)), ((6, 5, 7), (3, 3, 4)), ((4, 4, 4),)])
@pytest.mark.parametrize("ref_batch", [1, 2])
def test_fused_prologue_equals_the_three_step_prologue_plus_op(geom, ref_batch):
    """msda3d_*_fused against softmax -> ref + off / (W,H,D) -> MSDeformAttnFunction on the same raw offsets / logits:
    the sampled voxels are the same (identical location arithmetic), the softmax differs in the last ulp."""
    from transoar_b200.ops.functions import MSDeformAttnFusedFunction
    N, M, C, P = 2, 6, 64, 16 // len(geom) if len(geom) != 3 else 5
    L = len(geom)
    P = 16 // L
    gen = torch.Generator().manual_seed(L * 10 + ref_batch)
    shapes = torch.tensor(geom, dtype=torch.long)
    S = int(shapes.prod(1).sum())
    starts = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    Lq = S
    value = (torch.randn(N, S, M, C, generator=gen)).to(DEV).requires_grad_(True)
    ref = torch.rand(ref_batch, Lq, L, 3, generator=gen).to(DEV)
    off = (torch.randn(N, Lq, M, L, P, 3, generator=gen) * 2.0).to(DEV).requires_grad_(True)       # in voxels: some samples leave the volume
    logit = torch.randn(N, Lq, M, L, P, generator=gen).to(DEV).requires_grad_(True)
    g = torch.randn(N, Lq, M * C, generator=gen).to(DEV)
    shapes_d, starts_d = shapes.to(DEV), starts.to(DEV)
    out = MSDeformAttnFusedFunction.apply(value, shapes_d, starts_d, ref, off, logit)
    out.backward(g)
    got = (out.detach(), value.grad.clone(), off.grad.clone(), logit.grad.clone())
    value.grad = off.grad = logit.grad = None
    aw = torch.softmax(logit.view(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
    loc = ref[:, :, None, :, None, :] + off / shapes_d.flip(-1)[None, None, None, :, None, :]
    want = MSDeformAttnFunction.apply(value, shapes_d, starts_d, loc.contiguous(), aw, 64)
    want.backward(g)
    assert _relerr(_np(got[0]), _np(want)) < 2e-6
    assert _relerr(_np(got[1]), _np(value.grad)) < 1e-5
    assert _relerr(_np(got[2]), _np(off.grad)) < 2e-5
    assert _relerr(_np(got[3]), _np(logit.grad)) < 2e-5


def test_module_takes_the_fused_route_and_matches_the_unfused_one():
    from transoar_b200 import _lib
    torch.manual_seed(3)
    m = MSDeformAttn(d_model=384, n_levels=4, n_heads=6, n_points=4).to(DEV)
    with torch.no_grad():
        m.sampling_offsets.weight.normal_(0, 0.02)
        m.attention_weights.weight.normal_(0, 0.05)
    shapes = torch.tensor(((8, 8, 12), (4, 4, 6), (2, 2, 3), (1, 1, 2)), dtype=torch.long, device=DEV)
    starts = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    x = torch.randn(2, S, 384, device=DEV)
    ref = torch.rand(1, S, 4, 3, device=DEV)
    outs = []
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        for fused, tf32 in ((True, False), (False, False), (True, True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            m.fuse_prologue = fused
            m.zero_grad()
            xq = x.clone().requires_grad_(True)
            n0 = _lib.lib().msda3d_launch_count()
            y = m(xq, ref, xq, shapes, starts)
            y.square().mean().backward()
            outs.append((y.detach(), xq.grad.clone(), m.sampling_offsets.weight.grad.clone(), m.attention_weights.bias.grad.clone(),
                         _lib.lib().msda3d_launch_count() - n0))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    assert outs[0][4] == outs[1][4] == 2                                      # strict fp32: op forward + backward only, fused or not
    # TF32: value / merged offsets+logits / output projections x (fwd + 2 gradient GEMMs), + 2 column-sum kernels per bias gradient
    assert outs[2][4] == 2 + 3 * 3 + 3 * 2, outs[2][4]
    for a, b in zip(outs[0][:4], outs[1][:4]):
        assert _relerr(_np(a), _np(b)) < 5e-5
    # merged-projection route under TF32: the forward agrees to TF32's accuracy; its gradients through the sampling locations are
    # piecewise constant in the offsets, so a 1e-3 perturbation moves a few samples across cell borders -- the exact-arithmetic check
    # of that route is test_fused_prologue_equals_the_three_step_prologue_plus_op, here only finiteness is asserted
    assert _relerr(_np(outs[2][0]), _np(outs[1][0])) < 3e-2
    assert all(bool(torch.isfinite(t).all()) for t in outs[2][1:4])


@pytest.mark.parametrize("ref_batch", [1, 2])
def test_merged_layout_of_the_fused_op_equals_the_dense_one(ref_batch):
    """msda3d_*_fused_ld: offsets and logits as columns of ONE [N*Lq, 4*M*L*P] tensor (the output of the concatenated projection)
    must give bit-identical results and gradients to the two-array form."""
    from transoar_b200.ops.functions import MSDeformAttnFusedFunction, MSDeformAttnMergedFunction
    geom = ((8, 8, 16), (4, 4, 8), (2, 2, 4), (1, 1, 2))
    N, M, C, L, P = 2, 6, 64, 4, 4
    gen = torch.Generator().manual_seed(17 + ref_batch)
    shapes = torch.tensor(geom, dtype=torch.long).to(DEV)
    S = int(shapes.prod(1).sum())
    starts = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    value = torch.randn(N, S, M, C, generator=gen).to(DEV).requires_grad_(True)
    ref = torch.rand(ref_batch, S, L, 3, generator=gen).to(DEV)
    merged = (torch.randn(N, S, 4 * M * L * P, generator=gen) * 1.5).to(DEV).requires_grad_(True)
    g = torch.randn(N, S, M * C, generator=gen).to(DEV)
    out = MSDeformAttnMergedFunction.apply(value, shapes, starts, ref, merged, L, P)
    out.backward(g)
    got = (out.detach().clone(), value.grad.clone(), merged.grad.clone())
    value.grad = None
    off = merged.detach()[..., :3 * M * L * P].reshape(N, S, M, L, P, 3).contiguous().requires_grad_(True)
    logit = merged.detach()[..., 3 * M * L * P:].reshape(N, S, M, L, P).contiguous().requires_grad_(True)
    want = MSDeformAttnFusedFunction.apply(value, shapes, starts, ref, off, logit)
    want.backward(g)
    assert torch.equal(got[0], want.detach())
    assert _relerr(_np(got[1]), _np(value.grad)) < 1e-5                          # atomics: order differs run to run
    assert torch.equal(got[2][..., :3 * M * L * P].reshape(off.shape), off.grad)
    assert torch.equal(got[2][..., 3 * M * L * P:].reshape(logit.shape), logit.grad)


@pytest.mark.parametrize("ctas", [2, 3, 4])
@pytest.mark.parametrize("widen", [False, True])
def test_tma_staged_forward_experiment_is_bit_identical(ctas, widen):
    """msda3d_set_tuning("stage", n): the coarsest level gathered from a shared-memory slab filled by TMA bulk copies
    (fwd_stage_kernel) must give exactly what the shipped LDG kernel gives -- it is an A/B experiment on the data path only."""
    g = synth.Geometry("stage", ((8, 8, 16), (4, 4, 8), (2, 2, 4), (1, 1, 2)), 6, 64, 4)       # brick order (queries = voxels), C = 64
    x = synth.make_inputs(g, 2, "B", seed=11, device=DEV)
    if widen:
        x["loc"] = (x["loc"] * 1.5 - 0.25).contiguous()
    want = _run_fwd(x)
    lib = _lib.lib()
    assert lib.msda3d_set_tuning(b"stage", ctas) == 0
    try:
        got = _run_fwd(x)
    finally:
        lib.msda3d_set_tuning(b"stage", 0)
    assert torch.equal(got, want)


@pytest.mark.parametrize("dist,jitter", [("B0", 0.0), ("B0", 0.05), ("B", 0.0), ("A", 0.0)])
@pytest.mark.parametrize("widen", [False, True])
def test_backward_pair_kernel_joint_and_single_paths(dist, jitter, widen):
    """bwd_duo_kernel (two w-neighbouring queries per lane group; the default for fp32 / 64 channels / brick order): on the model's
    sampling pattern (dist B0 = the untrained module's offsets, optionally with a small jitter) nearly every pair takes the joint
    same-cell / adjacent-cell path, with dist A (uniform locations) nearly every pair takes the one-unit path.  All of them must match
    the oracle to 1e-4, and grad_loc / grad_attn_weight must be BIT-identical to the one-unit kernel (msda3d_set_tuning("duo", 0)):
    the pair kernel shares gathers and reductions, not arithmetic."""
    g = synth.Geometry("duo", ((8, 8, 16), (4, 4, 8), (2, 2, 4), (1, 1, 2)), 6, 64, 4)       # brick order (queries = voxels), C = 64
    x = synth.make_inputs(g, 2, dist, seed=23, device=DEV)
    if jitter:
        gen = torch.Generator().manual_seed(5)
        norm = torch.tensor([[w, h, d] for d, h, w in g.shapes], dtype=torch.float32)
        x["loc"] = (x["loc"].cpu() + jitter * torch.randn(x["loc"].shape, generator=gen) / norm[None, None, None, :, None, :]).to(DEV).contiguous()
    if widen:
        x["loc"] = (x["loc"] * 1.5 - 0.25).contiguous()
    lib = _lib.lib()
    n0 = lib.msda3d_launch_count()
    gv, gl, ga = _run_bwd(x)
    assert lib.msda3d_launch_count() == n0 + 1
    _, wv, wl, wa = _oracle(x)
    assert _relerr(_np(gv), wv) < 1e-4 and _relerr(_np(gl), wl) < 1e-4 and _relerr(_np(ga), wa) < 1e-4
    assert lib.msda3d_set_tuning(b"duo", 0) == 0
    try:
        gv1, gl1, ga1 = _run_bwd(x)
    finally:
        lib.msda3d_set_tuning(b"duo", 1)
    assert torch.equal(gl, gl1) and torch.equal(ga, ga1)
    assert _relerr(_np(gv), _np(gv1)) < 2e-5
