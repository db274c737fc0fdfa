"""Host-side mirror of the reference interface: error behaviour, module surface, synthetic input generators."""
import math
import os

import pytest
import torch

import transoar_b200
from transoar_b200 import MultiScaleDeformableAttention as MSDA
from transoar_b200 import synth
from transoar_b200.ops.functions import MSDeformAttnFunction
from transoar_b200.ops.modules import MSDeformAttn


def _cpu_inputs():
    g = synth.GEOMETRIES["test_small"]
    return synth.make_inputs(g, batch=2, dist="A", seed=1)


def test_cpu_tensors_raise_like_the_reference_dispatcher():
    x = _cpu_inputs()
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):        # ms_deform_attn.h:38
        MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], 2)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):        # ms_deform_attn.h:60
        MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], x["grad_out"], 2)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDeformAttnFunction.apply(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], 2)


def test_non_contiguous_inputs_raise():
    x = _cpu_inputs()
    bad = x["value"].transpose(1, 2)
    with pytest.raises(RuntimeError, match="value tensor has to be contiguous"):  # ms_deform_attn_cuda.cu:28
        MSDA.ms_deform_attn_forward(bad, x["shapes"], x["starts"], x["loc"], x["aw"], 2)


def test_no_python_fallback_is_exported():
    import transoar_b200.ops.functions.ms_deform_attn_func as f
    assert not hasattr(f, "ms_deform_attn_core_pytorch")
    m = MSDeformAttn(48, 2, 6, 2, use_cuda=False)
    q = torch.zeros(1, 9, 48)
    shapes, starts = synth.level_tensors([(2, 2, 2), (1, 1, 1)])
    ref = torch.zeros(1, 9, 2, 3)
    with pytest.raises(RuntimeError, match="use_cuda=True"):
        m(q, ref, q, shapes, starts)


def test_module_surface_matches_reference():
    m = MSDeformAttn(d_model=384, n_levels=4, n_heads=6, n_points=4)
    assert list(m.state_dict()) == ["sampling_offsets.weight", "sampling_offsets.bias", "attention_weights.weight",
                                    "attention_weights.bias", "value_proj.weight", "value_proj.bias",
                                    "output_proj.weight", "output_proj.bias"]
    assert sum(p.numel() for p in m.parameters()) == 443520
    assert m.im2col_step == 64 and m.use_cuda is True
    b = m.sampling_offsets.bias.view(6, 4, 4, 3)
    assert b.requires_grad
    # six face directions, scaled by (p+1), identical on every level (ms_deform_attn.py:63-79)
    assert torch.equal(b[:, 0, 0], torch.tensor([[-1., 0, 0], [0, -1, 0], [0, 0, -1], [0, 0, 1], [0, 1, 0], [1, 0, 0]]))
    assert torch.equal(b[:, 2, 3], 4 * b[:, 0, 0])
    assert float(m.sampling_offsets.weight.abs().max()) == 0.0 and float(m.attention_weights.bias.abs().max()) == 0.0
    with pytest.raises(ValueError, match="Only nheads of value 26 or 6"):        # ms_deform_attn.py:72-73
        MSDeformAttn(64, 1, 4, 1)
    with pytest.raises(ValueError, match="divisible"):
        MSDeformAttn(100, 1, 6, 1)


def test_geometries_match_survey_section8():
    g = synth.GEOMETRIES["visceral_refine"]
    assert (g.spatial_size, g.levels, g.heads, g.channels, g.points) == (117000, 4, 6, 64, 4)
    a = synth.GEOMETRIES["amos_refine"]
    assert (a.spatial_size, a.levels) == (18688, 3)
    shapes, starts = synth.level_tensors(g.shapes)
    assert starts.tolist() == [0, 102400, 115200, 116800]


def test_reference_points_are_voxel_centres_in_xyz_order():
    pts = synth.reference_points([(2, 3, 4)])
    assert pts.shape == (24, 3)
    assert torch.allclose(pts[0], torch.tensor([0.125, 1 / 6, 0.25]))            # (x=W, y=H, z=D)
    assert torch.allclose(pts[1], torch.tensor([0.375, 1 / 6, 0.25]))            # W is the fastest axis
    assert torch.allclose(pts[-1], torch.tensor([0.875, 5 / 6, 0.75]))


@pytest.mark.parametrize("dist", ["A", "B"])
def test_inputs_are_seeded_and_normalised(dist):
    g = synth.GEOMETRIES["test_medium"]
    a = synth.make_inputs(g, 1, dist, seed=5)
    b = synth.make_inputs(g, 1, dist, seed=5)
    for k in a:
        assert torch.equal(a[k], b[k])
    assert a["loc"].shape == (1, 4860, 16, 3, 4, 3) and a["loc"].dtype == torch.float32
    assert torch.allclose(a["aw"].sum((-1, -2)), torch.ones(1, 4860, 16), atol=1e-5)
    assert a["value"].is_contiguous() and a["loc"].is_contiguous()


def test_install_into_reference_requires_the_reference_package():
    try:
        import transoar  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            transoar_b200.install_into_reference()
    else:  # pragma: no cover - only when the reference happens to be importable
        mod = transoar_b200.install_into_reference()
        assert mod.MSDA is MSDA


def test_tuning_switches_and_their_environment_hook(monkeypatch):
    """include/msda3d.h msda3d_set_tuning: known keys / ranges are accepted, anything else is MSDA3D_EINVAL; TRANSOAR_B200_TUNING applies
    the switches when the package loads the library and rejects a typo loudly (no GPU needed: the switches are host state)."""
    from transoar_b200 import _lib
    lib = _lib.lib()
    for key, good, bad in ((b"duo", 0, 2), (b"rot", 6, 3), (b"duo_cfg", 2, 3), (b"roi_splits", 16, 17), (b"pair", 1, 2), (b"order", 2, 3)):
        assert lib.msda3d_set_tuning(key, good) == 0
        assert lib.msda3d_set_tuning(key, bad) != 0
        assert lib.msda3d_set_tuning(key, 1 if key == b"duo" else 0) == 0                  # back to the defaults
    assert lib.msda3d_set_tuning(b"no_such_switch", 1) != 0
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("TRANSOAR_B200_TUNING", "duo=0, rot=4")
    assert _lib.lib().msda3d_set_tuning(b"duo", 1) == 0 and _lib.lib().msda3d_set_tuning(b"rot", 0) == 0
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("TRANSOAR_B200_TUNING", "duoo=1")
    with pytest.raises(RuntimeError, match="TRANSOAR_B200_TUNING"):
        _lib.lib()
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.delenv("TRANSOAR_B200_TUNING")
    _lib.lib()


def test_only_the_checkers_import_the_oracle():
    """oracle/ is test infrastructure: nothing under transoar_b200/ (the product) or tools/install_reference.py may import it; the only
    importers are tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference / ref_* legs) and measurement tools."""
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def imports_oracle(path):
        with open(path) as f:
            tree = ast.parse(f.read(), path)
        for node in ast.walk(tree):
            if isinstance(node, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in node.names):
                return True
            if isinstance(node, ast.ImportFrom) and node.level == 0 and (node.module or "").split(".")[0] == "oracle":
                return True
        return False

    product = []
    for d, _, files in os.walk(os.path.join(root, "transoar_b200")):
        product += [os.path.join(d, f) for f in files if f.endswith(".py")]
    assert len(product) >= 25
    offenders = [os.path.relpath(p, root) for p in product if imports_oracle(p)]
    assert not offenders, offenders
    # ... and the library itself links nothing from oracle/
    import shutil
    import subprocess
    if shutil.which("readelf") is not None:
        needed = subprocess.run(["readelf", "-d", os.path.join(root, "transoar_b200", "libmsda3d.so")], capture_output=True, text=True).stdout
        assert "oracle" not in needed and "torch" not in needed and "libcudart" not in needed      # static CUDA runtime, no torch types
    assert imports_oracle(os.path.join(root, "bench.py")) and imports_oracle(os.path.join(root, "__graft_entry__.py"))
