"""The fp64 yardstick of the model-level fixtures (tests/golden/make_golden_model_fp64.py; VERDICT r01 "weak" 1).

The model-level GPU tests (tests/test_gpu_focused.py::test_transoarnet_against_reference_fixture) allow the encoder's parameter
gradients a far looser bound than everything else.  This file pins WHY from committed data alone: the reference's own fp32 CPU run
(transoarnet.npz) against the reference run in float64 (transoarnet_fp64.npz) is exact to 1e-6 in every output and to 1e-4 in
every gradient outside the encoder, and off by percents of the tensor's maximum in the encoder's gradients -- sums over up to
8.4 M voxels behind six InstanceNorm stages under a loss that sees 14 queries.  No accumulation order can do better than the
reference's own there, so the bound of those tensors is set from this yardstick, not from north_star's 1e-4."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def _pair(name):
    return np.load(os.path.join(GOLDEN, name + ".npz")), np.load(os.path.join(GOLDEN, name + "_fp64.npz"))


@pytest.mark.parametrize("name", ["attn_fpn", "transoarnet"])
def test_stored_yardstick_is_the_distance_between_the_two_committed_fixtures(name):
    z32, z64 = _pair(name)
    keys = [k for k in z64.files if not k.startswith("e32.")]
    assert keys and all(k in z32.files for k in keys)
    for k in keys:
        assert z64[k].dtype == np.float64 and z64[k].shape == z32[k].shape, k
        assert float(z64["e32." + k]) == pytest.approx(_rel(z32[k], z64[k]), rel=1e-12, abs=1e-300), k
    # every gradient the fp32 fixture holds has its fp64 twin
    assert sorted(k for k in z32.files if k.startswith("pg.")) == sorted(k for k in keys if k.startswith("pg."))


def test_reference_fp32_outputs_are_exact_and_encoder_gradients_are_not():
    z32, z64 = _pair("transoarnet")
    e = {k[4:]: float(z64[k]) for k in z64.files if k.startswith("e32.")}
    for k in ("pred_logits", "pred_boxes", "aux0_logits", "aux0_boxes"):
        assert e[k] < 1e-6, (k, e[k])
    enc = {k: v for k, v in e.items() if k.startswith("pg._backbone._encoder")}
    rest = {k: v for k, v in e.items() if k.startswith("pg.") and k not in enc}
    assert len(enc) >= 20 and len(rest) >= 40
    assert max(rest.values()) < 1e-4, max(rest.items(), key=lambda kv: kv[1])
    # the ill-conditioned ones: the reference's own fp32 run is percents of the tensor's maximum away from the exact gradient
    assert max(enc.values()) > 1e-2 and float(np.median(list(enc.values()))) > 1e-4
    # ... and still inside the bound the GPU test uses for them (tests/test_gpu_focused.py: 1e-1), with the room the triangle
    # inequality needs: |gpu - cpu32| <= |gpu - exact| + |cpu32 - exact|
    assert 2.0 * max(enc.values()) < 1e-1


def test_attn_fpn_fixture_is_well_conditioned_everywhere():
    """The small AttnFPN fixture (32x32x16 volume) has no such cancellation: the reference's fp32 run is within 2e-4 of fp64 in
    every tensor, which is what the GPU test's 2e-4 / 1e-3 / 2e-3 bounds sit on."""
    _, z64 = _pair("attn_fpn")
    e = {k[4:]: float(z64[k]) for k in z64.files if k.startswith("e32.")}
    assert max(v for k, v in e.items() if k.startswith("out.")) < 1e-5
    assert e["grad_x"] < 1e-4
    assert max(v for k, v in e.items() if k.startswith("pg.")) < 2e-4
