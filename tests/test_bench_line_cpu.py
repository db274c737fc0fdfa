"""The bench line of the round's final GPU run (profiles/r04j_bench.json, printed by `python bench.py` on one B200) against the contract the
driver reads: required keys, internal consistency of the numbers, and the roofline arithmetic (SURVEY.md 8(d) bytes / measured time / measured
peak).  A committed artifact cannot prove the next run, but it pins what the line must keep carrying."""
import json
import os

import pytest

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINE = os.path.join(ROOT, "profiles", "r04j_bench.json")


@pytest.fixture(scope="module")
def line():
    with open(LINE) as f:
        return json.load(f)


def test_contract_keys(line):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in line, k
    assert line["metric"] == bench.METRIC == "ct_volumes_per_sec" and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["config"]["workload"] == "visceral_train_step" and "model" not in line["config"]
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert line["warmup"] >= 3


def test_numbers_are_consistent(line):
    batch = line["config"]["batch_per_gpu"]
    assert line["value"] == pytest.approx(1e3 * batch * line["n_gpus"] / line["ms_per_step"], rel=1e-6)
    # end to end copies every step's volumes from the host: batch x 160 x 160 x 256 fp32
    assert line["e2e"]["h2d_bytes_per_step"] == batch * 160 * 160 * 256 * 4 and line["e2e"]["d2h_bytes_per_step"] >= 4
    assert 0.9 * line["value"] <= line["e2e"]["value"] <= 1.001 * line["value"]
    # launches of this library's kernels: an integer number per step
    assert line["gpu_launches"] > 0 and line["gpu_launches"] % line["steps"] == 0
    assert line["gpu_launches"] // line["steps"] == int(line["kernels"]["launches_of_this_library_per_step"])
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert not bad & set(line["clocks"]["reasons"]) and line["clocks"]["sm_mhz"] >= 0.9 * line["clocks"]["sm_max_mhz"]


def test_roofline_is_algorithmic_bytes_over_measured_time_over_measured_peak(line):
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    _, bwd_bytes = bench.algorithmic_bytes(2, 117000, 6, 64, 4, 117000, 4)         # N = 2 volumes per launch
    assert bwd_bytes == 2156544000
    assert r["achieved"] == pytest.approx(bwd_bytes / (line["kernels"]["msda3d_bwd_ms"] * 1e-3) / 1e9, rel=1e-3)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-6)
    assert r["traffic"] is None or r["traffic"] >= 0.9 * bwd_bytes                  # measured DRAM bytes cannot be far below compulsory
    # the dominant kernel's two launches per step cannot exceed the step
    assert 2 * line["kernels"]["msda3d_bwd_ms"] < line["ms_per_step"]


def test_dominant_kernels_share_agrees_between_events_and_the_ncu_launch_list(line):
    """The contract's cross-check: ncu's per-launch times are cold-cache and serialised, so the roofline kernel's SHARE of the step -- not
    its absolute time -- must agree between the CUDA-event timing of the bench line and the launch list of the same code."""
    import csv
    rows = []
    with open(os.path.join(ROOT, "profiles", "r04j_launches.csv")) as f:
        for r in csv.reader(l for l in f if l.startswith('"')):
            rows.append(r)
    head, rows = rows[0], rows[1:]
    name, unit, val = head.index("Kernel Name"), head.index("Metric Unit"), head.index("Metric Value")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}
    ms = [(r[name], float(r[val].replace(",", "")) * scale[r[unit]]) for r in rows]
    total = sum(t for _, t in ms)
    duo = [t for n, t in ms if "bwd_duo_kernel" in n]
    assert len(ms) >= 700 and len(duo) == 2                               # one eager step: two refinement layers
    share_ncu = sum(duo) / total
    share_events = 2 * line["kernels"]["msda3d_bwd_ms"] / line["ms_per_step"]
    assert abs(share_ncu - share_events) < 0.02, (share_ncu, share_events)
    ours = sum(t for n, t in ms if any(ns in n for ns in ("msda3d::", "tcgemm::", "convgen::", "convtc::", "roiattn::", "instnorm::", "fusedln::",
                                                          "stemconv::", "crit::", "winattn::")))
    assert ours / total > 0.85                                            # the step is this library's kernels
    assert not any("cudnn" in n.lower() or "implicit_gemm" in n or "implicit_convolve" in n for n, _ in ms)      # no cuDNN convolution left in the step
