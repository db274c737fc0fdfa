#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout.

Metric      : CT volumes/s of a full training step at 160x160x256 (BASELINE.json `metric`, part 1), with the 3D multi-scale
              deformable attention kernels' achieved HBM GB/s as the `roofline` object (part 2).
Workload    : `visceral_train_step` = configs[1]: the Focused-Decoder model (AttnFPN backbone + deformable FPN refinement +
              Focused Decoder + heads, config/attn_fpn_foc_dec_visceral.yaml with use_decoder_attn / use_cuda on) on synthetic
              160x160x256 volumes, fp32 storage with TF32 tensor-core multiplies (the reference's torch 1.10 default), batch 2
              per GPU (yaml batch_size).  One step = forward + matcher + losses + backward + AdamW, nothing skipped.
Step        : one pass of the hot path over one batch, through `transoar_b200.engine.TrainStep.step`.
Multi-GPU   : one process per GPU, volumes sharded by rank (weak scaling); the only collective is DDP's NCCL gradient all-reduce.

`value`     : inputs (volumes, targets) resident in HBM before the timed region.
`e2e`       : the same call with the volumes in pinned HOST memory (H2D copy inside the step) and the loss read back every step.
`roofline`  : the dominant kernel of the step (msda3d backward) -- algorithmic bytes / CUDA-event time of its launches inside
              the timed region, against the measured HBM peak.
`msda3d_op` : the operator alone (forward + gradient of both refinement layers) incl. the reference's own CUDA op on the same
              inputs; `tc_gemm`, `roi_attention`, `instnorm_relu`, `conv3d` (every AttnFPN 3x3x3 layer against cuDNN): the other
              hand-written kernels of the path.
`ref_gpu_model`: the UNMODIFIED reference model (baseline/_ref) with its OWN compiled CUDA op (oracle/_ref) running the same
              training step on the same GPU in the same run -- the comparator BASELINE.json's north_star names.
`cpu_baseline` / `--impl reference`: the unmodified reference's use_cuda=False route of the same training step on the host
              cores (oracle/reference_model.py drives baseline/_ref), one real 160x160x256 volume per step, nothing extrapolated.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ct_volumes_per_sec"
UNIT = "volumes/s"
LAYERS = 2            # config/attn_fpn_foc_dec_visceral.yaml:83  `layers: 2`
BATCH = 2             # config/attn_fpn_foc_dec_visceral.yaml:25  `batch_size: 2`
VOLUME = (160, 160, 256)
GEOM = "visceral_refine"
DIST = "B"


# ---------------------------------------------------------------------------------------------------------------
# Algorithmic bytes (SURVEY.md 8d / DESIGN.md): compulsory HBM traffic of one msda3d launch
# ---------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(N, S, M, C, L, Lq, P, ev=4, eg=4):
    T = N * Lq * M * L * P
    val = min(N * S * M * C, 8 * T * C)
    out = N * Lq * M * C
    fwd = ev * (val + out) + eg * 4 * T
    bwd = ev * (val + out) + eg * 2 * val + eg * 4 * T + eg * 4 * T     # value + gOut reads, gV zero-fill + write-back, loc/aw, gLoc/gAw
    return fwd, bwd


def volume_ids(step, rank, world, batch=BATCH):
    """Volume sharding: global volume stream 0,1,2,...; step `step` hands `batch` consecutive volumes to every rank
    (rank-major inside the step), so ranks never share a volume and the union over ranks is contiguous."""
    base = (step * world + rank) * batch
    return list(range(base, base + batch))


def aggregate_throughput(local_ms, steps, world, batch=BATCH, all_reduce_max=None):
    """Whole-job volumes/s: every rank did `steps` steps of `batch` volumes; time = max over ranks of the device time."""
    ms = all_reduce_max(local_ms) if all_reduce_max is not None else local_ms
    return world * batch * steps / (ms / 1e3), ms


def load_ncu_traffic(kind):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (or None)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)[kind]
        return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            j = json.load(f)
        return {"hbm_gbs": float(j["hbm_gbs"]), "bf16_tflops": float(j.get("bf16_tflops", 1590.0)),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PFLOP/s)"}


# ---------------------------------------------------------------------------------------------------------------
# Clock sampling during the timed region
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def workload_config():
    from transoar_b200 import synth
    g = synth.GEOMETRIES[GEOM]
    return {"workload": "visceral_train_step", "volume": "x".join(map(str, VOLUME)), "batch_per_gpu": BATCH,
            "step": "forward + matcher + losses (cls/bbox/giou, aux layers) + backward + AdamW; nothing skipped",
            "precision": "fp32 tensors, TF32 tensor-core multiplies for convolutions and linears (torch 1.10 default of the reference)",
            "refine": {"layers": LAYERS, "levels": [list(s) for s in g.shapes], "S": g.spatial_size, "heads": g.heads,
                       "channels_per_head": g.channels, "points": g.points},
            "decoder": {"queries": 540, "organs": 20, "layers": 3, "kv_tokens": 102400},
            "l2": "the step touches > 20 GB per batch (activations of 629 MB each at full resolution): every tensor is far larger than the 126 MB L2; no explicit flush",
            "execution": "whole step captured once as a CUDA graph and replayed (transoar_b200.engine.TrainStep(graph=True)); --no-graph runs it eagerly",
            "parallelism": "volumes sharded over ranks; one NCCL gradient all-reduce per step (flat bucket inside the graph; DDP buckets when eager)"}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (baseline/_ref, installed by tools/install_reference.py) through its use_cuda=False route on the
# host cores -- transoar.models.transoarnet.TransoarNet + build_criterion + the AdamW step of scripts/train.py, at the REAL volume size
# (one 160x160x256 volume per step; no crop, no scaling).  If the reference is not installed, the port in oracle/model_oracle.py runs.
# ---------------------------------------------------------------------------------------------------------------
CPU_STEP_VOLUMES = 1          # volumes per CPU step: the reference's host step at batch 2 needs > 50 GB and a minute per step


class CpuArm:
    def __init__(self, threads):
        import torch
        torch.set_num_threads(threads)
        from oracle import reference_model as R
        self.threads = torch.get_num_threads()
        if R.available():
            self.kind = "reference"
            self.ts = R.ReferenceTrainStep("cpu", op=None, seed=0)
            self.targets = R.list_targets(self.ts.config, CPU_STEP_VOLUMES, 0, "cpu")
            self.what = ("UNMODIFIED reference (baseline/_ref: transoar.models.transoarnet.TransoarNet, build_criterion, AdamW of scripts/train.py:52-64, "
                         "step of trainer.py:50-87 without autocast) through its use_cuda=False route (ms_deform_attn_core_pytorch)")
        else:
            from oracle.model_oracle import CpuTrainStep
            from transoar_b200.engine import synthetic_targets, visceral_train_config
            cfg = visceral_train_config()
            self.kind = "port"
            self._port = CpuTrainStep(cfg, VOLUME, threads=threads)
            self.targets = synthetic_targets(cfg, CPU_STEP_VOLUMES, 0, "cpu")
            self.ts = None
            self.what = "port of the reference's use_cuda=False route (oracle/model_oracle.py; baseline/_ref not installed)"
        self.x = torch.rand(CPU_STEP_VOLUMES, 1, *VOLUME, generator=torch.Generator().manual_seed(1))

    def step(self):
        t0 = time.perf_counter()
        if self.ts is not None:
            self.ts.step(self.x, self.targets)
        else:
            self._port.step(self.x, self.targets)
        return time.perf_counter() - t0

    def close(self):
        if self.ts is not None:
            from oracle import reference_model as R
            R.leave_cpu_mode()
        self.ts = self._port = None

    def sample_text(self, times, warm):
        return (f"{self.what}; whole training step (fwd + matcher + losses + bwd + AdamW) on {CPU_STEP_VOLUMES} synthetic "
                f"{VOLUME[0]}x{VOLUME[1]}x{VOLUME[2]} volume per step at the real size (no crop, no extrapolation), {self.threads} threads: "
                f"{sum(times) / len(times):.2f} s per step (mean of {len(times)} timed step(s) after {warm} warm-up)")


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print to file descriptor 1 (NCCL's version banner, cuDNN notes) goes to stderr from here on; the JSON line
    is written to the real stdout by ``emit``, so stdout carries that one line and nothing else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def run_reference_arm(args):
    """`--impl reference`: rank 0 alone; K timed steps after W warm-up steps of the reference's CPU step at the real volume size.  A step
    takes 20-60 s of host time, so the run is bounded by a wall budget (TRANSOAR_REF_BUDGET_S, default 780 s): when K + W steps would not
    fit, warm-up is cut to one step first and then the timed steps; `steps` / `warmup` in the line are what was actually run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    budget = float(os.environ.get("TRANSOAR_REF_BUDGET_S", "780"))
    t_start = time.perf_counter()
    arm = CpuArm(os.cpu_count() or 1)
    first = arm.step()                                       # always a warm-up step (allocator, thread pool)
    warm_done, times = 1, []
    left = lambda: budget - (time.perf_counter() - t_start)
    want_warm = max(args.warmup, 1)
    if first * (want_warm - 1 + args.steps) > left():        # cannot afford the requested warm-up: keep the one already done
        want_warm = 1
    while warm_done < want_warm:
        arm.step()
        warm_done += 1
    while len(times) < args.steps and (len(times) < 2 or left() > 1.1 * max(times)):
        times.append(arm.step())
    sec = sum(times) / len(times)
    value = CPU_STEP_VOLUMES / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": warm_done, "steps_requested": args.steps, "warmup_requested": args.warmup, "ms_per_step": 1e3 * sec,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.threads, "kind": arm.kind, "sample": arm.sample_text(times, warm_done)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "volumes_per_step": CPU_STEP_VOLUMES, "wall_budget_s": budget,
        "note": "host-only arm: one process, all host cores, independent of --gpus (at N > 1 only rank 0 runs it)",
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------
# Extras: the hand-written kernels of the path one by one (reported beside the headline, never in it)
# ---------------------------------------------------------------------------------------------------------------
def _event_ms(fn, warm=2, reps=5):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def measure_msda_op(dev, rank, world, dist_name, with_ref):
    """The operator alone on the refinement workload (N=2, S=Lq=117000, 4 levels, 6 heads x 64, 4 points): per-launch times, HBM
    fractions, and the reference's own CUDA kernels (oracle/_ref) on the same inputs."""
    import torch
    from transoar_b200 import MultiScaleDeformableAttention as MSDA
    from transoar_b200 import synth
    g = synth.GEOMETRIES[GEOM]
    N, S, M, C, L, Lq, P = BATCH, g.spatial_size, g.heads, g.channels, g.levels, g.num_query, g.points
    layers = [synth.make_inputs(g, N, dist_name, seed=1234 + 7919 * volume_ids(0, rank, world)[0] + li, device=dev) for li in range(LAYERS)]
    ev = []

    def step(record):
        for x in layers:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], 64)
            e[1].record()
            MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], x["grad_out"], 64)
            e[2].record()
            if record:
                ev.append(e)

    for _ in range(3):
        step(False)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    t0.record()
    for _ in range(reps):
        step(True)
    t1.record()
    torch.cuda.synchronize()
    fwd = sum(e[0].elapsed_time(e[1]) for e in ev) / len(ev)
    bwd = sum(e[1].elapsed_time(e[2]) for e in ev) / len(ev)
    bf, bb = algorithmic_bytes(N, S, M, C, L, Lq, P)
    peak = load_peaks()["hbm_gbs"]
    out = {"what": f"operator alone: {LAYERS} layers x (forward + gradient), N={N}, S=Lq={S}, L={L}, M={M}, C={C}, P={P}, dist {dist_name}",
           "volumes_per_s": BATCH * reps / (t0.elapsed_time(t1) / 1e3), "ms_per_step": t0.elapsed_time(t1) / reps,
           "fwd_ms": fwd, "bwd_ms": bwd, "fwd_gbs": bf / fwd / 1e6, "bwd_gbs": bb / bwd / 1e6,
           "fwd_frac_of_hbm_peak": bf / fwd / 1e6 / peak, "bwd_frac_of_hbm_peak": bb / bwd / 1e6 / peak,
           "algorithmic_bytes": {"fwd": bf, "bwd": bb}}
    if with_ref:
        from oracle import msda3d_oracle as O
        if O.refcuda_available():
            x = layers[0]
            zeros = (torch.zeros_like(x["value"]), torch.zeros_like(x["loc"]), torch.zeros_like(x["aw"]))
            rf = _event_ms(lambda: O.refcuda_forward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"]), 2, 3)

            def ref_b():
                for z in zeros:
                    z.zero_()
                O.refcuda_backward(x["grad_out"], x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], out=zeros)

            rb = _event_ms(ref_b, 1, 3)
            out["ref_cuda_op"] = {"what": "the reference's own CUDA kernels (oracle/_ref, compiled from /root/reference for sm_100a), same inputs",
                                  "fwd_ms": rf, "bwd_ms": rb, "volumes_per_s": BATCH / (LAYERS * (rf + rb) * 1e-3),
                                  "speedup_fwd": rf / fwd, "speedup_bwd": rb / bwd}
    return out


def measure_kernels(dev, rank):
    import torch
    import torch.nn.functional as F
    from transoar_b200 import focused
    from transoar_b200.configs import visceral_config
    from transoar_b200.instnorm import instance_norm_relu
    from transoar_b200.linear import gemm
    peaks = load_peaks()
    out = {}
    gen = torch.Generator().manual_seed(7 + rank)
    cfg = visceral_config()

    # dense contractions: the FFN / projection GEMMs of one refinement layer at 2 x 117000 tokens (TF32 tcgen05 kernel vs cuBLAS TF32)
    T = BATCH * 117000
    tf32_peak = peaks["bf16_tflops"] / 2          # TF32 dense is half the bf16 rate on this part (1.1 vs 2.25 PFLOP/s nominal)
    rows = []
    for K, N in ((384, 384), (384, 1024), (1024, 384)):
        x = torch.randn(T, K, device=dev)
        w = torch.randn(N, K, device=dev) / K ** 0.5
        b = torch.randn(N, device=dev)
        dy = torch.randn(T, N, device=dev)
        y, dx, dw = torch.empty(T, N, device=dev), torch.empty(T, K, device=dev), torch.zeros(N, K, device=dev)
        fl = 2.0 * T * K * N / 1e12
        for name, ours, lib in (("fwd", lambda: gemm(x, 0, K, w, 0, K, y, T, N, K, bias=b), lambda: torch.addmm(b, x, w.t(), out=y)),
                                ("dgrad", lambda: gemm(dy, 0, N, w, 1, K, dx, T, K, N), lambda: torch.mm(dy, w, out=dx)),
                                ("wgrad", lambda: gemm(dy, 1, N, x, 1, K, dw, N, K, T, accumulate=True, split_k=0), lambda: torch.mm(dy.t(), x, out=dw))):
            a, c = _event_ms(ours, 2, 5), _event_ms(lib, 2, 5)
            rows.append({"gemm": f"{name} tokens={T} in={K} out={N}", "ms": a, "tflops": fl / a * 1e3, "frac_of_tf32_peak": fl / a * 1e3 / tf32_peak,
                         "cublas_tf32_ms": c})
        del x, w, dy, y, dx, dw
    out["tc_gemm"] = {"bound": "tensor", "peak_tflops": tf32_peak, "peak_source": peaks["source"] + ", bf16_tflops / 2 for TF32", "cases": rows}

    # RoI-restricted cross-attention of one Focused-Decoder layer, B = 2, 540 queries, P2 grid 40x40x64, 8 heads x 48
    grid = (40, 40, 64)
    boxes = focused.boxes_from_bbox_props(cfg["bbox_properties"], 540, grid)
    groups = focused.groups_from_boxes(boxes).to(dev)
    q = (torch.randn(BATCH, 540, 8, 48, generator=gen) * 0.3).to(dev).requires_grad_(True)
    k = torch.randn(BATCH, 102400, 8, 48, generator=gen).to(dev).requires_grad_(True)
    v = torch.randn(BATCH, 102400, 8, 48, generator=gen).to(dev).requires_grad_(True)
    g = torch.randn(BATCH, 540, 384, generator=gen).to(dev)

    def roi_fb():
        focused.RoIAttentionFunction.apply(q, k, v, groups, grid[1:]).backward(g)
        q.grad = k.grad = v.grad = None

    def roi_fb32():
        focused.RoIAttentionFunction.apply(q, k, v, groups, grid[1:], False).backward(g)
        q.grad = k.grad = v.grad = None

    with torch.no_grad():
        roi_f = _event_ms(lambda: focused.RoIAttentionFunction.apply(q, k, v, groups, grid[1:]))
        roi_f32 = _event_ms(lambda: focused.RoIAttentionFunction.apply(q, k, v, groups, grid[1:], False))
    vol = ((boxes[:, 3] - boxes[:, 0]) * (boxes[:, 4] - boxes[:, 1]) * (boxes[:, 5] - boxes[:, 2])).float()
    out["roi_attention"] = {"fwd_ms": roi_f, "fwd_bwd_ms": _event_ms(roi_fb), "kernels": "mma.sync m16n8k8 TF32 (what the step runs with TF32 on)",
                            "fp32_cuda_core_kernels": {"fwd_ms": roi_f32, "fwd_bwd_ms": _event_ms(roi_fb32)},
                            "batch": BATCH, "queries": 540, "kv_tokens": 102400,
                            "unmasked_kv_fraction": float(vol.mean()) / 102400,
                            "dense_score_tensor_avoided_gb": BATCH * 8 * 540 * 102400 * 4 / 1e9}
    del q, k, v, g

    # fused InstanceNorm3d + ReLU on the first encoder activation (B x 24 x 160 x 160 x 256 fp32)
    x = torch.randn(BATCH, 24, 160, 160, 256, device=dev, requires_grad=True)
    w = torch.ones(24, device=dev, requires_grad=True)
    b = torch.zeros(24, device=dev, requires_grad=True)
    dy = torch.randn_like(x)

    def in_fb():
        instance_norm_relu(x, w, b).backward(dy)
        x.grad = None

    def aten_fb():
        F.relu(F.instance_norm(x, weight=w, bias=b, eps=1e-5)).backward(dy)
        x.grad = None

    with torch.no_grad():
        in_f = _event_ms(lambda: instance_norm_relu(x, w, b))
    in_fbm, aten_fbm = _event_ms(in_fb), _event_ms(aten_fb, 1, 2)
    nbytes = x.numel() * 4
    out["instnorm_relu"] = {"fwd_ms": in_f, "fwd_bwd_ms": in_fbm, "aten_cudnn_fwd_bwd_ms": aten_fbm,
                            "fwd_gbs": 3 * nbytes / in_f / 1e6, "fwd_frac_of_hbm_peak": 3 * nbytes / in_f / 1e6 / peaks["hbm_gbs"],
                            "fwd_bwd_gbs": 8 * nbytes / in_fbm / 1e6, "fwd_bwd_frac_of_hbm_peak": 8 * nbytes / in_fbm / 1e6 / peaks["hbm_gbs"],
                            "bytes_model": "fwd: 2 reads + 1 write of the activation; bwd: 4 reads + 1 write"}
    del x, dy

    # the AttnFPN 3x3x3 convolutions of the step (batch 2 x 160x160x256) on the general tcgen05 kernels (include/conv3d_gen.h) against cuDNN
    # (autotuned, TF32) on the same tensors: forward, input gradient, weight gradient
    try:
        out["conv3d"] = measure_conv_layers(dev, tf32_peak)
    except Exception as exc:                                                                # keep the headline if an extra fails
        out["conv3d"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    torch.cuda.empty_cache()
    return out


CONV_LAYERS = [("enc1.conv1", 24, 48, 2, (160, 160, 256), False), ("enc1.conv2", 48, 48, 1, (80, 80, 128), False),
               ("enc2.conv1", 48, 96, 2, (80, 80, 128), False), ("enc2.conv2", 96, 96, 1, (40, 40, 64), False),
               ("enc3.conv1", 96, 192, 2, (40, 40, 64), False), ("enc3.conv2", 192, 192, 1, (20, 20, 32), False),
               ("enc4.conv1", 192, 384, 2, (20, 20, 32), False), ("enc4.conv2", 384, 384, 1, (10, 10, 16), False),
               ("enc5.conv1", 384, 768, 2, (10, 10, 16), False), ("enc5.conv2", 768, 768, 1, (5, 5, 8), False),
               ("out.P2", 96, 384, 1, (40, 40, 64), True), ("out.P3", 192, 384, 1, (20, 20, 32), True),
               ("out.P4", 384, 384, 1, (10, 10, 16), True), ("out.P5", 384, 384, 1, (5, 5, 8), True)]


def measure_conv_layers(dev, tf32_peak):
    """name, CI, CO, stride, input volume, bias of every 3x3x3 convolution of the VISCERAL AttnFPN besides the 1-channel stem and the 24 -> 24
    layer (their own kernels): encoder_blocks.py:28-46 stages 1-5, attn_fpn.py:65-74 output convolutions."""
    import ctypes
    import torch
    import torch.nn.functional as F
    from transoar_b200 import _lib
    from transoar_b200.conv3d_gen import fold_stride2_weights
    lib = _lib.lib()
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    cl = lambda t: t.contiguous(memory_format=torch.channels_last_3d)
    was = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    rows, tot = [], {k: [0.0, 0.0] for k in ("fwd", "dgrad", "wgrad")}
    try:
        for name, ci, co, s, (D, H, W), bias in CONV_LAYERS:
            N = BATCH
            x = cl(torch.randn(N, ci, D, H, W, device=dev))
            w = cl(torch.randn(co, ci, 3, 3, 3, device=dev) / (27 * ci) ** 0.5)
            wt = w.permute(2, 3, 4, 0, 1).reshape(27, co, ci).contiguous()
            b = torch.randn(co, device=dev) if bias else None
            od, oh, ow = ((v + s - 1) // s for v in (D, H, W))
            y, dy = cl(torch.empty(N, co, od, oh, ow, device=dev)), cl(torch.randn(N, co, od, oh, ow, device=dev))
            dx, dw = torch.empty_like(x), torch.empty_like(w)
            ours = {"fwd": lambda: lib.conv3d_gen_forward(st(), p(x), p(wt), p(b), N, D, H, W, ci, co, s, p(y)),
                    "dgrad": lambda: lib.conv3d_gen_dgrad(st(), p(dy), p(wt), N, D, H, W, ci, co, s, p(dx)),
                    "wgrad": lambda: lib.conv3d_gen_wgrad(st(), p(x), p(dy), N, D, H, W, ci, co, s, p(dw))}
            if s == 2 and ci <= 64:
                wf = fold_stride2_weights(wt)
                ours["dgrad"] = lambda: lib.conv3d_gen_dgrad_s2_folded(st(), p(dy), p(wf), N, D, H, W, ci, co, p(dx))
            cb = lambda mask: torch.ops.aten.convolution_backward(dy, x, w, None, [s] * 3, [1] * 3, [1] * 3, False, [0] * 3, 1, mask)
            lib_ = {"fwd": lambda: F.conv3d(x, w, b, s, 1), "dgrad": lambda: cb([True, False, False]), "wgrad": lambda: cb([False, True, False])}
            gf = 2.0 * N * od * oh * ow * 27 * ci * co / 1e9
            row = {"layer": name, "ci": ci, "co": co, "stride": s, "gflop": round(gf, 1)}
            for k in ("fwd", "dgrad", "wgrad"):
                if ours[k]() != 0:
                    raise RuntimeError(f"conv3d_gen {k} failed on {name}")
                a, c = _event_ms(ours[k], 2, 5), _event_ms(lib_[k], 2, 5)
                tot[k][0] += a
                tot[k][1] += c
                row[k] = {"ms": round(a, 4), "cudnn_ms": round(c, 4), "frac_of_tf32_peak": round(gf / a / tf32_peak, 3)}
            rows.append(row)
            del x, w, wt, y, dy, dx, dw
    finally:
        torch.backends.cudnn.benchmark = was
    return {"bound": "tensor", "peak_tflops": tf32_peak, "layers": rows,
            "total_ms": {k: {"ours": round(v[0], 3), "cudnn": round(v[1], 3)} for k, v in tot.items()}}


def measure_reference_gpu_model(dev, rank, steps=6, warm=3):
    """The comparator BASELINE.json's north_star names: the UNMODIFIED reference model (baseline/_ref) with its OWN compiled CUDA op
    (oracle/_ref, bound as `MSDA` -- oracle/reference_model.py) running the same training step on this GPU: visceral yaml with
    use_decoder_attn / use_cuda on, batch 2, same synthetic volumes and targets, fp32 + TF32 (torch 1.10's default, the reference's pin;
    no autocast: the reference op cannot run under it).  Three variants: cuDNN as scripts/train.py:113-114 sets it
    (benchmark off, deterministic on), cuDNN autotuned (the favourable setting; our arm autotunes too), and the same unmodified model
    on THIS repository's op through install_into_reference().  Every step ends with the reference's own `.item()` host read."""
    import torch
    from oracle import msda3d_oracle as O
    from oracle import reference_model as R
    if not R.available():
        return {"unavailable": "baseline/_ref not installed (python tools/install_reference.py in the build container)"}
    if not O.refcuda_available():
        return {"unavailable": "oracle/_ref/libmsda3d_refcuda.so not built"}
    saved = (torch.backends.cudnn.benchmark, torch.backends.cudnn.deterministic)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    gen = torch.Generator().manual_seed(100 + volume_ids(0, rank, 1)[0])
    vols = [torch.rand(BATCH, 1, *VOLUME, generator=gen).to(dev) for _ in range(2)]
    out = {"what": "unmodified reference TransoarNet + TransoarCriterion + AdamW (baseline/_ref) on cuda, visceral yaml with use_decoder_attn / use_cuda = True, "
                   f"batch {BATCH} x {VOLUME[0]}x{VOLUME[1]}x{VOLUME[2]}, fp32 + TF32, whole training step, CUDA events, {warm} warm-up + {steps} timed steps",
           "variants": {}}
    try:
        for label, op, bench_flag, det in (("reference_op_cudnn_as_shipped", "reference", False, True),
                                           ("reference_op_cudnn_autotuned", "reference", True, False),
                                           ("this_repo_op_in_reference_model", "ours", True, False)):
            torch.backends.cudnn.benchmark, torch.backends.cudnn.deterministic = bench_flag, det
            ts = R.ReferenceTrainStep(dev, op=op, seed=0)
            tgs = [R.list_targets(ts.config, BATCH, 1000 * rank + i, dev) for i in range(2)]
            for i in range(warm):
                ts.step(vols[i % 2], tgs[i % 2])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                loss = ts.step(vols[i % 2], tgs[i % 2])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out["variants"][label] = {"ms_per_step": ms, "volumes_per_s": BATCH / (ms / 1e3), "final_loss": loss,
                                      "cudnn_benchmark": bench_flag, "cudnn_deterministic": det, "op": op}
            del ts
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.deterministic = saved
    best = max((v["volumes_per_s"] for k, v in out["variants"].items() if k.startswith("reference_op")), default=None)
    out["volumes_per_s"] = best
    out["volumes_per_s_is"] = "the faster of the two reference_op variants"
    return out


# ---------------------------------------------------------------------------------------------------------------
# The other BASELINE.json configurations as bench workloads (`--workload`): configs[2] and configs[3], bf16.  Same contract line; the
# headline (`python bench.py` with no flags) stays configs[1].
# ---------------------------------------------------------------------------------------------------------------
def other_workloads():
    import torch
    from transoar_b200.engine import defdetr_train_config, swin_focused_train_config
    return {
        "defdetr_300_amos_bf16": dict(
            cfg=defdetr_train_config, volume=(256, 256, 128), amp=torch.bfloat16, dtype="bf16",
            what="BASELINE configs[2]: 3D Deformable-DETR style detector (transoar_b200/def_detr.py: AttnFPN + 2 multi-level deformable encoder layers over "
                 "P2..P5 + 3 deformable decoder layers, 300 queries), synthetic AMOS-shape 256x256x128 volumes, bf16 autocast; neck restated (not in the "
                 "reference tree, SURVEY D5: parity unpinned at neck level, operator pinned)"),
        "swin_focused_192_bf16": dict(
            cfg=swin_focused_train_config, volume=(192, 192, 384), amp=torch.bfloat16, dtype="bf16",
            what="BASELINE configs[3]: SwinFPN backbone (use_encoder_attn=True, stages 2-5 Swin3D) + deformable FPN refinement + Focused Decoder on synthetic "
                 "192x192x384 volumes, bf16 autocast; RoI grid derived from the P2 map (48x48x96; no row in the reference's shape table, SURVEY D4)"),
    }


def run_other_workload(args):
    import torch
    import torch.distributed as dist
    from transoar_b200 import MultiScaleDeformableAttention as MSDA
    from transoar_b200 import _lib
    from transoar_b200.engine import TrainStep, synthetic_targets
    wl = other_workloads()[args.workload]
    lib = _lib.lib()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cfg = wl["cfg"](volume=wl["volume"])
    torch.manual_seed(0)
    use_graph = bool(args.graph)
    ts = TrainStep(cfg, dev, world=world, graph=use_graph, graph_warmup=3, amp_dtype=wl["amp"])
    torch.manual_seed(1 + 7919 * rank)
    gen = torch.Generator().manual_seed(100 + volume_ids(0, rank, world)[0])
    vols_host = [torch.rand(BATCH, 1, *wl["volume"], generator=gen).pin_memory() for _ in range(2)]
    vols_dev = [v.to(dev) for v in vols_host]
    targets = [synthetic_targets(cfg, BATCH, 1000 * rank + i, dev) for i in range(2)]
    warm = max(3, args.warmup) + (4 if use_graph else 0)
    for i in range(warm):
        ts.step(vols_dev[i % 2], targets[i % 2])
    fence()
    if args.profile_one_step:                                     # ncu --profile-from-start off: one step between cudaProfilerStart / Stop
        torch.cuda.profiler.start()
        ts.step(vols_dev[0], targets[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return 0
    ev_log = []
    if not use_graph:
        MSDA.set_event_log(ev_log)
    launches0 = lib.msda3d_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        fence()
        start.record()
        for i in range(args.steps):
            loss = ts.step(vols_dev[i % 2], targets[i % 2])
        stop.record()
        fence()
    MSDA.set_event_log(None)
    launches = lib.msda3d_launch_count() - launches0
    value, ms_total = aggregate_throughput(start.elapsed_time(stop), args.steps, world, all_reduce_max=reduce_max)
    final_loss = float(loss.item())
    peaks = load_peaks()
    roofline = None
    if ev_log:                                                    # the op's launches inside the timed region, by kind; algorithmic bytes from each launch's own dims
        tot = {"fwd": [0.0, 0.0, 0], "bwd": [0.0, 0.0, 0]}
        for kind, a, b, d in ev_log:
            N, S, M, C, L, Lq, P, ev = d
            bf, bb = algorithmic_bytes(N, S, M, C, L, Lq, P, ev=ev)
            tot[kind][0] += bf if kind == "fwd" else bb
            tot[kind][1] += a.elapsed_time(b)
            tot[kind][2] += 1
        dom = max(tot, key=lambda k: tot[k][1])
        gbs = {k: (v[0] / v[1] / 1e6 if v[1] else None) for k, v in tot.items()}
        roofline = {"bound": "hbm", "kernel": f"msda3d {dom} launches of the step (encoder Lq = S and decoder Lq = {cfg['neck']['num_queries']} calls together)",
                    "achieved": gbs[dom], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs[dom] / peaks["hbm_gbs"], "traffic": None,
                    "peak_source": peaks["source"], "launches_timed": tot[dom][2], "share_of_step": tot[dom][1] / ms_total,
                    "forward_gbs": gbs["fwd"], "backward_gbs": gbs["bwd"],
                    "timed_in": "CUDA events on the launching stream around every launch inside the timed region (eager steps)"}
    e2e_steps = max(3, min(args.steps, 6))
    for i in range(2):
        float(ts.step(vols_host[i % 2], targets[i % 2]).item())
    fence()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        float(ts.step(vols_host[i % 2], targets[i % 2]).item())
    fence()
    dt = reduce_max(time.perf_counter() - t0)
    e2e = {"value": world * BATCH * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": vols_host[0].numel() * 4, "d2h_bytes_per_step": 4,
           "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
           "api": "transoar_b200.engine.TrainStep.step(volumes_in_pinned_host_memory, targets) -> loss; float(loss) on the host every step"}
    if rank == 0:
        emit({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_total / args.steps,
              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
              "config": {"workload": args.workload, "volume": "x".join(map(str, wl["volume"])), "batch_per_gpu": BATCH, "what": wl["what"],
                         "step": "forward + matcher + losses + backward + AdamW; nothing skipped",
                         "precision": "torch.autocast(bfloat16): bf16 tcgen05 GEMMs (kind::f16) for every Linear, bf16 `value` into the msda3d kernels with fp32 "
                                      "locations / weights; the convolutional backbone stays on this library's fp32-storage / TF32-multiply tcgen05 convolutions and fp32 "
                                      "InstanceNorm kernels (higher precision than bf16; no cuDNN launch); fp32 master weights and optimiser",
                         "l2": "activations of hundreds of MB per tensor: far larger than the 126 MB L2; no explicit flush",
                         "execution": "CUDA graph replay" if use_graph else "eager launches",
                         "parallelism": "volumes sharded over ranks; one NCCL gradient all-reduce per step"},
              "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches), "cuda_graph": use_graph,
              "model": {"params": sum(p.numel() for p in ts.net.parameters()), "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30, "final_loss": final_loss},
              "clocks": clocks.summary()})
    if world > 1:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~20 s host baseline (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-kernel extras (operator alone, GEMM, RoI attention, InstanceNorm)")
    ap.add_argument("--workload", default="visceral_train_step", choices=["visceral_train_step", "defdetr_300_amos_bf16", "swin_focused_192_bf16"],
                    help="visceral_train_step = BASELINE configs[1] (the headline); the other two are configs[2] / configs[3]")
    ap.add_argument("--graph", action="store_true", help="other workloads only: capture the step into a CUDA graph (default eager)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-model-on-GPU comparator (ref_gpu_model)")
    ap.add_argument("--no-graph", action="store_true", help="run every step eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--profile-one-step", action="store_true",
                    help="for ncu --profile-from-start off: warm up, bracket ONE training step with cudaProfilerStart/Stop, exit")
    ap.add_argument("--dist", default=DIST, choices=["A", "B"], help="sampling-location distribution of the operator-alone extra")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload != "visceral_train_step":
        return run_other_workload(args)

    import torch
    import torch.distributed as dist
    from transoar_b200 import MultiScaleDeformableAttention as MSDA
    from transoar_b200 import _lib, synth
    from transoar_b200.engine import TrainStep, synthetic_targets, visceral_train_config

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: transoar_b200 has no CPU path")
    lib = _lib.lib()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"              # keep NCCL's version banner off stdout: the JSON line is the only output
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the headline: whole training step, inputs resident on the device
    cfg = visceral_train_config()
    torch.manual_seed(0)                                   # identical initial weights on every rank (DDP also broadcasts rank 0's)
    use_graph = not (args.no_graph or args.profile_one_step)
    eager_warm = 4 if use_graph else 0                     # eager steps before the capture; the msda3d kernels are timed in them
    ts = TrainStep(cfg, dev, world=world, graph=use_graph, graph_warmup=eager_warm)
    torch.manual_seed(1 + 7919 * rank)                     # per-rank dropout streams (weights are already identical / broadcast)
    gen = torch.Generator().manual_seed(100 + volume_ids(0, rank, world)[0])
    n_sets = 2                                             # alternate between two resident batches
    vols_host = [torch.rand(BATCH, 1, *VOLUME, generator=gen).pin_memory() for _ in range(n_sets)]
    vols_dev = [v.to(dev) for v in vols_host]
    targets = [synthetic_targets(cfg, BATCH, 1000 * rank + i, dev) for i in range(n_sets)]
    warm = max(3, args.warmup)
    ev_log = []
    launches_per_step = None
    for i in range(eager_warm):                            # eager steps (not counted in `warmup`): autotuning, lazy init, then events
        if i == 1:
            torch.cuda.synchronize()
            MSDA.set_event_log(ev_log)
            launches0 = lib.msda3d_launch_count()
        ts.step(vols_dev[i % n_sets], targets[i % n_sets])
    if use_graph:
        torch.cuda.synchronize()
        MSDA.set_event_log(None)
        launches_per_step = (lib.msda3d_launch_count() - launches0) / (eager_warm - 1)
    for i in range(warm):                                  # the first of these captures the graph, the rest replay it
        ts.step(vols_dev[i % n_sets], targets[i % n_sets])
    fence()
    if args.profile_one_step:
        torch.cuda.profiler.start()
        ts.step(vols_dev[0], targets[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return 0
    if not use_graph:
        MSDA.set_event_log(ev_log)
        launches0 = lib.msda3d_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        fence()
        start.record()
        for i in range(args.steps):
            loss = ts.step(vols_dev[i % n_sets], targets[i % n_sets])
        stop.record()
        fence()
    if use_graph:                                          # a replay re-issues, from the graph, exactly the launches of one eager step
        launches = int(round(launches_per_step * args.steps))
    else:
        launches = lib.msda3d_launch_count() - launches0
        MSDA.set_event_log(None)
    value, ms_total = aggregate_throughput(start.elapsed_time(stop), args.steps, world, all_reduce_max=reduce_max)
    ms_step = ms_total / args.steps
    final_loss = float(loss.item())
    if use_graph:
        # Kernel-level timing of THIS run: the replayed graph hides individual launches from CUDA events, so right after the timed region
        # (same process, same weights / optimiser state / inputs, clocks still under load) three steps are launched eagerly with an event
        # pair around every msda3d launch.  These replace the events of the pre-capture warm-up steps.
        ev_log.clear()
        MSDA.set_event_log(ev_log)
        for i in range(3):
            ts.eager_step(vols_dev[i % n_sets], targets[i % n_sets])
        torch.cuda.synchronize()
        MSDA.set_event_log(None)

    # per-launch times of the msda3d kernels inside the timed region (events on the launching stream)
    fwd_ms = [a.elapsed_time(b) for k, a, b, _ in ev_log if k == "fwd"]
    bwd_ms = [a.elapsed_time(b) for k, a, b, _ in ev_log if k == "bwd"]
    g = synth.GEOMETRIES[GEOM]
    N, S, M, C, L, Lq, P = BATCH, g.spatial_size, g.heads, g.channels, g.levels, g.num_query, g.points
    bf, bb = algorithmic_bytes(N, S, M, C, L, Lq, P)
    peaks = load_peaks()
    peak = peaks["hbm_gbs"]
    fwd_avg, bwd_avg = sum(fwd_ms) / len(fwd_ms), sum(bwd_ms) / len(bwd_ms)
    roofline = {"bound": "hbm", "kernel": "msda3d backward: bwd_duo_kernel<FUSED=1> (two w-neighbouring queries per lane group) + the cudaMemsetAsync zero-fill of grad_value "
                                          f"({LAYERS} launches per step, the largest single kernel of the step)",
                "achieved": bb / bwd_avg / 1e6, "peak": peak, "unit": "GB/s", "frac": bb / bwd_avg / 1e6 / peak,
                "traffic": load_ncu_traffic("backward"), "peak_source": peaks["source"], "algorithmic_bytes": bb,
                "launches_timed": len(bwd_ms), "share_of_step": LAYERS * bwd_avg / ms_step,
                "timed_in": ("CUDA events on the launching stream around every launch of three steps launched eagerly right AFTER the timed graph "
                             "replays of this run (same process, weights, inputs and clocks; events cannot be read inside a replayed graph)") if use_graph else
                            "CUDA events on the launching stream around every launch inside the timed region",
                "forward": {"ms": fwd_avg, "bytes": bf, "gbs": bf / fwd_avg / 1e6, "frac": bf / fwd_avg / 1e6 / peak,
                            "traffic": load_ncu_traffic("forward"), "share_of_step": LAYERS * fwd_avg / ms_step},
                "backward": {"ms": bwd_avg, "bytes": bb, "gbs": bb / bwd_avg / 1e6, "frac": bb / bwd_avg / 1e6 / peak},
                "note": "gather / scatter path: 8 corner reads (and 8 reductions) per sample go through L1/L2, requested bytes = "
                        f"{8 * N * Lq * M * L * P * C * 4 / 1e9:.1f} GB per launch vs {bf / 1e9:.2f} GB compulsory; the binding resources are the "
                        "L1 data pipe (forward) and the L2 atomic units (backward), not HBM -- see DESIGN.md"}

    # ---- end to end: same call, volumes in pinned host memory (H2D inside the step), loss read back every step
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 10))
        def e2e_step(i):
            # what a training loop with a data loader does: launch the step, start the NEXT batch's host -> device copy (it runs under
            # this step's kernels), then block on this step's loss.  Every step copies its 52 MB from pinned host memory and reads its loss.
            loss = ts.step(vols_host[i % n_sets], targets[i % n_sets])
            ts.prefetch(vols_host[(i + 1) % n_sets])
            return float(loss.item())
        for i in range(2):
            e2e_step(i)
        fence()
        t0 = time.perf_counter()
        for i in range(2, 2 + e2e_steps):
            e2e_step(i)
        fence()
        dt = reduce_max(time.perf_counter() - t0)
        e2e = {"value": world * BATCH * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": vols_host[0].numel() * 4, "d2h_bytes_per_step": 4,
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "api": "loss = TrainStep.step(volumes_in_pinned_host_memory, targets); TrainStep.prefetch(next_volumes_in_pinned_host_memory); "
                      "float(loss) on the host every step -- every step's volumes cross PCIe inside the timed region, the copy of step i + 1 "
                      "overlapping the kernels of step i"}
    params = sum(p.numel() for p in ts.net.parameters())
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30
    del ts, vols_dev, vols_host
    torch.cuda.empty_cache()

    # ---- the kernels of the path one by one
    extras = {}
    if not args.no_extras:
        try:
            extras["msda3d_op"] = measure_msda_op(dev, rank, world, args.dist, with_ref=(rank == 0 and world == 1))
            torch.cuda.empty_cache()
            if rank == 0:
                extras.update(measure_kernels(dev, rank))
        except Exception as exc:            # never lose the headline line because an extra failed
            extras["extras_error"] = f"{type(exc).__name__}: {exc}"[:300]
    if world > 1:
        dist.barrier()

    # ---- the other shipped yaml (attn_fpn_foc_dec_amos.yaml: refinement over P3..P5, 405 queries / 15 organs) on the same volumes
    if not args.no_extras and "extras_error" not in extras:
        try:
            from transoar_b200.engine import amos_train_config
            acfg = amos_train_config(volume=VOLUME)
            torch.manual_seed(0)
            ats = TrainStep(acfg, dev, world=1)
            av = torch.rand(BATCH, 1, *VOLUME, device=dev)
            atg = synthetic_targets(acfg, BATCH, 5, dev)
            a_ms = _event_ms(lambda: ats.step(av, atg), 4, 8)
            extras["amos_yaml_train_step"] = {"what": "config/attn_fpn_foc_dec_amos.yaml (feature_levels P3-P5, 405 queries, 15 organs) on 160x160x256 volumes, "
                                                      "same step, this rank only", "volumes_per_s_per_gpu": BATCH / (a_ms / 1e3), "ms_per_step": a_ms}
            del ats, av
            torch.cuda.empty_cache()
        except Exception as exc:
            extras["amos_yaml_error"] = f"{type(exc).__name__}: {exc}"[:300]
    if world > 1:
        dist.barrier()

    # ---- the reference's own model + its own compiled op on this GPU (rank 0, single-GPU runs only): the north-star comparator
    if rank == 0 and world == 1 and not args.no_extras and not args.no_ref_gpu:
        try:
            extras["ref_gpu_model"] = measure_reference_gpu_model(dev, rank)
            best = extras["ref_gpu_model"].get("volumes_per_s")
            if best:
                extras["vs_reference_gpu_model"] = {"value_over_reference": value / best, "e2e_over_reference": (e2e["value"] / best) if e2e else None,
                                                    "reference_volumes_per_s": best,
                                                    "what": "this arm's volumes/s divided by the unmodified reference model running its own compiled CUDA op on the same GPU, same run"}
        except Exception as exc:
            extras["ref_gpu_model"] = {"error": f"{type(exc).__name__}: {exc}"[:400]}
        torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, single-GPU runs only): the reference's CPU route of the same step, bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(os.cpu_count() or 1)
        times = [arm.step()]                                # ONE full-size step, no warm-up: 20-60 s of host time (bounded sample)
        cpu_baseline = {"value": CPU_STEP_VOLUMES / times[0], "unit": UNIT, "cores": arm.threads, "kind": arm.kind,
                        "sample": arm.sample_text(times, 0)}
        arm.close()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "kernels": {"msda3d_fwd_ms": fwd_avg, "msda3d_bwd_ms": bwd_avg, "msda3d_fwd_ms_min": min(fwd_ms), "msda3d_bwd_ms_min": min(bwd_ms),
                        "launches_of_this_library_per_step": launches / args.steps},
            "model": {"params": params, "peak_mem_gib": peak_mem, "final_loss": final_loss},
            "cuda_graph": use_graph, "clocks": clocks.summary(), **extras,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
