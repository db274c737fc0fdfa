"""N > 1 host logic on CPU: two gloo ranks on 127.0.0.1 exercise bench.py's volume sharding, the max-over-ranks
timing reduction and the whole-job aggregation (the data path itself has no collective: volumes are independent)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bench


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = [bench.volume_ids(step, rank, world) for step in range(3)]
        gathered = [None] * world
        dist.all_gather_object(gathered, ids)

        def reduce_max(ms):
            t = torch.tensor([ms], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        local_ms = 100.0 * (rank + 1)                 # rank 1 is the slow one
        value, ms = bench.aggregate_throughput(local_ms, steps=5, world=world, all_reduce_max=reduce_max)
        dist.barrier()
        if rank == 0:
            out.put((gathered, value, ms))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_max_over_ranks_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, value, ms = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = sorted(v for per_rank in gathered for step in per_rank for v in step)
    assert flat == list(range(3 * world * bench.BATCH)), "ranks must cover a contiguous volume stream without overlap"
    assert ms == 200.0                                 # max over ranks, not the mean
    assert value == pytest.approx(world * bench.BATCH * 5 / 0.2)


def test_algorithmic_bytes_match_survey_8d():
    # VISCERAL fp32 N=1: B_f = 0.539 GB, B_b ~ 1.08 GB
    bf, bb = bench.algorithmic_bytes(1, 117000, 6, 64, 4, 117000, 4)
    assert bf == 4 * (117000 * 384 + 117000 * 384) + 16 * 117000 * 6 * 16 == 539136000
    assert bb == 2 * bf
    # decoder-style call (Lq = 300): the value term is capped by what 8 corners per sample can touch
    bf300, _ = bench.algorithmic_bytes(1, 117000, 6, 64, 4, 300, 4)
    assert bf300 == 4 * (8 * 300 * 6 * 16 * 64 + 300 * 384) + 16 * 300 * 6 * 16


def test_cpu_arm_times_the_real_volume():
    # the CPU arm runs whole 160x160x256 volumes (no crop, nothing extrapolated): value = volumes per step / seconds per step
    assert bench.CPU_STEP_VOLUMES == 1 and bench.VOLUME == (160, 160, 256)
    assert not hasattr(bench, "cpu_value") and not hasattr(bench, "CPU_SAMPLE_SHAPES")


def test_cpu_reference_route_of_the_train_step_runs_and_learns():
    """oracle/model_oracle.py (the CPU arm of bench.py): two steps on the smallest sample volume, loss finite and decreasing."""
    times = []
    import torch
    from oracle.model_oracle import CpuTrainStep
    from transoar_b200.engine import synthetic_targets, visceral_train_config
    cfg = visceral_train_config()
    ts = CpuTrainStep(cfg, (32, 32, 64))
    x = torch.rand(1, 1, 32, 32, 64, generator=torch.Generator().manual_seed(1))
    tg = synthetic_targets(cfg, 1, 0, "cpu")
    losses = [ts.step(x, tg)[1] for _ in range(2)]
    assert all(l == l and l < 1e3 for l in losses) and losses[1] < losses[0]


def test_reference_arm_runs_on_rank_zero_only(monkeypatch, capfd):
    """`--impl reference` under torchrun (N > 1): rank 0 alone times the host arm and prints the line; every other rank exits 0 without
    work and without output."""
    import argparse
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setattr(bench, "CpuArm", lambda *a, **k: (_ for _ in ()).throw(AssertionError("rank 1 must not build the CPU arm")))
    assert bench.run_reference_arm(argparse.Namespace(gpus=2, steps=3, warmup=1)) == 0
    assert capfd.readouterr().out == ""


def test_reference_arm_line_and_wall_budget(monkeypatch):
    """The host arm's JSON line (contract keys of `--impl reference`) and its wall budget: with steps that would overrun the budget the
    warm-up is cut to the one step already done and the timed steps stop early, and `steps` / `warmup` say what was actually run."""
    import argparse

    class FakeArm:
        threads, kind = 16, "reference"

        def __init__(self, threads):
            self.calls = 0

        def step(self):
            self.calls += 1
            clock.now += self.sec                           # the arm's wall budget is read from the clock, not from the returned times
            return self.sec

        def sample_text(self, times, warm):
            return f"{len(times)} timed / {warm} warm"

    class clock:
        now = 0.0
        perf_counter = staticmethod(lambda: clock.now)

    lines = []
    monkeypatch.setattr(bench, "time", clock)
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setattr(bench, "CpuArm", FakeArm)
    monkeypatch.setattr(bench, "emit", lines.append)
    # cheap steps: everything requested is run
    FakeArm.sec = 0.001
    assert bench.run_reference_arm(argparse.Namespace(gpus=1, steps=4, warmup=2)) == 0
    line = lines.pop()
    assert line["impl"] == "reference" and line["steps"] == 4 and line["warmup"] == 2 and line["gpu_launches"] == 0
    assert line["metric"] == bench.METRIC and line["unit"] == bench.UNIT and line["config"] == bench.workload_config()
    assert line["value"] == pytest.approx(bench.CPU_STEP_VOLUMES / 0.001) and line["ms_per_step"] == pytest.approx(1.0)
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == 16 and line["cpu_baseline"]["value"] == line["value"]
    # steps that claim 400 s each against a 780 s budget: one warm-up, then no more than the two timed steps the mean needs
    monkeypatch.setenv("TRANSOAR_REF_BUDGET_S", "780")
    FakeArm.sec = 400.0
    assert bench.run_reference_arm(argparse.Namespace(gpus=1, steps=20, warmup=5)) == 0
    line = lines.pop()
    assert line["warmup"] == 1 and line["steps"] == 2 and line["steps_requested"] == 20 and line["warmup_requested"] == 5
