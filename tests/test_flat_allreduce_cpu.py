"""engine.all_reduce_gradients (the gradient exchange of the graph-captured step at world > 1) on gloo, world_size 2: every rank ends
with the mean gradient, in each parameter's own memory layout, and parameters without a gradient are skipped on every rank alike."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from transoar_b200.engine import all_reduce_gradients
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(4, 3, 3, 3, 3).contiguous(memory_format=torch.channels_last_3d)),
              torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(5, 2)), torch.nn.Parameter(torch.zeros(3))]
    g = torch.Generator().manual_seed(10 + rank)
    for p in params[:3]:
        p.grad = torch.randn(p.shape, generator=g).contiguous(memory_format=torch.channels_last_3d) if p.dim() == 5 else torch.randn(p.shape, generator=g)
    all_reduce_gradients(params, world)
    assert params[3].grad is None
    assert params[0].grad.is_contiguous(memory_format=torch.channels_last_3d)
    if rank == 0:
        torch.save([p.grad for p in params[:3]], out)
    dist.destroy_process_group()


def test_flat_bucket_all_reduce_averages_gradients(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    shapes = [(4, 3, 3, 3, 3), (7,), (5, 2)]
    want = []
    for shp_i, shp in enumerate(shapes):
        acc = torch.zeros(shp)
        for rank in range(2):
            g = torch.Generator().manual_seed(10 + rank)
            for j, s in enumerate(shapes):
                t = torch.randn(s, generator=g)
                if j == shp_i:
                    acc += t
        want.append(acc / 2)
    for a, b in zip(got, want):
        assert torch.allclose(a, b, atol=1e-6)


def _overlap_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from transoar_b200.engine import OverlappedGradientAverage
    torch.manual_seed(0)
    net = torch.nn.ModuleDict({"stem": torch.nn.Linear(6, 6), "body": torch.nn.Linear(6, 6), "head": torch.nn.Linear(6, 2),
                               "unused": torch.nn.Linear(3, 3)})
    avg = OverlappedGradientAverage(net, world, late_prefixes=("stem.",))
    assert len(avg.late) == 2 and len(avg.early) == 6
    log = []
    for step in range(3):
        x = torch.randn(5, 6, generator=torch.Generator().manual_seed(100 * step + rank))
        avg.zero()
        avg.before_backward()
        net["head"](net["body"](net["stem"](x))).square().sum().backward()
        fired_in_backward = avg.fired
        avg.after_backward()
        log.append((fired_in_backward, [None if p.grad is None else p.grad.clone() for p in net.parameters()]))
        # after the calibration step every gradient is a view into the one persistent flat buffer: no cat, no copy back
        lo, hi = avg.flat.data_ptr(), avg.flat.data_ptr() + avg.flat.numel() * 4
        assert all(lo <= p.grad.data_ptr() < hi for p in net.parameters() if p.grad is not None)
        assert avg.flat.numel() == sum(p.numel() for n, p in net.named_parameters() if not n.startswith("unused"))
    assert avg.expected == 4                                             # body + head; the unused layer never gets a gradient
    assert all(p.grad is None for p in net["unused"].parameters())       # ... and keeps grad None (AdamW skips it, as in the reference)
    assert [f for f, _ in log] == [False, True, True]                    # calibration step, then the early bucket fires inside backward
    if rank == 0:
        torch.save([g for _, g in log], out)
    dist.destroy_process_group()


def test_two_bucket_average_fires_inside_backward_and_matches_the_mean(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_overlap_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    net = torch.nn.ModuleDict({"stem": torch.nn.Linear(6, 6), "body": torch.nn.Linear(6, 6), "head": torch.nn.Linear(6, 2),
                               "unused": torch.nn.Linear(3, 3)})
    for step in range(3):
        want = None
        for rank in range(2):
            x = torch.randn(5, 6, generator=torch.Generator().manual_seed(100 * step + rank))
            for p in net.parameters():
                p.grad = None
            net["head"](net["body"](net["stem"](x))).square().sum().backward()
            grads = [None if p.grad is None else p.grad.clone() for p in net.parameters()]
            want = grads if want is None else [None if a is None else a + b for a, b in zip(want, grads)]
        for a, b in zip(got[step], want):
            assert (a is None) == (b is None)
            if a is not None:
                assert torch.allclose(a, b / 2, atol=1e-6)


def test_dense_view_keeps_channels_last_strides():
    from transoar_b200.engine import _dense_view
    flat = torch.zeros(1000)
    p = torch.nn.Parameter(torch.randn(4, 3, 2, 3, 3).contiguous(memory_format=torch.channels_last_3d))
    v = _dense_view(flat, 10, p)
    assert v.shape == p.shape and v.stride() == p.stride() and v.data_ptr() == flat.data_ptr() + 40
    v.copy_(p.detach())
    assert torch.equal(v, p.detach()) and flat[:10].abs().sum() == 0 and flat[10 + p.numel():].abs().sum() == 0
    q = torch.nn.Parameter(torch.randn(6, 5))
    w = _dense_view(flat, 300, q)
    assert w.is_contiguous() and w.shape == q.shape
    import pytest
    with pytest.raises(ValueError):
        _dense_view(flat, 0, torch.nn.Parameter(torch.randn(4, 8)[:, ::2]))
