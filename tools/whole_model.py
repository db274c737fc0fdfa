"""Whole training step (transoar_b200.engine.TrainStep, VISCERAL config) timing + kernel-time breakdown (diagnostic).
usage: whole_model.py [batch] [cl|nocl] [tf32|fp32] [topN]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200.engine import TrainStep, synthetic_targets, visceral_train_config
dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cl = (sys.argv[2] if len(sys.argv) > 2 else "cl") == "cl"
tf32 = (sys.argv[3] if len(sys.argv) > 3 else "tf32") == "tf32"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
torch.backends.cudnn.benchmark = (sys.argv[5] == "bench") if len(sys.argv) > 5 else False
torch.manual_seed(0)
cfg = visceral_train_config()
ts = TrainStep(cfg, dev, tf32=tf32, channels_last=cl)
if not tf32:
    torch.backends.cuda.matmul.allow_tf32 = False
x = torch.rand(B, 1, 160, 160, 256, device=dev)
tg = synthetic_targets(cfg, B, 0, dev)
for _ in range(4): ts.step(x, tg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ts.step(x, tg)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"B={B} channels_last={cl} tf32={tf32}: {ms:.1f} ms/step  {B / ms * 1e3:.2f} volumes/s  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ts.step(x, tg); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:top]
tot = sum(r.device_time_total for r in prof.key_averages())
print(f"sum of kernel time {tot / 1e3:.1f} ms")
for r in rows:
    print(f"{r.device_time_total / 1e3:8.2f} ms {100 * r.device_time_total / tot:5.1f}%  x{r.count:<4d} {r.key[:150]}")
