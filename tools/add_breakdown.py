"""Which elementwise adds / copies of the training step are the large ones (diagnostic): ATen op x input shapes, by device time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from transoar_b200.engine import TrainStep, synthetic_targets, visceral_train_config
dev = "cuda:0"
cfg = visceral_train_config()
ts = TrainStep(cfg, dev)
x = torch.rand(2, 1, 160, 160, 256, device=dev)
tg = synthetic_targets(cfg, 2, 0, dev)
for _ in range(4): ts.step(x, tg)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    ts.step(x, tg); torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    t = getattr(e, "self_device_time_total", 0) or 0
    if t > 0:
        rows.append((t / 1e3, e.count, e.key, str(e.input_shapes)[:110]))
rows.sort(reverse=True)
names = sys.argv[1].split(",") if len(sys.argv) > 1 else None
n = 0
for t, c, k, shp in rows:
    if names is None or any(k == nm for nm in names):
        print(f"{t:8.3f} ms x{c:<4d} {k:38s} {shp}")
        n += 1
        if n >= int(sys.argv[2]) if len(sys.argv) > 2 else n >= 40: break
