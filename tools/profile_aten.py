"""Which ATen elementwise kernels does one eager training step launch, from where?  (torch.profiler with shapes + python stacks; the top entries by CUDA time)"""
import sys, collections
import torch
sys.path.insert(0, ".")
from transoar_b200.engine import TrainStep, synthetic_targets, visceral_train_config
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
cfg = visceral_train_config()
ts = TrainStep(cfg, "cuda:0", graph=False)
x = torch.rand(2, 1, 160, 160, 256, device="cuda:0")
tg = synthetic_targets(cfg, 2, 0, "cuda:0")
for _ in range(3): ts.step(x, tg)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    ts.step(x, tg); torch.cuda.synchronize()
rows = [k for k in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6) if k.key.startswith("aten::")]
tm = lambda k: getattr(k, "self_device_time_total", None) or getattr(k, "self_cuda_time_total", 0)
for k in sorted(rows, key=lambda k: -tm(k))[:40]:
    stack = [s_ for s_ in (k.stack or []) if "transoar_b200" in s_][:2]
    print(f"{tm(k)/1e3:7.3f} ms x{k.count:<3d} {k.key:30s} {str(k.input_shapes)[:80]:80s} {' <- '.join(x.split('/')[-1][:55] for x in stack)}")
