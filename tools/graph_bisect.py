"""Which modules of the model survive a CUDA-graph capture of their forward (diagnostic): records every module call of one eager
forward down to `depth`, then captures each call on its own.   usage: graph_bisect.py [depth] [name-prefix]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200.engine import TrainStep, synthetic_targets, visceral_train_config

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 1
prefix = sys.argv[2] if len(sys.argv) > 2 else ""
dev = "cuda:0"
shape = (64, 64, 128)
cfg = visceral_train_config()
cfg["neck_input_shape"] = tuple(s // 4 for s in shape)
ts = TrainStep(cfg, dev)
x = torch.rand(1, 1, *shape, device=dev)
calls = []
def pre(name):
    def hook(mod, args, kwargs):
        calls.append((name, mod, args, kwargs))
    return hook
for name, m in ts.net.named_modules():
    if name and name.count(".") < depth and name.startswith(prefix):
        m.register_forward_pre_hook(pre(name), with_kwargs=True)
with torch.no_grad():
    for _ in range(2):
        calls.clear()
        ts.net(x)
torch.cuda.synchronize()
recorded = list(calls)
for name, mod, args, kwargs in recorded:
    mod._forward_pre_hooks.clear()
stream = torch.cuda.Stream()
for name, mod, args, kwargs in recorded:
    g = torch.cuda.CUDAGraph()
    try:
        with torch.no_grad(), torch.cuda.graph(g, stream=stream):
            mod(*args, **kwargs)
        g.replay(); torch.cuda.synchronize()
        print(f"ok      {name} ({type(mod).__name__})", flush=True)
    except Exception as e:
        print(f"BROKEN  {name} ({type(mod).__name__}): {str(e).splitlines()[0][:120]}", flush=True)
