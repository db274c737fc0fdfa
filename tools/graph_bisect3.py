"""Stage-wise capture of the training step (diagnostic): forward / + criterion / + backward / + optimizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200.engine import TrainStep, synthetic_targets, visceral_train_config
from transoar_b200.criterion import total_loss
dev = "cuda:0"
shape = (64, 64, 128)
cfg = visceral_train_config()
cfg["neck_input_shape"] = tuple(s // 4 for s in shape)
ts = TrainStep(cfg, dev, graph=True, graph_warmup=10 ** 9)
x = torch.rand(1, 1, *shape, device=dev)
tg = synthetic_targets(cfg, 1, 0, dev)
stream = torch.cuda.Stream()
stream.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(stream):
    x = ts.to_device(x)
    for _ in range(2): ts._run(x, tg, None)
torch.cuda.synchronize()

def fwd():
    with torch.no_grad():
        return ts.model(x)["pred_logits"].sum()
def crit():
    with torch.no_grad():
        return total_loss(ts.criterion(ts.model(x), tg, None, ts.net._anchors), cfg["loss_coefs"])
def bwd():
    ts.optim.zero_grad(set_to_none=True)
    loss = total_loss(ts.criterion(ts.model(x), tg, None, ts.net._anchors), cfg["loss_coefs"])
    loss.backward()
    return loss.detach()
def full():
    loss = bwd()
    ts.optim.step()
    return loss
def bwd_surrogate():
    ts.optim.zero_grad(set_to_none=True)
    out = ts.model(x)
    loss = out["pred_logits"].square().mean() + out["pred_boxes"].square().mean()
    loss.backward()
    return loss.detach()
for name, fn in (("forward", fwd), ("forward+criterion", crit), ("fwd+surrogate loss+backward", bwd_surrogate), ("fwd+criterion+backward", bwd), ("full step", full)):
    ts.optim.zero_grad(set_to_none=True)
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, stream=stream):
            res = fn()
        g.replay(); g.replay(); torch.cuda.synchronize()
        print(f"ok      {name}: {float(res):.4f}", flush=True)
    except Exception as e:
        print(f"BROKEN  {name}: {type(e).__name__}: {str(e).splitlines()[0][:150]}", flush=True)
