"""Whole-model (TransoarNet mirror, VISCERAL config, refine on) fwd+bwd+AdamW timing + kernel-time breakdown (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200.configs import visceral_config
from transoar_b200.transoarnet import TransoarNet
dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
amp = sys.argv[2] if len(sys.argv) > 2 else "fp32"
torch.manual_seed(0)
net = TransoarNet(visceral_config()).to(dev).train()
opt = torch.optim.AdamW(net.parameters(), lr=2e-4, weight_decay=1e-4)
x = torch.rand(B, 1, 160, 160, 256, device=dev)
tgt_boxes = torch.rand(B, 540, 6, device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(amp == "bf16")):
        out = net(x)
        loss = torch.nn.functional.l1_loss(out["pred_boxes"].float(), tgt_boxes) + out["pred_logits"].float().sigmoid().mean()
        for a in out["aux_outputs"]:
            loss = loss + torch.nn.functional.l1_loss(a["pred_boxes"].float(), tgt_boxes)
    loss.backward()
    opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"B={B} {amp}: {ms:.1f} ms/step  {B / ms * 1e3:.2f} volumes/s  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:45]
tot = sum(r.device_time_total for r in prof.key_averages())
for r in rows:
    print(f"{r.device_time_total / 1e3:8.2f} ms {100 * r.device_time_total / tot:5.1f}%  x{r.count:<4d} {r.key[:110]}")
