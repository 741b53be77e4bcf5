"""Kernel-time breakdown of one training step (config 2) — GPU box."""
import os, sys, importlib.util
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch.profiler import profile, ProfilerActivity
import lewin_b200 as L
from lewin_b200 import parallel
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(dev).train()
parallel.freeze_dead_parameters(model)
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=2e-4, weight_decay=0.02)
x = torch.rand(32, 3, 128, 128, device=dev); y = torch.rand(32, 3, 128, 128, device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", torch.bfloat16):
        out = model(x)
    d = out.float() - y
    loss = torch.mean(torch.sqrt(d * d + 1e-6))
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[ev.name[:110]]
        a[0] += 1; a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
lew = sum(v[1] for k, v in agg.items() if "lewin" in k)
print(f"training step (eager, Charbonnier only): kernel time {tot / 1e3:.2f} ms in {sum(v[0] for v in agg.values())} kernels; lewin:: kernels {lew / 1e3:.2f} ms")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{us:9.1f} us  x{n:4d}  {name}")
