"""Per-launch kernel times of one full-resolution fp32 tiled forward (169 tiles), in launch order — GPU box."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch.profiler import profile, ProfilerActivity
import lewin_b200 as L
from lewin_b200 import fullres
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(dev).eval()
img = torch.rand(1, 3, 1200, 1600, device=dev)
idx = model.draw_index_samples()
def fwd():
    with torch.no_grad():
        return fullres.dehaze_tiled(model, img, ps=128, index_samples=idx)
for _ in range(2): fwd()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fwd(); torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
tot = 0.0
agg = {}
for e in evs:
    us = e.device_time if hasattr(e, "device_time") else e.cuda_time
    tot += us
    nm = e.name.replace("void ", "").replace("lewin::", "")[:70]
    a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += us
    if "lewin" in e.name and us > 150:
        print(f"{us:9.1f} us  {nm}")
print(f"total kernel time {tot/1e3:.2f} ms")
for nm, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{us:9.1f} us x{n:3d}  {nm}")
