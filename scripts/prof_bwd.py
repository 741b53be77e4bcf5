"""Kernel-time breakdown of one LeWin block forward+backward at a training shape (GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lewin_b200 as L
from torch.profiler import profile, ProfilerActivity
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 128
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dt = torch.bfloat16 if (len(sys.argv) <= 4 or sys.argv[4] == "bf16") else torch.float32
dev = torch.device("cuda:0")
blk = L.LeWinTransformerBlock(dim=C, input_resolution=(hw, hw), num_heads=C // 32, win_size=8, shift_size=4 if hw > 8 else 0).to(dev)
x = torch.randn(B, hw * hw, C, device=dev, dtype=dt)
dy = torch.randn_like(x)
idx = torch.randint(64, (64, 25))
def step():
    xr = x.detach().requires_grad_(True)
    blk.zero_grad(set_to_none=True)
    blk(xr, None, idx).backward(dy)
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[ev.name[:100]]
        a[0] += 1; a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"C={C} map {hw}x{hw} batch {B}: forward + backward kernel time {tot:.1f} us, {sum(v[0] for v in agg.values())} kernels")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{us:9.1f} us  x{n:3d}  {name}")
