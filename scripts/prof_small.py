"""Where the per-rank forward time goes at the 8-GPU shard size (22 tiles): kineto kernel table of CUDA-graph replays
(device durations, no ncu serialisation) + sum of kernel time vs the replay's elapsed time (launch gaps) — GPU box."""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lewin_b200 as L
from lewin_b200 import fullres
from torch.profiler import profile, ProfilerActivity

T = int(sys.argv[1]) if len(sys.argv) > 1 else 22
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(dev).eval()
idx = model.draw_index_samples()
x = torch.rand(T, 3, 128, 128, device=dev)
g = fullres.GraphedForward(model, x, idx, torch.bfloat16)
for _ in range(3):
    g(x, idx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    g.graph.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
N = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        g.graph.replay()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[ev.name[:90]]
        a[0] += 1; a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values()) / N
print(f"tiles {T}: replay {ms * 1e3:.1f} us; sum of kernel durations {tot:.1f} us per replay; {sum(v[0] for v in agg.values()) // N} kernels per replay")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{us / N:9.1f} us  x{n // N:3d}  {name}")
