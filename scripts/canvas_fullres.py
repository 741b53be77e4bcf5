"""Canvas mode at the reference's real size (test_long_GPU.py:74-93): ONE forward over the 1664^2 wrap-padded canvas of a
1200 x 1600 image on one GPU (43 264 windows at level 0).  Prints images/s (CUDA events), peak memory, and how far the
result is from tiled mode (different computations by construction, SURVEY finding 7)."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lewin_b200 as L
from lewin_b200 import fullres

dt = sys.argv[1] if len(sys.argv) > 1 else "bf16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(dev).eval()
g = torch.Generator().manual_seed(4321)
img = torch.rand(1, 3, 1200, 1600, generator=g).to(dev)
idx = model.draw_index_samples()

def run():
    if dt == "bf16":
        with torch.autocast("cuda", torch.bfloat16):
            return fullres.dehaze_canvas(model, img, ps=128, index_samples=idx)
    return fullres.dehaze_canvas(model, img, ps=128, index_samples=idx)

torch.cuda.reset_peak_memory_stats()
for _ in range(2):
    y = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    y = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
with torch.no_grad():
    if dt == "bf16":
        with torch.autocast("cuda", torch.bfloat16):
            yt = fullres.dehaze_tiled(model, img, ps=128, index_samples=idx)
    else:
        yt = fullres.dehaze_tiled(model, img, ps=128, index_samples=idx)
d = (y.float() - yt.float()).abs()
print(json.dumps(dict(mode="canvas 1664^2 (one forward)", dtype=dt, ms_per_image=ms, images_per_s=1e3 / ms,
                      peak_mem_gb=torch.cuda.max_memory_allocated() / 2**30,
                      vs_tiled_max_abs=float(d.max()), vs_tiled_mean_abs=float(d.mean()))))
