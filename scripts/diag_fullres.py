"""Diagnostic (GPU box): golden error distribution with TF32 convs off, and a kernel-time breakdown of the
169-tile full-resolution forward."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lewin_b200 as L
from oracle import param_fill

dev = torch.device("cuda:0")
z = np.load("tests/golden/uformer32_b2.npz")
model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff")
param_fill.fill_module(model, int(z["seed"]))
model = model.to(dev).eval()
x = torch.from_numpy(z["x"]).to(dev)
idx = torch.from_numpy(z["idx"].astype(np.int64))
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    with torch.no_grad():
        y = model(x, index_samples=idx).cpu().numpy()
    e = np.abs(y - z["y"]).ravel()
    print(f"tf32_convs={tf32}: max {e.max():.3e} median {np.median(e):.3e} p99 {np.percentile(e,99):.3e} p99.9 {np.percentile(e,99.9):.3e} frac>1e-3 {(e>1e-3).mean():.4f} mse {np.mean(e**2):.3e}")

dt = sys.argv[1] if len(sys.argv) > 1 else "f32"
tiles = torch.rand(169, 3, 128, 128, device=dev)
idx = model.draw_index_samples()
def run():
    with torch.no_grad():
        if dt == "bf16":
            with torch.autocast("cuda", torch.bfloat16):
                return model(tiles, index_samples=idx)
        return model(tiles, index_samples=idx)
torch.backends.cudnn.allow_tf32 = True
for _ in range(3): run()
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5): run()
t1.record(); torch.cuda.synchronize()
print(f"[{dt}] 169-tile forward: {t0.elapsed_time(t1)/5:.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
