"""Diagnostic: canvas row bands vs single-device canvas forward, block by block (selection sets)."""
import os, sys, contextlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lewin_b200 as L
from lewin_b200 import canvas_bands, fullres, ops
from oracle import param_fill
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dt = sys.argv[2] if len(sys.argv) > 2 else "f32"
z = np.load("tests/golden/uformer32_canvas_200x300.npz")
dev = torch.device("cuda:0")
model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff")
param_fill.fill_module(model, int(z["seed"]))
model = model.to(dev).eval()
img = torch.from_numpy(z["x"]).to(dev)
idx = torch.from_numpy(z["idx"].astype(np.int64))
torch.backends.cudnn.allow_tf32 = False
with (torch.autocast("cuda", torch.bfloat16) if dt == "bf16" else contextlib.nullcontext()):
    with ops.TopRecorder() as a:
        ref = fullres.dehaze_canvas(model, img, ps=128, index_samples=idx)
    with ops.TopRecorder() as b:
        got = canvas_bands.dehaze_canvas_bands(model, img, ps=128, index_samples=idx, virtual_world=world)
for i in range(18):
    t_ref = np.sort(a.tops[i].cpu().numpy().astype(np.int64), -1)
    t_b = np.sort(torch.cat([b.tops[i * world + r] for r in range(world)], 0).cpu().numpy().astype(np.int64), -1)
    bad = (t_ref != t_b).any(-1)
    print(f"block {i:2d}: windows {t_ref.shape[0]:6d} heads {t_ref.shape[1]:2d}  rows differing {int(bad.sum()):5d}", (np.argwhere(bad)[:4].tolist() if bad.any() else ""))
d = (got.float() - ref.float()).abs()
print("max diff", float(d.max()), "frac > 1e-3", float((d > 1e-3).float().mean()))
