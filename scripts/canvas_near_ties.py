"""CPU: how many (window, head) rows of ONE canvas-mode forward (200 x 300 golden image, 384^2 canvas) are near-ties in the top-u
selection?  fp64 whole-model oracle, rank-25/26 gap of M relative to the row's M range.  Result (round 1): 24 of 26 208 rows are
below the fp32 tie threshold 1e-5 (4 below 1e-6, one at 1.2e-8), among them one bottleneck row (C = 512) and one C = 256 row -
any fp32 implementation with a different summation order flips some of them.  Usage: python scripts/canvas_near_ties.py"""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import param_fill, uformer_oracle as U, lewin_oracle as O
import lewin_b200 as L
from lewin_b200 import fullres
z = np.load(os.path.join(ROOT, "tests", "golden", "uformer32_canvas_200x300.npz"))
model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff")
param_fill.fill_module(model, int(z["seed"]))
sd = {k: v.numpy() for k, v in model.state_dict().items()}
canvas = fullres.wrap_pad(torch.from_numpy(z["x"]), ps=128).numpy()
gaps = []
orig = O.lewin_block
def spy(x, p, shift, idx, *a, **k):
    out, aux = orig(x, p, shift, idx, *a, return_aux=True, **k)
    gaps.append((x.shape, aux["rel_gap"].copy()))
    return out
U.O.lewin_block = spy
U.uformer_forward(canvas, sd, z["idx"].astype(np.int64), img_size=128, dtype=np.float64)
tot = sum(g.size for _, g in gaps)
for t in (1e-4, 1e-5, 1e-6, 1e-7):
    print(f"rows with rank-25/26 gap < {t:g} of the M range: {sum(int((g < t).sum()) for _, g in gaps)} of {tot}")
for i, (sh, g) in enumerate(gaps):
    if (g < 1e-5).any(): print("  block", i, "tokens x C", sh[1:], "rows < 1e-5:", int((g < 1e-5).sum()), "min gap", float(g.min()))
