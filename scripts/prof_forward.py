"""Kernel-time breakdown of one full-resolution bf16 tiled forward (169 tiles) — GPU box."""
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
    with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
        return fullres.dehaze_tiled(model, img, ps=128, index_samples=idx)
for _ in range(3): fwd()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fwd(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
