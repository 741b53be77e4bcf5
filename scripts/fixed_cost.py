"""Per-rank forward time vs tile count (graph replay): how much of the step is fixed cost — GPU box."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lewin_b200 as L
from lewin_b200 import fullres
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(dev).eval()
idx = model.draw_index_samples()
for T in (169, 85, 43, 22, 8):
    x = torch.rand(T, 3, 128, 128, device=dev)
    g = fullres.GraphedForward(model, x, idx, torch.bfloat16)
    for _ in range(3): g(x, idx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.graph.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"tiles {T:4d}: {ms:7.3f} ms  ({ms / T * 1e3:6.1f} us per tile)")
