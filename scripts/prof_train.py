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
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=32, max_name_column_width=80))
