"""Small single-op workloads for ncu captures (GPU box): one LeWin block at a given level, bf16 or f32."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lewin_b200 as L

C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
hw = int(sys.argv[3]) if len(sys.argv) > 3 else 128
dt = sys.argv[4] if len(sys.argv) > 4 else "bf16"
dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = L.LeWinTransformerBlock(dim=C, input_resolution=(128, 128), num_heads=C // 32, win_size=8, shift_size=4).to(dev).eval()
for p in blk.parameters():
    torch.nn.init.normal_(p, std=0.1)
x = torch.randn(B, hw * hw, C, device=dev, dtype=torch.bfloat16 if dt == "bf16" else torch.float32)
idx = torch.randint(64, (64, 25))
with torch.no_grad():
    for _ in range(3):
        y = blk(x, None, idx)
torch.cuda.synchronize()
print("ok", y.float().abs().mean().item())
