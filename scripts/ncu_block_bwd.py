"""One LeWin block forward + backward at a training shape (batch 32 unless given), for `ncu --set full` captures of the backward
kernels (GPU box):
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 -o gpurun_out/<name> python scripts/ncu_block_bwd.py dec3
Levels as bench.LEVELS (name, C, map)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lewin_b200 as L
LEVELS = {"enc0": (32, 128), "enc1": (64, 64), "enc2": (128, 32), "enc3": (256, 16), "bottleneck": (512, 8),
          "dec0": (512, 16), "dec1": (256, 32), "dec2": (128, 64), "dec3": (64, 128)}
name = sys.argv[1] if len(sys.argv) > 1 else "dec3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
C, hw = LEVELS[name]
dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = L.LeWinTransformerBlock(dim=C, input_resolution=(hw, hw), num_heads=C // 32, win_size=8, shift_size=4 if hw > 8 else 0).to(dev).train()
x = torch.randn(B, hw * hw, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
dout = torch.randn(B, hw * hw, C, device=dev, dtype=torch.bfloat16)
idx = torch.randint(64, (64, 25))
for _ in range(2):
    blk(x, None, idx).backward(dout)
torch.cuda.synchronize()
print("ok", name, B, float(x.grad.float().abs().mean()))
