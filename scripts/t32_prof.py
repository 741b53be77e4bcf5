"""Per-role cycle counters of the fp32 tcgen05 GEMMs (LEWIN_T32_PROF=1) for one LeWin block at a benchmark shape - GPU box.
usage (library built with -DLEWIN_T32_PROF_BUILD added to NVCC_FLAGS in __graft_entry__.py): LEWIN_T32_PROF=1 python scripts/t32_prof.py dec3 [tiles]"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lewin_b200 as L
from lewin_b200 import _lib
LEVELS = {"enc0": (32, 128), "enc1": (64, 64), "enc2": (128, 32), "enc3": (256, 16), "bottleneck": (512, 8),
          "dec0": (512, 16), "dec1": (256, 32), "dec2": (128, 64), "dec3": (64, 128)}
lib = _lib.load()
lib.lewin_debug_t32_prof.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(128, dtype=np.uint64)
dev = torch.device("cuda:0")
for name in sys.argv[1].split(","):
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 169
    C, hw = LEVELS[name]
    torch.manual_seed(0)
    blk = L.LeWinTransformerBlock(dim=C, input_resolution=(hw, hw), num_heads=C // 32, win_size=8, shift_size=4 if hw > 8 else 0).to(dev).eval()
    x = torch.randn(B, hw * hw, C, device=dev)
    idx = torch.randint(64, (64, 25))
    with torch.no_grad():
        blk(x, None, idx)
        lib.lewin_debug_t32_prof(buf.ctypes.data, 1)
        blk(x, None, idx)
    assert lib.lewin_debug_t32_prof(buf.ctypes.data, 1) == 1, "run with LEWIN_T32_PROF=1"
    print(f"{name}: C={C}, tokens={B * hw * hw}")
    for i, g in enumerate(["qkv", "out", "fc1", "fc2"]):
        v = buf[16 * i:16 * i + 16].astype(np.float64)
        n = max(v[10], 1)
        print(f"  {g}: CTA cycles {v[8]/n:9.0f} | producer total {v[0]/n:9.0f} (table barrier {v[1]/n:8.0f}, wait free stage {v[2]/n:8.0f}), stages/CTA {v[3]/n:6.1f}"
              f" | MMA wait accumulator {v[5]/n:8.0f}, wait stage {v[6]/n:8.0f}, tiles/CTA {v[7]/n:6.1f} | epilogue wait accumulator {v[9]/n:8.0f}")
