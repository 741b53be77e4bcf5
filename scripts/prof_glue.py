"""Which aten ops (with shapes) launch the non-LeWin kernels of the full-resolution bf16 forward — GPU box."""
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    fwd(); torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.device_time_total > 20 and e.key.startswith("aten::")]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:22]:
    print(f"{e.key:34s} {e.device_time_total:9.1f} us  x{e.count:3d}  {str(e.input_shapes)[:110]}")
