import sys, torch
sys.path.insert(0, "/root/repo")
import lewin_b200 as L
from lewin_b200 import uformer as U
dev = torch.device("cuda:0")
m = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(dev).eval()
x = torch.rand(8, 3, 128, 128, device=dev)
orig = U._nchw_to_tokens
def probe(y):
    print("conv out", tuple(y.shape), y.dtype, "channels_last:", y.is_contiguous(memory_format=torch.channels_last), "contig:", y.is_contiguous())
    return orig(y)
U._nchw_to_tokens = probe
with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
    m(x)
