"""Experiment: per-rank forward throughput with ONE vs L images in flight (L CUDA graphs replayed on L streams; L = argv[1],
default 2), for several tile counts (169 = the 1-GPU step, 85 / 43 / 22 = the 2 / 4 / 8-GPU shards) - GPU box.
Result on one B200 (L = 2): +3.6 % at 169 tiles, +4.9 % at 85, +6.8 % at 43, +9.3 % at 22 -> fullres.TiledPipeline lanes."""
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
N = 12
LANES = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for T in (169, 85, 43, 22):
    x = torch.rand(T, 3, 128, 128, device=dev)
    gs = [fullres.GraphedForward(model, x, idx, torch.bfloat16) for _ in range(LANES)]
    ss = [torch.cuda.Stream(device=dev) for _ in range(LANES)]
    for g in gs:
        for _ in range(2): g(x, idx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(N): gs[0].graph.replay()
    e1.record(); torch.cuda.synchronize()
    ms1 = e0.elapsed_time(e1) / N
    cur = torch.cuda.current_stream(dev)
    e0.record()
    for s in ss: s.wait_stream(cur)
    for i in range(N):
        with torch.cuda.stream(ss[i % LANES]):
            gs[i % LANES].graph.replay()
    for s in ss: cur.wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / N
    print(f"tiles {T:4d}: one in flight {ms1:7.3f} ms / image-shard, {LANES} in flight {ms2:7.3f} ms  ({ms1 / ms2:5.3f}x)", flush=True)
