"""Canvas mode over N GPUs by row bands (torchrun): result vs the single-GPU canvas forward, and images/s of both.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/canvas_bands_multi.py [bf16|f32]"""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lewin_b200 as L
from lewin_b200 import canvas_bands, fullres

dt = sys.argv[1] if len(sys.argv) > 1 else "bf16"
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(1234)
model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(dev).eval()
g = torch.Generator().manual_seed(4321)
img = torch.rand(1, 3, 1200, 1600, generator=g).to(dev)
idx = model.draw_index_samples()
ac = (lambda: torch.autocast("cuda", torch.bfloat16)) if dt == "bf16" else (lambda: torch.autocast("cuda", enabled=False))

def timed(fn, n):
    for _ in range(2):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms), out

use_graph = len(sys.argv) > 2 and sys.argv[2] == "graph"      # NCCL point-to-point inside a captured graph: opt-in
with ac():
    ms_b, y_b = timed(lambda: canvas_bands.dehaze_canvas_bands(model, img, index_samples=idx), 5)
    ms_1, y_1 = timed(lambda: fullres.dehaze_canvas(model, img, index_samples=idx), 3)      # every rank, redundantly
d = (y_b.float() - y_1.float()).abs()
res = dict(mode="canvas 1664^2 by row bands", n_gpus=world, dtype=dt, bands_ms=ms_b, bands_images_per_s=1e3 / ms_b,
           single_gpu_ms=ms_1, speedup=ms_1 / ms_b, max_abs_vs_single=float(d.max()), frac_gt_2e2=float((d > 2e-2).float().mean()),
           units_per_rank=[canvas_bands.band_units(13, r, world)[1] - canvas_bands.band_units(13, r, world)[0] for r in range(world)])
if rank == 0:
    print(json.dumps(res), flush=True)
if use_graph:
    gb = canvas_bands.GraphedCanvasBands(model, img, idx.to(dev), torch.bfloat16 if dt == "bf16" else None)
    ms_g, y_g = timed(lambda: gb(img, idx.to(dev)), 10)
    if rank == 0:
        print(json.dumps(dict(graph_ms=ms_g, graph_images_per_s=1e3 / ms_g, graph_captured=gb.graph is not None,
                              graph_error=getattr(gb, "error", None), graph_equal_single=bool(torch.equal(y_g, y_1)))), flush=True)
if world > 1:
    dist.destroy_process_group()
