"""BASELINE configs 2 / 4: one Uformer_ProbSparse training step (batch 32 x 3 x 128 x 128 per GPU, bf16 autocast,
restored = clamp(model(input), 0, 1), loss = Charbonnier(eps 1e-3) + ContrastLoss(VGG19, random init: no weight file
offline), AdamW 2e-4 / wd 0.02; My_train.py:221-250) on the sm_100a LeWin ops; under torchrun the model is wrapped with
parallel.wrap_ddp (NCCL gradient all-reduce).  `contrast=False` drops the VGG term (the LeWin path alone).
Prints one JSON line (rank 0)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def charbonnier(x, y, eps=1e-3):
    d = x.float() - y.float()
    return torch.mean(torch.sqrt(d * d + eps * eps))


def run(batch=32, steps=10, warmup=3, dtype="bf16", graph=True, contrast=True):
    global torch
    import torch
    import torch.distributed as dist
    import lewin_b200 as L
    from lewin_b200 import parallel
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)
    model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(dev).train()
    parallel.freeze_dead_parameters(model)
    net = parallel.wrap_ddp(model, dev) if world > 1 else model
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.02)
    g = torch.Generator().manual_seed(1234 + rank)
    x = torch.rand(batch, 3, 128, 128, generator=g).to(dev)
    y = torch.rand(batch, 3, 128, 128, generator=g).to(dev)

    idx_static = model.draw_index_samples().to(dev, dtype=torch.int32)       # refreshed before every step (attn.py:91 draws)
    crit_cr = None
    if contrast:
        from lewin_b200.losses import ContrastLoss
        torch.manual_seed(0)
        crit_cr = ContrastLoss(ablation=False, pretrained=False, device=dev)  # frozen VGG19, seed 0

    def loss_fn():
        # My_train.py:224-238: forward, clamp and both criteria inside autocast; w_loss_* = 1 (options.py:16-17)
        with torch.autocast("cuda", torch.bfloat16, enabled=(dtype == "bf16")):
            restored = torch.clamp(net(x, index_samples=idx_static), 0, 1)
            loss = charbonnier(restored, y)
            if crit_cr is not None:
                loss = loss + crit_cr(restored, y, x)[0]
        return loss

    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn()
        loss.backward()
        opt.step()
        return loss

    eager_step = step
    mode = "eager"
    if graph and world == 1:
        # whole-step CUDA graph (forward + backward + AdamW): the deep levels are launch-bound in eager mode
        try:
            opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=2e-4, betas=(0.9, 0.999), eps=1e-8,
                                    weight_decay=0.02, capturable=True)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    eager_step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            gph = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(gph):
                loss_s = loss_fn()
                loss_s.backward()
                opt.step()

            def step():
                idx_static.copy_(model.draw_index_samples(), non_blocking=True)   # fresh key samples every step, as the reference
                gph.replay()
                return loss_s
            mode = "cuda-graph (forward + backward + optimizer)"
        except Exception as e:
            step = eager_step
            mode = "eager (graph capture failed: %s)" % repr(e)[:120]

    for _ in range(warmup):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res = dict(step_ms=float(ms.item()), patches_per_s=batch * world / (float(ms.item()) / 1e3), batch_per_gpu=batch, n_gpus=world,
               dtype=dtype, loss=float(loss.item()),
               loss_terms="Charbonnier + ContrastLoss (VGG19 random init, p / n passes batched under no_grad, channels-last bf16)"
               if contrast else "Charbonnier only",
               optimizer="AdamW(2e-4, wd 0.02)", parallelism="ddp" if world > 1 else "single", launch=mode)
    return res, rank


if __name__ == "__main__":
    r, rank = run(batch=int(sys.argv[1]) if len(sys.argv) > 1 else 32)
    if rank == 0:
        print(json.dumps(r), flush=True)
