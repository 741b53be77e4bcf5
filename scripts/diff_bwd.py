"""GPU differential check of backward kernel variants (bf16): parameter / input gradients of the attention half and the
LeFF half with environment knob A vs B (default: LEWIN_NO_WGRAD2=1 = previous weight-gradient kernel vs current)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = [(32, 1, 2, 32, 32, 4), (64, 2, 2, 32, 32, 4), (128, 4, 2, 32, 32, 4), (256, 8, 2, 16, 16, 4), (512, 16, 3, 16, 16, 4),
         (64, 2, 32, 128, 128, 4), (128, 4, 32, 64, 64, 4), (32, 1, 32, 128, 128, 4), (256, 8, 32, 32, 32, 4), (512, 16, 32, 16, 16, 4)]


def child(tag, out_dir):
    import torch
    from lewin_b200 import ops
    dev = torch.device("cuda:0")
    for ci, (C, nH, B, H, W, shift) in enumerate(CASES):
        g = torch.Generator().manual_seed(200 + ci)
        r = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
        x = r(B, H * W, C).to(torch.bfloat16).requires_grad_(True)
        p = dict(ln_w=1 + r(C, sc=0.1), ln_b=r(C, sc=0.1), w_qkv=r(3 * C, C, sc=C ** -0.5), b_qkv=r(3 * C, sc=0.1),
                 w_out=r(C, C, sc=C ** -0.5), b_out=r(C, sc=0.1), rpb_table=r(225, nH, sc=0.2))
        q = dict(ln_w=1 + r(C, sc=0.1), ln_b=r(C, sc=0.1), w1=r(4 * C, C, sc=C ** -0.5), b1=r(4 * C, sc=0.1),
                 w_dw=r(4 * C, 1, 3, 3, sc=0.3), b_dw=r(4 * C, sc=0.1), w2=r(C, 4 * C, sc=(4 * C) ** -0.5), b2=r(C, sc=0.1))
        for d in (p, q):
            for v in d.values():
                v.requires_grad_(True)
        idx = torch.randint(0, 64, (64, 25), generator=g)
        ds = (0.5 + torch.rand(B, generator=g)).to(dev)
        dout = r(B, H * W, C).to(torch.bfloat16)

        def run():
            y = ops.lewin_attn(x, B=B, H=H, W=W, num_heads=nH, shift=shift, index_sample=idx, drop_scale=ds, **p)
            o = ops.lewin_leff(y, B=B, H=H, W=W, drop_scale=ds, **q)
            o.backward(dout)
        run()
        torch.cuda.synchronize()
        grads = {"x": x.grad.float().cpu()}
        grads.update({"a." + k: v.grad.float().cpu() for k, v in p.items() if v.grad is not None})
        grads.update({"l." + k: v.grad.float().cpu() for k, v in q.items() if v.grad is not None})
        if B >= 32:
            for t in [x] + list(p.values()) + list(q.values()):
                t.grad = None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            run(); torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                run()
            e1.record(); torch.cuda.synchronize()
            print(f"TIME {tag} case {ci} {(C, nH, B, H, W)}: fwd+bwd {e0.elapsed_time(e1) / 5 * 1e3:.0f} us", flush=True)
            if os.environ.get("DIFF_BWD_PROFILE"):
                from torch.profiler import profile, ProfilerActivity
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    run(); torch.cuda.synchronize()
                print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70), flush=True)
        else:
            torch.save(grads, os.path.join(out_dir, f"{tag}_{ci}.pt"))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child(sys.argv[2], sys.argv[3])
    import torch
    out_dir = "/tmp/diff_bwd"
    os.makedirs(out_dir, exist_ok=True)
    variants = {"old": {k: "1" for k in os.environ.get("DIFF_OLD", "LEWIN_NO_BWD2,LEWIN_NO_WGRAD2,LEWIN_NO_CORE_BWD2").split(",") if k},
                "new": dict(kv.split("=") for kv in os.environ.get("DIFF_NEW", "").split(",") if kv)}
    for tag, envx in variants.items():
        r = subprocess.run([sys.executable, __file__, "--child", tag, out_dir], env=dict(os.environ, **envx), timeout=300,
                           capture_output=True, text=True)
        print(f"[{tag}] rc={r.returncode}", r.stdout[-9000:], r.stderr[-1500:], flush=True)
    for ci, case in enumerate(CASES):
        try:
            a = torch.load(os.path.join(out_dir, f"old_{ci}.pt")); b = torch.load(os.path.join(out_dir, f"new_{ci}.pt"))
        except FileNotFoundError:
            continue
        worst = max(((a[k] - b[k]).abs().max() / (a[k].abs().max() + 1e-12)).item() for k in a)
        detail = " ".join(f"{k}:{((a[k] - b[k]).abs().max() / (a[k].abs().max() + 1e-12)).item():.1e}" for k in a)
        print(f"case {case}: worst rel-to-max grad diff {worst:.2e} | {detail}")


if __name__ == "__main__":
    main()
