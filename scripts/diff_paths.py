"""GPU differential check + timing of alternative kernel paths behind the same ABI call.

Runs the attention half and the LeFF half of a LeWin block (bf16) in child processes that differ only in the
environment knobs selecting the kernel path (e.g. LEWIN_NO_WS_GEMM=1 = previous kernels), saves the outputs and
compares them; prints per-call CUDA-event times at a bench-sized shape.  Usage (on the GPU box):
    python scripts/diff_paths.py                 # parent: spawns the children, prints the comparison
"""
import os
import subprocess
import sys
import json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [  # (C, nH, B, H, W, shift)
    (32, 1, 1, 24, 24, 0), (32, 1, 3, 64, 64, 4), (64, 2, 2, 40, 24, 4), (64, 2, 1, 64, 64, 0),
    (128, 4, 2, 32, 32, 4), (128, 4, 1, 24, 40, 0), (256, 8, 2, 32, 32, 4), (512, 16, 3, 16, 16, 4), (256, 8, 1, 24, 24, 0),
]
TIMED = [(32, 1, 16, 128, 128, 4), (64, 2, 16, 128, 128, 4), (64, 2, 16, 64, 64, 4), (128, 4, 16, 64, 64, 4),
         (128, 4, 16, 32, 32, 4), (256, 8, 169, 16, 16, 4), (512, 16, 169, 16, 16, 4)]


def child(tag, out_dir, timed):
    import torch
    import lewin_b200 as L
    from lewin_b200 import ops
    dev = torch.device("cuda:0")
    res = {}
    for ci, (C, nH, B, H, W, shift) in enumerate(TIMED if timed else CASES):
        g = torch.Generator().manual_seed(100 + ci)
        r = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
        x = r(B, H * W, C).to(torch.bfloat16)
        p = dict(ln_w=1 + r(C, sc=0.1), ln_b=r(C, sc=0.1), w_qkv=r(3 * C, C, sc=C ** -0.5), b_qkv=r(3 * C, sc=0.1),
                 w_out=r(C, C, sc=C ** -0.5), b_out=r(C, sc=0.1), rpb_table=r(225, nH, sc=0.2))
        q = dict(ln_w=1 + r(C, sc=0.1), ln_b=r(C, sc=0.1), w1=r(4 * C, C, sc=C ** -0.5), b1=r(4 * C, sc=0.1),
                 w_dw=r(4 * C, 1, 3, 3, sc=0.3), b_dw=r(4 * C, sc=0.1), w2=r(C, 4 * C, sc=(4 * C) ** -0.5), b2=r(C, sc=0.1))
        idx = torch.randint(0, 64, (64, 25), generator=g)
        ds = (0.5 + torch.rand(B, generator=g)).to(dev)

        def run():
            with torch.no_grad():
                y, top = ops.lewin_attn(x, B=B, H=H, W=W, num_heads=nH, shift=shift, index_sample=idx, drop_scale=ds,
                                        return_top=True, **p)
                o = ops.lewin_leff(y, B=B, H=H, W=W, drop_scale=ds, **q)
            return y, top, o
        y, top, o = run()
        torch.cuda.synchronize()
        if timed:
            for _ in range(3):
                run()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            n = 10
            with torch.no_grad():
                torch.cuda.synchronize()
                ev[0].record()
                for _ in range(n):
                    y, top = ops.lewin_attn(x, B=B, H=H, W=W, num_heads=nH, shift=shift, index_sample=idx, drop_scale=ds,
                                            return_top=True, **p)
                ev[1].record()
                for _ in range(n):
                    o = ops.lewin_leff(y, B=B, H=H, W=W, drop_scale=ds, **q)
                ev[2].record()
            torch.cuda.synchronize()
            res[str(ci)] = dict(shape=[C, nH, B, H, W, shift], attn_us=1e3 * ev[0].elapsed_time(ev[1]) / n,
                                leff_us=1e3 * ev[1].elapsed_time(ev[2]) / n)
        else:
            torch.save(dict(y=y.float().cpu(), top=top.cpu(), o=o.float().cpu()), os.path.join(out_dir, f"{tag}_{ci}.pt"))
    if timed:
        print("TIMES " + tag + " " + json.dumps(res), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3], sys.argv[4] == "1")
        return
    import torch
    out_dir = "/tmp/diff_paths"
    os.makedirs(out_dir, exist_ok=True)
    variants = {"old": {"LEWIN_NO_WS_GEMM": "1", "LEWIN_NO_STREAM_DWCONV": "1", "LEWIN_NO_CORE_V3": "1"}, "new": {}, "oldcore": {"LEWIN_NO_CORE_V3": "1"}, "notma": {"LEWIN_NO_TMA": "1"}, "notmastore": {"LEWIN_NO_TMA_STORE": "1"}, "mmadw": {"LEWIN_MMA_DWCONV": "1"}, "nowcache": {"LEWIN_NO_WEIGHT_CACHE": "1"},
                "old_unfused": {"LEWIN_NO_WS_GEMM": "1", "LEWIN_NO_FUSED_LEFF": "1", "LEWIN_NO_STREAM_DWCONV": "1"}}
    sel = sys.argv[1:] or list(variants)
    for timed in ("0", "1"):
        for tag in sel:
            env = dict(os.environ, **variants[tag])
            try:
                r = subprocess.run([sys.executable, __file__, "--child", tag, out_dir, timed], env=env, timeout=240,
                                   capture_output=True, text=True)
                print(f"[{tag} timed={timed}] rc={r.returncode}", r.stdout[-1500:], r.stderr[-1500:], flush=True)
            except subprocess.TimeoutExpired:
                print(f"[{tag} timed={timed}] TIMEOUT (hang?)", flush=True)
    ref = sel[0]
    for tag in sel[1:]:
        for ci, case in enumerate(CASES):
            try:
                a = torch.load(os.path.join(out_dir, f"{ref}_{ci}.pt"))
                b = torch.load(os.path.join(out_dir, f"{tag}_{ci}.pt"))
            except FileNotFoundError:
                print(f"{tag} case {ci}: missing output")
                continue
            dy = (a["y"] - b["y"]).abs()
            do = (a["o"] - b["o"]).abs()
            same_top = bool((a["top"].sort(-1).values == b["top"].sort(-1).values).all())
            print(f"{ref} vs {tag} case {case}: attn max {dy.max():.4g} (frac>0: {(dy > 0).float().mean():.4f}) "
                  f"leff max {do.max():.4g} (frac>0: {(do > 0).float().mean():.4f}) |o|max {a['o'].abs().max():.3g} top-u equal: {same_top}")


if __name__ == "__main__":
    main()
