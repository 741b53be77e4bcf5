"""Generate tests/golden/*.npz by running the UNMODIFIED reference here (build container).

TEST INFRASTRUCTURE ONLY.  Run:  python oracle/make_golden.py
It imports /root/reference/Uformer_ProbSparse/My_model_1.py through oracle/ref_shim.py,
runs LeWinTransformerBlock / Uformer forward (+ backward) on CPU in fp32 with fixed seeds,
records the exact ``index_sample`` draws (attn.py:91), the selected ``M_top`` (attn.py:122)
and DropPath masks, cross-checks the numpy oracle against the reference, and writes small
fixtures.  The GPU box has no reference tree; it only reads the committed fixtures.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import lewin_oracle as O          # noqa: E402
from oracle import param_fill, ref_shim       # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class Recorder:
    """Records torch.randint draws, ProbAttention top-u indices and DropPath masks."""

    def __init__(self, ref_mod):
        import ProbSparse.attn as attn_mod
        self.attn_mod = attn_mod
        self.idx, self.top, self.drop = [], [], []
        self._randint = torch.randint
        self._prob_qk = attn_mod.ProbAttention._prob_QK

    def __enter__(self):
        rec = self

        def randint(*a, **k):
            t = rec._randint(*a, **k)
            rec.idx.append(t.clone())
            return t

        def prob_qk(self_, Q, K, sample_k, n_top):
            qk, top = rec._prob_qk(self_, Q, K, sample_k, n_top)
            rec.top.append(top.clone())
            return qk, top

        torch.randint = randint
        self.attn_mod.ProbAttention._prob_QK = prob_qk
        return self

    def __exit__(self, *exc):
        torch.randint = self._randint
        self.attn_mod.ProbAttention._prob_QK = self._prob_qk


def _hook_droppath(block, store):
    from timm.models.layers import DropPath
    if not isinstance(block.drop_path, DropPath):
        return
    orig = block.drop_path.forward

    def fwd(x):
        if block.drop_path.drop_prob == 0.0 or not block.drop_path.training:
            return x
        keep = 1.0 - block.drop_path.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep).div_(keep)
        store.append(mask.reshape(-1).clone())
        return x * mask

    block.drop_path.forward = fwd
    return orig


BLOCK_CASES = [
    # name, C, nH, map, B, shift, input_mask, drop_path, train
    ("block_c32_h1_s0", 32, 1, 16, 2, 0, False, 0.0, False),
    ("block_c32_h1_s4", 32, 1, 16, 2, 4, False, 0.0, False),
    ("block_c64_h2_s4", 64, 2, 16, 2, 4, False, 0.0, False),
    ("block_c64_h2_s4_inmask", 64, 2, 24, 1, 4, True, 0.0, False),
    ("block_c64_h2_s4_droppath", 64, 2, 16, 4, 4, False, 0.4, True),
    ("block_c128_h4_s0", 128, 4, 8, 3, 0, False, 0.0, False),
]


# Deep levels (C = 256, 512): the same recording, stored compactly - parameters are regenerated from the seed
# (param_fill.fill_value) and every parameter gradient is kept as its L2 norm plus a seeded sample of <= 4096 elements.
COMPACT_CASES = [
    ("block_c256_h8_s4_compact", 256, 8, 16, 1, 4, False, 0.0, False),
    ("block_c512_h16_s0_compact", 512, 16, 8, 2, 0, False, 0.0, False),
]


# head_dim = embed_dim 64 / 128 (My_model_1.py:962 makes head_dim = embed_dim; BASELINE config 5 sweeps embed_dim 32-128):
# the same compact recording for a one-head C = 64 block, a two-head C = 128 block (head_dim 64) and a one-head C = 128 block
# (head_dim 128), so those kernel instances are pinned to the unmodified reference as well, forward and backward.
HEAD_DIM_CASES = [
    ("block_c64_h1_s4_hd64_compact", 64, 1, 16, 2, 4, False, 0.0, False),
    ("block_c128_h2_s0_hd64_compact", 128, 2, 8, 3, 0, False, 0.0, False),
    ("block_c128_h1_s4_hd128_compact", 128, 1, 16, 1, 4, False, 0.0, False),
]


def grad_sample_ids(key, n, seed, k=4096):
    import zlib
    rng = np.random.default_rng([seed, zlib.crc32(key.encode()), 7])
    return np.sort(rng.choice(n, size=min(n, k), replace=False))


def make_block_case(ref, name, C, nH, hw, B, shift, use_inmask, drop_path, train, seed, compact=False):
    torch.manual_seed(seed)
    blk = ref.LeWinTransformerBlock(dim=C, input_resolution=(128, 128), num_heads=nH, win_size=8,
                                    shift_size=shift, token_mlp="leff", drop_path=drop_path)
    param_fill.fill_module(blk, seed)
    blk.train(train)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, hw * hw, C, generator=g, requires_grad=True)
    dout = torch.randn(B, hw * hw, C, generator=g)
    inmask = None
    if use_inmask:
        inmask = torch.zeros(1, 1, hw * 2, hw * 2)
        inmask[:, :, -10:, :] = 1.0
        inmask[:, :, :, -6:] = 1.0
    drops = []
    _hook_droppath(blk, drops)
    torch.manual_seed(seed + 2)
    with Recorder(ref) as rec:
        out = blk(x, inmask)
    out.backward(dout)
    assert len(rec.idx) == 1 and len(rec.top) == 1
    idx = rec.idx[0].numpy()
    top = np.sort(rec.top[0].numpy(), -1)
    drop_scale = np.stack([d.numpy() for d in drops]) if drops else None

    p = {k: v.detach().numpy() for k, v in blk.state_dict().items()}
    grads = {k: v.grad.detach().numpy() for k, v in blk.named_parameters() if v.grad is not None}
    dead = sorted(k for k, v in blk.named_parameters() if v.grad is None)
    assert sorted(grads) == sorted(O.GRAD_KEYS), (sorted(grads), dead)

    # ---- cross-check the numpy oracle against the reference (fp32 and fp64)
    xn, don = x.detach().numpy(), dout.numpy()
    inm = None if inmask is None else inmask.numpy()
    report = {}
    for dt, tol in ((np.float32, 2e-4), (np.float64, 2e-4)):
        pp = O.as_dtype(p, dt)
        o, aux = O.lewin_block(xn.astype(dt), pp, shift, idx, inm, True, drop_scale, return_aux=True)
        same_top = np.array_equal(aux["top"], top)
        if not same_top:   # near-tie rows: evaluate the oracle with the reference's selection
            nbad = int((aux["top"] != top).any(-1).sum())
            print(f"  [{name}/{dt.__name__}] {nbad} (window,head) rows differ in top-u (near ties); forcing reference selection")
            o = O.lewin_block(xn.astype(dt), pp, shift, idx, inm, True, drop_scale, top=top)
        err = np.abs(o - out.detach().numpy()).max()
        dx, go = O.lewin_block_bwd(don.astype(dt), xn.astype(dt), pp, shift, idx, inm, True, drop_scale, top=top)
        # key_projection.bias has a mathematically zero gradient (softmax shift invariance): floor the scale
        gscale = max(np.abs(grads[k]).max() for k in O.GRAD_KEYS)
        gerr = max(np.abs(go[k] - grads[k]).max() / max(np.abs(grads[k]).max(), 1e-4 * gscale) for k in O.GRAD_KEYS)
        dxerr = np.abs(dx - x.grad.numpy()).max() / np.abs(x.grad.numpy()).max()
        report[dt.__name__] = (err, dxerr, gerr, same_top)
        assert err < tol * max(1.0, np.abs(out.detach().numpy()).max()), (name, dt, err)
        assert dxerr < 1e-3 and gerr < 1e-3, (name, dt, dxerr, gerr)
    print(f"{name}: " + "  ".join(f"{k}: out {v[0]:.2e} dx {v[1]:.2e} dparam {v[2]:.2e} top_equal={v[3]}" for k, v in report.items()))

    save = dict(x=xn, dout=don, out=out.detach().numpy(), dx=x.grad.numpy(), idx=idx.astype(np.int64),
                top=top.astype(np.int64), shift=np.int64(shift), nH=np.int64(nH), hw=np.int64(hw),
                seed=np.int64(seed))
    if inm is not None:
        save["input_mask"] = inm
    if drop_scale is not None:
        save["drop_scale"] = drop_scale
    if compact:
        for k, v in p.items():                 # the stored seed must reproduce the reference block's parameters exactly
            if np.issubdtype(v.dtype, np.floating):
                assert np.array_equal(v, param_fill.fill_value(k, v.shape, seed)), k
        save["params_from_seed"] = np.int64(1)
        for k, v in grads.items():
            save["gn:" + k] = np.float64(np.linalg.norm(v.astype(np.float64)))
            save["gs:" + k] = v.ravel()[grad_sample_ids(k, v.size, seed)]
    else:
        for k, v in p.items():
            save["p:" + k] = v
        for k, v in grads.items():
            save["g:" + k] = v
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **save)


# bf16: the reference itself under torch.autocast("cpu", dtype=torch.bfloat16) - the only way to run the UNMODIFIED
# reference at BASELINE's compute dtype in this container (no GPU).  CPU autocast keeps LayerNorm / softmax outputs in bf16
# where CUDA autocast keeps them fp32, so agreement with the CUDA-semantics bf16 oracle is expected to one bf16 ulp, not
# bit for bit; the top-u sets may differ on rows whose rank-25/26 gap is below the bf16 resolution (2^-7 of the M range).
BF16_CASES = [
    ("block_c64_h2_s4_bf16cpu", 64, 2, 16, 2, 4),
    ("block_c256_h8_s4_bf16cpu", 256, 8, 16, 1, 4),
]


def make_bf16_case(ref, name, C, nH, hw, B, shift, seed):
    torch.manual_seed(seed)
    blk = ref.LeWinTransformerBlock(dim=C, input_resolution=(128, 128), num_heads=nH, win_size=8,
                                    shift_size=shift, token_mlp="leff", drop_path=0.0)
    param_fill.fill_module(blk, seed)
    blk.eval()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, hw * hw, C, generator=g).to(torch.bfloat16).requires_grad_(True)
    dout = torch.randn(B, hw * hw, C, generator=g).to(torch.bfloat16)
    torch.manual_seed(seed + 2)
    with Recorder(ref) as rec, torch.autocast("cpu", dtype=torch.bfloat16):
        out = blk(x)
    out.backward(dout)                                       # fp32 parameter gradients, bf16 dx (autocast backward)
    out = out.detach()
    assert out.dtype == torch.bfloat16 and len(rec.idx) == 1
    idx, top = rec.idx[0].numpy(), np.sort(rec.top[0].numpy(), -1)
    p = {k: v.detach().numpy() for k, v in blk.state_dict().items()}
    for k, v in p.items():
        if np.issubdtype(v.dtype, np.floating):
            assert np.array_equal(v, param_fill.fill_value(k, v.shape, seed)), k
    xn, on = x.detach().float().numpy(), out.float().numpy()
    grads = {k: v.grad.detach().float().numpy() for k, v in blk.named_parameters() if v.grad is not None}
    assert sorted(grads) == sorted(O.GRAD_KEYS)
    o, aux = O.lewin_block(xn, O.as_dtype(p, np.float32), shift, idx, None, True, None, return_aux=True, bf16=True)
    bad = (aux["top"] != top).any(-1)
    assert (aux["rel_gap"][bad] < 2.0 ** -7).all(), name
    o2 = O.lewin_block(xn, O.as_dtype(p, np.float32), shift, idx, None, True, None, top=top, bf16=True)
    # one bf16 ulp at the magnitude of the block's activations (the residual sums cancel, so an element's own magnitude
    # is not the scale of its rounding error)
    ulp = 2.0 ** (np.floor(np.log2(np.abs(on).max())) - 7)
    d = np.abs(o2 - on)
    print(f"{name}: top-u rows differing {int(bad.sum())}/{bad.size} (all below the bf16 tie threshold); oracle(bf16) vs reference: "
          f"max {d.max():.4f} (ulp at |out|max = {ulp:.4f}), mean {d.mean():.2e}, {100 * (d > 0).mean():.1f} % of elements differ")
    assert d.max() <= ulp and d.mean() < 2e-3, name
    # backward: the fp64 oracle backward (reference selection) must describe the reference's autocast gradients up to bf16 noise
    dx64, g64 = O.lewin_block_bwd(dout.float().numpy().astype(np.float64), xn.astype(np.float64), O.as_dtype(p, np.float64),
                                  shift, idx, None, True, None, top=top)
    gscale = max(np.abs(v).max() for v in g64.values())
    worst = 1.0
    for k in O.GRAD_KEYS:
        if np.abs(g64[k]).max() < 1e-9 * gscale:
            continue
        a, b = grads[k].ravel().astype(np.float64), g64[k].ravel()
        worst = min(worst, float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b))))
    print(f"   backward: min cosine(reference autocast grads, fp64 oracle grads) over the parameters = {worst:.5f}")
    assert worst > 0.999, name
    save = dict(x=xn, out=on, dout=dout.float().numpy(), dx=x.grad.float().numpy(), idx=idx.astype(np.int64),
                top=top.astype(np.int64), shift=np.int64(shift), nH=np.int64(nH), hw=np.int64(hw), seed=np.int64(seed),
                params_from_seed=np.int64(1))
    for k, v in grads.items():
        save["gn:" + k] = np.float64(np.linalg.norm(v.astype(np.float64)))
        save["gs:" + k] = v.ravel()[grad_sample_ids(k, v.size, seed)]
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **save)


def make_model_case(ref, name, B, seed, mask=False, embed_dim=32):
    """Uformer(img_size=128, embed_dim=32) forward, config 1 of BASELINE.json (B tiles); embed_dim 64: the head_dim 64 variant
    (My_model_1.py:962), C = 64 ... 1024."""
    torch.manual_seed(seed)
    model = ref.Uformer(img_size=128, embed_dim=embed_dim, win_size=8, token_projection="linear", token_mlp="leff")
    param_fill.fill_module(model, seed)
    model.eval()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(B, 3, 128, 128, generator=g)
    torch.manual_seed(seed + 2)
    with Recorder(ref) as rec, torch.no_grad():
        y = model(x)
    assert len(rec.idx) == 18
    keys = sorted(model.state_dict().keys())
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        x=x.numpy(), y=y.numpy(), idx=np.stack([i.numpy() for i in rec.idx]).astype(np.int8),
        seed=np.int64(seed), n_keys=np.int64(len(keys)),
        key_crc=np.int64(__import__("zlib").crc32("\n".join(
            f"{k}:{tuple(model.state_dict()[k].shape)}" for k in keys).encode())))
    if embed_dim == 32:
        with open(os.path.join(GOLD, "uformer32_state_dict_keys.txt"), "w") as f:
            for k in keys:
                f.write(f"{k} {tuple(model.state_dict()[k].shape)} {str(model.state_dict()[k].dtype).replace('torch.', '')}\n")
    print(f"{name}: out range [{y.min():.3f}, {y.max():.3f}], |y-x| max {np.abs((y - x).numpy()).max():.3f}")


def make_canvas_case(ref, name, H, W, seed, ps=128):
    """Canvas mode = the reference's full-resolution computation, test_long_GPU.py:74-93, restated line by line around the
    unmodified model (the script itself is not importable): wrap-pad to L = (max(H, W) // ps + 1) * ps (the script hard-codes
    1664 for its 1200 x 1600 inputs, :82), ONE forward over the canvas, crop, clamp."""
    torch.manual_seed(seed)
    model = ref.Uformer(img_size=ps, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff")
    param_fill.fill_module(model, seed)
    model.eval()
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.rand(1, 3, H, W, generator=g)
    B, C = 1, 3
    L = (max(H, W) // ps + 1) * ps                                    # :80-81
    L_H, L_W = L - H, L - W                                           # :83-84
    big = torch.zeros((B, C, L, L))                                   # :86
    big[:, :, :H, :W] = img[:, :, :H, :W]                             # :87
    big[:, :, :H, W:W + L_W] = img[:, :, :, :L_W]                     # :88
    big[:, :, H:H + L_H, :] = big[:, :, :L_H, :]                      # :89
    torch.manual_seed(seed + 2)
    with Recorder(ref) as rec, torch.no_grad():
        restored = model(big)                                         # :91
    raw = restored[:, :, :H, :W].clone()                              # unclamped crop: the random-init model saturates 79 % of the pixels
    restored = torch.clamp(restored[:, :, :H, :W], 0, 1)              # :92-93
    assert len(rec.idx) == 18
    idx = np.stack([i.numpy() for i in rec.idx])
    from oracle import uformer_oracle as U
    sd = {k: v.numpy() for k, v in model.state_dict().items()}
    rec_o = []
    yo_raw = U.uformer_forward(big.numpy(), sd, idx, img_size=ps, dtype=np.float64, record=rec_o)[:, :, :H, :W]
    # per-block selection check: the oracle's top-u sets against the reference's M_top, block by block
    ref_tops = [np.sort(t.numpy(), -1) for t in rec.top]
    assert len(ref_tops) == len(rec_o) == 18
    nrows = ndiff = 0
    for r, t in zip(rec_o, ref_tops):
        bad = (r["top"] != t).any(-1)
        assert (r["rel_gap"][bad] < 1e-5).all(), (name, r["block"])
        nrows += bad.size; ndiff += int(bad.sum())
    print(f"{name}: top-u sets, oracle vs reference over the 18 blocks: {ndiff} of {nrows} rows differ (all near-ties)")
    er = np.abs(yo_raw - raw.numpy())
    print(f"{name}: raw output range [{raw.min():.2f}, {raw.max():.2f}]; numpy oracle vs reference (raw): max {er.max():.2e}, "
          f"median {np.median(er):.2e}, frac > 1e-3 {(er > 1e-3).mean():.4f}")
    yo = np.clip(yo_raw, 0, 1)
    e = np.abs(yo - restored.numpy())
    print(f"{name}: canvas {L}^2; numpy oracle vs reference: max {e.max():.2e}, median {np.median(e):.2e}, frac > 1e-3 {(e > 1e-3).mean():.4f}")
    assert np.median(e) < 1e-4 and (e > 1e-3).mean() < 0.02
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), x=img.numpy(), y=restored.numpy(), y_raw=raw.numpy(),
                        **{f"top{i:02d}": t.astype(np.int8) for i, t in enumerate(ref_tops)}, idx=idx.astype(np.int8), seed=np.int64(seed))


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_shim.import_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    only = os.environ.get("GOLDEN_ONLY", "")            # e.g. GOLDEN_ONLY=compact regenerates just the compact cases
    if only in ("", "blocks"):
        for i, case in enumerate(BLOCK_CASES):
            make_block_case(ref, *case, seed=100 + 10 * i)
    if only in ("", "compact"):
        for i, case in enumerate(COMPACT_CASES):
            make_block_case(ref, *case, seed=500 + 10 * i, compact=True)
    if only in ("", "headdim"):
        for i, case in enumerate(HEAD_DIM_CASES):
            make_block_case(ref, *case, seed=700 + 10 * i, compact=True)
    if only in ("", "bf16"):
        for i, case in enumerate(BF16_CASES):
            make_bf16_case(ref, *case, seed=900 + 10 * i)
    if only in ("", "canvas"):
        make_canvas_case(ref, "uformer32_canvas_200x300", 200, 300, seed=4321)
    if only in ("", "model"):
        make_model_case(ref, "uformer32_b2", B=2, seed=1234)
    if only in ("", "model64"):
        make_model_case(ref, "uformer64_b1", B=1, seed=2345, embed_dim=64)


if __name__ == "__main__":
    main()
