"""CPU oracle for the LeWin hot path (ProbSparse window attention + LeFF).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker / CPU baseline.

This is a numpy restatement of the reference algorithm (no torch autograd, no
shared code with the CUDA path).  Every function cites the reference file:line it
follows; paths are relative to ``/root/reference/Uformer_ProbSparse``.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4).  The restatement is pinned against the reference itself,
executed in the build container by ``oracle/make_golden.py`` (fixtures committed
under ``tests/golden/``) and re-checked by ``tests/test_oracle_golden.py``: fp32 blocks
(shift, input mask, DropPath, C = 32 ... 512) forward + backward, the whole model on
tiles and in canvas mode, and - for the ``bf16=True`` emulation of the CUDA autocast
rounding points - the reference run under ``torch.autocast("cpu", bfloat16)`` (agreement
to one bf16 ulp of the activation scale, tie-aware top-u; the CPU autocast policy is not
bit-identical to the CUDA one).

Conventions: tokens are channel-last, ``x[B, L=H*W, C]``; windows are 8x8 (N=64);
``idx`` is the reference's ``index_sample`` int array ``[64, sample_k]`` drawn by the
caller with ``torch.randint(64, (64, 25))`` (attn.py:91).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import erf as _erf

WIN = 8
N_TOK = WIN * WIN


# --------------------------------------------------------------------------- bf16 emulation
def rbf(x, on=True):
    """Round-to-nearest-even to bfloat16 (returned in the input float dtype).  Used with ``bf16=True`` to
    restate the rounding points of the reference under torch.autocast(cuda, bfloat16):
    linear / matmul / conv outputs and their operands are bf16, layer_norm / softmax / sums are fp32."""
    if not on:
        return x
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    u = a.view(np.uint32)
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    out = ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)
    return out.astype(np.asarray(x).dtype) if np.asarray(x).dtype != np.float32 else out


# --------------------------------------------------------------------------- basics
def prob_sizes(L_K: int = N_TOK, L_Q: int = N_TOK, factor: int = 5):
    """U_part and u, attn.py:310-315 (both 25 for 8x8 windows)."""
    U_part = factor * int(np.ceil(np.log(L_K)))
    u = factor * int(np.ceil(np.log(L_Q)))
    return min(U_part, L_K), min(u, L_Q)


def layer_norm(x, g, b, eps=1e-5):
    """nn.LayerNorm(C), My_model_1.py:769,776,839,873 (biased variance, eps 1e-5)."""
    mu = x.mean(-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    return xc * rstd * g + b


def layer_norm_bwd(dy, x, g, eps=1e-5):
    """Gradient of layer_norm wrt (x, g, b)."""
    C = x.shape[-1]
    mu = x.mean(-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xh = xc * rstd
    dg = (dy * xh).reshape(-1, C).sum(0)
    db = dy.reshape(-1, C).sum(0)
    dxh = dy * g
    dx = rstd * (dxh - dxh.mean(-1, keepdims=True) - xh * (dxh * xh).mean(-1, keepdims=True))
    return dx, dg, db


def gelu(x):
    """nn.GELU() exact erf form, My_model_1.py:487-491."""
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def gelu_grad(x):
    return 0.5 * (1.0 + _erf(x / math.sqrt(2.0))) + x * np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


def softmax(x):
    m = x.max(-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(-1, keepdims=True)


def window_partition(x, ws=WIN):
    """My_model_1.py:550-574 (dilation_rate == 1 branch). x[B,H,W,C] -> [B*nW, ws, ws, C]."""
    B, H, W, C = x.shape
    x = x.reshape(B, H // ws, ws, W // ws, ws, C)
    return x.transpose(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)


def window_reverse(windows, ws, H, W):
    """My_model_1.py:577-601. [B*nW, ws, ws, C] -> [B,H,W,C]."""
    B = windows.shape[0] // ((H // ws) * (W // ws))
    x = windows.reshape(B, H // ws, W // ws, ws, ws, -1)
    return x.transpose(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def relative_position_index(ws=WIN):
    """My_model_1.py:366-381: rel(n,m) = (ty_n-ty_m+ws-1)*(2ws-1) + (tx_n-tx_m+ws-1)."""
    ty, tx = np.meshgrid(np.arange(ws), np.arange(ws), indexing="ij")
    ty = ty.reshape(-1)
    tx = tx.reshape(-1)
    return ((ty[:, None] - ty[None, :] + ws - 1) * (2 * ws - 1) + (tx[:, None] - tx[None, :] + ws - 1)).astype(np.int64)


def shift_attn_mask(H, W, ws, shift, dtype=np.float32):
    """My_model_1.py:803-836: region ids on the (shifted-frame) map, -100 where regions differ."""
    img = np.zeros((1, H, W, 1), dtype=dtype)
    slices = (slice(0, -ws), slice(-ws, -shift), slice(-shift, None))
    cnt = 0
    for hs in slices:
        for wsl in slices:
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = window_partition(img, ws).reshape(-1, ws * ws)
    diff = mw[:, None, :] - mw[:, :, None]
    return np.where(diff != 0, dtype(-100.0), dtype(0.0)).astype(dtype)


def input_attn_mask(mask_img, H, W, ws, dtype=np.float32):
    """My_model_1.py:791-798: nearest-resized input mask -> [nW,64,64] in {0,-100}.

    ``mask_img`` is [1,1,h,w]; F.interpolate default mode is 'nearest'
    (src index = floor(dst * in/out))."""
    h, w = mask_img.shape[-2:]
    ri = np.floor(np.arange(H) * (h / H)).astype(np.int64)
    ci = np.floor(np.arange(W) * (w / W)).astype(np.int64)
    m = mask_img[0, 0][ri][:, ci].astype(dtype)[None, :, :, None]
    mw = window_partition(m, ws).reshape(-1, ws * ws)
    am = mw[:, :, None] * mw[:, None, :]
    return np.where(am != 0, dtype(-100.0), dtype(0.0)).astype(dtype)


# ------------------------------------------------------------------ ProbSparse attention
def sparsity_measure(q, k, idx, bf16=False):
    """attn.py:88-117.  q,k [B_,nH,64,D]; idx [64,U] -> M [B_,nH,64].

    S~[n,t] = q_n . k_{idx[n,t]} (unscaled); M_n = max_t S~ - sum_t S~ / L_K (L_K = 64, not U)."""
    L_K = k.shape[2]
    k_sample = k[:, :, idx, :]                                # [B_,nH,64,U,D]  (attn.py:104)
    qk_sample = rbf(np.einsum("bhnd,bhntd->bhnt", q, k_sample), bf16)   # attn.py:110 (bf16 matmul output)
    return qk_sample.max(-1) - qk_sample.sum(-1) / L_K         # attn.py:117


def select_top(M, u):
    """attn.py:122 ``M.topk(u, sorted=False)[1]``; returned ascending (set semantics),
    ties broken towards the lower query index.  Also returns the rank-u / rank-(u+1) gap
    relative to the row's M range, used by the tie-aware index comparison."""
    order = np.argsort(-M, axis=-1, kind="stable")
    top = np.sort(order[..., :u], axis=-1)
    Ms = np.take_along_axis(M, order, -1)
    rng = Ms[..., 0] - Ms[..., -1]
    gap = (Ms[..., u - 1] - Ms[..., u]) if M.shape[-1] > u else np.full(M.shape[:-1], np.inf)
    rel_gap = gap / np.maximum(rng, 1e-30)
    return top.astype(np.int64), rel_gap


def prob_attention(q, k, v, rpb, mask, idx, use_rpb=True, top=None, return_aux=False, bf16=False):
    """ProbAttention.forward, attn.py:287-342 (with _prob_QK :71-152,
    _get_initial_context :154-176, _update_context :178-281).

    q,k,v [B_,64,nH,D]; rpb [nH,64,64]; mask [nW,64,64] or None; idx [64,U].
    Returns ctx [B_,64,nH,D].  ``top`` may be forced (used to evaluate near-tie rows)."""
    B_, L, nH, D = q.shape
    q = q.transpose(0, 2, 1, 3)
    k = k.transpose(0, 2, 1, 3)
    v = v.transpose(0, 2, 1, 3)                                # attn.py:301-303
    U_part, u = prob_sizes(L, L)
    assert idx.shape == (L, U_part)
    M = sparsity_measure(q, k, idx, bf16)
    sel, rel_gap = select_top(M, u)
    if top is None:
        top = sel
    bi = np.arange(B_)[:, None, None]
    hi = np.arange(nH)[None, :, None]
    q_red = q[bi, hi, top]                                     # attn.py:129-131  [B_,nH,u,D]
    scores = rbf(np.matmul(q_red, k.transpose(0, 1, 3, 2)), bf16)   # attn.py:150
    scores = rbf(scores * q.dtype.type(1.0 / math.sqrt(D)), bf16)   # attn.py:327-329
    ctx = np.broadcast_to(rbf(v.mean(2, keepdims=True), bf16), v.shape).copy()   # attn.py:168-172
    p1 = softmax(scores)                                       # attn.py:195  (first softmax)
    a = p1
    if use_rpb:                                                # attn.py:227-230
        a = a + rpb[None][np.zeros_like(bi), hi, top]
    if mask is not None:                                       # attn.py:236-261
        nW = mask.shape[0]
        wi = (np.arange(B_) % nW)[:, None, None]               # batch-major window order, attn.py:250
        a = a + mask[wi, top]
    p2 = softmax(a)                                            # attn.py:262/264 (second softmax)
    ctx[bi, hi, top] = rbf(np.matmul(rbf(p2, bf16), v), bf16)  # attn.py:271-272 (autocast: P2 cast to bf16)
    out = np.ascontiguousarray(ctx.transpose(0, 2, 1, 3))      # attn.py:342
    if return_aux:
        return out, dict(M=M, top=top, sel=sel, rel_gap=rel_gap, p1=p1, p2=p2, q=q, k=k, v=v)   # sel: the oracle's own selection
    return out


def rpb_from_table(table, ws=WIN):
    """WindowAttention.forward, My_model_1.py:408-410: table[225,nH] -> rpb [nH,64,64]."""
    ri = relative_position_index(ws)
    n = ws * ws
    return np.ascontiguousarray(table[ri.reshape(-1)].reshape(n, n, -1).transpose(2, 0, 1))


def window_attention(xw, p, mask, idx, use_rpb=True, top=None, return_aux=False, bf16=False):
    """WindowAttention.forward (My_model_1.py:400-415) -> AttentionLayer.forward (attn.py:385-461).

    xw [B_,64,C]; ``p`` holds the block parameters under their state_dict names."""
    B_, L, C = xw.shape
    table = p["attn.relative_position_bias_table"]
    nH = table.shape[1]
    rpb = rpb_from_table(table)
    pre = "attn.ProbSpare."
    x2 = rbf(xw.reshape(-1, C), bf16)
    lin = lambda a, n: rbf(a @ rbf(p[pre + n + ".weight"], bf16).T + rbf(p[pre + n + ".bias"], bf16), bf16)
    q = lin(x2, "query_projection").reshape(B_, L, nH, -1)
    k = lin(x2, "key_projection").reshape(B_, L, nH, -1)
    v = lin(x2, "value_projection").reshape(B_, L, nH, -1)
    res = prob_attention(q, k, v, rpb, mask, idx, use_rpb, top=top, return_aux=return_aux, bf16=bf16)
    ctx, aux = res if return_aux else (res, None)
    out = lin(ctx.reshape(-1, C), "out_projection")
    out = out.reshape(B_, L, C)
    if return_aux:
        aux["ctx"] = ctx
        return out, aux
    return out


# ------------------------------------------------------------------------------ LeFF
def dwconv3x3(x, w, b):
    """nn.Conv2d(hid, hid, groups=hid, 3, 1, 1), My_model_1.py:490 — cross-correlation,
    zero padding, channel-last here.  x [B,H,W,Ch]; w [Ch,1,3,3]; b [Ch]."""
    B, H, W, Ch = x.shape
    xp = np.zeros((B, H + 2, W + 2, Ch), dtype=x.dtype)
    xp[:, 1:-1, 1:-1] = x
    out = np.zeros_like(x)
    for ky in range(3):
        for kx in range(3):
            out += xp[:, ky:ky + H, kx:kx + W] * w[:, 0, ky, kx]
    return out + b


def leff(x, p, return_aux=False, bf16=False):
    """LeFF.forward, My_model_1.py:496-534.  x [B,L,C] -> [B,L,C]."""
    B, L, C = x.shape
    hh = int(math.sqrt(L))
    r = lambda t: rbf(t, bf16)
    a1 = r(r(x.reshape(-1, C)) @ r(p["mlp.linear1.0.weight"]).T + r(p["mlp.linear1.0.bias"]))      # :508
    h1 = r(gelu(a1)).reshape(B, hh, hh, -1)
    a2 = r(dwconv3x3(h1, r(p["mlp.dwconv.0.weight"]), r(p["mlp.dwconv.0.bias"])))                   # :517
    h2 = r(gelu(a2))
    out = r(h2.reshape(B * L, -1) @ r(p["mlp.linear2.0.weight"]).T + r(p["mlp.linear2.0.bias"]))   # :529
    out = out.reshape(B, L, C)
    if return_aux:
        return out, dict(a1=a1.reshape(B, hh, hh, -1), h1=h1, a2=a2, h2=h2)
    return out


# ------------------------------------------------------------------------ LeWin block
def lewin_block(x, p, shift, idx, input_mask=None, use_rpb=True, drop_scale=None, top=None,
                return_aux=False, bf16=False):
    """LeWinTransformerBlock.forward, My_model_1.py:785-875.

    x [B,L,C]; ``shift`` in {0,4}; ``input_mask`` [1,1,h,w] or None (test_in_any_resolution.py:106);
    ``drop_scale`` [2,B] per-sample DropPath factors (0 or 1/keep) for the two residual branches,
    None in eval (My_model_1.py:872-873)."""
    B, L, C = x.shape
    H = W = int(math.sqrt(L))
    ws = WIN
    dt = x.dtype.type
    attn_mask = None
    if input_mask is not None:
        attn_mask = input_attn_mask(input_mask, H, W, ws, x.dtype.type)
    if shift > 0:
        sm = shift_attn_mask(H, W, ws, shift, x.dtype.type)
        attn_mask = sm if attn_mask is None else attn_mask + sm                  # :836
    xn = layer_norm(x, p["norm1.weight"], p["norm1.bias"]).reshape(B, H, W, C)   # :839
    if shift > 0:
        xn = np.roll(xn, (-shift, -shift), axis=(1, 2))                           # :846
    xw = window_partition(xn, ws).reshape(-1, ws * ws, C)                         # :851-852
    res = window_attention(xw, p, attn_mask, idx, use_rpb, top=top, return_aux=return_aux, bf16=bf16)
    aw, aux = res if return_aux else (res, None)
    sx = window_reverse(aw.reshape(-1, ws, ws, C), ws, H, W)                      # :861
    if shift > 0:
        sx = np.roll(sx, (shift, shift), axis=(1, 2))                             # :866
    a = sx.reshape(B, L, C)
    s0 = dt(1.0) if drop_scale is None else np.asarray(drop_scale[0], dtype=x.dtype)[:, None, None]
    s1 = dt(1.0) if drop_scale is None else np.asarray(drop_scale[1], dtype=x.dtype)[:, None, None]
    y = rbf(x + s0 * a, bf16)                                                     # :872
    z = layer_norm(y, p["norm2.weight"], p["norm2.bias"])
    out = rbf(y + s1 * leff(z, p, bf16=bf16), bf16)                               # :873
    if return_aux:
        aux["y"] = y
        aux["mask"] = attn_mask
        return out, aux
    return out


# --------------------------------------------------------------------------- backward
GRAD_KEYS = (
    "norm1.weight", "norm1.bias", "attn.relative_position_bias_table",
    "attn.ProbSpare.query_projection.weight", "attn.ProbSpare.query_projection.bias",
    "attn.ProbSpare.key_projection.weight", "attn.ProbSpare.key_projection.bias",
    "attn.ProbSpare.value_projection.weight", "attn.ProbSpare.value_projection.bias",
    "attn.ProbSpare.out_projection.weight", "attn.ProbSpare.out_projection.bias",
    "norm2.weight", "norm2.bias",
    "mlp.linear1.0.weight", "mlp.linear1.0.bias", "mlp.dwconv.0.weight", "mlp.dwconv.0.bias",
    "mlp.linear2.0.weight", "mlp.linear2.0.bias",
)


def _softmax_bwd(p, dp):
    return p * (dp - (dp * p).sum(-1, keepdims=True))


def leff_bwd(dout, z, p):
    """Gradient of ``leff`` wrt its input z and its six parameters (autograd restated by hand;
    SURVEY 3.4).  dout [B,L,C]."""
    B, L, C = z.shape
    hh = int(math.sqrt(L))
    _, aux = leff(z, p, return_aux=True)
    a1, h1, a2, h2 = aux["a1"], aux["h1"], aux["a2"], aux["h2"]
    Ch = a1.shape[-1]
    g = {}
    do2 = dout.reshape(-1, C)
    g["mlp.linear2.0.weight"] = do2.T @ h2.reshape(-1, Ch)
    g["mlp.linear2.0.bias"] = do2.sum(0)
    dh2 = (do2 @ p["mlp.linear2.0.weight"]).reshape(B, hh, hh, Ch)
    da2 = dh2 * gelu_grad(a2)
    w = p["mlp.dwconv.0.weight"]
    h1p = np.zeros((B, hh + 2, hh + 2, Ch), dtype=z.dtype)
    h1p[:, 1:-1, 1:-1] = h1
    da2p = np.zeros_like(h1p)
    da2p[:, 1:-1, 1:-1] = da2
    dw = np.zeros_like(w)
    dh1 = np.zeros_like(h1)
    for ky in range(3):
        for kx in range(3):
            dw[:, 0, ky, kx] = (da2 * h1p[:, ky:ky + hh, kx:kx + hh]).reshape(-1, Ch).sum(0)
            # dh1[y,x] += da2[y-ky+1, x-kx+1] * w[ky,kx]
            dh1 += da2p[:, 2 - ky:2 - ky + hh, 2 - kx:2 - kx + hh] * w[:, 0, ky, kx]
    g["mlp.dwconv.0.weight"] = dw
    g["mlp.dwconv.0.bias"] = da2.reshape(-1, Ch).sum(0)
    da1 = (dh1 * gelu_grad(a1)).reshape(-1, Ch)
    g["mlp.linear1.0.weight"] = da1.T @ z.reshape(-1, C)
    g["mlp.linear1.0.bias"] = da1.sum(0)
    dz = (da1 @ p["mlp.linear1.0.weight"]).reshape(B, L, C)
    return dz, g


def prob_attention_bwd(dctx, aux, use_rpb=True):
    """Gradient of ``prob_attention`` (what autograd derives from attn.py:287-342) for the selection in ``aux['top']``.

    dctx [B_,nH,64,D] (head-major like aux['q']).  Returns dq, dk, dv [B_,nH,64,D] and the gradient w.r.t. the gathered
    bias ``rpb`` [nH,64,64].  Gradient paths: index_sample, M and topk carry none; dq is non-zero only at the
    selected rows; dv = P2^T dctx[top] + (1/64) * sum over the NON-selected rows of dctx."""
    q, k, v, p1, p2, top = aux["q"], aux["k"], aux["v"], aux["p1"], aux["p2"], aux["top"]
    B_, nH, L, D = q.shape
    bi = np.arange(B_)[:, None, None]
    hi = np.arange(nH)[None, :, None]
    dctx_top = dctx[bi, hi, top]                                  # [B_,nH,u,D]
    sel = np.zeros((B_, nH, L), dtype=bool)
    sel[bi, hi, top] = True
    dmean = (dctx * (~sel)[..., None]).sum(2, keepdims=True) / L   # mean-fill rows
    dv = np.matmul(p2.transpose(0, 1, 3, 2), dctx_top) + dmean
    dp2 = np.matmul(dctx_top, v.transpose(0, 1, 3, 2))
    da = _softmax_bwd(p2, dp2)
    drpb = np.zeros((nH, L, L), dtype=q.dtype)
    if use_rpb:
        np.add.at(drpb, (np.broadcast_to(hi, top.shape).reshape(-1), top.reshape(-1)), da.reshape(-1, L))
    ds = _softmax_bwd(p1, da) * q.dtype.type(1.0 / math.sqrt(D))
    dq = np.zeros_like(q)
    dq[bi, hi, top] = np.matmul(ds, k)
    dk = np.matmul(ds.transpose(0, 1, 3, 2), q[bi, hi, top])
    return dq, dk, dv, drpb, da


def window_attention_bwd(dout, xw, p, mask, idx, use_rpb=True, top=None):
    """Gradient of ``window_attention`` wrt xw and the 9 live attention parameters."""
    B_, L, C = xw.shape
    _, aux = window_attention(xw, p, mask, idx, use_rpb, top=top, return_aux=True)
    q, top, ctx = aux["q"], aux["top"], aux["ctx"]
    nH, D = q.shape[1], q.shape[3]
    pre = "attn.ProbSpare."
    g = {}
    do = dout.reshape(-1, C)
    g[pre + "out_projection.weight"] = do.T @ ctx.reshape(-1, C)
    g[pre + "out_projection.bias"] = do.sum(0)
    dctx = (do @ p[pre + "out_projection.weight"]).reshape(B_, L, nH, D).transpose(0, 2, 1, 3)
    dq, dk, dv, _, da = prob_attention_bwd(dctx, aux, use_rpb)
    dtab = np.zeros_like(p["attn.relative_position_bias_table"])
    if use_rpb:
        ri = relative_position_index()
        rel = ri[top]                                             # [B_,nH,u,64]
        hh = np.broadcast_to(np.arange(nH)[None, :, None, None], rel.shape)
        np.add.at(dtab, (rel.reshape(-1), hh.reshape(-1)), da.reshape(-1))
    g["attn.relative_position_bias_table"] = dtab
    x2 = xw.reshape(-1, C)
    dx = np.zeros_like(x2)
    for name, d in (("query", dq), ("key", dk), ("value", dv)):
        d2 = d.transpose(0, 2, 1, 3).reshape(-1, C)
        g[pre + name + "_projection.weight"] = d2.T @ x2
        g[pre + name + "_projection.bias"] = d2.sum(0)
        dx = dx + d2 @ p[pre + name + "_projection.weight"]
    return dx.reshape(B_, L, C), g


def lewin_block_bwd(dout, x, p, shift, idx, input_mask=None, use_rpb=True, drop_scale=None, top=None):
    """Gradient of ``lewin_block`` wrt x and the 19 live parameters (keys = GRAD_KEYS)."""
    B, L, C = x.shape
    H = W = int(math.sqrt(L))
    ws = WIN
    _, aux = lewin_block(x, p, shift, idx, input_mask, use_rpb, drop_scale, top=top, return_aux=True)
    y, mask, top = aux["y"], aux["mask"], aux["top"]
    dt = x.dtype.type
    s0 = dt(1.0) if drop_scale is None else np.asarray(drop_scale[0], dtype=x.dtype)[:, None, None]
    s1 = dt(1.0) if drop_scale is None else np.asarray(drop_scale[1], dtype=x.dtype)[:, None, None]
    g = {}
    # out = y + s1 * leff(LN2(y))
    z = layer_norm(y, p["norm2.weight"], p["norm2.bias"])
    dz, g_leff = leff_bwd(dout * s1, z, p)
    g.update(g_leff)
    dy_ln, g["norm2.weight"], g["norm2.bias"] = layer_norm_bwd(dz, y, p["norm2.weight"])
    dy = dout + dy_ln
    # y = x + s0 * unwindow(attn(window(LN1(x))))
    da = (dy * s0).reshape(B, H, W, C)
    if shift > 0:
        da = np.roll(da, (-shift, -shift), axis=(1, 2))
    daw = window_partition(da, ws).reshape(-1, ws * ws, C)
    xn = layer_norm(x, p["norm1.weight"], p["norm1.bias"]).reshape(B, H, W, C)
    if shift > 0:
        xn = np.roll(xn, (-shift, -shift), axis=(1, 2))
    xw = window_partition(xn, ws).reshape(-1, ws * ws, C)
    dxw, g_attn = window_attention_bwd(daw, xw, p, mask, idx, use_rpb, top=top)
    g.update(g_attn)
    dxn = window_reverse(dxw.reshape(-1, ws, ws, C), ws, H, W)
    if shift > 0:
        dxn = np.roll(dxn, (shift, shift), axis=(1, 2))
    dx_ln, g["norm1.weight"], g["norm1.bias"] = layer_norm_bwd(dxn.reshape(B, L, C), x, p["norm1.weight"])
    return dy + dx_ln, g


# ------------------------------------------------------------------ helpers for tests
def as_dtype(p: dict, dtype):
    return {k: (np.asarray(v).astype(dtype) if np.issubdtype(np.asarray(v).dtype, np.floating) else np.asarray(v))
            for k, v in p.items()}


def random_block_params(C, nH, rng: np.random.Generator, std=0.2, dtype=np.float32):
    """Random live parameters of one block, state_dict-named (Appendix B of SURVEY.md)."""
    p = {}
    def w(*s):
        return (rng.standard_normal(s) * std).astype(dtype)
    p["norm1.weight"] = (1.0 + 0.1 * rng.standard_normal(C)).astype(dtype)
    p["norm1.bias"] = w(C)
    p["attn.relative_position_bias_table"] = w(225, nH)
    for n in ("query", "key", "value", "out"):
        p[f"attn.ProbSpare.{n}_projection.weight"] = w(C, C)
        p[f"attn.ProbSpare.{n}_projection.bias"] = w(C)
    p["norm2.weight"] = (1.0 + 0.1 * rng.standard_normal(C)).astype(dtype)
    p["norm2.bias"] = w(C)
    p["mlp.linear1.0.weight"] = w(4 * C, C)
    p["mlp.linear1.0.bias"] = w(4 * C)
    p["mlp.dwconv.0.weight"] = w(4 * C, 1, 3, 3)
    p["mlp.dwconv.0.bias"] = w(4 * C)
    p["mlp.linear2.0.weight"] = w(C, 4 * C)
    p["mlp.linear2.0.bias"] = w(C)
    return p
