"""Multi-threaded CPU port of the reference forward, op for op, in torch eager — the CPU BASELINE TIMER.

TEST INFRASTRUCTURE ONLY (see oracle/lewin_oracle.py).  The numpy oracle is the checker; numpy's
element-wise kernels are single-threaded, so timing it would understate what the reference achieves on
the host.  This file restates the reference's ATen op sequence (including the materialised
K_sample[B_,nH,64,25,D] gather of ProbSparse/attn.py:104 that dominates its CPU time) so that
`bench.py --impl reference` / `cpu_baseline` measure the reference's own CPU cost with all host threads.
It is pinned to the UNMODIFIED reference: tests/test_oracle_golden.py::test_torch_port_matches_reference_golden_whole_model
compares its whole-model forward with the reference's recorded outputs (embed_dim 32: 8.6e-6, embed_dim 64: bit-identical).
Forward only, fp32.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _partition(x, ws=8):                                   # My_model_1.py:550-574
    B, H, W, C = x.shape
    return x.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def _reverse(w, ws, H, W):                                 # My_model_1.py:577-601
    B = w.shape[0] // ((H // ws) * (W // ws))
    return w.view(B, H // ws, W // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def _shift_mask(H, W, ws, s):                              # My_model_1.py:803-836
    m = torch.zeros((1, H, W, 1))
    sl = (slice(0, -ws), slice(-ws, -s), slice(-s, None))
    c = 0
    for h in sl:
        for w in sl:
            m[:, h, w, :] = c
            c += 1
    mw = _partition(m, ws).view(-1, ws * ws)
    d = mw.unsqueeze(1) - mw.unsqueeze(2)
    return d.masked_fill(d != 0, -100.0).masked_fill(d == 0, 0.0)


_REL = None


def _rel_index(ws=8):                                      # My_model_1.py:366-381
    global _REL
    if _REL is None:
        c = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
        r = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0).contiguous()
        r[:, :, 0] += ws - 1
        r[:, :, 1] += ws - 1
        r[:, :, 0] *= 2 * ws - 1
        _REL = r.sum(-1)
    return _REL


def prob_attention(q, k, v, rpb, mask, idx):               # attn.py:287-342
    B, L, H, D = q.shape
    q, k, v = q.transpose(2, 1), k.transpose(2, 1), v.transpose(2, 1)
    u = 25
    k_exp = k.unsqueeze(-3).expand(B, H, L, L, D)
    k_sample = k_exp[:, :, torch.arange(L).unsqueeze(1), idx, :]                      # attn.py:104
    qk_s = torch.matmul(q.unsqueeze(-2), k_sample.transpose(-2, -1)).squeeze(-2)      # attn.py:110
    M = qk_s.max(-1)[0] - torch.div(qk_s.sum(-1), L)                                  # attn.py:117
    top = M.topk(u, sorted=False)[1]
    bi, hi = torch.arange(B)[:, None, None], torch.arange(H)[None, :, None]
    s = torch.matmul(q[bi, hi, top, :], k.transpose(-2, -1)) * (1.0 / math.sqrt(D))   # attn.py:150, 329
    ctx = v.mean(dim=-2).unsqueeze(-2).expand(B, H, L, D).clone()                      # attn.py:168-172
    a = torch.softmax(s, dim=-1)                                                       # attn.py:195
    a = a + rpb.unsqueeze(0).repeat(B, 1, 1, 1)[bi, hi, top, :]                        # attn.py:229
    if mask is not None:                                                               # attn.py:236-261
        nW = mask.shape[0]
        mm = mask.unsqueeze(1).unsqueeze(0).repeat(B // nW, 1, H, 1, 1)
        ti = top.unsqueeze(1).view(B // nW, nW, H, u)
        a = a.view(B // nW, nW, H, u, L) + mm[torch.arange(B // nW)[:, None, None, None],
                                               torch.arange(nW)[None, :, None, None],
                                               torch.arange(H)[None, None, :, None], ti, :]
        a = a.view(-1, H, u, L)
    a = torch.softmax(a, dim=-1)                                                       # attn.py:262/264
    ctx[bi, hi, top, :] = torch.matmul(a, v)                                           # attn.py:271
    return ctx.transpose(2, 1).contiguous()


def lewin_block(x, p, shift, idx):                         # My_model_1.py:785-875 (eval, no input mask)
    B, L, C = x.shape
    H = W = int(math.sqrt(L))
    nH = p["attn.relative_position_bias_table"].shape[1]
    mask = _shift_mask(H, W, 8, shift) if shift > 0 else None
    xn = F.layer_norm(x, (C,), p["norm1.weight"], p["norm1.bias"]).view(B, H, W, C)
    if shift > 0:
        xn = torch.roll(xn, shifts=(-shift, -shift), dims=(1, 2))
    xw = _partition(xn).view(-1, 64, C)
    rpb = p["attn.relative_position_bias_table"][_rel_index().view(-1)].view(64, 64, -1).permute(2, 0, 1).contiguous()
    pre = "attn.ProbSpare."
    x2 = xw.reshape(-1, C)
    q = F.linear(x2, p[pre + "query_projection.weight"], p[pre + "query_projection.bias"]).view(-1, 64, nH, C // nH)
    k = F.linear(x2, p[pre + "key_projection.weight"], p[pre + "key_projection.bias"]).view(-1, 64, nH, C // nH)
    v = F.linear(x2, p[pre + "value_projection.weight"], p[pre + "value_projection.bias"]).view(-1, 64, nH, C // nH)
    ctx = prob_attention(q, k, v, rpb, mask, idx).view(-1, C)
    aw = F.linear(ctx, p[pre + "out_projection.weight"], p[pre + "out_projection.bias"]).view(-1, 8, 8, C)
    sx = _reverse(aw, 8, H, W)
    if shift > 0:
        sx = torch.roll(sx, shifts=(shift, shift), dims=(1, 2))
    y = x + sx.view(B, L, C)
    z = F.layer_norm(y, (C,), p["norm2.weight"], p["norm2.bias"])
    h = F.gelu(F.linear(z.reshape(-1, C), p["mlp.linear1.0.weight"], p["mlp.linear1.0.bias"])).view(B, L, -1)
    h = h.view(B, H, W, -1).permute(0, 3, 1, 2)                                        # My_model_1.py:514
    h = F.gelu(F.conv2d(h, p["mlp.dwconv.0.weight"], p["mlp.dwconv.0.bias"], padding=1, groups=h.shape[1]))
    h = h.permute(0, 2, 3, 1).reshape(B * L, -1)                                       # My_model_1.py:520
    return y + F.linear(h, p["mlp.linear2.0.weight"], p["mlp.linear2.0.bias"]).view(B, L, C)


STAGES = ["encoderlayer_0", "encoderlayer_1", "encoderlayer_2", "encoderlayer_3", "conv",
          "decoderlayer_0", "decoderlayer_1", "decoderlayer_2", "decoderlayer_3"]


@torch.no_grad()
def uformer_forward(x, sd, idx, depths=(2,) * 9, img_size=128, win=8):
    """x [B,3,H,W] torch fp32 (CPU), sd {key: tensor}, idx [18,64,25] int64."""
    it = iter(range(sum(depths)))

    def tok2img(t):
        B, L, C = t.shape
        H = int(math.sqrt(L))
        return t.transpose(1, 2).contiguous().view(B, C, H, H)

    def img2tok(t):
        return t.flatten(2).transpose(1, 2).contiguous()

    def stage(name, tok, res_div):
        for i in range(depths[STAGES.index(name)]):
            shift = 0 if (i % 2 == 0 or img_size // res_div <= win) else win // 2
            pre = f"{name}.blocks.{i}."
            p = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
            tok = lewin_block(tok, p, shift, idx[next(it)])
        return tok

    y = F.leaky_relu(F.conv2d(x, sd["input_proj.proj.0.weight"], sd["input_proj.proj.0.bias"], padding=1), 0.01)
    tok = img2tok(y)
    skips = []
    for lvl in range(4):
        tok = stage(f"encoderlayer_{lvl}", tok, 2 ** lvl)
        skips.append(tok)
        tok = img2tok(F.conv2d(tok2img(tok), sd[f"dowsample_{lvl}.conv.0.weight"], sd[f"dowsample_{lvl}.conv.0.bias"],
                               stride=2, padding=1))
    tok = stage("conv", tok, 16)
    for lvl in range(4):
        up = F.conv_transpose2d(tok2img(tok), sd[f"upsample_{lvl}.deconv.0.weight"], sd[f"upsample_{lvl}.deconv.0.bias"],
                                stride=2)
        tok = torch.cat([img2tok(up), skips[3 - lvl]], -1)
        tok = stage(f"decoderlayer_{lvl}", tok, 2 ** (3 - lvl))
    y = F.conv2d(tok2img(tok), sd["output_proj.proj.0.weight"], sd["output_proj.proj.0.bias"], padding=1)
    return x + y
