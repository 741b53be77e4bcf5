"""CPU restatement of the whole Uformer forward (My_model_1.py:1169-1207) around the numpy LeWin oracle.

TEST INFRASTRUCTURE ONLY (see oracle/lewin_oracle.py).  Used as the checker for model-level parity and
as the CPU baseline of bench.py (`cpu_baseline`, `--impl reference`).  The 18 LeWin blocks run through
oracle.lewin_oracle (numpy, BLAS threads = host cores); the out-of-scope convolutions around them
(InputProj / Downsample / Upsample / OutputProj, My_model_1.py:606-733) use torch CPU conv2d exactly as
the reference does.  Parameters come as a {state_dict key: ndarray} mapping.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import lewin_oracle as O

STAGES = ["encoderlayer_0", "encoderlayer_1", "encoderlayer_2", "encoderlayer_3", "conv",
          "decoderlayer_0", "decoderlayer_1", "decoderlayer_2", "decoderlayer_3"]


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _tok2img(x):
    B, L, C = x.shape
    H = int(round(L ** 0.5))
    return _t(x).transpose(1, 2).reshape(B, C, H, H)


def _img2tok(t):
    return t.flatten(2).transpose(1, 2).contiguous().numpy()


def _block_params(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def uformer_forward(x, sd, idx, depths=(2,) * 9, img_size=128, win=8, dtype=np.float32, record=None, force_tops=None):
    """x [B,3,H,W] ndarray, sd state_dict as ndarrays, idx [18,64,25] index_sample draws in module order.
    record: optional list; one dict per LeWin block (module order) with the block's input tokens, selected query sets
    ``top`` and the rank-u / rank-(u+1) gaps ``rel_gap`` is appended (per-block, tie-aware model-level comparisons).
    force_tops: optional list of 18 arrays [B_, nH, 25] (sorted): block i attends with that selection instead of its own
    (``record`` then also carries ``sel``, the oracle's own choice on the same input) - how a device run whose near-tie rows
    fell the other way is followed block by block."""
    sd = {k: (np.asarray(v).astype(dtype) if np.issubdtype(np.asarray(v).dtype, np.floating) else np.asarray(v))
          for k, v in sd.items()}
    x = x.astype(dtype)
    it = iter(range(sum(depths)))

    def stage(name, tok, res_div):
        for i in range(depths[STAGES.index(name)]):
            shift = 0 if i % 2 == 0 else win // 2
            if img_size // res_div <= win:                      # My_model_1.py:764-766
                shift = 0
            p = _block_params(sd, f"{name}.blocks.{i}.")
            bi = next(it)
            forced = None if force_tops is None else np.asarray(force_tops[bi]).astype(np.int64)
            if record is None:
                tok = O.lewin_block(tok, p, shift, np.asarray(idx[bi]).astype(np.int64), top=forced)
            else:
                x_in = tok
                tok, aux = O.lewin_block(tok, p, shift, np.asarray(idx[bi]).astype(np.int64), top=forced, return_aux=True)
                record.append(dict(block=bi, stage=name, shift=shift, x=x_in, top=aux["top"], sel=aux["sel"],
                                   rel_gap=aux["rel_gap"]))
        return tok

    def conv(t, w, b, **kw):
        return F.conv2d(t, _t(sd[w]), _t(sd[b]), **kw)

    y = F.leaky_relu(conv(_t(x), "input_proj.proj.0.weight", "input_proj.proj.0.bias", padding=1), 0.01)
    tok = _img2tok(y)
    skips = []
    for lvl in range(4):
        tok = stage(f"encoderlayer_{lvl}", tok, 2 ** lvl)
        skips.append(tok)
        tok = _img2tok(conv(_tok2img(tok), f"dowsample_{lvl}.conv.0.weight", f"dowsample_{lvl}.conv.0.bias",
                            stride=2, padding=1))
    tok = stage("conv", tok, 16)
    for lvl in range(4):
        up = F.conv_transpose2d(_tok2img(tok), _t(sd[f"upsample_{lvl}.deconv.0.weight"]),
                                _t(sd[f"upsample_{lvl}.deconv.0.bias"]), stride=2)
        tok = np.concatenate([_img2tok(up), skips[3 - lvl]], -1)
        tok = stage(f"decoderlayer_{lvl}", tok, 2 ** (3 - lvl))
    y = conv(_tok2img(tok), "output_proj.proj.0.weight", "output_proj.proj.0.bias", padding=1).numpy()
    return x + y
