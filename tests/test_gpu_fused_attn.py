"""The single-kernel attention half (csrc/attn_fused.cuh: LN1 -> q|k|v tcgen05.mma -> TMEM -> ProbSparse core -> out-projection
tcgen05.mma -> residual store) against the three-kernel pipeline it replaces at the C <= 64 levels (bf16 inference).

Both paths run the reference arithmetic of My_model_1.py:803-872 / ProbSparse/attn.py:287-461 with the SAME rounding points
(the parity of the three-kernel path against the oracle / the reference goldens is tests/test_gpu_bf16.py and
test_gpu_block_forward.py, which now go through the fused kernel wherever it applies), so they must agree BIT FOR BIT,
selections included.  The three-kernel path is reached in the same process by asking for gradients (save_for_backward)."""
import numpy as np
import pytest
import torch

from oracle import lewin_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(C, nH, B, hw, shift, seed, use_rpb=True, drop=False):
    import lewin_b200 as L
    from lewin_b200 import _lib
    rng = np.random.default_rng(seed)
    p = O.random_block_params(C, nH, rng)
    blk = L.LeWinTransformerBlock(dim=C, input_resolution=(128, 128), num_heads=nH, win_size=8, shift_size=shift)
    sd = blk.state_dict()
    for k, v in p.items():
        sd[k].copy_(torch.from_numpy(v))
    blk = blk.to(DEV).eval()
    x = torch.from_numpy(rng.standard_normal((B, hw * hw, C)).astype(np.float32)).to(DEV, torch.bfloat16)
    idx = torch.from_numpy(rng.integers(0, 64, size=(64, 25)).astype(np.int64))
    ds = torch.tensor([0.0, 1.25, 1.25, 0.0, 1.25][:B] + [1.25] * max(0, B - 5), device=DEV) if drop else None
    w_qkv, b_qkv = blk.attn.ProbSpare.qkv_weights()
    kw = dict(B=B, H=hw, W=hw, num_heads=nH, shift=shift, ln_w=blk.norm1.weight, ln_b=blk.norm1.bias,
              w_qkv=w_qkv, b_qkv=b_qkv, w_out=blk.attn.ProbSpare.out_projection.weight,
              b_out=blk.attn.ProbSpare.out_projection.bias, rpb_table=blk.attn.relative_position_bias_table,
              index_sample=idx, drop_scale=ds, use_rpb=use_rpb, return_top=True)
    lib = _lib.load()
    n0 = lib.lewin_launch_count()
    with torch.no_grad():
        y_f, top_f = L.ops.lewin_attn(x, **kw)                       # inference: the fused kernel
    n_fused = lib.lewin_launch_count() - n0
    xr = x.detach().clone().requires_grad_(True)
    n0 = lib.lewin_launch_count()
    y_u, top_u = L.ops.lewin_attn(xr, **kw)                          # gradients wanted: q|k|v, ctx saved -> three kernels
    n_unfused = lib.lewin_launch_count() - n0
    torch.cuda.synchronize()
    return y_f, top_f, y_u.detach(), top_u, n_fused, n_unfused


@pytest.mark.parametrize("C,nH,B,hw,shift", [(32, 1, 2, 16, 0), (32, 1, 3, 32, 4), (64, 2, 2, 16, 4), (64, 2, 1, 64, 0),
                                             (64, 2, 1, 24, 4), (32, 1, 1, 24, 4), (64, 2, 5, 16, 4)])
def test_fused_attention_half_is_bit_identical_to_the_three_kernel_path(C, nH, B, hw, shift):
    y_f, top_f, y_u, top_u, n_f, n_u = _run(C, nH, B, hw, shift, seed=C + hw + shift + B)
    assert n_f == 1 and n_u == 3, (n_f, n_u)          # one launch instead of three (q|k|v GEMM, core, out GEMM)
    assert torch.equal(top_f, top_u), "fused kernel selected different top-u queries"
    assert torch.equal(y_f, y_u), float((y_f.float() - y_u.float()).abs().max())


def test_fused_attention_half_options():
    """DropPath factors per sample (My_model_1.py:872) and the relative-position-bias ablation switch (options.py:5)."""
    y_f, top_f, y_u, top_u, _, _ = _run(64, 2, 5, 16, 4, seed=3, drop=True)
    assert torch.equal(top_f, top_u) and torch.equal(y_f, y_u)
    y_f, top_f, y_u, top_u, _, _ = _run(32, 1, 2, 16, 4, seed=4, use_rpb=False)
    assert torch.equal(top_f, top_u) and torch.equal(y_f, y_u)


def test_fused_attention_half_large_map_all_sms():
    """A map with more tiles than 2 x 148 groups (persistent loop, prefetch of the next tile) at both widths."""
    for C, nH in ((32, 1), (64, 2)):
        y_f, top_f, y_u, top_u, _, _ = _run(C, nH, 4, 128, 4, seed=C)
        assert torch.equal(top_f, top_u)
        assert torch.equal(y_f, y_u), float((y_f.float() - y_u.float()).abs().max())
