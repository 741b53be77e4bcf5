"""The single-kernel LeFF tail (csrc/leff_tail.cuh: depthwise 3x3 + GELU tile -> tcgen05.mma A operand -> linear2 -> DropPath *
residual store; h2 never reaches HBM) against the two kernels it replaces at the C <= 64 levels (bf16 inference).

Both paths run the reference arithmetic of LeFF.forward (My_model_1.py:512-531) and the residual of
LeWinTransformerBlock.forward (:873) with the SAME rounding points and summation orders (the parity of the three-kernel path
against the oracle / the reference goldens is tests/test_gpu_bf16.py and test_gpu_block_forward.py, which now go through the
fused tail wherever it applies), so they must agree BIT FOR BIT.  The three-kernel path is reached in the same process by
asking for gradients (save_for_backward)."""
import numpy as np
import pytest
import torch

from oracle import lewin_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(C, B, H, W, seed, drop=False, scale=1.0):
    import lewin_b200 as L
    from lewin_b200 import _lib
    rng = np.random.default_rng(seed)
    p = O.random_block_params(C, C // 32, rng)
    t = lambda k: torch.from_numpy(p[k]).to(DEV)
    y = torch.from_numpy((scale * rng.standard_normal((B, H * W, C))).astype(np.float32)).to(DEV, torch.bfloat16)
    ds = torch.tensor(([0.0, 1.25, 1.25, 0.0, 1.25] * B)[:B], device=DEV) if drop else None
    kw = dict(B=B, H=H, W=W, ln_w=t("norm2.weight"), ln_b=t("norm2.bias"), w1=t("mlp.linear1.0.weight"), b1=t("mlp.linear1.0.bias"),
              w_dw=t("mlp.dwconv.0.weight"), b_dw=t("mlp.dwconv.0.bias"), w2=t("mlp.linear2.0.weight"), b2=t("mlp.linear2.0.bias"),
              drop_scale=ds, fused=True)
    lib = _lib.load()
    n0 = lib.lewin_launch_count()
    with torch.no_grad():
        o_f = L.ops.lewin_leff(y, **kw)                              # inference: linear1 kernel + the fused tail
    n_fused = lib.lewin_launch_count() - n0
    yr = y.detach().clone().requires_grad_(True)
    n0 = lib.lewin_launch_count()
    o_u = L.ops.lewin_leff(yr, **kw)                                 # gradients wanted: h2, a1, a2 saved -> three kernels
    n_unfused = lib.lewin_launch_count() - n0
    torch.cuda.synchronize()
    return o_f, o_u.detach(), n_fused, n_unfused


@pytest.mark.parametrize("C,B,H,W", [(32, 2, 16, 16), (32, 3, 32, 32), (64, 2, 16, 16), (64, 1, 64, 64), (64, 1, 24, 32),
                                     (32, 1, 40, 48), (64, 5, 16, 16), (32, 1, 18, 32), (64, 1, 13, 64)])
def test_leff_tail_is_bit_identical_to_the_two_kernel_path(C, B, H, W):
    o_f, o_u, n_f, n_u = _run(C, B, H, W, seed=C + H + W + B)
    assert n_f == 2 and n_u == 3, (n_f, n_u)          # linear1 + tail instead of linear1 + dwconv + linear2
    assert torch.equal(o_f, o_u), float((o_f.float() - o_u.float()).abs().max())


def test_leff_tail_options():
    """DropPath factors per sample (My_model_1.py:873) and activations far outside the GELU table's hot range."""
    o_f, o_u, _, _ = _run(64, 5, 16, 16, seed=3, drop=True)
    assert torch.equal(o_f, o_u)
    o_f, o_u, _, _ = _run(32, 2, 16, 32, seed=4, scale=40.0)
    assert torch.equal(o_f, o_u)


def test_leff_tail_large_map_all_sms():
    """A map with more tiles than 2 x 148 teams (persistent loop, two-units-ahead halo loads, accumulator ping-pong)."""
    for C in (32, 64):
        o_f, o_u, _, _ = _run(C, 4, 128, 128, seed=C)
        assert torch.equal(o_f, o_u), float((o_f.float() - o_u.float()).abs().max())
