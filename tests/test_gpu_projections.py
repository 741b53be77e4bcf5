"""Downsample / OutputProj on the implicit-GEMM tcgen05 kernel (csrc/conv_igemm.cuh) against the torch
convolution they replace (My_model_1.py:606-630 Conv2d(C, 2C, 4, stride 2, pad 1); :696-733 Conv2d(2C, 3, 3, pad 1) and the
`x + y` of Uformer.forward :1207).  Operands are bf16 values, accumulation is fp32, and the result carries torch's two autocast
roundings (conv -> bf16, + bias -> bf16); the only freedom left is the fp32 summation order, so the comparison allows one bf16
ulp on a small fraction of the elements and nothing else."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rb(t):
    return t.to(torch.bfloat16).float()


def _close_bf16(got, ref, what, pre=None):
    """pre: the bf16-rounded convolution result before the bias add - a one-ulp difference THERE (summation order) is the
    allowed unit, also where the bias cancels most of the value."""
    got, ref = got.float(), ref.float()
    d = (got - ref).abs()
    mag = ref.abs() if pre is None else torch.maximum(ref.abs(), pre.float().abs())
    ulp = torch.maximum(mag, torch.full_like(ref, 2.0 ** -6)) * 2.0 ** -7      # (at most) one bf16 ulp at that magnitude
    bad = d > 1.01 * ulp
    frac = float((d > 0).float().mean())
    print(f"{what}: max |diff| {float(d.max()):.3e}, elements differing {frac:.2e}, beyond one ulp {int(bad.sum())}")
    # two roundings (conv -> bf16, + bias -> bf16): a one-ulp difference of the first can be doubled by the second on a few
    # elements (a binade crossing); nothing may be further off, and those elements must stay rare
    assert not (d > 2.02 * ulp).any(), (what, float(d.max()))
    assert int(bad.sum()) <= max(2, ref.numel() // 5000), (what, int(bad.sum()))
    assert frac < 0.02, (what, frac)


@pytest.mark.parametrize("B,H,W,Cin,ld_x,pad_h", [(2, 32, 32, 32, 32, True), (1, 64, 64, 64, 64, True), (3, 16, 16, 128, 128, True),
                                                   (2, 16, 16, 256, 256, True), (1, 18, 48, 64, 64, True), (2, 32, 32, 32, 64, True),
                                                   (1, 32, 32, 64, 128, True), (1, 34, 32, 32, 32, False), (1, 10, 48, 128, 128, False),
                                                   (5, 8, 16, 256, 256, True)])
def test_downsample_matches_torch_conv(B, H, W, Cin, ld_x, pad_h):
    from lewin_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + Cin)
    buf = _rb(torch.randn(B, H * W, ld_x, generator=g)).to(DEV)
    x = buf[..., ld_x - Cin:] if ld_x > Cin else buf                 # the right half of a wider (concat) buffer
    w = (torch.randn(2 * Cin, Cin, 4, 4, generator=g) / (4.0 * Cin ** 0.5)).to(DEV)
    b = (0.1 * torch.randn(2 * Cin, generator=g)).to(DEV)
    got = ops.lewin_downsample(x.to(torch.bfloat16), w, b, B=B, H=H, W=W, pad_h=pad_h)
    xi = x.float().reshape(B, H, W, Cin).permute(0, 3, 1, 2)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        y = F.conv2d(xi, _rb(w), None, stride=2, padding=(1 if pad_h else 0, 1))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    tok = lambda t: t.permute(0, 2, 3, 1).reshape(B, -1, 2 * Cin)
    ref = tok(_rb(_rb(y) + _rb(b).view(1, -1, 1, 1)))
    assert got.shape == ref.shape and got.dtype == torch.bfloat16
    _close_bf16(got, ref, f"downsample B{B} {H}x{W} C{Cin} ld{ld_x} pad{int(pad_h)}", pre=tok(_rb(y)))


@pytest.mark.parametrize("B,H,W,Cin,pad_h,resid", [(2, 32, 32, 64, True, True), (1, 20, 48, 64, True, False), (1, 34, 24, 64, False, True),
                                                    (3, 16, 16, 128, True, True), (1, 130, 128, 64, False, False)])
def test_output_proj_matches_torch_conv(B, H, W, Cin, pad_h, resid):
    from lewin_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + Cin)
    x = _rb(torch.randn(B, H * W, Cin, generator=g)).to(DEV)
    w = (torch.randn(3, Cin, 3, 3, generator=g) / (3.0 * Cin ** 0.5)).to(DEV)
    b = (0.1 * torch.randn(3, generator=g)).to(DEV)
    Hout = H if pad_h else H - 2
    res = torch.rand(B, 3, Hout, W, generator=g).to(DEV) if resid else None
    got = ops.lewin_output_proj(x.to(torch.bfloat16), w, b, B=B, H=H, W=W, residual=res, pad_h=pad_h)
    xi = x.reshape(B, H, W, Cin).permute(0, 3, 1, 2)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        y = F.conv2d(xi, _rb(w), None, padding=(1 if pad_h else 0, 1))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    pre = _rb(y)
    y = _rb(pre + _rb(b).view(1, -1, 1, 1))
    assert got.shape == y.shape and got.dtype == torch.float32
    if resid:
        got = got - res          # fp32 add of a bf16 value to a [0, 1) image: subtracting it back is exact to ~1e-7
        d = (got - y).abs()
        assert float(d.max()) < 2.0 ** -7 * max(1.0, float(y.abs().max())), float(d.max())
    else:
        _close_bf16(got, y, f"output_proj B{B} {H}x{W} C{Cin} pad{int(pad_h)}", pre=pre)


def test_uformer_forward_uses_own_projection_kernels():
    """bf16 inference of the whole model: no cuDNN convolution is left (all ten projections run on this library's kernels)."""
    import lewin_b200 as L
    from torch.profiler import profile, ProfilerActivity
    torch.manual_seed(0)
    model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(DEV).eval()
    x = torch.rand(4, 3, 128, 128, device=DEV)
    idx = model.draw_index_samples()
    with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
        model(x, index_samples=idx)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            y = model(x, index_samples=idx)
            torch.cuda.synchronize()
    names = {ev.name for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA}
    assert not [n for n in names if "cudnn" in n.lower() or "cutlass" in n.lower() or "implicit_gemm" in n.lower()], names
    assert any("conv_igemm_kernel" in n for n in names), names
    assert y.shape == x.shape and y.dtype == torch.float32 and torch.isfinite(y).all()
