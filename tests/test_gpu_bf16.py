"""GPU parity of the bf16 path (reference semantics under torch.autocast(cuda, bfloat16))
against the numpy oracle with bf16 rounding emulation.  Tolerance: max-abs 2e-2 (north_star); top-u sets must
match on every row whose rank-25/26 gap exceeds 2^-7 of the row's M range (bf16 resolution)."""
import numpy as np
import pytest
import torch

from oracle import lewin_oracle as O
from tests.util import BF16_REF_FIXTURES, TOL_BF16, check_top

pytestmark = pytest.mark.gpu
TAU_BF16 = 2.0 ** -7


def _mk(C, nH, shift, p, dev):
    import lewin_b200 as L
    blk = L.LeWinTransformerBlock(dim=C, input_resolution=(128, 128), num_heads=nH, win_size=8, shift_size=shift)
    sd = blk.state_dict()
    for k, v in p.items():
        sd[k].copy_(torch.from_numpy(v))
    return blk.to(dev).eval()


@pytest.mark.parametrize("C,nH,hw,B,shift", [(32, 1, 16, 2, 0), (32, 1, 16, 2, 4), (64, 2, 32, 1, 4),
                                             (128, 4, 16, 2, 4), (512, 16, 8, 2, 0),
                                             (256, 8, 16, 2, 4), (512, 16, 16, 2, 4), (64, 2, 48, 1, 4)])
def test_block_forward_bf16_matches_bf16_oracle(C, nH, hw, B, shift):
    import lewin_b200 as L
    rng = np.random.default_rng(77 + C + shift)
    p = O.random_block_params(C, nH, rng, std=0.1)
    x = O.rbf(rng.standard_normal((B, hw * hw, C)).astype(np.float32))
    idx = rng.integers(0, 64, size=(64, 25)).astype(np.int64)
    ref, aux = O.lewin_block(x, p, shift, idx, return_aux=True, bf16=True)
    dev = torch.device("cuda:0")
    blk = _mk(C, nH, shift, p, dev)
    xs = torch.from_numpy(x).to(dev).to(torch.bfloat16)
    ps = blk.attn.ProbSpare
    with torch.no_grad():
        y, top = L.ops.lewin_attn(
            xs, B=B, H=hw, W=hw, num_heads=nH, shift=shift, ln_w=blk.norm1.weight, ln_b=blk.norm1.bias,
            w_qkv=ps.qkv_weights()[0], b_qkv=ps.qkv_weights()[1], w_out=ps.out_projection.weight,
            b_out=ps.out_projection.bias, rpb_table=blk.attn.relative_position_bias_table,
            index_sample=torch.from_numpy(idx), return_top=True)
        out = blk(xs, None, torch.from_numpy(idx))
    assert out.dtype == torch.bfloat16
    nbad, namb, nhard = check_top(top.cpu().numpy(), aux["top"], aux["rel_gap"], TAU_BF16)
    assert nhard == 0, f"{nhard} non-ambiguous rows differ ({nbad} differ, {namb} ambiguous of {aux['rel_gap'].size})"
    if nbad:
        ref, aux = O.lewin_block(x, p, shift, idx, top=np.sort(top.cpu().numpy().astype(np.int64), -1),
                                 return_aux=True, bf16=True)
    e_y = np.abs(y.float().cpu().numpy() - aux["y"]).max()
    e_o = np.abs(out.float().cpu().numpy() - ref).max()
    print(f"bf16 C={C}: top rows differing {nbad} (ambiguous {namb}/{aux['rel_gap'].size}); err attn-half {e_y:.3e}, block {e_o:.3e}, |ref| max {np.abs(ref).max():.2f}")
    # 2e-2 absolute for O(1) outputs; outputs are stored in bf16, so allow 3 bf16 ulps of the output scale
    assert e_y < max(TOL_BF16, 3 * 2.0 ** -8 * np.abs(aux["y"]).max())
    assert e_o < max(TOL_BF16, 3 * 2.0 ** -8 * np.abs(ref).max())


@pytest.mark.parametrize("name", BF16_REF_FIXTURES)
def test_block_forward_bf16_matches_reference_cpu_autocast_golden(name):
    """bf16 kernels against the UNMODIFIED reference run under torch.autocast("cpu", bfloat16) (tests/golden/*_bf16cpu):
    same top-u sets on every row whose rank-25/26 gap exceeds the bf16 resolution, outputs within 2 bf16 ulps of the
    activation scale on 99.9 % of the elements (1 ulp kernel vs oracle + 1 ulp CPU- vs CUDA-autocast policy; the rest are
    the few tokens of rows where a near-tie was resolved differently), mean error < 3e-3."""
    import lewin_b200.ops as ops
    from tests.util import load_fixture, make_block
    fx = load_fixture(name)
    dev = torch.device("cuda:0")
    blk = make_block(fx, dev).eval()
    xs = torch.from_numpy(fx["x"]).to(dev).to(torch.bfloat16)
    captured = {}
    orig = ops.lewin_attn

    def spy(*a, **k):
        y, top = orig(*a, return_top=True, **k)
        captured["top"] = top
        return y

    ops.lewin_attn = spy
    try:
        with torch.no_grad():
            out = blk(xs, None, torch.from_numpy(fx["idx"]))
    finally:
        ops.lewin_attn = orig
    assert out.dtype == torch.bfloat16
    _, aux = O.lewin_block(fx["x"].astype(np.float64), O.as_dtype(fx["params"], np.float64), fx["shift"], fx["idx"], None, True, None,
                           return_aux=True)
    nbad, namb, nhard = check_top(captured["top"].cpu().numpy(), fx["top"], aux["rel_gap"], TAU_BF16)
    assert nhard == 0, f"{nhard} non-ambiguous rows differ from the reference's selection ({nbad} differ, {namb} ambiguous)"
    ulp = 2.0 ** (np.floor(np.log2(np.abs(fx["out"]).max())) - 7)
    d = np.abs(out.float().cpu().numpy() - fx["out"])
    # A near-tie resolved differently swaps one query token between "attended" and "mean(V)": that token's attention output
    # changes and LeFF's 3x3 depthwise conv carries it to the 8 neighbouring pixels.  Those pixels are excluded; everything
    # else must agree with the reference to 2 bf16 ulps of the activation scale.
    B, L, C = fx["x"].shape
    hw, sh, nWw = fx["hw"], fx["shift"], fx["hw"] // 8
    affected = np.zeros((B, hw, hw), dtype=bool)
    tg = np.sort(captured["top"].cpu().numpy().astype(np.int64), -1)
    for w_ in range(tg.shape[0]):
        for h_ in range(tg.shape[1]):
            for n in set(tg[w_, h_]) ^ set(fx["top"][w_, h_]):
                b, w = divmod(w_, nWw * nWw)
                wy, wx = divmod(w, nWw)
                y, x_ = (wy * 8 + n // 8 + sh) % hw, (wx * 8 + n % 8 + sh) % hw
                affected[b, max(y - 1, 0):y + 2, max(x_ - 1, 0):x_ + 2] = True
    keep = ~affected.reshape(B, L)
    print(f"{name}: rows differing {nbad}, pixels excluded {int(affected.sum())} of {affected.size}; max over the rest "
          f"{d[keep].max():.4f}, mean {d[keep].mean():.2e}, ulp {ulp:.4f}")
    assert affected.mean() < 0.25           # 2 swapped tokens x 9 pixels each on a 16 x 16 map are already 12 %
    assert d[keep].max() <= 2 * ulp and d[keep].mean() < 3e-3


def test_autocast_routes_to_bf16_kernels_and_trains():
    """Under torch.autocast(cuda, bfloat16) the block runs the bf16 entry points, returns bf16, and the
    backward produces finite fp32 parameter gradients close to the fp32 path's."""
    import lewin_b200 as L
    rng = np.random.default_rng(3)
    C, nH, hw, B, shift = 64, 2, 16, 2, 4
    p = O.random_block_params(C, nH, rng, std=0.1)
    dev = torch.device("cuda:0")
    blk = _mk(C, nH, shift, p, dev)
    x = torch.from_numpy(rng.standard_normal((B, hw * hw, C)).astype(np.float32)).to(dev)
    dout = torch.from_numpy(rng.standard_normal((B, hw * hw, C)).astype(np.float32)).to(dev)
    idx = torch.from_numpy(rng.integers(0, 64, size=(64, 25)).astype(np.int64))
    blk.zero_grad()
    x32 = x.clone().requires_grad_(True)
    blk(x32, None, idx).backward(dout)
    g32 = {k: v.grad.clone() for k, v in blk.named_parameters() if v.grad is not None}
    blk.zero_grad()
    xb = x.clone().requires_grad_(True)
    with torch.autocast("cuda", torch.bfloat16):
        out = blk(xb, None, idx)
    assert out.dtype == torch.bfloat16
    out.backward(dout.to(torch.bfloat16))
    assert xb.grad is not None and torch.isfinite(xb.grad).all()
    for k, g in g32.items():
        gb = dict(blk.named_parameters())[k].grad
        assert gb is not None and gb.dtype == torch.float32 and torch.isfinite(gb).all(), k
    # the two precisions select different top-u sets on ~30 % of rows (SURVEY finding 9); the bulk statistics
    # of the gradients still have to agree
    for k in ("mlp.linear2.0.weight", "mlp.linear1.0.weight", "attn.ProbSpare.value_projection.weight"):
        a, b = g32[k].flatten(), dict(blk.named_parameters())[k].grad.flatten()
        cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
        assert cos > 0.9, (k, cos)


def test_fp16_autocast_runs_on_bf16_kernels():
    """My_train.py:224 wraps the forward in torch.cuda.amp.autocast() (fp16) with a loss scaler.  The library has no fp16
    kernels: under fp16 autocast the LeWin ops compute in bf16, so the reference's training step runs unchanged - the result
    equals the bf16-autocast result bit for bit, and a GradScaler step produces finite, unscaled fp32 gradients."""
    import warnings
    import lewin_b200 as L
    rng = np.random.default_rng(5)
    C, nH, hw, B, shift = 64, 2, 16, 2, 4
    p = O.random_block_params(C, nH, rng, std=0.1)
    dev = torch.device("cuda:0")
    blk = _mk(C, nH, shift, p, dev)
    x = torch.from_numpy(rng.standard_normal((B, hw * hw, C)).astype(np.float32)).to(dev)
    idx = torch.from_numpy(rng.integers(0, 64, size=(64, 25)).astype(np.int64))
    with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
        ref = blk(x, None, idx)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with torch.no_grad(), torch.autocast("cuda", torch.float16):
            out = blk(x, None, idx)
        assert out.dtype == torch.bfloat16 and torch.equal(out, ref)
        scaler = torch.amp.GradScaler("cuda")
        opt = torch.optim.SGD(blk.parameters(), lr=0.0)
        blk.zero_grad()
        with torch.autocast("cuda", torch.float16):
            loss = blk(x, None, idx).float().pow(2).mean()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
    live = [v.grad for v in blk.parameters() if v.grad is not None]
    assert len(live) == 19 and all(torch.isfinite(g).all() for g in live)
    L.modules._FP16_NOTE[0] = False


def test_gelu_table_edges_through_leff():
    """The branch-free GELU pair lookup defers its range test; inputs outside the 2^-28 <= |x| < 16 table (exact zeros,
    huge and tiny values) must take the exact slow path.  A LeFF whose depthwise kernel is zero makes the dwconv output
    equal its per-channel bias, so GELU sees exactly the values planted in b_dw."""
    import lewin_b200 as L
    dev = torch.device("cuda:0")
    C, B, hw = 32, 2, 16
    g = torch.Generator().manual_seed(5)
    y = torch.randn(B, hw * hw, C, generator=g).to(dev).to(torch.bfloat16)
    planted = torch.tensor([0.0, -0.0, 20.0, -20.0, 1e-10, -1e-10, 3e-5, -3e-5, 15.9, -15.9, 16.0, 1e30, -1e30, 0.5, -0.5, 2.0 ** -28])
    b_dw = planted.repeat(8)                                   # 128 hidden channels
    w1 = torch.randn(4 * C, C, generator=g) * 0.2
    w2 = torch.randn(C, 4 * C, generator=g) * 0.1
    args = dict(ln_w=torch.ones(C), ln_b=torch.zeros(C), w1=w1, b1=torch.zeros(4 * C), w_dw=torch.zeros(4 * C, 1, 3, 3),
                b_dw=b_dw, w2=w2, b2=torch.zeros(C))
    args = {k: v.to(dev) for k, v in args.items()}
    with torch.no_grad():
        out = L.ops.lewin_leff(y, B=B, H=hw, W=hw, **args)
    bb = b_dw.to(dev).to(torch.bfloat16)
    h2 = torch.nn.functional.gelu(bb.float()).to(torch.bfloat16)          # exact (erf) GELU of the bf16-rounded bias
    lin = (h2.float() @ w2.to(dev).to(torch.bfloat16).float().t()).to(torch.bfloat16)
    ref = (y.float() + lin.float()).to(torch.bfloat16)
    err = (out.float() - ref.float()).abs()
    tol = 2.0 ** -7 * ref.float().abs().clamp(min=1.0)
    assert torch.isfinite(out.float()).all()
    assert (err <= tol).all(), float((err / tol).max())


@pytest.mark.parametrize("Cin,Cout,B,hw", [(128, 32, 2, 16), (256, 64, 2, 16), (512, 128, 3, 16), (512, 256, 8, 8)])
def test_upsample_gemm_matches_conv_transpose(Cin, Cout, B, hw):
    """Upsample.forward (My_model_1.py:633-648) as a token GEMM with pixel-shuffle addressing vs torch's ConvTranspose2d in
    fp32 on the same bf16-rounded operands; also the fused torch.cat([up, skip], -1) form."""
    import lewin_b200 as L
    from lewin_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn(B, hw * hw, Cin, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(Cin, Cout, 2, 2, generator=g) * Cin ** -0.5).to(dev)
    b = (torch.randn(Cout, generator=g) * 0.1).to(dev)
    skip = torch.randn(B, 4 * hw * hw, Cout, generator=g).to(dev).to(torch.bfloat16)
    assert ops.upsample_supported(x, Cin, Cout, B * hw * hw)
    y = ops.lewin_upsample(x, w, b, B=B, H=hw, W=hw)
    xr = x.float().reshape(B, hw, hw, Cin).permute(0, 3, 1, 2)
    ref = torch.nn.functional.conv_transpose2d(xr, w.to(torch.bfloat16).float(), b.to(torch.bfloat16).float(), stride=2)
    ref = ref.permute(0, 2, 3, 1).reshape(B, 4 * hw * hw, Cout)
    err = (y.float() - ref).abs()
    tol = 2.0 ** -7 * ref.abs().clamp(min=1.0)             # bf16 output: 2 ulps
    assert (err <= tol).all(), float((err / tol).max())
    up = L.uformer.Upsample(Cin, Cout).to(dev)
    with torch.no_grad():
        up.deconv[0].weight.copy_(w); up.deconv[0].bias.copy_(b)
        cat = up(x, skip)
    assert cat.shape == (B, 4 * hw * hw, 2 * Cout)
    assert torch.equal(cat[..., :Cout], y) and torch.equal(cat[..., Cout:], skip)


def test_input_proj_fused_matches_autocast_reference():
    """InputProj (My_model_1.py:659-682): fused conv3x3 + bias + LeakyReLU kernel vs the stock torch path under bf16 autocast."""
    import lewin_b200 as L
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    ip = L.uformer.InputProj(in_channel=3, out_channel=32, kernel_size=3, stride=1, act_layer=torch.nn.LeakyReLU).to(dev)
    x = torch.rand(5, 3, 40, 24, device=dev)
    with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
        y = ip(x)                                                           # fused kernel
        ref = ip.proj(x.contiguous(memory_format=torch.channels_last))      # cuDNN + bias + LeakyReLU
    ref = ref.permute(0, 2, 3, 1).reshape(5, 40 * 24, 32)
    assert y.dtype == torch.bfloat16 and y.shape == ref.shape
    err = (y.float() - ref.float()).abs()
    tol = 2.0 ** -7 * ref.float().abs().clamp(min=0.25)
    assert (err <= tol).all(), float((err / tol).max())
