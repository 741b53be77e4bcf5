"""GPU parity of the reference's inner drop-in surfaces, each through its own C-ABI entry point:

  lewin_probsparse_core_{fwd,bwd}_{f32,bf16}   ProbAttention.forward + autograd      ProbSparse/attn.py:287-342
  modules.ProbAttention.forward                same, module surface (q, k, v [B_,64,nH,D], gathered bias, SW mask)
  modules.AttentionLayer.forward               q/k/v/out linears around it           ProbSparse/attn.py:385-461

against oracle.lewin_oracle (prob_attention / window_attention and their hand-derived backward), including head_dim 64 and
128 (BASELINE config 5: embed_dim 32-128; head_dim = embed_dim, My_model_1.py:962) and the gradient w.r.t. the GATHERED
relative-position bias that AttentionLayer.forward receives (it used to come back as None).
Tolerances: fp32 max-abs 1e-3, bf16 2e-2 (north_star); selections tie-aware (tests/util.check_top)."""
import numpy as np
import pytest
import torch

from oracle import lewin_oracle as O
from tests.util import TIE_TAU_F32, TOL_BF16, TOL_F32, check_top

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TIE_TAU_BF16 = 2.0 ** -7


def _qkv_case(seed, B_, nH, D, nW_mask=None, scale=1.0):
    rng = np.random.default_rng(seed)
    q = (rng.standard_normal((B_, 64, nH, D)) * scale).astype(np.float32)
    k = (rng.standard_normal((B_, 64, nH, D)) * scale).astype(np.float32)
    v = rng.standard_normal((B_, 64, nH, D)).astype(np.float32)
    rpb = (rng.standard_normal((nH, 64, 64)) * 0.5).astype(np.float32)
    idx = rng.integers(0, 64, size=(64, 25)).astype(np.int64)
    mask = None
    if nW_mask:
        mask = np.where(rng.random((nW_mask, 64, 64)) < 0.2, -100.0, 0.0).astype(np.float32)
    return q, k, v, rpb, mask, idx


def _forced_oracle(q, k, v, rpb, mask, idx, top_gpu, dtype, bf16):
    """Oracle output with the device's selection forced (ambiguous rows may legitimately fall either way)."""
    top = np.sort(np.asarray(top_gpu).astype(np.int64), -1)
    return O.prob_attention(q.astype(dtype), k.astype(dtype), v.astype(dtype), rpb.astype(dtype),
                            None if mask is None else mask.astype(dtype), idx, top=top, return_aux=True, bf16=bf16)


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("B_,nH,D,nW", [(6, 2, 32, None), (8, 1, 32, 4), (4, 2, 64, 2), (3, 1, 128, None), (2, 16, 32, None)])
def test_probsparse_core_entry_forward_matches_oracle(dt, B_, nH, D, nW):
    """lewin_probsparse_core_fwd_* == ProbAttention.forward (attn.py:287-342) with a gathered bias and a dense mask."""
    import lewin_b200 as L
    bf = dt == "bf16"
    q, k, v, rpb, mask, idx = _qkv_case(100 + B_ + D, B_, nH, D, nW, scale=0.7)
    if bf:      # the bf16 entry takes bf16 activations: start from bf16-representable values on both sides
        q, k, v = (O.rbf(t) for t in (q, k, v))
    tdt = torch.bfloat16 if bf else torch.float32
    C = nH * D
    qkv = torch.from_numpy(np.concatenate([q.reshape(B_, 64, C), k.reshape(B_, 64, C), v.reshape(B_, 64, C)], -1)).to(DEV, tdt)
    out, top = L.ops.probsparse_core(qkv, num_heads=nH, index_sample=torch.from_numpy(idx), rpb_dense=torch.from_numpy(rpb).to(DEV),
                                     mask=None if mask is None else torch.from_numpy(mask).to(DEV), return_top=True)
    torch.cuda.synchronize()
    wide = np.float64 if not bf else np.float32
    _, aux = O.prob_attention(q.astype(wide), k.astype(wide), v.astype(wide), rpb.astype(wide),
                              None if mask is None else mask.astype(wide), idx, return_aux=True, bf16=bf)
    nbad, namb, nhard = check_top(top.cpu().numpy(), aux["top"], aux["rel_gap"], TIE_TAU_BF16 if bf else TIE_TAU_F32)
    assert nhard == 0, f"{nhard} non-ambiguous rows differ ({nbad} differ, {namb} ambiguous)"
    ref, _ = _forced_oracle(q, k, v, rpb, mask, idx, top.cpu().numpy(), wide, bf)
    err = np.abs(out.float().cpu().numpy().reshape(B_, 64, nH, D) - ref).max()
    assert err < (TOL_BF16 if bf else TOL_F32), err


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("B_,nH,D,table", [(5, 2, 32, False), (4, 1, 32, True), (3, 2, 64, False), (2, 1, 128, True)])
def test_probsparse_core_entry_backward_matches_oracle(dt, B_, nH, D, table):
    """lewin_probsparse_core_bwd_*: dq | dk | dv and the bias gradient (gathered [nH,64,64] or table [225,nH]) for the
    selection the forward saved, against the oracle's hand-derived backward with the same selection."""
    import lewin_b200 as L
    bf = dt == "bf16"
    q, k, v, rpb, mask, idx = _qkv_case(7 + D + B_, B_, nH, D, None, scale=0.7)
    rng = np.random.default_rng(3)
    tab = (rng.standard_normal((225, nH)) * 0.5).astype(np.float32)
    if table:
        rpb = O.rpb_from_table(tab)
    if bf:
        q, k, v = (O.rbf(t) for t in (q, k, v))
    dctx = rng.standard_normal((B_, 64, nH, D)).astype(np.float32)
    if bf:
        dctx = O.rbf(dctx)
    tdt = torch.bfloat16 if bf else torch.float32
    C = nH * D
    qkv = torch.from_numpy(np.concatenate([q.reshape(B_, 64, C), k.reshape(B_, 64, C), v.reshape(B_, 64, C)], -1)).to(DEV, tdt)
    qkv.requires_grad_(True)
    bias = torch.from_numpy(tab if table else rpb).to(DEV).requires_grad_(True)
    kw = dict(rpb_table=bias) if table else dict(rpb_dense=bias)
    out, top = L.ops.probsparse_core(qkv, num_heads=nH, index_sample=torch.from_numpy(idx), return_top=True, **kw)
    out.backward(torch.from_numpy(dctx.reshape(B_, 64, C)).to(DEV, tdt))
    torch.cuda.synchronize()
    _, aux = _forced_oracle(q, k, v, rpb, None, idx, top.cpu().numpy(), np.float64, False)
    dq, dk, dv, drpb, da = O.prob_attention_bwd(dctx.astype(np.float64).transpose(0, 2, 1, 3), aux)
    ref = np.concatenate([t.transpose(0, 2, 1, 3).reshape(B_, 64, C) for t in (dq, dk, dv)], -1)
    got = qkv.grad.float().cpu().numpy()
    rel = np.abs(got - ref).max() / np.abs(ref).max()
    assert rel < (3e-2 if bf else 1e-3), rel
    if table:
        ri = O.relative_position_index()
        dref = np.zeros((225, nH))
        np.add.at(dref, (np.broadcast_to(ri[None], (nH, 64, 64)).reshape(-1), np.repeat(np.arange(nH), 4096)), drpb.reshape(-1))
    else:
        dref = drpb
    gb = bias.grad.cpu().numpy()
    assert gb.shape == dref.shape
    relb = np.abs(gb - dref).max() / np.abs(dref).max()
    assert relb < (3e-2 if bf else 1e-3), relb


@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_prob_attention_module_matches_oracle_and_reference_signature(dt):
    """modules.ProbAttention(mask_flag, factor, scale, attention_dropout, output_attention).forward(queries, keys, values,
    relative_position_bias, SW_mask, attn_mask) -> (context [B_,64,nH,D] contiguous, None)  (attn.py:55, 287, 342); the key
    samples are drawn from the CPU generator with the reference's own call when not passed (attn.py:91)."""
    import lewin_b200 as L
    bf = dt == "bf16"
    B_, nH, D = 8, 2, 32
    q, k, v, rpb, mask, idx = _qkv_case(11, B_, nH, D, 4, scale=0.7)
    pa = L.ProbAttention(False, 5, None, 0.1, False)
    assert len(list(pa.parameters())) == 0 and len(pa.state_dict()) == 0       # contributes no state_dict keys (Appendix B)
    tq, tk, tv = (torch.from_numpy(t).to(DEV).requires_grad_(True) for t in (q, k, v))
    tb = torch.from_numpy(rpb).to(DEV).requires_grad_(True)
    torch.manual_seed(5)
    idx_ref = torch.randint(64, (64, 25)).numpy()          # what attn.py:91 draws at this point of the CPU stream
    after_ref = torch.randint(64, (3,))                    # ... and where the stream stands afterwards
    torch.manual_seed(5)
    import contextlib
    ctxm = torch.autocast("cuda", torch.bfloat16) if bf else contextlib.nullcontext()
    with L.ops.TopRecorder() as _unused, ctxm:
        ctx, none = pa(tq, tk, tv, tb, torch.from_numpy(mask).to(DEV))
    assert none is None and ctx.shape == (B_, 64, nH, D) and ctx.is_contiguous()
    assert ctx.dtype == (torch.bfloat16 if bf else torch.float32)
    assert torch.equal(torch.randint(64, (3,)), after_ref), "the module must consume the CPU RNG exactly like attn.py:91"
    wide = np.float32 if bf else np.float64
    qq, kk, vv = ((O.rbf(t) if bf else t) for t in (q, k, v))
    ref, aux = O.prob_attention(qq.astype(wide), kk.astype(wide), vv.astype(wide), rpb.astype(wide), mask.astype(wide), idx_ref,
                                return_aux=True, bf16=bf)
    got = ctx.detach().float().cpu().numpy()
    bad = np.abs(got - ref).reshape(B_, 64, nH, D).max(axis=(1, 3)) > (TOL_BF16 if bf else TOL_F32)     # per (window, head)
    amb = aux["rel_gap"] < (TIE_TAU_BF16 if bf else TIE_TAU_F32)
    assert not (bad & ~amb).any(), "a non-ambiguous (window, head) row is off: wrong selection or arithmetic"
    # gradients flow to q, k, v AND the gathered bias
    dctx = torch.randn_like(ctx)
    ctx.backward(dctx)
    for t in (tq, tk, tv, tb):
        assert t.grad is not None and torch.isfinite(t.grad).all() and float(t.grad.abs().max()) > 0


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("C,nH", [(64, 2), (64, 1), (128, 1)])
def test_attention_layer_module_forward_backward_matches_oracle(dt, C, nH):
    """modules.AttentionLayer(d_model, n_heads).forward(x, x, x, relative_position_bias, SW_mask) -> (out [B_,64,C], None)
    (attn.py:357, 385-461) fed the way the reference's WindowAttention feeds it (My_model_1.py:408-413): the bias is GATHERED
    from the table by the caller, so the table's gradient arrives through the op's d(relative_position_bias)."""
    import lewin_b200 as L
    bf = dt == "bf16"
    rng = np.random.default_rng(C + nH)
    p = O.random_block_params(C, nH, rng)
    B_ = 6
    xw = rng.standard_normal((B_, 64, C)).astype(np.float32)
    idx = rng.integers(0, 64, size=(64, 25)).astype(np.int64)
    mask = np.where(rng.random((3, 64, 64)) < 0.15, -100.0, 0.0).astype(np.float32)
    layer = L.AttentionLayer(C, nH)
    pre = "attn.ProbSpare."
    sd = {k[len(pre):]: torch.from_numpy(v) for k, v in p.items() if k.startswith(pre)}
    layer.load_state_dict(sd, strict=True)
    layer = layer.to(DEV)
    table = torch.from_numpy(p["attn.relative_position_bias_table"]).to(DEV).requires_grad_(True)
    rel_index = torch.from_numpy(O.relative_position_index()).to(DEV)
    x = torch.from_numpy(O.rbf(xw) if bf else xw).to(DEV).requires_grad_(True)
    import contextlib
    ctxm = torch.autocast("cuda", torch.bfloat16) if bf else contextlib.nullcontext()
    with L.ops.TopRecorder() as rec, ctxm:
        rpb = table[rel_index.view(-1)].view(64, 64, -1).permute(2, 0, 1).contiguous()        # My_model_1.py:408-410
        out, none = layer(x, x, x, rpb, torch.from_numpy(mask).to(DEV), index_sample=torch.from_numpy(idx))
    assert none is None and out.shape == (B_, 64, C)
    top = np.sort(rec.tops[0].cpu().numpy().astype(np.int64), -1)
    wide = np.float32 if bf else np.float64
    pw = O.as_dtype(p, wide)
    xin = (O.rbf(xw) if bf else xw).astype(wide)
    _, aux = O.window_attention(xin, pw, mask.astype(wide), idx, return_aux=True, bf16=bf)
    nbad, namb, nhard = check_top(top, aux["top"], aux["rel_gap"], TIE_TAU_BF16 if bf else TIE_TAU_F32)
    assert nhard == 0
    ref = O.window_attention(xin, pw, mask.astype(wide), idx, top=top, bf16=bf)
    err = np.abs(out.detach().float().cpu().numpy() - ref).max()
    assert err < (TOL_BF16 if bf else TOL_F32), err
    # backward: all 8 linear parameters, x, and the relative_position_bias_table through the gathered bias
    dout = rng.standard_normal((B_, 64, C)).astype(np.float32)
    out.backward(torch.from_numpy(dout).to(DEV, out.dtype))
    dx_ref, g_ref = O.window_attention_bwd(dout.astype(np.float64), xin.astype(np.float64), O.as_dtype(p, np.float64),
                                           mask.astype(np.float64), idx, top=top)
    tol = 5e-2 if bf else 1e-3
    assert np.abs(x.grad.float().cpu().numpy() - dx_ref).max() / np.abs(dx_ref).max() < tol
    assert table.grad is not None, "relative_position_bias_table got no gradient through AttentionLayer"
    gt = g_ref["attn.relative_position_bias_table"]
    assert np.abs(table.grad.cpu().numpy() - gt).max() / np.abs(gt).max() < tol
    gscale = max(np.abs(v).max() for v in g_ref.values())
    for name in ("query", "key", "value", "out"):
        for wb in ("weight", "bias"):
            got = getattr(getattr(layer, name + "_projection"), wb).grad.cpu().numpy()
            ref_g = g_ref[pre + f"{name}_projection.{wb}"]
            assert np.abs(got - ref_g).max() < tol * max(np.abs(ref_g).max(), 1e-3 * gscale), (name, wb)


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("C,nH,hw,shift", [(64, 1, 16, 4), (128, 1, 16, 0), (128, 2, 24, 4), (256, 2, 16, 4)])
def test_block_forward_backward_head_dim_64_128(dt, C, nH, hw, shift):
    """BASELINE config 5 sweeps embed_dim 32-128 and head_dim == embed_dim in this model (My_model_1.py:962;
    attn.py:370-372 d_keys = d_model // n_heads): the LeWin block at head_dim 64 and 128, forward and backward."""
    import lewin_b200 as L
    bf = dt == "bf16"
    rng = np.random.default_rng(C * 7 + nH + hw)
    p = O.random_block_params(C, nH, rng)
    B = 2
    x = rng.standard_normal((B, hw * hw, C)).astype(np.float32)
    if bf:
        x = O.rbf(x)
    idx = rng.integers(0, 64, size=(64, 25)).astype(np.int64)
    blk = L.LeWinTransformerBlock(dim=C, input_resolution=(128, 128), num_heads=nH, win_size=8, shift_size=shift)
    sd = blk.state_dict()
    for k_, v_ in p.items():
        sd[k_].copy_(torch.from_numpy(v_))
    blk = blk.to(DEV).eval()
    xt = torch.from_numpy(x).to(DEV, torch.bfloat16 if bf else torch.float32).requires_grad_(True)
    with L.ops.TopRecorder() as rec:
        out = blk(xt, None, torch.from_numpy(idx))
    top = np.sort(rec.tops[0].cpu().numpy().astype(np.int64), -1)
    wide = np.float32 if bf else np.float64
    _, aux = O.lewin_block(x.astype(wide), O.as_dtype(p, wide), shift, idx, return_aux=True, bf16=bf)
    nbad, namb, nhard = check_top(top, aux["top"], aux["rel_gap"], TIE_TAU_BF16 if bf else TIE_TAU_F32)
    assert nhard == 0, (nbad, namb, nhard)
    ref = O.lewin_block(x.astype(wide), O.as_dtype(p, wide), shift, idx, top=top, bf16=bf)
    scale = max(1.0, np.abs(ref).max()) if bf else 1.0          # bf16: one ulp grows with the activation magnitude
    err = np.abs(out.detach().float().cpu().numpy() - ref).max()
    assert err < (TOL_BF16 * scale if bf else TOL_F32), err
    dout = rng.standard_normal(x.shape).astype(np.float32)
    out.backward(torch.from_numpy(dout).to(DEV, out.dtype))
    dx_ref, g_ref = O.lewin_block_bwd(dout.astype(np.float64), x.astype(np.float64), O.as_dtype(p, np.float64), shift, idx, top=top)
    tol = 6e-2 if bf else 1e-3
    dxg = xt.grad.float().cpu().numpy()
    if bf:      # bf16 backward against the fp64 oracle: direction and worst element (rounding of ~10 chained bf16 operands)
        cos = float((dxg.ravel() @ dx_ref.ravel()) / (np.linalg.norm(dxg) * np.linalg.norm(dx_ref)))
        assert cos > 0.999 and np.abs(dxg - dx_ref).max() / np.abs(dx_ref).max() < 0.12, (cos, np.abs(dxg - dx_ref).max())
    else:
        assert np.abs(dxg - dx_ref).max() / np.abs(dx_ref).max() < tol
    grads = {k_: v_.grad for k_, v_ in blk.named_parameters() if v_.grad is not None}
    assert sorted(grads) == sorted(O.GRAD_KEYS)
    if not bf:
        gscale = max(np.abs(v_).max() for v_ in g_ref.values())
        for k_ in O.GRAD_KEYS:
            assert np.abs(grads[k_].cpu().numpy() - g_ref[k_]).max() < tol * max(np.abs(g_ref[k_]).max(), 1e-3 * gscale), k_
    else:
        for k_ in O.GRAD_KEYS:
            if k_.endswith("key_projection.bias"):
                continue                                   # analytically zero
            a, b = grads[k_].cpu().numpy().ravel().astype(np.float64), g_ref[k_].ravel()
            cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
            assert cos > 0.99, (k_, cos)
