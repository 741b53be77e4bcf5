"""The block halves as torch.library custom ops (compile_ops.py; SURVEY 8(b) "Python glue": register as custom ops so
torch.compile / export see ONE opaque node per half of LeWinTransformerBlock.forward, My_model_1.py:785-875).

CPU part: schemas, shape-only (meta) forward, and the registered autograd formula end to end on meta tensors.
GPU part: a compiled block is traced onto torch.ops.lewin_b200.{attn_fwd, leff_fwd}, runs the SAME kernels as the eager
autograd.Function path (forward bit-identical; gradients equal up to the atomics' summation order) and works under aot_eager
(which exercises the fake implementations of the backward ops as well)."""
import pytest
import torch


def _params(C, nH, dev, req=True, seed=0):
    g = torch.Generator().manual_seed(seed)

    def r(*s, sc=1.0):
        return (torch.randn(*s, generator=g) * sc).to(dev).requires_grad_(req)
    P = dict(ln_w=r(C, sc=0.1), ln_b=r(C, sc=0.1), w_qkv=r(3 * C, C, sc=C ** -0.5), b_qkv=r(3 * C, sc=0.1),
             w_out=r(C, C, sc=C ** -0.5), b_out=r(C, sc=0.1), rpb_table=r(225, nH, sc=0.2))
    Q = dict(ln_w=r(C, sc=0.1), ln_b=r(C, sc=0.1), w1=r(4 * C, C, sc=C ** -0.5), b1=r(4 * C, sc=0.1),
             w_dw=r(4 * C, 1, 3, 3, sc=0.3), b_dw=r(4 * C, sc=0.1), w2=r(C, 4 * C, sc=(4 * C) ** -0.5), b2=r(C, sc=0.1))
    return P, Q


def test_custom_ops_are_registered_with_the_documented_schemas():
    import lewin_b200  # noqa: F401
    for name, n_out in (("attn_fwd", 4), ("attn_bwd", 8), ("leff_fwd", 5), ("leff_bwd", 9)):
        op = getattr(torch.ops.lewin_b200, name).default
        assert len(op._schema.returns) == n_out, (name, op._schema)
        assert not any(a.alias_info is not None for a in op._schema.arguments), name      # functional: nothing mutated


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_meta_forward_and_registered_autograd_shapes(dtype):
    """Meta tensors go through the fake forward, the registered autograd formula and the fake backward: every live parameter
    gets a gradient of its own shape in fp32, x a gradient of its own dtype; absent parameters (windowed mode has no LN, the
    dense-bias form has no table) get none."""
    from lewin_b200 import compile_ops
    C, nH, B, H, W = 64, 2, 2, 16, 16
    P, Q = _params(C, nH, "meta")
    x = torch.empty(B, H * W, C, device="meta", dtype=dtype, requires_grad=True)
    idx = torch.empty(64, 25, dtype=torch.int64, device="meta")
    y, top = compile_ops.lewin_attn(x, B=B, H=H, W=W, num_heads=nH, shift=4, rpb_dense=None, index_sample=idx, mask=None,
                                    drop_scale=None, windowed=False, use_rpb=True, analytic_shift_mask=True, need=True, **P)
    assert y.shape == x.shape and y.dtype == dtype and top.shape == (B * H * W // 64, nH, 25) and top.dtype == torch.uint8
    out = compile_ops.lewin_leff(y, B=B, H=H, W=W, drop_scale=None, fused=True, need=True, **Q)
    leaves = [x] + list(P.values()) + list(Q.values())
    grads = torch.autograd.grad(out.float().sum(), leaves)
    assert grads[0].shape == x.shape and grads[0].dtype == dtype
    for t, g in zip(leaves[1:], grads[1:]):
        assert g.shape == t.shape and g.dtype == torch.float32

    # WindowAttention.forward form: pre-partitioned windows, no LayerNorm, gathered dense bias instead of the table
    xw = torch.empty(8, 64, C, device="meta", dtype=dtype, requires_grad=True)
    dense = torch.empty(nH, 64, 64, device="meta", requires_grad=True)
    Pw = {k: v for k, v in P.items() if k not in ("ln_w", "ln_b", "rpb_table")}
    yw, _ = compile_ops.lewin_attn(xw, B=8, H=8, W=8, num_heads=nH, shift=0, ln_w=None, ln_b=None, rpb_table=None,
                                   rpb_dense=dense, index_sample=idx, mask=None, drop_scale=None, windowed=True, use_rpb=True,
                                   analytic_shift_mask=False, need=True, **Pw)
    gx, gd = torch.autograd.grad(yw.float().sum(), [xw, dense])
    assert gx.shape == xw.shape and gd.shape == dense.shape

    # an inference call (save=False) keeps nothing and refuses a backward with the reason
    yi, _ = compile_ops.lewin_attn(x, B=B, H=H, W=W, num_heads=nH, shift=4, rpb_dense=None, index_sample=idx, mask=None,
                                   drop_scale=None, windowed=False, use_rpb=True, analytic_shift_mask=True, need=False, **P)
    with pytest.raises(RuntimeError, match="save=False"):
        yi.float().sum().backward()


def test_dynamo_traces_a_whole_block_into_one_graph_with_one_node_per_half():
    """torch.compile of LeWinTransformerBlock on meta tensors (no GPU needed to trace): no graph break, and the captured graph
    holds exactly one attn_fwd and one leff_fwd node; 19 live parameters receive gradients through the registered formula."""
    import torch._dynamo
    from torch._dynamo.utils import counters
    import lewin_b200 as L
    blk = L.LeWinTransformerBlock(dim=64, input_resolution=(16, 16), num_heads=2, win_size=8, shift_size=4).to("meta")
    idx = torch.randint(64, (64, 25))
    x = torch.empty(2, 256, 64, device="meta", requires_grad=True)
    graphs = []

    def backend(gm, example_inputs):
        graphs.append([str(n.target) for n in gm.graph.nodes if n.op == "call_function"])
        return gm.forward

    torch._dynamo.reset()
    counters.clear()
    out = torch.compile(blk, backend=backend)(x, None, idx)
    assert len(graphs) == 1 and not dict(counters["graph_break"]), (graphs, dict(counters["graph_break"]))
    assert sum("lewin_b200.attn_fwd" in t for t in graphs[0]) == 1 and sum("lewin_b200.leff_fwd" in t for t in graphs[0]) == 1
    out.backward(torch.empty_like(out))
    assert x.grad is not None and sum(p.grad is not None for p in blk.parameters()) == 19
    torch._dynamo.reset()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_compiled_block_is_traced_onto_the_custom_ops_and_matches_eager(dtype):
    import lewin_b200 as L
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    blk = L.LeWinTransformerBlock(dim=64, input_resolution=(16, 16), num_heads=2, win_size=8, shift_size=4).to(dev)
    with torch.no_grad():
        blk.attn.relative_position_bias_table.normal_(std=0.3)
    idx = torch.randint(64, (64, 25))
    x0 = torch.randn(2, 256, 64, device=dev).to(dtype)
    dout = torch.randn(2, 256, 64, device=dev).to(dtype)

    def run(mod):
        for p in blk.parameters():
            p.grad = None
        x = x0.clone().requires_grad_(True)
        out = mod(x, None, idx)
        out.backward(dout)
        return out.detach(), x.grad, {k: p.grad.clone() for k, p in blk.named_parameters() if p.grad is not None}

    out_e, dx_e, g_e = run(blk)
    seen = []

    def backend(gm, example_inputs):
        seen.extend(str(n.target) for n in gm.graph.nodes if n.op == "call_function")
        return gm.forward

    torch._dynamo.reset()
    out_c, dx_c, g_c = run(torch.compile(blk, backend=backend))
    assert any("lewin_b200.attn_fwd" in s for s in seen) and any("lewin_b200.leff_fwd" in s for s in seen), seen
    assert torch.equal(out_c, out_e)                       # the same kernels ran
    assert sorted(g_c) == sorted(g_e)
    tol = 1e-4 if dtype == torch.float32 else 2e-2       # weight gradients are summed with atomics (order varies)
    assert float((dx_c.float() - dx_e.float()).abs().max()) <= tol * float(dx_e.float().abs().max())
    gscale = max(float(v.abs().max()) for v in g_e.values())

    def close(a, b, k):          # the key bias' gradient is analytically zero (rounding noise): floor every scale at 1e-3 of the largest
        return float((a - b).abs().max()) <= tol * max(float(b.abs().max()), 1e-3 * gscale)
    for k in g_e:
        assert close(g_c[k], g_e[k], k), k

    # aot_eager traces forward AND backward ahead of time through the fake implementations
    torch._dynamo.reset()
    out_a, dx_a, g_a = run(torch.compile(blk, backend="aot_eager"))
    assert torch.equal(out_a, out_e)
    assert float((dx_a.float() - dx_e.float()).abs().max()) <= tol * float(dx_e.float().abs().max())
    for k in g_e:
        assert close(g_a[k], g_e[k], k), k


@pytest.mark.gpu
def test_compiled_bf16_inference_block_matches_eager_bitwise():
    """Inference (no grad, bf16): the compiled call reaches the fused attention / LeFF-tail kernels through attn_fwd / leff_fwd
    with save=False and returns what the eager call returns."""
    import lewin_b200 as L
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    blk = L.LeWinTransformerBlock(dim=32, input_resolution=(32, 32), num_heads=1, win_size=8, shift_size=0).to(dev).eval()
    idx = torch.randint(64, (64, 25))
    x = torch.randn(3, 1024, 32, device=dev).to(torch.bfloat16)
    torch._dynamo.reset()
    with torch.no_grad():
        ref = blk(x, None, idx)
        got = torch.compile(blk, backend="aot_eager")(x, None, idx)
    assert torch.equal(ref, got)
