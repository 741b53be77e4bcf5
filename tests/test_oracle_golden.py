"""CPU: the numpy oracle is pinned against the golden vectors generated from the UNMODIFIED reference
(oracle/make_golden.py, run in the build container)."""
import numpy as np
import pytest

from oracle import lewin_oracle as O
from tests.util import (BF16_REF_FIXTURES, BLOCK_FIXTURES, COMPACT_FIXTURES, check_compact_grads, check_compact_grads_statistical,
                        load_fixture)


@pytest.mark.parametrize("name", BLOCK_FIXTURES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_oracle_forward_backward_matches_reference(name, dtype):
    fx = load_fixture(name)
    p = O.as_dtype(fx["params"], dtype)
    x = fx["x"].astype(dtype)
    out, aux = O.lewin_block(x, p, fx["shift"], fx["idx"], fx.get("input_mask"), True, fx.get("drop_scale"),
                             return_aux=True)
    same = np.array_equal(aux["top"], fx["top"])
    if not same:   # only near-tie rows may differ
        bad = (aux["top"] != fx["top"]).any(-1)
        assert (aux["rel_gap"][bad] < 1e-5).all()
        out = O.lewin_block(x, p, fx["shift"], fx["idx"], fx.get("input_mask"), True, fx.get("drop_scale"), top=fx["top"])
    scale = max(1.0, np.abs(fx["out"]).max())
    assert np.abs(out - fx["out"]).max() < 2e-4 * scale
    dx, g = O.lewin_block_bwd(fx["dout"].astype(dtype), x, p, fx["shift"], fx["idx"], fx.get("input_mask"), True,
                              fx.get("drop_scale"), top=fx["top"])
    assert np.abs(dx - fx["dx"]).max() < 1e-3 * np.abs(fx["dx"]).max()
    gscale = max(np.abs(v).max() for v in fx["grads"].values())
    assert sorted(g) == sorted(fx["grads"]) == sorted(O.GRAD_KEYS)
    for k in O.GRAD_KEYS:
        ref = fx["grads"][k]
        assert np.abs(g[k] - ref).max() < 1e-3 * max(np.abs(ref).max(), 1e-4 * gscale), k


@pytest.mark.parametrize("name", COMPACT_FIXTURES)
def test_oracle_matches_reference_at_deep_levels(name):
    """C = 256 / 512 blocks (8 / 16 heads) and head_dim 64 / 128 blocks recorded from the unmodified reference, compact fixtures: forward, dx, the
    selected query sets and all 19 parameter gradients (sampled elements + norms)."""
    fx = load_fixture(name)
    p = O.as_dtype(fx["params"], np.float64)
    x = fx["x"].astype(np.float64)
    out, aux = O.lewin_block(x, p, fx["shift"], fx["idx"], None, True, None, return_aux=True)
    bad = (aux["top"] != fx["top"]).any(-1)
    assert (aux["rel_gap"][bad] < 1e-5).all()
    if bad.any():
        out = O.lewin_block(x, p, fx["shift"], fx["idx"], None, True, None, top=fx["top"])
    assert np.abs(out - fx["out"]).max() < 2e-4 * max(1.0, np.abs(fx["out"]).max())
    dx, g = O.lewin_block_bwd(fx["dout"].astype(np.float64), x, p, fx["shift"], fx["idx"], None, True, None, top=fx["top"])
    assert np.abs(dx - fx["dx"]).max() < 1e-3 * np.abs(fx["dx"]).max()
    check_compact_grads(fx, g, 1e-3)


@pytest.mark.parametrize("name", BF16_REF_FIXTURES)
def test_bf16_oracle_matches_reference_under_cpu_autocast(name):
    """Pins the bf16 (autocast-emulating) oracle to the unmodified reference run under torch.autocast("cpu", bfloat16):
    the selected query sets agree on every row whose rank-25/26 gap exceeds the bf16 resolution (2^-7 of the M range), and
    with the reference's selection the block output agrees to ONE bf16 ulp of the activation scale (CPU autocast keeps
    LayerNorm / softmax results in bf16 where the CUDA policy the oracle follows keeps fp32, hence not bit for bit)."""
    fx = load_fixture(name)
    p = O.as_dtype(fx["params"], np.float32)
    out, aux = O.lewin_block(fx["x"], p, fx["shift"], fx["idx"], None, True, None, return_aux=True, bf16=True)
    bad = (aux["top"] != fx["top"]).any(-1)
    assert (aux["rel_gap"][bad] < 2.0 ** -7).all()
    if bad.any():
        out = O.lewin_block(fx["x"], p, fx["shift"], fx["idx"], None, True, None, top=fx["top"], bf16=True)
    ulp = 2.0 ** (np.floor(np.log2(np.abs(fx["out"]).max())) - 7)
    d = np.abs(out - fx["out"])
    assert d.max() <= ulp and d.mean() < 2e-3 and (d > 0).mean() < 0.3
    # backward: the reference's autocast gradients (fp32 parameter grads, bf16 dx) are the fp64 oracle's up to bf16 noise
    dx, g = O.lewin_block_bwd(fx["dout"].astype(np.float64), fx["x"].astype(np.float64), O.as_dtype(fx["params"], np.float64),
                              fx["shift"], fx["idx"], None, True, None, top=fx["top"])
    a, b = dx.ravel(), fx["dx"].astype(np.float64).ravel()
    assert a @ b / (np.linalg.norm(a) * np.linalg.norm(b)) > 0.9999
    assert np.abs(a - b).max() < 2e-2 * np.abs(b).max()
    check_compact_grads_statistical(fx, g, 0.999, 2e-2)


def test_whole_model_oracle_matches_reference_in_canvas_mode(golden_dir):
    """Canvas mode (test_long_GPU.py:74-93: wrap-pad, ONE forward over the 384^2 canvas of a 200 x 300 image, crop, clamp)
    recorded from the unmodified reference: the numpy whole-model oracle reproduces it - shift masks, LeFF borders and the
    construction-time shift rule (My_model_1.py:764-766) at a resolution other than the model's img_size."""
    import os
    import torch
    from oracle import param_fill, uformer_oracle as U
    import lewin_b200 as L
    from lewin_b200 import fullres
    z = np.load(os.path.join(golden_dir, "uformer32_canvas_200x300.npz"))
    model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff")
    param_fill.fill_module(model, int(z["seed"]))
    sd = {k: v.numpy() for k, v in model.state_dict().items()}
    canvas = fullres.wrap_pad(torch.from_numpy(z["x"]), ps=128).numpy()
    assert canvas.shape == (1, 3, 384, 384)
    rec = []
    raw = U.uformer_forward(canvas, sd, z["idx"].astype(np.int64), img_size=128, dtype=np.float32, record=rec)[:, :, :200, :300]
    # block by block: the selected query sets equal the reference's M_top on every row that is not a near-tie
    assert len(rec) == 18
    rows = 0
    for r in rec:
        ref_top = z[f"top{r['block']:02d}"].astype(np.int64)
        bad = (r["top"] != ref_top).any(-1)
        assert (r["rel_gap"][bad] < 1e-5).all(), r["block"]
        rows += bad.size
    assert rows == 26208
    e = np.abs(raw - z["y_raw"])
    assert np.median(e) < 1e-4 and (e > 1e-3).mean() < 0.02
    assert np.abs(np.clip(raw, 0, 1) - z["y"]).mean() < 1e-4


def test_prob_sizes():
    assert O.prob_sizes(64, 64) == (25, 25)      # attn.py:310-315


def test_shift_mask_only_on_last_window_row_and_column():
    m = O.shift_attn_mask(32, 32, 8, 4)
    nz = np.abs(m).reshape(16, -1).sum(1) > 0
    expect = np.array([(w // 4 == 3) or (w % 4 == 3) for w in range(16)])
    assert np.array_equal(nz, expect)
    assert set(np.unique(m)) <= {0.0, -100.0}


def test_mask_none_equals_shift0_and_window_permutation_equivariance():
    rng = np.random.default_rng(0)
    p = O.random_block_params(32, 1, rng, dtype=np.float64)
    idx = rng.integers(0, 64, (64, 25))
    xw = rng.standard_normal((6, 64, 32))
    out = O.window_attention(xw, p, None, idx)
    perm = rng.permutation(6)
    out_p = O.window_attention(xw[perm], p, None, idx)
    assert np.allclose(out[perm], out_p, atol=1e-12)
    zero_mask = np.zeros((3, 64, 64))
    assert np.allclose(O.window_attention(xw, p, zero_mask, idx), out, atol=1e-12)


def test_numeric_gradient_spot_check():
    """Finite differences on the fp64 oracle at a few coordinates (independent of the reference)."""
    rng = np.random.default_rng(3)
    C, nH, hw = 32, 1, 8
    p = O.random_block_params(C, nH, rng, dtype=np.float64)
    idx = rng.integers(0, 64, (64, 25))
    x = rng.standard_normal((1, hw * hw, C))
    dout = rng.standard_normal(x.shape)
    _, aux = O.lewin_block(x, p, 0, idx, return_aux=True)
    top = aux["top"]
    dx, g = O.lewin_block_bwd(dout, x, p, 0, idx, top=top)
    eps = 1e-6
    for (i, j) in [(3, 5), (40, 17), (63, 31)]:
        xp = x.copy(); xp[0, i, j] += eps
        xm = x.copy(); xm[0, i, j] -= eps
        num = ((O.lewin_block(xp, p, 0, idx, top=top) - O.lewin_block(xm, p, 0, idx, top=top)) * dout).sum() / (2 * eps)
        assert abs(num - dx[0, i, j]) < 1e-5 * max(1.0, abs(num))
    k = "attn.relative_position_bias_table"
    for r in (0, 112, 224):
        pp = dict(p); pp[k] = p[k].copy(); pp[k][r, 0] += eps
        pm = dict(p); pm[k] = p[k].copy(); pm[k][r, 0] -= eps
        num = ((O.lewin_block(x, pp, 0, idx, top=top) - O.lewin_block(x, pm, 0, idx, top=top)) * dout).sum() / (2 * eps)
        assert abs(num - g[k][r, 0]) < 1e-5 * max(1.0, abs(num))


@pytest.mark.parametrize("embed_dim,name", [(32, "uformer32_b2"), (64, "uformer64_b1")])
def test_torch_port_matches_reference_golden_whole_model(golden_dir, embed_dim, name):
    """(embed_dim 64 = the reference's head_dim 64 variant, My_model_1.py:962: C = 64 ... 1024, one recorded tile.)
    oracle/torch_port.py is the denominator of every cpu_baseline / --impl reference number bench.py prints: its whole-model
    forward must reproduce the UNMODIFIED reference's recorded output (tests/golden/uformer32_b2.npz, written by
    oracle/make_golden.py from /root/reference) on the same weights and the same 18 key-sample draws.  The port runs the
    reference's own ATen op sequence, so the only slack is near-tie top-u rows (none at fp32 accuracy on this fixture)."""
    import os
    import torch
    import lewin_b200 as L
    from oracle import param_fill, torch_port
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    model = L.Uformer(img_size=128, embed_dim=embed_dim, win_size=8, token_projection="linear", token_mlp="leff")
    param_fill.fill_module(model, int(z["seed"]))
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    with torch.no_grad():
        out = torch_port.uformer_forward(torch.from_numpy(z["x"]), sd, torch.from_numpy(z["idx"].astype(np.int64)))
    e = (out - torch.from_numpy(z["y"])).abs()
    scale = max(1.0, float(np.abs(z["y"]).max()))
    print(f"torch_port vs reference ({name}): max {float(e.max()):.3e} median {float(e.median()):.3e} (output scale {scale:.2f})")
    assert float(e.median()) < 1e-5 * scale
    assert float((e > 1e-3 * scale).float().mean()) < 1e-3
