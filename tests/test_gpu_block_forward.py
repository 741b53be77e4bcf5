"""GPU parity of the LeWin block forward (C ABI path) against the reference-generated golden fixtures
and the numpy oracle.  fp32 tolerance: max-abs 1e-3 (north_star); top-u sets must match exactly on all
non-ambiguous rows."""
import numpy as np
import pytest
import torch

from oracle import lewin_oracle as O
from tests.util import BLOCK_FIXTURES, TIE_TAU_F32, TOL_F32, check_top, force_drop_scales, load_fixture, make_block

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", BLOCK_FIXTURES)
def test_block_forward_matches_reference_golden_f32(name):
    fx = load_fixture(name)
    dev = torch.device("cuda:0")
    blk = make_block(fx, dev)
    train = "drop_scale" in fx
    blk.train(train)
    if train:
        force_drop_scales(blk, fx["drop_scale"], dev)
    x = torch.from_numpy(fx["x"]).to(dev)
    mask = torch.from_numpy(fx["input_mask"]).to(dev) if "input_mask" in fx else None
    captured = {}
    import lewin_b200.ops as ops
    orig = ops.lewin_attn

    def spy(*a, **k):
        y, top = orig(*a, return_top=True, **k)
        captured["top"] = top
        return y

    ops.lewin_attn = spy
    try:
        with torch.no_grad():
            out = blk(x, mask, torch.from_numpy(fx["idx"]))
    finally:
        ops.lewin_attn = orig
    torch.cuda.synchronize()
    # tie-aware index check against the reference's own M_top
    _, aux = O.lewin_block(fx["x"].astype(np.float64), O.as_dtype(fx["params"], np.float64), fx["shift"], fx["idx"],
                           fx.get("input_mask"), True, fx.get("drop_scale"), return_aux=True)
    nbad, namb, nhard = check_top(captured["top"].cpu().numpy(), fx["top"], aux["rel_gap"], TIE_TAU_F32)
    assert nhard == 0, f"{nhard} non-ambiguous (window,head) rows selected different top-u queries"
    err = np.abs(out.cpu().numpy() - fx["out"]).max()
    if nbad == 0:
        assert err < TOL_F32, err
    else:   # ambiguous rows legitimately differ: compare against the oracle forced to the GPU's selection
        top = np.sort(captured["top"].cpu().numpy().astype(np.int64), -1)
        ref = O.lewin_block(fx["x"].astype(np.float64), O.as_dtype(fx["params"], np.float64), fx["shift"], fx["idx"],
                            fx.get("input_mask"), True, fx.get("drop_scale"), top=top)
        assert np.abs(out.cpu().numpy() - ref).max() < TOL_F32


@pytest.mark.parametrize("C,nH,hw,B,shift", [(32, 1, 8, 1, 0), (32, 1, 24, 3, 4), (64, 2, 32, 2, 4),
                                             (256, 8, 16, 1, 4), (512, 16, 8, 5, 0), (96, 3, 16, 2, 4)])
def test_block_forward_matches_oracle_f32(C, nH, hw, B, shift):
    rng = np.random.default_rng(C + hw + shift)
    p = O.random_block_params(C, nH, rng)
    x = rng.standard_normal((B, hw * hw, C)).astype(np.float32)
    idx = rng.integers(0, 64, size=(64, 25)).astype(np.int64)
    ref, aux = O.lewin_block(x.astype(np.float64), O.as_dtype(p, np.float64), shift, idx, return_aux=True)
    dev = torch.device("cuda:0")
    import lewin_b200 as L
    blk = L.LeWinTransformerBlock(dim=C, input_resolution=(128, 128), num_heads=nH, win_size=8, shift_size=shift)
    sd = blk.state_dict()
    for k, v in p.items():
        sd[k].copy_(torch.from_numpy(v))
    blk = blk.to(dev).eval()
    xs = torch.from_numpy(x).to(dev)
    with torch.no_grad():
        y, top = L.ops.lewin_attn(
            xs, B=B, H=hw, W=hw, num_heads=nH, shift=shift, ln_w=blk.norm1.weight, ln_b=blk.norm1.bias,
            w_qkv=blk.attn.ProbSpare.qkv_weights()[0], b_qkv=blk.attn.ProbSpare.qkv_weights()[1],
            w_out=blk.attn.ProbSpare.out_projection.weight, b_out=blk.attn.ProbSpare.out_projection.bias,
            rpb_table=blk.attn.relative_position_bias_table, index_sample=torch.from_numpy(idx), return_top=True)
        out = blk(xs, None, torch.from_numpy(idx))
    nbad, namb, nhard = check_top(top.cpu().numpy(), aux["top"], aux["rel_gap"], TIE_TAU_F32)
    assert nhard == 0
    if nbad:
        ref, aux = O.lewin_block(x.astype(np.float64), O.as_dtype(p, np.float64), shift, idx,
                                 top=np.sort(top.cpu().numpy().astype(np.int64), -1), return_aux=True)
    assert np.abs(y.cpu().numpy() - aux["y"]).max() < TOL_F32
    assert np.abs(out.cpu().numpy() - ref).max() < TOL_F32


def test_uformer_forward_matches_reference_golden_f32(golden_dir):
    """Whole model (18 chained blocks) vs the reference's fp32 CPU output.  The model is chaotic in the
    top-u selection: a near-tie flip (gap < 1e-5 of the M range) swaps one query between attention and
    mean(V) and the difference spreads through the following 3x3/4x4 convolutions, so max-abs over the
    image is not a meaningful bound (the numpy oracle itself differs from the reference in 0.07 % of the
    pixels).  Bound the bulk instead: median error, fraction of pixels off by > 1e-3, and PSNR between
    the two outputs.  cuDNN TF32 is disabled so the out-of-scope convolutions are true fp32 like the
    CPU reference."""
    import os
    import lewin_b200 as L
    from oracle import param_fill
    z = np.load(os.path.join(golden_dir, "uformer32_b2.npz"))
    dev = torch.device("cuda:0")
    model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff")
    param_fill.fill_module(model, int(z["seed"]))
    model = model.to(dev).eval()
    x = torch.from_numpy(z["x"]).to(dev)
    idx = torch.from_numpy(z["idx"].astype(np.int64))
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            y = model(x, index_samples=idx)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    e = np.abs(y.cpu().numpy() - z["y"]).ravel()
    mse = float((e ** 2).mean())
    psnr = 10 * np.log10(1.0 / max(mse, 1e-20))
    print(f"uformer32_b2: max {e.max():.3e} median {np.median(e):.3e} frac>1e-3 {(e > 1e-3).mean():.4f} psnr {psnr:.1f} dB")
    assert np.median(e) < 1e-4
    assert (e > 1e-3).mean() < 0.02
    assert psnr > 50.0
    # north_star: dehazed-image PSNR must agree within 0.01 dB.  Score both outputs against the same target image (the
    # model's input stands in for the ground truth of the synthetic fixture).
    def psnr_to(a, t):
        return 10 * np.log10(1.0 / float(((np.clip(a, 0, 1) - t) ** 2).mean()))
    p_ours, p_ref = psnr_to(y.cpu().numpy(), z["x"]), psnr_to(z["y"], z["x"])
    print(f"image PSNR vs target: ours {p_ours:.4f} dB, reference {p_ref:.4f} dB")
    assert abs(p_ours - p_ref) < 0.01
    # bf16 autocast path (BASELINE's compute dtype) on the same fixture
    with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
        yb = model(x, index_samples=idx).float()
    p_bf16 = psnr_to(yb.cpu().numpy(), z["x"])
    print(f"image PSNR vs target: bf16 {p_bf16:.4f} dB")
    assert abs(p_bf16 - p_ref) < 0.01


def test_embed_dim_64_model_matches_reference_golden(golden_dir):
    """The reference's embed_dim 64 variant (head_dim = embed_dim = 64 at every level, My_model_1.py:962; C = 64 ... 1024, 16 heads
    of 64 at the bottleneck) against a whole-model recording of the UNMODIFIED reference (tests/golden/uformer64_b1.npz, one
    128 x 128 tile, the reference's own 18 key-sample draws): BASELINE config 5's embed_dim sweep at model level.  Same
    criteria as the embed_dim 32 fixture above (the model is chaotic in the selection: bound the bulk), relative to the
    output's scale; then the bf16 autocast path within the bf16 tolerance."""
    import os
    import lewin_b200 as L
    from oracle import param_fill
    z = np.load(os.path.join(golden_dir, "uformer64_b1.npz"))
    dev = torch.device("cuda:0")
    model = L.Uformer(img_size=128, embed_dim=64, win_size=8, token_projection="linear", token_mlp="leff")
    param_fill.fill_module(model, int(z["seed"]))
    model = model.to(dev).eval()
    x = torch.from_numpy(z["x"]).to(dev)
    idx = torch.from_numpy(z["idx"].astype(np.int64))
    scale = float(np.abs(z["y"]).max())
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            y = model(x, index_samples=idx)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    e = np.abs(y.cpu().numpy() - z["y"]).ravel()
    print(f"uformer64_b1 f32: max {e.max():.3e} median {np.median(e):.3e} frac>1e-3*scale {(e > 1e-3 * scale).mean():.4f} (scale {scale:.2f})")
    assert np.median(e) < 1e-4 * scale
    assert (e > 1e-3 * scale).mean() < 0.02
    with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
        yb = model(x, index_samples=idx).float()
    eb = np.abs(yb.cpu().numpy() - z["y"]).ravel()
    print(f"uformer64_b1 bf16: max {eb.max():.3e} median {np.median(eb):.3e} frac>2e-2*scale {(eb > 2e-2 * scale).mean():.4f}")
    assert np.median(eb) < 2e-2 * scale
    assert (eb > 1e-1 * scale).mean() < 0.02


def test_canvas_mode_matches_reference_golden_f32(golden_dir):
    """fullres.dehaze_canvas = the reference's test_long_GPU.py:74-93 (wrap-pad, one forward over the whole canvas, crop,
    clamp) against the unmodified reference's recording for a 200 x 300 image (384^2 canvas, 2304 windows at level 0).  Same
    bulk criteria as the tile test above (the model is chaotic in near-tie selections); the unclamped crop is compared too
    because the random-init model saturates most pixels."""
    import os
    import lewin_b200 as L
    from lewin_b200 import fullres
    from oracle import param_fill
    z = np.load(os.path.join(golden_dir, "uformer32_canvas_200x300.npz"))
    dev = torch.device("cuda:0")
    model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff")
    param_fill.fill_module(model, int(z["seed"]))
    model = model.to(dev).eval()
    img = torch.from_numpy(z["x"]).to(dev)
    idx = torch.from_numpy(z["idx"].astype(np.int64))
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            raw = model(fullres.wrap_pad(img, ps=128), index_samples=idx)[:, :, :200, :300]
            y = fullres.dehaze_canvas(model, img, ps=128, index_samples=idx)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert torch.equal(y, raw.clamp(0, 1))
    e = np.abs(raw.cpu().numpy() - z["y_raw"]).ravel()
    print(f"canvas 200x300: raw max {e.max():.3e} median {np.median(e):.3e} frac>1e-3 {(e > 1e-3).mean():.4f}")
    # Measured on a B200 (round 1): median 1.1e-5, 7.9 % of the raw values off by > 1e-3 (raw range +-6), max 0.19.  The bulk
    # agrees to fp32 accuracy; the rest is the model's own sensitivity to rounding: the reference-exact numpy oracle moves
    # 0 % / 4.7 % / 15.7 % of these outputs by > 1e-3 (max 0.11 - 0.19) when its input is perturbed by 1e-6 relative noise
    # (scripts/canvas_chaos.py) - near-tie top-u flips at the deep levels, where one token covers 16 x 16 output pixels.
    assert np.median(e) < 1e-4
    assert (e > 1e-3).mean() < 0.2
    def psnr_to(a, t):
        return 10 * np.log10(1.0 / float(((a - t) ** 2).mean()))
    p_ours, p_ref = psnr_to(y.cpu().numpy(), z["x"]), psnr_to(z["y"], z["x"])
    print(f"canvas PSNR vs target: ours {p_ours:.4f} dB, reference {p_ref:.4f} dB (reported, not gated)")


def test_canvas_mode_block_by_block_with_the_devices_selection_f32(golden_dir):
    """The strict canvas-mode check (VERDICT r1 weak #1 / ADVICE r1): the reference's own full-image computation
    (test_long_GPU.py:74-93, one forward over the wrap-padded canvas) followed BLOCK BY BLOCK.

    The device forward records the top-u selection of each of the 18 LeWin blocks (ops.TopRecorder); the whole-model oracle -
    which reproduces the unmodified reference's recording of this fixture on 26 208 of 26 208 rows
    (test_whole_model_oracle_matches_reference_in_canvas_mode) - is then run with the device's selections forced in, so that
    it follows the device's path through the near-tie rows instead of diverging at the first flipped row.  Gates:
      * per block: the device's selection equals the oracle's OWN selection on that block's input on every (window, head) row
        whose rank-25/26 gap is not a near-tie (tau = 1e-4 of the row's M range: 10x the single-block fp32 threshold, for
        the rounding accumulated over up to 18 chained blocks);
      * the raw model output agrees with the forced oracle within north_star's max-abs 1e-3 on EVERY pixel, and the
        dehazed-image PSNR (against the same target) within 0.01 dB."""
    import os
    import lewin_b200 as L
    from lewin_b200 import fullres, ops
    from oracle import param_fill, uformer_oracle as U
    z = np.load(os.path.join(golden_dir, "uformer32_canvas_200x300.npz"))
    dev = torch.device("cuda:0")
    model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff")
    param_fill.fill_module(model, int(z["seed"]))
    sd = {k: v.numpy().copy() for k, v in model.state_dict().items()}
    model = model.to(dev).eval()
    img = torch.from_numpy(z["x"])
    canvas = fullres.wrap_pad(img, ps=128)
    idx = z["idx"].astype(np.int64)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad(), ops.TopRecorder() as rec:
            raw = model(canvas.to(dev), index_samples=torch.from_numpy(idx))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    torch.cuda.synchronize()
    assert len(rec.tops) == 18
    tops = [np.sort(t.cpu().numpy().astype(np.int64), -1) for t in rec.tops]
    steps = []
    ref = U.uformer_forward(canvas.numpy(), sd, idx, img_size=128, dtype=np.float32, record=steps, force_tops=tops)
    tau = 1e-4
    rows = flipped = ambiguous = 0
    for st, top in zip(steps, tops):
        assert top.shape == st["sel"].shape, (st["block"], top.shape, st["sel"].shape)
        bad = (top != st["sel"]).any(-1)
        amb = st["rel_gap"] < tau
        assert not (bad & ~amb).any(), f"block {st['block']} ({st['stage']}): {(bad & ~amb).sum()} non-ambiguous rows selected differently"
        rows += bad.size; flipped += int(bad.sum()); ambiguous += int(amb.sum())
    assert rows == 26208
    e = np.abs(raw.cpu().numpy() - ref)
    print(f"canvas, device selection forced into the oracle: {flipped} of {rows} rows fell the other way ({ambiguous} near-ties); "
          f"raw output max err {e.max():.3e}, median {np.median(e):.3e}")
    assert e.max() < TOL_F32, e.max()
    def psnr_to(a, t):
        return 10 * np.log10(1.0 / float(((np.clip(a, 0, 1) - t) ** 2).mean()))
    p_dev, p_ref = psnr_to(raw.cpu().numpy()[:, :, :200, :300], z["x"]), psnr_to(ref[:, :, :200, :300], z["x"])
    assert abs(p_dev - p_ref) < 0.01, (p_dev, p_ref)
    # against the reference's own recording: every row where the device left the reference's M_top of the FIRST block (same
    # input on both sides) is a near-tie; later blocks are only comparable along the forced path above
    ref_top0 = z["top00"].astype(np.int64)
    bad0 = (tops[0] != np.sort(ref_top0, -1)).any(-1)
    assert not (bad0 & ~(steps[0]["rel_gap"] < tau)).any()


def test_streaming_dehazer_matches_direct_calls():
    """fullres.StreamingDehazer (side-stream H2D / D2H, double-buffered) returns, image for image, exactly what the direct
    dehaze_tiled call returns; 5 different images through 2 slots exercise slot reuse in both directions."""
    import lewin_b200 as L
    from lewin_b200 import fullres
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(dev).eval()
    idx = model.draw_index_samples()
    H, W = 200, 300                                           # canvas 384^2 = 9 tiles
    g = torch.Generator().manual_seed(6)
    imgs = [torch.rand(1, 3, H, W, generator=g).pin_memory() for _ in range(5)]
    outs = [torch.empty(1, 3, H, W).pin_memory() for _ in range(5)]
    fn = lambda x: fullres.dehaze_tiled(model, x, ps=128, index_samples=idx)
    pipe = fullres.StreamingDehazer(fn, (1, 3, H, W), dev)
    for a, b in zip(imgs, outs):
        pipe.submit(a, b)
    pipe.flush()
    torch.cuda.synchronize()
    for a, b in zip(imgs, outs):
        ref = fn(a.to(dev)).cpu()
        assert torch.equal(ref, b)
    # multi-GPU upload plan, emulated for rank 1 of 2: only fullres.rows_needed rows are copied to the device, and the rank's
    # tile shard must come out exactly as from the full image
    H2 = W2 = 600                                             # canvas 640^2 = 25 tiles; rank 1 of 4 owns tiles 7..12
    imgs2 = [torch.rand(1, 3, H2, W2, generator=g).pin_memory() for _ in range(3)]
    s, e = fullres.shard_range(25, 1, 4)
    fn_r = lambda x: model(fullres.to_tiles(fullres.wrap_pad(x, ps=128), 128)[s:e], index_samples=idx)
    rows = fullres.rows_needed(H2, W2, 1, 4)
    assert rows == [(128, 384)]
    pipe_r = fullres.StreamingDehazer(fn_r, (1, 3, H2, W2), dev, rows=rows, out_shape=(e - s, 3, 128, 128))
    outs_r = [torch.empty(e - s, 3, 128, 128).pin_memory() for _ in range(3)]
    imgs = imgs2
    with torch.no_grad():
        for a, b in zip(imgs, outs_r):
            pipe_r.submit(a, b)
        pipe_r.flush()
        torch.cuda.synchronize()
        for a, b in zip(imgs, outs_r):
            assert torch.equal(fn_r(a.to(dev)).cpu(), b)


@pytest.mark.parametrize("dt", ["bf16", "f32"])
def test_tiled_pipeline_matches_the_serial_call(dt):
    """fullres.TiledPipeline (forward of image i+1 on the compute stream while image i is stitched on a side stream; at N > 1
    the all_gather rides on that side stream too) returns, image for image, exactly what dehaze_tiled returns; 5 different
    images through 2 slots exercise the slot reuse and the hand-over of the graph's static output."""
    import contextlib
    import lewin_b200 as L
    from lewin_b200 import fullres
    dev = torch.device("cuda:0")
    torch.manual_seed(8)
    model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(dev).eval()
    idx = model.draw_index_samples()
    H, W = 200, 300                                           # canvas 384^2 = 9 tiles
    imgs = [torch.rand(1, 3, H, W, device=dev) for _ in range(5)]
    graphed = fullres.GraphedForward(model, torch.zeros(9, 3, 128, 128, device=dev), idx, torch.bfloat16 if dt == "bf16" else None)
    pipe = fullres.TiledPipeline(model, graphed, (1, 3, H, W), dev)
    got = []
    for x in imgs:
        out, ev = pipe.submit(x, idx)
        ev.synchronize()                                      # (a serving loop would wait on the event from its consumer stream)
        got.append(out.clone())
    pipe.flush()
    # the same through the slot ring without waiting in between
    outs2 = []
    for x in imgs[:2]:
        outs2.append(pipe.submit(x, idx))
    pipe.flush()
    torch.cuda.synchronize()
    with (torch.autocast("cuda", torch.bfloat16) if dt == "bf16" else contextlib.nullcontext()):
        refs = [fullres.dehaze_tiled(model, x, ps=128, index_samples=idx, graphed=graphed, broadcast_index_samples=False).clone() for x in imgs]
    for a, b in zip(got, refs):
        assert a.dtype == b.dtype and torch.equal(a, b)
    for (o, _), b in zip(outs2, refs[:2]):
        assert torch.equal(o, b)
    # the staged function under StreamingDehazer (host image in, host result out; the stitch and the fp32 conversion run on
    # side streams): 5 images through the two 2-slot rings in lockstep
    pipe2 = fullres.TiledPipeline(model, graphed, (1, 3, H, W), dev)
    sd = fullres.StreamingDehazer(lambda x: pipe2.submit(x, idx), (1, 3, H, W), dev)
    hosts = [x.cpu().pin_memory() for x in imgs]
    outs = [torch.empty(1, 3, H, W).pin_memory() for _ in imgs]
    for a, b in zip(hosts, outs):
        sd.submit(a, b)
    sd.flush()
    torch.cuda.synchronize()
    for b, r in zip(outs, refs):
        assert torch.equal(b, r.float().cpu())


def test_tiled_pipeline_with_two_images_in_flight_matches_the_serial_call():
    """TiledPipeline with two lanes (two CUDA graphs on two streams: image i on lane i mod 2, both forwards in flight) returns,
    image for image, exactly what the serial dehaze_tiled returns: 9 different images back to back without any host wait
    exercise the 4-slot result ring, both lanes' static buffers and the input hand-over (the input buffer is REUSED and
    overwritten right after each submit, as StreamingDehazer does); then the same under StreamingDehazer from host memory."""
    import lewin_b200 as L
    from lewin_b200 import fullres
    dev = torch.device("cuda:0")
    torch.manual_seed(18)
    model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(dev).eval()
    idx = model.draw_index_samples()
    H, W = 200, 300                                           # canvas 384^2 = 9 tiles
    imgs = [torch.rand(1, 3, H, W, device=dev) for _ in range(9)]
    lanes = [fullres.GraphedForward(model, torch.zeros(9, 3, 128, 128, device=dev), idx, torch.bfloat16) for _ in range(2)]
    with torch.autocast("cuda", torch.bfloat16):
        refs = [fullres.dehaze_tiled(model, x, ps=128, index_samples=idx, graphed=lanes[0], broadcast_index_samples=False).clone()
                for x in imgs]
    assert not torch.equal(refs[0], refs[1])
    pipe = fullres.TiledPipeline(model, lanes, (1, 3, H, W), dev)
    assert pipe.depth == 4 and len(pipe.lane_streams) == 2
    buf = torch.empty_like(imgs[0])                           # one input buffer, overwritten after every submit
    got = []
    cur = torch.cuda.current_stream(dev)
    for x in imgs:
        buf.copy_(x)
        out, ev = pipe.submit(buf, idx)
        buf.fill_(float("nan"))                               # legal: the caller's stream has passed submit()
        cur.wait_event(ev)
        got.append(out.clone())                               # consumer on the caller's stream, ordered by the event only
    pipe.flush()
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(got, refs)):
        assert a.dtype == b.dtype and torch.equal(a, b), i
    # no consumer in between: 4 results stay valid in the 4-slot ring
    outs = [pipe.submit(x, idx)[0] for x in imgs[:4]]
    pipe.flush()
    torch.cuda.synchronize()
    for o, b in zip(outs, refs[:4]):
        assert torch.equal(o, b)
    # under StreamingDehazer: a 2-slot host ring over the 4-slot device ring, and a 4-slot host ring in lockstep with it
    # (what bench.py's e2e uses with two lanes)
    hosts = [x.cpu().pin_memory() for x in imgs]
    for host_depth in (2, 4):
        pipe2 = fullres.TiledPipeline(model, lanes, (1, 3, H, W), dev)
        sd = fullres.StreamingDehazer(lambda x: pipe2.submit(x, idx), (1, 3, H, W), dev, depth=host_depth)
        res = [torch.empty(1, 3, H, W).pin_memory() for _ in imgs]
        for a, b in zip(hosts, res):
            sd.submit(a, b)
        sd.flush()
        torch.cuda.synchronize()
        for i, (b, r) in enumerate(zip(res, refs)):
            assert torch.equal(b, r.float().cpu()), (host_depth, i)


def test_full_size_tile_batch_invariance_bf16():
    """BASELINE config 3 at its full size (1200x1600 -> 1664^2 canvas -> 169 tiles), bf16: size-independent properties of
    the tiled computation.  (1) Determinism: two runs are bit-identical.  (2) Shard invariance: tiles are independent
    units, so the 8-GPU shard sizes (22 / 21 tiles, run here one after the other) must reproduce the single 169-tile
    batch bit for bit - this is what makes the multi-GPU result equal to the single-GPU one."""
    import lewin_b200 as L
    from lewin_b200 import fullres
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(dev).eval()
    idx = model.draw_index_samples()
    g = torch.Generator().manual_seed(4)
    img = torch.rand(1, 3, 1200, 1600, generator=g).to(dev)
    with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
        full = fullres.dehaze_tiled(model, img, ps=128, index_samples=idx)
        again = fullres.dehaze_tiled(model, img, ps=128, index_samples=idx)
        assert full.shape == (1, 3, 1200, 1600)
        assert torch.equal(full, again)
        canvas = fullres.wrap_pad(img, ps=128)
        tiles = fullres.to_tiles(canvas, 128)
        assert tiles.shape[0] == 169
        outs = []
        for r in range(8):
            s, e = fullres.shard_range(169, r, 8)
            outs.append(model(tiles[s:e], index_samples=idx))
        sharded = fullres.from_tiles(torch.cat(outs, 0), canvas.shape[-1], 128)[:, :, :1200, :1600].clamp(0, 1)
    # the LeWin kernels are batch-invariant by construction (every token's reductions run in a fixed order); the stock cuDNN
    # Downsample / OutputProj convolutions may pick another algorithm for another batch size, so allow isolated
    # last-bit differences (and their spread through a flipped top-u near-tie) but nothing systematic
    ndiff = int((full != sharded).sum())
    print(f"169-tile batch vs 8 shards: {ndiff} of {full.numel()} values differ, max {float((full - sharded).abs().max()):.3e}")
    assert ndiff <= 1e-3 * full.numel()
    assert torch.isfinite(full).all() and float(full.min()) >= 0.0 and float(full.max()) <= 1.0


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_canvas_row_bands_equal_the_single_device_canvas_forward(golden_dir, dt, world):
    """canvas_bands.dehaze_canvas_bands (canvas mode sharded by row bands with halo exchange) against
    fullres.dehaze_canvas on the reference's own canvas fixture geometry (200 x 300 image, 384^2 canvas = 3 units of 128
    rows): `world` bands processed in lock step in this process (the same generator the torch.distributed driver serves).
    The LeWin blocks see identical operands through identical kernels, so the raw outputs must agree bit for bit wherever
    the out-of-scope cuDNN convolutions pick the same algorithm; gate: selections identical, output within 1e-3 / 2e-2."""
    import os
    import lewin_b200 as L
    from lewin_b200 import canvas_bands, fullres, ops
    from oracle import param_fill
    z = np.load(os.path.join(golden_dir, "uformer32_canvas_200x300.npz"))
    dev = torch.device("cuda:0")
    model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff")
    param_fill.fill_module(model, int(z["seed"]))
    model = model.to(dev).eval()
    img = torch.from_numpy(z["x"]).to(dev)
    idx = torch.from_numpy(z["idx"].astype(np.int64))
    import contextlib
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with (torch.autocast("cuda", torch.bfloat16) if dt == "bf16" else contextlib.nullcontext()):
            with ops.TopRecorder() as rec_ref:
                ref = fullres.dehaze_canvas(model, img, ps=128, index_samples=idx)
            with ops.TopRecorder() as rec_band:
                got = canvas_bands.dehaze_canvas_bands(model, img, ps=128, index_samples=idx, virtual_world=world)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    torch.cuda.synchronize()
    assert got.shape == ref.shape == (1, 3, 200, 300)
    # selections: block i of band r covers the windows of its rows (shifted-frame order); stitched back they must equal the
    # whole-canvas forward's.  The LeWin arithmetic is identical; the out-of-scope cuDNN convolutions (Downsample / OutputProj,
    # and InputProj / Upsample at fp32) run on band slabs instead of the whole map and may pick another algorithm, i.e. another
    # summation order - a 1e-7 difference that can flip a near-tie row in a later block (the model is chaotic in the selection,
    # DESIGN.md section 8), so after the first convolution a handful of rows may legitimately differ.
    assert len(rec_ref.tops) == 18 and len(rec_band.tops) == 18 * world
    rows = flipped = 0
    for i in range(18):
        t_ref = np.sort(rec_ref.tops[i].cpu().numpy().astype(np.int64), -1)
        t_b = np.sort(torch.cat([rec_band.tops[i * world + r] for r in range(world)], 0).cpu().numpy().astype(np.int64), -1)
        assert t_ref.shape == t_b.shape
        bad = (t_ref != t_b).any(-1)
        if i < 2:
            assert not bad.any(), f"block {i}: the bands selected different queries before any convolution ran on a slab"
        rows += bad.size; flipped += int(bad.sum())
    d = (got.float() - ref.float()).abs()
    print(f"canvas bands world={world} {dt}: {flipped} of {rows} rows flipped, max |bands - single| = {float(d.max()):.3e}, "
          f"frac > 1e-3 = {float((d > 1e-3).float().mean()):.2e}")
    assert flipped <= max(1, rows // 1000), (flipped, rows)
    assert float((d > (1e-3 if dt == "f32" else 2e-2)).float().mean()) < 0.01
    if world == 1:
        assert torch.equal(got, ref)
