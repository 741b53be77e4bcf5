"""Canvas mode sharded over the GPUs of one box by ROW BANDS.

``test_long_GPU.py:85-92`` runs ONE forward over the 1664^2 wrap-padded canvas; ``fullres.dehaze_canvas`` is that computation
on one GPU.  Here the canvas is cut into contiguous bands of whole 128-row units (13 units for 1664 rows: the granularity at
which every U-Net level still holds whole 8-row windows, 128 / 2^4 = 8 rows at the bottleneck); rank r keeps its band of the
residual stream at every level and exchanges only the rows a neighbour's arithmetic reaches into:

  * shifted-window attention (``torch.roll(x, (-4, -4))``, My_model_1.py:846): windows of the shifted frame straddle the band
    edge, so the band takes the first 4 rows of the NEXT band in (cyclically: the last band wraps to the first, exactly the
    roll), runs the attention half on rows [a+4, b+4) in shifted-frame order (``band`` mode of lewin_attn: column shift only,
    analytic shift mask evaluated at the global row), and hands the 4 result rows that belong to the next band back;
  * LeFF's depthwise 3x3 (My_model_1.py:489-491, zero padded at the IMAGE border): one row of the attention output from each
    neighbour; LN2 + linear1 + GELU are recomputed on those two rows, the halo results are dropped;
  * Downsample (4x4, stride 2, pad 1) and OutputProj (3x3): one row from each neighbour; InputProj reads the canvas itself;
    Upsample (2x2, stride 2) and the skip concatenations are local.

Every (window, head) item and every token sees the operands it sees in the single-GPU canvas forward and goes through the
same kernels, so the LeWin arithmetic is identical; the out-of-scope convolutions run on the band slabs (same cuDNN
arithmetic per output pixel).  Messages are small (4 rows x W x C) and latency bound: 18 blocks x (2 or 4) + 5 convolutions.

The forward is written once as a generator that yields its halo exchanges; ``dehaze_canvas_bands`` drives it either with
torch.distributed point-to-point operations (one process per GPU, NCCL) or, with ``virtual_world=N`` in a single process,
for all N bands in lock step on one device (copies instead of messages) - the same code path, used by the parity tests.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops
from .fullres import canvas_size, wrap_pad
from .modules import _act_dtype, _use_rpb

UNIT = 128          # band granularity in canvas rows (= 8 rows at the bottleneck level)


def band_units(n_units, rank, world):
    """Contiguous unit range [u0, u1) of `rank` (the first n_units % world ranks hold one unit more)."""
    base, extra = divmod(n_units, world)
    u0 = rank * base + min(rank, extra)
    return u0, u0 + base + (1 if rank < extra else 0)


class _Xchg:
    """One halo exchange: rows for the upper / lower neighbour and the row counts expected back."""
    __slots__ = ("to_up", "to_down", "want_up", "want_down", "cyclic")

    def __init__(self, to_up=None, to_down=None, want_up=0, want_down=0, cyclic=False):
        self.to_up, self.to_down, self.want_up, self.want_down, self.cyclic = to_up, to_down, want_up, want_down, cyclic


def _attn_kwargs(blk):
    ps = blk.attn.ProbSpare
    w_qkv, b_qkv = ps.qkv_weights() if hasattr(ps, "qkv_weights") else (
        torch.cat([ps.query_projection.weight, ps.key_projection.weight, ps.value_projection.weight], 0),
        torch.cat([ps.query_projection.bias, ps.key_projection.bias, ps.value_projection.bias], 0))
    return dict(num_heads=blk.num_heads, ln_w=blk.norm1.weight, ln_b=blk.norm1.bias, w_qkv=w_qkv, b_qkv=b_qkv,
                w_out=ps.out_projection.weight, b_out=ps.out_projection.bias,
                rpb_table=blk.attn.relative_position_bias_table, use_rpb=_use_rpb())


def _block(blk, x, Hb, W, idx, a, Hg, rank, world):
    """One LeWin block (My_model_1.py:785-875) on the band rows [a, a + Hb) of an Hg-row map; x [Hb, W, C]."""
    C = x.shape[-1]
    s = blk.shift_size
    kw = _attn_kwargs(blk)
    if s == 0:
        y = ops.lewin_attn(x.reshape(1, Hb * W, C), B=1, H=Hb, W=W, shift=0, index_sample=idx, **kw).view(Hb, W, C)
    else:
        # rows [a + s, b + s) in shifted-frame order: mine from s on, then the next band's first s rows (cyclic == the roll)
        _, from_down = yield _Xchg(to_up=x[:s].contiguous(), want_down=s, cyclic=True)
        local = torch.cat([x[s:], from_down], 0)
        yl = ops.lewin_attn(local.reshape(1, Hb * W, C), B=1, H=Hb, W=W, shift=s, index_sample=idx, band=(a, Hg), **kw).view(Hb, W, C)
        from_up, _ = yield _Xchg(to_down=yl[Hb - s:].contiguous(), want_up=s, cyclic=True)
        y = torch.cat([from_up, yl[:Hb - s]], 0)
    # LeFF: the depthwise 3x3 reaches one row into each neighbour (zero padding only at the image border)
    top, bot = rank > 0, rank < world - 1
    from_up, from_down = yield _Xchg(to_up=y[:1].contiguous() if top else None, to_down=y[Hb - 1:].contiguous() if bot else None,
                                     want_up=1 if top else 0, want_down=1 if bot else 0)
    parts = ([from_up] if top else []) + [y] + ([from_down] if bot else [])
    slab = torch.cat(parts, 0) if len(parts) > 1 else y
    Hl = slab.shape[0]
    mlp = blk.mlp
    out = ops.lewin_leff(slab.reshape(1, Hl * W, C), B=1, H=Hl, W=W, ln_w=blk.norm2.weight, ln_b=blk.norm2.bias,
                         w1=mlp.linear1[0].weight, b1=mlp.linear1[0].bias, w_dw=mlp.dwconv[0].weight, b_dw=mlp.dwconv[0].bias,
                         w2=mlp.linear2[0].weight, b2=mlp.linear2[0].bias, fused=True).view(Hl, W, C)
    o0 = 1 if top else 0
    return out[o0:o0 + Hb]


def _halo_slab(x, rank, world):
    """[1 row from above | x | 1 row from below] with zero rows at the image border (the convolutions' zero padding)."""
    Hb = x.shape[0]
    top, bot = rank > 0, rank < world - 1
    from_up, from_down = yield _Xchg(to_up=x[:1].contiguous() if top else None, to_down=x[Hb - 1:].contiguous() if bot else None,
                                     want_up=1 if top else 0, want_down=1 if bot else 0)
    z = x.new_zeros((1,) + tuple(x.shape[1:]))
    return torch.cat([from_up if top else z, x, from_down if bot else z], 0)


def _nchw(t):
    """[H, W, C] token map -> [1, C, H, W] channels-last VIEW (no copy)."""
    return t.permute(2, 0, 1).unsqueeze(0)


def _band_forward(model, canvas, idx, rank, world):
    """Generator: the whole Uformer forward (My_model_1.py:1169-1207) for this rank's row band of `canvas` [1, 3, L, L].
    Yields _Xchg requests, receives (from_up, from_down); returns the band [3, rows, L] of x + output_proj(...)."""
    L = canvas.shape[-1]
    n_units = L // UNIT
    u0, u1 = band_units(n_units, rank, world)
    a0, b0 = u0 * UNIT, u1 * UNIT
    dev = canvas.device
    idx = idx.to(device=dev, dtype=torch.int32)
    d = model.depths
    offs = [sum(d[:i]) for i in range(len(d) + 1)]

    def stage(layer, li, x, lvl):
        Hb, W, a, Hg = (b0 - a0) >> lvl, L >> lvl, a0 >> lvl, L >> lvl
        for i, blk in enumerate(layer.blocks):
            x = yield from _block(blk, x, Hb, W, idx[offs[li] + i], a, Hg, rank, world)
        return x

    # InputProj (3x3 conv + LeakyReLU, My_model_1.py:659-682) on the canvas rows [a0 - 1, b0 + 1)
    r0, r1 = max(a0 - 1, 0), min(b0 + 1, L)
    tok = model.input_proj(canvas[:, :, r0:r1, :].contiguous())                 # [1, rows * L, C]
    C = tok.shape[-1]
    x = tok.view(r1 - r0, L, C)[a0 - r0:a0 - r0 + (b0 - a0)]
    x = x.to(_act_dtype(x))
    skips = []
    for lvl in range(4):
        x = yield from stage(getattr(model, f"encoderlayer_{lvl}"), lvl, x, lvl)
        skips.append(x)
        slab = yield from _halo_slab(x, rank, world)
        Hl, Wl, Cl = slab.shape
        y = getattr(model, f"dowsample_{lvl}")(slab.reshape(1, Hl * Wl, Cl), hw=(Hl, Wl), pad_h=False)
        x = y.reshape(Hl // 2 - 1, Wl // 2, 2 * Cl)                                # [Hb / 2, W / 2, 2C]
    x = yield from stage(model.conv, 4, x, 4)
    for lvl in range(4):
        up = getattr(model, f"upsample_{lvl}")
        Hb, W, Cin = x.shape
        skip = skips[3 - lvl]
        cat = up(x.reshape(1, Hb * W, Cin), skip.reshape(1, -1, skip.shape[-1]), hw=(Hb, W))     # cat([up, skip], -1)
        x = cat.view(2 * Hb, 2 * W, -1)
        x = yield from stage(getattr(model, f"decoderlayer_{lvl}"), 5 + lvl, x, 3 - lvl)
    slab = yield from _halo_slab(x, rank, world)
    Hl, Wl, Cl = slab.shape
    y = model.output_proj(slab.reshape(1, Hl * Wl, Cl), residual=canvas[:, :, a0:b0, :].contiguous(), hw=(Hl, Wl), pad_h=False)
    return y[0]                                                                  # [3, rows, L] = canvas band + projection


# ---------------------------------------------------------------------------------------------- drivers
def _serve_virtual(gens):
    """All bands in one process, in lock step: route every exchange by copying (parity tests / single-GPU check)."""
    n = len(gens)
    reqs = [next(g) for g in gens]
    outs = [None] * n
    while True:
        replies = []
        for r in range(n):
            q = reqs[r]
            up, down = (r - 1) % n, (r + 1) % n
            fu = fd = None
            if q.want_up and (q.cyclic or r > 0):
                fu = reqs[up].to_down
            if q.want_down and (q.cyclic or r < n - 1):
                fd = reqs[down].to_up
            replies.append((fu, fd))
        nxt, done = [], 0
        for r, g in enumerate(gens):
            try:
                nxt.append(g.send(replies[r]))
            except StopIteration as e:
                outs[r] = e.value
                done += 1
                nxt.append(None)
        if done:
            assert done == n, "bands must make the same sequence of exchanges"
            return outs
        reqs = nxt


def _serve_dist(gen, rank, world, group, dev, dtype_hint):
    import torch.distributed as dist
    try:
        q = next(gen)
        while True:
            up, down = (rank - 1) % world, (rank + 1) % world
            opsl, fu, fd = [], None, None
            has_up, has_down = (q.cyclic or rank > 0), (q.cyclic or rank < world - 1)
            if world == 1:                                   # the ring closes on myself
                fu, fd = (q.to_down if q.want_up and has_up else None), (q.to_up if q.want_down and has_down else None)
            else:
                ref = q.to_up if q.to_up is not None else q.to_down
                if q.to_up is not None and has_up:
                    opsl.append(dist.P2POp(dist.isend, q.to_up, dist.get_global_rank(group, up) if group is not None else up, group))
                if q.to_down is not None and has_down:
                    opsl.append(dist.P2POp(dist.isend, q.to_down, dist.get_global_rank(group, down) if group is not None else down, group))
                if q.want_up and has_up:
                    fu = torch.empty((q.want_up,) + tuple(ref.shape[1:]), dtype=ref.dtype, device=ref.device)
                    opsl.append(dist.P2POp(dist.irecv, fu, dist.get_global_rank(group, up) if group is not None else up, group))
                if q.want_down and has_down:
                    fd = torch.empty((q.want_down,) + tuple(ref.shape[1:]), dtype=ref.dtype, device=ref.device)
                    opsl.append(dist.P2POp(dist.irecv, fd, dist.get_global_rank(group, down) if group is not None else down, group))
                if opsl:
                    for w in dist.batch_isend_irecv(opsl):
                        w.wait()
            q = gen.send((fu, fd))
    except StopIteration as e:
        return e.value


@torch.no_grad()
def dehaze_canvas_bands(model, img, ps=128, index_samples=None, group=None, virtual_world=None, _broadcast=True):
    """Canvas mode (test_long_GPU.py:74-93: wrap-pad, ONE forward over the canvas, crop, clamp) with the canvas split into
    row bands over the ranks of `group` (torch.distributed; every rank passes the full image and ends with the full result),
    or - ``virtual_world=N`` - over N bands processed in lock step inside this process.  Identical to fullres.dehaze_canvas."""
    import torch.distributed as dist
    B, C, H, W = img.shape
    assert B == 1
    canvas = wrap_pad(img, ps=ps)
    L = canvas.shape[-1]
    assert L % UNIT == 0 and ps == UNIT
    if index_samples is None:
        index_samples = model.draw_index_samples()
    if virtual_world is not None:
        world = int(virtual_world)
        assert 1 <= world <= L // UNIT
        bands = _serve_virtual([_band_forward(model, canvas, index_samples, r, world) for r in range(world)])
        full = torch.cat(bands, 1)
    else:
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        rank = dist.get_rank(group) if distributed else 0
        world = dist.get_world_size(group) if distributed else 1
        assert world <= L // UNIT, "more ranks than 128-row units"
        if distributed and _broadcast:          # one set of key-sample draws for the whole image, as the reference's single call
            idx = index_samples.to(img.device)
            dist.broadcast(idx, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            index_samples = idx
        band = _serve_dist(_band_forward(model, canvas, index_samples, rank, world), rank, world, group, img.device, None)
        if distributed:
            per = -(-(L // UNIT) // world) * UNIT
            padded = band.new_zeros((C, per, L))
            padded[:, :band.shape[1]] = band
            gathered = band.new_empty((world, C, per, L))
            dist.all_gather_into_tensor(gathered, padded, group=group)
            parts = []
            for r in range(world):
                u0, u1 = band_units(L // UNIT, r, world)
                parts.append(gathered[r, :, :(u1 - u0) * UNIT])
            full = torch.cat(parts, 1)
        else:
            full = band
    return full[None, :, :H, :W].clamp(0, 1)


class GraphedCanvasBands:
    """CUDA-graph replay of ``dehaze_canvas_bands`` for one fixed image shape on this rank: the band forward is ~300 kernel
    launches plus ~100 small point-to-point messages, and at 2+ GPUs the per-rank GPU time drops below the time the host
    needs to issue them.  The NCCL sends / receives are captured into the graph together with the kernels (every rank
    captures the same sequence); the image and the 18 key-sample draws are copied into static buffers before each replay.
    Falls back to eager calls if the capture is refused."""

    def __init__(self, model, img, index_samples, autocast_dtype=None, group=None, warmup=2):
        from . import ops
        self.model, self.group, self.autocast_dtype = model, group, autocast_dtype
        self.img = img.clone()
        self.idx = index_samples.to(device=img.device, dtype=torch.int64).clone()
        self.graph, self.out = None, None
        try:
            with ops.weight_images.pin() as held:
                side = torch.cuda.Stream(device=img.device)
                side.wait_stream(torch.cuda.current_stream(img.device))
                with torch.cuda.stream(side):
                    for _ in range(warmup):
                        self._run()
                torch.cuda.current_stream(img.device).wait_stream(side)
                torch.cuda.synchronize(img.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.out = self._run()
                self.graph = g
            self._images = list(held)
        except Exception as e:              # capture is an optimisation
            self.graph, self.error = None, repr(e)[:200]
            torch.cuda.synchronize(img.device)

    def _run(self):
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", self.autocast_dtype):
                return dehaze_canvas_bands(self.model, self.img, index_samples=self.idx, group=self.group, _broadcast=False)
        return dehaze_canvas_bands(self.model, self.img, index_samples=self.idx, group=self.group, _broadcast=False)

    @torch.no_grad()
    def __call__(self, img, index_samples):
        self.img.copy_(img, non_blocking=True)
        self.idx.copy_(index_samples.to(dtype=torch.int64), non_blocking=True)
        if self.graph is None:
            return self._run()
        self.graph.replay()
        return self.out
