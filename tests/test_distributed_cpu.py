"""CPU, world_size 2, gloo: the host-side multi-GPU logic (tile sharding + final gather of config 3, index_sample
broadcast, DDP with the dead parameters frozen).  The CUDA ops have no CPU fallback, so a stand-in per-tile function
is used where a model forward is needed; what is under test is the sharding / collective plumbing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _TileModel(nn.Module):
    """Per-tile stand-in: depends on the tile content AND on index_samples (so a missing broadcast is caught)."""

    def draw_index_samples(self):
        return torch.stack([torch.randint(64, (64, 25)) for _ in range(18)])

    def forward(self, x, index_samples=None):
        k = index_samples.float().mean() / 64.0
        return x.flip(-1) * 0.5 + k + x.mean(dim=(1, 2, 3), keepdim=True)


def _tiled_worker(rank, world, port, ref_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lewin_b200 import fullres
    torch.manual_seed(7)
    img = torch.rand(1, 3, 250, 380)
    torch.manual_seed(100 + rank)                 # ranks would draw DIFFERENT index_samples without the broadcast
    out = fullres.dehaze_tiled(_TileModel(), img, ps=128)
    ref = torch.load(ref_path)
    assert torch.equal(out, ref), f"rank {rank}: sharded result differs from the single-process result"
    # the caller-guaranteed form (bench.py): identical draws on every rank from a shared seed, no per-image broadcast;
    # and the upload plan: NaN outside fullres.rows_needed must not reach this rank's tiles (rank 0's result is the image)
    torch.manual_seed(100)
    idx = _TileModel().draw_index_samples()
    torch.manual_seed(8)
    img2 = torch.rand(1, 3, 600, 380)                        # 640^2 canvas, 25 tiles: 13 / 12 per rank
    rows = fullres.rows_needed(600, 380, rank, world)
    assert rows == ([(0, 384)] if rank == 0 else [(0, 40), (256, 600)])
    part = torch.full_like(img2, float("nan"))
    for r0, r1 in rows:
        part[:, :, r0:r1] = img2[:, :, r0:r1]
    out2 = fullres.dehaze_tiled(_TileModel(), part, ps=128, index_samples=idx, broadcast_index_samples=False)
    canvas = fullres.wrap_pad(img2, ps=128)
    ref2 = fullres.from_tiles(_TileModel()(fullres.to_tiles(canvas, 128), index_samples=idx), 640, 128)[:, :, :600, :380].clamp(0, 1)
    assert torch.equal(out2, ref2), f"rank {rank}: partial-upload / no-broadcast result differs"
    dist.destroy_process_group()


def test_tiled_sharding_matches_single_process(tmp_path):
    from lewin_b200 import fullres
    torch.manual_seed(7)
    img = torch.rand(1, 3, 250, 380)
    torch.manual_seed(100)                        # == rank 0's generator state in the workers
    ref = fullres.dehaze_tiled(_TileModel(), img, ps=128)
    assert ref.shape == img.shape
    path = str(tmp_path / "ref.pt")
    torch.save(ref, path)
    mp.spawn(_tiled_worker, args=(2, _free_port(), path), nprocs=2, join=True)


class _BlockLike(nn.Module):
    """Module tree with the reference's dead-parameter layout (attn.qkv.to_q / to_kv / attn.proj never used)."""

    def __init__(self):
        super().__init__()
        self.attn = nn.Module()
        self.attn.qkv = nn.Module()
        self.attn.qkv.to_q = nn.Linear(8, 8)
        self.attn.qkv.to_kv = nn.Linear(8, 16)
        self.attn.proj = nn.Linear(8, 8)
        self.live = nn.Linear(8, 8)

    def forward(self, x):
        return self.live(x)


def _ddp_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lewin_b200 import parallel
    torch.manual_seed(0)
    model = _BlockLike()
    ddp = parallel.wrap_ddp(model)               # would hang / raise on the unused parameters without the freeze
    torch.manual_seed(0)
    x = torch.randn(4, 8)
    ddp(x[rank * 2:(rank + 1) * 2]).sum().backward()
    torch.manual_seed(0)
    single = _BlockLike()
    single(x).sum().backward()
    assert torch.allclose(model.live.weight.grad * world, single.live.weight.grad, atol=1e-6)
    assert model.attn.proj.weight.grad is None and not model.attn.proj.weight.requires_grad
    assert "attn.proj.weight" in model.state_dict()
    dist.destroy_process_group()


def test_ddp_with_dead_parameters_gloo():
    mp.spawn(_ddp_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_dead_parameter_inventory():
    import lewin_b200 as L
    from lewin_b200 import parallel
    blk = L.LeWinTransformerBlock(dim=32, input_resolution=(128, 128), num_heads=1)
    dead = parallel.dead_parameter_names(blk)
    assert len(dead) == 6
    assert parallel.freeze_dead_parameters(blk) == 6
    assert sum(p.requires_grad for p in blk.parameters()) == 19      # the 19 live parameters of SURVEY Appendix B


# ---------------------------------------------------------------- canvas row bands: the halo-exchange driver
def _band_stand_in(x, rank, world):
    """A generator with the exchange pattern of canvas_bands._block: a cyclic 'take the next band's first rows / hand the
    last rows back' pair followed by a non-cyclic one-row halo on each side.  Returns what a rank can only know if every
    message arrived from the right neighbour."""
    from lewin_b200.canvas_bands import _Xchg
    s = 2
    _, from_down = yield _Xchg(to_up=x[:s].contiguous(), want_down=s, cyclic=True)
    local = torch.cat([x[s:], from_down], 0)
    from_up, _ = yield _Xchg(to_down=local[-s:].contiguous() * 2, want_up=s, cyclic=True)
    y = torch.cat([from_up, local[:-s]], 0)
    top, bot = rank > 0, rank < world - 1
    fu, fd = yield _Xchg(to_up=y[:1].contiguous() if top else None, to_down=y[-1:].contiguous() if bot else None,
                         want_up=1 if top else 0, want_down=1 if bot else 0)
    z = x.new_zeros((1,) + tuple(x.shape[1:]))
    return torch.cat([fu if top else z, y, fd if bot else z], 0)


def _bands_worker(rank, world, port, ref_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lewin_b200 import canvas_bands
    full = torch.arange(world * 6 * 5 * 3, dtype=torch.float32).view(world * 6, 5, 3)
    mine = full[rank * 6:(rank + 1) * 6]
    out = canvas_bands._serve_dist(_band_stand_in(mine, rank, world), rank, world, None, torch.device("cpu"), None)
    ref = torch.load(ref_path)[rank]
    assert torch.equal(out, ref), f"rank {rank}: halo exchange over torch.distributed differs from the in-process routing"
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_canvas_band_halo_exchange_gloo_matches_in_process_routing(tmp_path, world):
    """canvas_bands: the torch.distributed driver (_serve_dist: batched isend / irecv with the ring neighbours) delivers the
    same rows as the in-process lock-step driver (_serve_virtual) that the GPU parity test runs the whole model through."""
    from lewin_b200 import canvas_bands
    full = torch.arange(world * 6 * 5 * 3, dtype=torch.float32).view(world * 6, 5, 3)
    ref = canvas_bands._serve_virtual([_band_stand_in(full[r * 6:(r + 1) * 6], r, world) for r in range(world)])
    # ring semantics: band 0's first two rows came back from the LAST band (its handed-back rows, doubled by the stand-in),
    # and those are band 0's own first rows that travelled up the ring: the roll closes on itself
    assert torch.equal(ref[0][1:3], 2 * full[0:2])
    assert canvas_bands.band_units(13, 0, 8) == (0, 2) and canvas_bands.band_units(13, 7, 8) == (12, 13)
    assert [canvas_bands.band_units(13, r, 4) for r in range(4)] == [(0, 4), (4, 7), (7, 10), (10, 13)]
    path = str(tmp_path / "ref.pt")
    torch.save(ref, path)
    mp.spawn(_bands_worker, args=(world, _free_port(), path), nprocs=world, join=True)
