"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

BLOCK_FIXTURES = ["block_c32_h1_s0", "block_c32_h1_s4", "block_c64_h2_s4", "block_c64_h2_s4_inmask",
                  "block_c64_h2_s4_droppath", "block_c128_h4_s0"]

# deep levels, stored compactly (oracle/make_golden.py COMPACT_CASES): parameters regenerated from the seed, parameter
# gradients as L2 norm + a seeded sample of <= 4096 elements
COMPACT_FIXTURES = ["block_c256_h8_s4_compact", "block_c512_h16_s0_compact",
                    # head_dim = embed_dim 64 / 128 (My_model_1.py:962; BASELINE config 5), oracle/make_golden.py HEAD_DIM_CASES
                    "block_c64_h1_s4_hd64_compact", "block_c128_h2_s0_hd64_compact", "block_c128_h1_s4_hd128_compact"]

# the UNMODIFIED reference under torch.autocast("cpu", bfloat16) (oracle/make_golden.py BF16_CASES): x, out, idx, top
BF16_REF_FIXTURES = ["block_c64_h2_s4_bf16cpu", "block_c256_h8_s4_bf16cpu"]

TOL_F32 = 1e-3     # BASELINE.json north_star: max-abs 1e-3 (fp32)
TOL_BF16 = 2e-2    # max-abs 2e-2 (bf16)
TIE_TAU_F32 = 1e-5  # rows whose rank-25/26 gap < tau * range(M) are ambiguous


def load_fixture(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    fx = {k: z[k] for k in z.files if not k.startswith(("p:", "g:"))}
    fx["params"] = {k[2:]: z[k] for k in z.files if k.startswith("p:")}
    fx["grads"] = {k[2:]: z[k] for k in z.files if k.startswith("g:")}
    fx["shift"] = int(fx["shift"]); fx["nH"] = int(fx["nH"]); fx["hw"] = int(fx["hw"])
    if "params_from_seed" in fx:
        import lewin_b200 as L
        from oracle import param_fill
        C = fx["x"].shape[-1]
        sd = L.LeWinTransformerBlock(dim=C, input_resolution=(128, 128), num_heads=fx["nH"], win_size=8,
                                     shift_size=fx["shift"]).state_dict()
        fx["params"] = {k: (param_fill.fill_value(k, tuple(v.shape), int(fx["seed"])) if v.is_floating_point() else v.numpy())
                        for k, v in sd.items()}
        fx["grad_norm"] = {k[3:]: float(z[k]) for k in z.files if k.startswith("gn:")}
        fx["grad_sample"] = {k[3:]: z[k] for k in z.files if k.startswith("gs:")}
        fx = {k: v for k, v in fx.items() if not k.startswith(("gn:", "gs:"))}
    return fx


def grad_sample_ids(key, n, seed, k=4096):
    """The element sample of a compact fixture's parameter gradient (same rule as oracle/make_golden.py)."""
    import zlib
    rng = np.random.default_rng([int(seed), zlib.crc32(key.encode()), 7])
    return np.sort(rng.choice(n, size=min(n, k), replace=False))


def check_compact_grads(fx, grads, rtol):
    """Parameter gradients against a compact fixture: sampled elements and the L2 norm, relative to each gradient's scale
    (analytically-zero gradients - the key bias - are floored at 1e-4 of the largest gradient)."""
    gscale = max(np.abs(v).max() for v in fx["grad_sample"].values())
    assert sorted(grads) == sorted(fx["grad_sample"])
    for k, ref in fx["grad_sample"].items():
        got = np.asarray(grads[k], dtype=np.float64)
        ids = grad_sample_ids(k, got.size, fx["seed"])
        floor = max(np.abs(ref).max(), 1e-4 * gscale)
        assert np.abs(got.ravel()[ids] - ref).max() < rtol * floor, k
        assert abs(np.linalg.norm(got) - fx["grad_norm"][k]) < rtol * max(fx["grad_norm"][k], 1e-4 * gscale * np.sqrt(got.size)), k


def check_compact_grads_statistical(fx, grads, cos_min, norm_rtol):
    """bf16 gradients against a compact fixture: cosine similarity on the sampled elements and the L2 norm; the key
    bias, whose gradient is analytically zero, only has to stay at noise level."""
    nmax = max(fx["grad_norm"].values())
    assert sorted(grads) == sorted(fx["grad_sample"])
    for k, ref in fx["grad_sample"].items():
        got = np.asarray(grads[k], dtype=np.float64)
        if k.endswith("key_projection.bias"):     # analytically zero (softmax shift invariance): rounding noise on both sides
            assert np.linalg.norm(got) < 1e-2 * nmax and fx["grad_norm"][k] < 1e-2 * nmax, k
            continue
        a, b = got.ravel()[grad_sample_ids(k, got.size, fx["seed"])], ref.astype(np.float64)
        cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
        assert cos > cos_min, (k, cos)
        assert abs(np.linalg.norm(got) - fx["grad_norm"][k]) < norm_rtol * fx["grad_norm"][k], (k, np.linalg.norm(got), fx["grad_norm"][k])


def make_block(fx, device=None):
    """This package's LeWinTransformerBlock with the fixture's (reference) state_dict, strict load."""
    import torch
    import lewin_b200 as L
    C = fx["x"].shape[-1]
    dp = 0.4 if "drop_scale" in fx else 0.0
    blk = L.LeWinTransformerBlock(dim=C, input_resolution=(128, 128), num_heads=fx["nH"], win_size=8,
                                  shift_size=fx["shift"], drop_path=dp)
    blk.load_state_dict({k: torch.from_numpy(v) for k, v in fx["params"].items()}, strict=True)
    if device is not None:
        blk = blk.to(device)
    return blk


def force_drop_scales(blk, scales, device):
    """Make blk.drop_path hand out the recorded per-sample factors (attention branch, then LeFF branch)."""
    import torch
    it = iter([torch.from_numpy(np.asarray(s, dtype=np.float32)).to(device) for s in scales])
    blk.drop_path.sample_scale = lambda x: next(it)


def check_top(top_gpu, top_ref, rel_gap, tau):
    """Tie-aware comparison of selected query sets: exact equality required on every (window, head) row whose
    rank-u / rank-(u+1) gap (relative to the row's M range, from the fp64 oracle) is >= tau."""
    a = np.sort(np.asarray(top_gpu).astype(np.int64), -1)
    b = np.sort(np.asarray(top_ref).astype(np.int64), -1)
    bad = (a != b).any(-1)
    ambiguous = rel_gap < tau
    hard_fail = bad & ~ambiguous
    return int(bad.sum()), int(ambiguous.sum()), int(hard_fail.sum())
