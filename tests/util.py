"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

BLOCK_FIXTURES = ["block_c32_h1_s0", "block_c32_h1_s4", "block_c64_h2_s4", "block_c64_h2_s4_inmask",
                  "block_c64_h2_s4_droppath", "block_c128_h4_s0"]

TOL_F32 = 1e-3     # BASELINE.json north_star: max-abs 1e-3 (fp32)
TOL_BF16 = 2e-2    # max-abs 2e-2 (bf16)
TIE_TAU_F32 = 1e-5  # rows whose rank-25/26 gap < tau * range(M) are ambiguous


def load_fixture(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    fx = {k: z[k] for k in z.files if not k.startswith(("p:", "g:"))}
    fx["params"] = {k[2:]: z[k] for k in z.files if k.startswith("p:")}
    fx["grads"] = {k[2:]: z[k] for k in z.files if k.startswith("g:")}
    fx["shift"] = int(fx["shift"]); fx["nH"] = int(fx["nH"]); fx["hw"] = int(fx["hw"])
    return fx


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
