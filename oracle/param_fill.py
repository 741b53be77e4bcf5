"""Deterministic, name-keyed parameter fill shared by the golden generator and the tests.

TEST INFRASTRUCTURE ONLY.  Fills any module whose state_dict keys follow the reference
layout with values that depend only on (seed, key name, shape) —
numpy PCG64, so independent of the torch version and of constructor RNG order.  Used so
that the reference model (in the build container) and the B200 host model (on the GPU
box) carry identical weights without committing a 105 MB checkpoint.
"""
from __future__ import annotations

import zlib

import numpy as np


def _std_for(name: str, shape) -> tuple[float, float]:
    """(mean, std) by parameter role; scaled so activations stay O(1) and softmaxes are non-flat."""
    if name.endswith("relative_position_bias_table"):
        return 0.0, 1.0
    if ".norm" in name or name.startswith("norm"):
        return (1.0, 0.1) if name.endswith("weight") else (0.0, 0.1)
    if name.endswith("bias"):
        return 0.0, 0.1
    if len(shape) == 2:                       # Linear [out, in]
        return 0.0, 1.0 / np.sqrt(shape[1])
    if len(shape) == 4:
        if "deconv" in name:                  # ConvTranspose2d [in, out, 2, 2]: one tap per output pixel
            return 0.0, 1.0 / np.sqrt(shape[0])
        return 0.0, 1.0 / np.sqrt(shape[1] * shape[2] * shape[3])
    return 0.0, 0.1


def fill_value(name: str, shape, seed: int) -> np.ndarray:
    rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
    mean, std = _std_for(name, tuple(shape))
    return (mean + std * rng.standard_normal(tuple(shape))).astype(np.float32)


def fill_module(module, seed: int, prefix_strip: str = ""):
    """In-place fill of every floating-point entry of ``module.state_dict()``."""
    import torch

    sd = module.state_dict()
    with torch.no_grad():
        for name, t in sd.items():
            if not t.is_floating_point():
                continue
            key = name[len(prefix_strip):] if prefix_strip and name.startswith(prefix_strip) else name
            t.copy_(torch.from_numpy(fill_value(key, t.shape, seed)).to(t.dtype))
    return module
