"""Drop the sm_100a ops into an UNMODIFIED reference model.

``patch(model)`` rebinds ``forward`` of every reference ``LeWinTransformerBlock``, ``WindowAttention`` and
``LeFF`` instance (matched by class name, My_model_1.py:738 / :336 / :477) to the implementations of
``modules.py``.  Parameter objects are untouched, so optimizers, checkpoints (utils/model_utils.py:28-40),
``nn.DataParallel`` replication and ``state_dict()`` keep working; ``unpatch(model)`` restores the
reference forwards.  The reference keeps drawing nothing itself afterwards: the patched block draws
``index_sample`` with the same ``torch.randint(64, (64, 25))`` call at the same point (attn.py:91).
"""
from __future__ import annotations

import types

import torch.nn as nn

from . import modules as M


def _block_forward(self, x, mask=None):
    return M.lewin_block_forward(self, x, mask)


def _window_attention_forward(self, x, attn_kv=None, mask=None):
    return M.WindowAttention.forward(self, x, attn_kv, mask)


def _leff_forward(self, x):
    return M.LeFF.forward(self, x)


def _attention_layer_qkv_weights(self):
    return M._cat_qkv(self)


_TARGETS = {
    "LeWinTransformerBlock": _block_forward,
    "WindowAttention": _window_attention_forward,
    "LeFF": _leff_forward,
}


def patch(model: nn.Module) -> nn.Module:
    """In-place; returns ``model``.  Raises if no LeWin block is found (nothing to accelerate)."""
    n = 0
    for mod in model.modules():
        fn = _TARGETS.get(type(mod).__name__)
        if fn is None or isinstance(mod, (M.LeWinTransformerBlock, M.WindowAttention, M.LeFF)):
            continue
        if type(mod).__name__ == "LeWinTransformerBlock" and getattr(mod, "token_mlp", "leff") != "leff":
            raise NotImplementedError("lewin_b200.patch: token_mlp must be 'leff'")
        if "_lewin_b200_orig_forward" not in mod.__dict__:
            mod.__dict__["_lewin_b200_orig_forward"] = mod.__dict__.get("forward")
        mod.forward = types.MethodType(fn, mod)
        if type(mod).__name__ == "WindowAttention" and not hasattr(mod.ProbSpare, "qkv_weights"):
            mod.ProbSpare.qkv_weights = types.MethodType(_attention_layer_qkv_weights, mod.ProbSpare)
        n += 1
    if n == 0:
        raise RuntimeError("lewin_b200.patch: no LeWinTransformerBlock / WindowAttention / LeFF modules found")
    return model


def unpatch(model: nn.Module) -> nn.Module:
    for mod in model.modules():
        if "_lewin_b200_orig_forward" in mod.__dict__:
            orig = mod.__dict__.pop("_lewin_b200_orig_forward")
            if orig is None:
                mod.__dict__.pop("forward", None)
            else:
                mod.forward = orig
            if hasattr(mod, "ProbSpare"):
                mod.ProbSpare.__dict__.pop("qkv_weights", None)
    return model
