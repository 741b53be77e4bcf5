"""B200-native (sm_100a) LeWin hot path of Uformer_ProbSparse: shifted-window ProbSparse attention + LeFF.

Importable as ``lewin_b200`` (alias module at the repo root).  The CUDA library is loaded lazily on the
first op call and there is no CPU fallback (``_lib.load`` raises if ``csrc/liblewin_b200.so`` is missing).
"""
from . import _lib, options, fullres, parallel, losses, training, canvas_bands  # noqa: F401
from .modules import (AttentionLayer, DropPath, LeFF, LeWinTransformerBlock, LinearProjection,  # noqa: F401
                      ProbAttention, WindowAttention, draw_index_sample, lewin_block_forward)
from .ops import lewin_attn, lewin_leff, probsparse_core  # noqa: F401
from .patch import patch, unpatch  # noqa: F401
from .uformer import BasicUformerLayer, Downsample, InputProj, OutputProj, Upsample, Uformer  # noqa: F401

__all__ = [
    "AttentionLayer", "DropPath", "LeFF", "LeWinTransformerBlock", "LinearProjection", "ProbAttention",
    "WindowAttention", "draw_index_sample", "lewin_block_forward", "lewin_attn", "lewin_leff", "probsparse_core",
    "patch", "unpatch", "BasicUformerLayer", "Downsample", "InputProj", "OutputProj", "Upsample", "Uformer",
]
