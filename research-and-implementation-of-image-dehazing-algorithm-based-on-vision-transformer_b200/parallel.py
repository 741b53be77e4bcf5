"""Data-parallel training glue (BASELINE config 4): one process per GPU, NCCL gradient all-reduce.

The reference uses single-process ``nn.DataParallel`` (My_train.py:97).  The B200 path is
``torch.nn.parallel.DistributedDataParallel`` over NCCL/NVLink; the only hot-path specific issue is that every
LeWin block carries 6 DEAD parameters (attn.qkv.to_q / attn.qkv.to_kv / attn.proj, My_model_1.py:384-393) that are
in the state_dict but never used in forward (SURVEY.md finding 6): they never receive a gradient, so DDP's
reducer must not wait for them.  ``freeze_dead_parameters`` takes them out of autograd (they stay in the
state_dict, optimizers skip them exactly as they skip ``grad is None`` today) so that DDP needs neither
``find_unused_parameters`` nor a second graph traversal.
"""
from __future__ import annotations

import torch
import torch.nn as nn

DEAD_SUFFIXES = ("attn.qkv.to_q.weight", "attn.qkv.to_q.bias", "attn.qkv.to_kv.weight", "attn.qkv.to_kv.bias",
                 "attn.proj.weight", "attn.proj.bias")


def dead_parameter_names(model: nn.Module):
    return [n for n, _ in model.named_parameters() if n.endswith(DEAD_SUFFIXES)]


def freeze_dead_parameters(model: nn.Module) -> int:
    names = set(dead_parameter_names(model))
    for n, p in model.named_parameters():
        if n in names:
            p.requires_grad_(False)
    return len(names)


def wrap_ddp(model: nn.Module, device=None, bucket_cap_mb: int = 25, **kw):
    """DistributedDataParallel over the initialised process group, gradients as bucket views, dead parameters
    excluded.  ~20.6 M live parameters (82.5 MB fp32) -> 4 buckets overlapped with the backward."""
    freeze_dead_parameters(model)
    ids = None if device is None else [device.index if isinstance(device, torch.device) else int(device)]
    return nn.parallel.DistributedDataParallel(model, device_ids=ids, bucket_cap_mb=bucket_cap_mb,
                                               gradient_as_bucket_view=True, **kw)
