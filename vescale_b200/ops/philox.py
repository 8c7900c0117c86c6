"""Front-end of ``csrc/philox_shard.cu``: fill the boxes of a local shard with the single-device-equivalent stream."""
from __future__ import annotations

from typing import Sequence

import torch

from . import _ext


def available() -> bool:
    return _ext.available() and hasattr(torch.ops.vescale_b200, "philox_fill_box")


def philox_fill_boxes(local: torch.Tensor, global_shape: Sequence[int], boxes, seed: int, offset: int, kind: str, low: float, high: float, mean: float, std: float, ragged: bool = False) -> None:
    gstride = [1] * len(global_shape)
    for i in range(len(global_shape) - 2, -1, -1):
        gstride[i] = gstride[i + 1] * global_shape[i + 1]
    normal = kind != "uniform"
    a, b = (mean, std) if normal else (low, high)
    for off, sz, loc in boxes:
        if ragged:
            # the local tensor is flat; the box occupies a contiguous run starting at loc[0], row-major inside the box
            ls = [1] * len(sz)
            for i in range(len(sz) - 2, -1, -1):
                ls[i] = ls[i + 1] * sz[i + 1]
            lbase = loc[0]
        else:
            ls = list(local.stride())
            lbase = sum(o * s for o, s in zip(loc, ls))
        _ext.count_launch("philox_fill")
        torch.ops.vescale_b200.philox_fill_box(local, list(sz), list(off), gstride, ls, lbase, int(seed), int(offset), normal, float(a), float(b))


def dropout_available() -> bool:
    return _ext.available() and hasattr(torch.ops.vescale_b200, "philox_dropout_box")


def philox_dropout_boxes(x: torch.Tensor, global_shape: Sequence[int], boxes, seed: int, offset: int, p: float, ragged: bool = False):
    """Fused sharded dropout (``csrc/philox_shard.cu``): returns ``(out, mask)`` for the contiguous local shard ``x`` whose
    elements sit at the global positions described by ``boxes``; the mask is the single-device one."""
    out = torch.empty_like(x)
    mask = torch.empty(x.shape, dtype=torch.bool, device=x.device)
    gstride = [1] * len(global_shape)
    for i in range(len(global_shape) - 2, -1, -1):
        gstride[i] = gstride[i + 1] * global_shape[i + 1]
    for off, sz, loc in boxes:
        if ragged:
            ls = [1] * len(sz)
            for i in range(len(sz) - 2, -1, -1):
                ls[i] = ls[i + 1] * sz[i + 1]
            lbase = loc[0]
        else:
            ls = list(x.stride())
            lbase = sum(o * s for o, s in zip(loc, ls))
        _ext.count_launch("philox_dropout")
        torch.ops.vescale_b200.philox_dropout_box(x, out, mask, list(sz), list(off), gstride, ls, lbase, int(seed), int(offset), float(p))
    return out, mask
