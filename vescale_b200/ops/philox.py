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
