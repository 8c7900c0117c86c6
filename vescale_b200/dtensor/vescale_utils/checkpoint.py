"""Box decomposition of a ragged shard for checkpoint planning, under the reference's entry point
(``vescale/dtensor/vescale_utils/checkpoint.py:70-172`` ``_break_ragged_box``): the general form, where any *contiguous range* of
tensor dims is flattened into one "ragged" dim and the shard is an n-d box in that ragged view.

Built on ``layout.break_ragged_box`` (flat interval -> at most 2k-1 axis-aligned boxes over k dims): the flattened range is
decomposed, and every piece is extended by the box's extents in the untouched leading / trailing dims.  The DCP hooks of this
framework (``DTensor.__create_write_items__`` etc.) use ``layout.local_boxes`` directly."""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

from ...layout import break_ragged_box

__all__ = ["_break_ragged_box"]


def _break_ragged_box(ragged_sizes: Sequence[int], ragged_offsets: Sequence[int], ragged_dims: Tuple[int, ...], unragged_tensor_shape: Sequence[int],
                      original_tensor_shape: Sequence[int], global_offsets: Sequence[int]) -> Tuple[List[Tuple[int, ...]], List[Tuple[int, ...]]]:
    """``ragged_sizes`` / ``ragged_offsets``: the box in the ragged view (dims ``ragged_dims`` of ``unragged_tensor_shape``
    collapsed into one, at position ``ragged_dims[0]``); ``global_offsets``: where the un-ragged local tensor sits inside
    ``original_tensor_shape`` (the ragged-dim offset is given in global flat numbering).  Returns ``(sizes_list, offsets_list)``
    of axis-aligned boxes in un-ragged global coordinates that tile the box exactly, without overlap."""
    ragged_sizes, ragged_offsets = tuple(int(x) for x in ragged_sizes), tuple(int(x) for x in ragged_offsets)
    if ragged_sizes == (0,) and ragged_offsets == ():
        return [], []
    d0, d1 = ragged_dims[0], ragged_dims[-1]
    if tuple(ragged_dims) != tuple(range(d0, d1 + 1)):
        raise ValueError(f"ragged dims must be contiguous, got {ragged_dims}")
    if len(ragged_sizes) - 1 + len(ragged_dims) != len(unragged_tensor_shape):
        raise ValueError("ragged box rank does not match the un-ragged shape")
    sub_shape = tuple(int(unragged_tensor_shape[d]) for d in ragged_dims)
    # the flat offset of the local tensor's origin inside the flattened global dims
    g_flat = 0
    for i, d in enumerate(ragged_dims):
        g_flat += int(global_offsets[d]) * math.prod(int(original_tensor_shape[e]) for e in ragged_dims[i + 1 :])
    start = ragged_offsets[d0] - g_flat
    end = start + ragged_sizes[d0]
    lead_sz, lead_off = ragged_sizes[:d0], ragged_offsets[:d0]
    tail_sz, tail_off = ragged_sizes[d0 + 1 :], ragged_offsets[d0 + 1 :]
    if any(s == 0 for s in lead_sz + tail_sz):
        return [], []
    sizes_list, offsets_list = [], []
    for off, sz in break_ragged_box(sub_shape, start, end):
        off = tuple(o + int(global_offsets[d]) for o, d in zip(off, ragged_dims))
        sizes_list.append((*lead_sz, *sz, *tail_sz))
        offsets_list.append((*lead_off, *off, *tail_off))
    return sizes_list, offsets_list
