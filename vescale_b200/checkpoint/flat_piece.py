"""``FlatPiece``: a flat range of a LOCAL shard, checkpointed in the GLOBAL coordinates of the tensor it belongs to.

ZeRO-style optimizers own, per rank, a contiguous range ``[lo, hi)`` of the row-major flattening of each parameter's local shard
(bucket shards ignore parameter edges).  When the parameter itself is sharded over a model-parallel mesh (a DModule / TP
``Shard``), that range is a piece of a piece.  The reference gathers such 1-D ranges to one writer rank over P2P before saving
(``legacy/vescale/checkpoint/planner/vescale/vescale_planner.py:152-242``); here the range is written IN PLACE as the <= 2^n - 1
axis-aligned boxes it decomposes into (``layout.break_ragged_box``), shifted by the local shard's offset in the global tensor.
Every (data-parallel rank, model-parallel rank) writes disjoint boxes of one N-D tensor per key, so the state reloads under any
other DP size, bucket size or TP degree through torch-DCP's ordinary chunk intersection — no communication at save time.

The object is a ``torch.Tensor`` wrapper subclass only because DCP resolves ``__get_tensor_shard__`` on tensors
(``torch/distributed/checkpoint/utils.py::find_state_dict_object``); it supports no arithmetic."""
from __future__ import annotations

import math
from typing import Any, List, Sequence, Tuple

import torch

from ..layout import break_ragged_box

__all__ = ["FlatPiece"]


class FlatPiece(torch.Tensor):
    @staticmethod
    def __new__(cls, local: torch.Tensor, local_shape: Sequence[int], lo: int, hi: int, global_shape: Sequence[int], global_offset: Sequence[int]):
        r = torch.Tensor._make_wrapper_subclass(cls, tuple(int(s) for s in global_shape), dtype=local.dtype, device=local.device, requires_grad=False)
        return r

    def __init__(self, local: torch.Tensor, local_shape: Sequence[int], lo: int, hi: int, global_shape: Sequence[int], global_offset: Sequence[int]):
        if local.dim() != 1 or local.numel() != max(0, hi - lo):
            raise ValueError(f"FlatPiece: local tensor must be 1-D with hi - lo = {hi - lo} elements, got {tuple(local.shape)}")
        self._local_tensor = local
        self._geom = (tuple(int(s) for s in local_shape), int(lo), int(hi), tuple(int(s) for s in global_shape), tuple(int(s) for s in global_offset))

    def with_local(self, local: torch.Tensor) -> "FlatPiece":
        """Same geometry over other storage (the pinned staging copy of an asynchronous save)."""
        return FlatPiece(local, *self._geom)

    def __repr__(self) -> str:  # noqa: D105
        ls, lo, hi, gs, go = self._geom
        return f"FlatPiece([{lo}, {hi}) of local {ls} at {go} in {gs}, dtype={self.dtype})"

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        raise NotImplementedError(f"FlatPiece is a checkpoint descriptor, not a tensor to compute with ({func})")

    # ------------------------------------------------------------------ geometry
    def boxes(self) -> List[Tuple[Tuple[int, ...], Tuple[int, ...], int]]:
        """[(global offsets, sizes, start inside the flat local piece)] in flat order."""
        ls, lo, hi, _gs, go = self._geom
        out, pos = [], 0
        for off, sz in break_ragged_box(ls, lo, hi):
            out.append((tuple(o + g for o, g in zip(off, go)), tuple(sz), pos))
            pos += math.prod(sz)
        return out

    # ------------------------------------------------------------------ torch-DCP protocol (_Checkpointable)
    def __create_write_items__(self, fqn: str, object: Any):
        from torch.distributed.checkpoint.metadata import ChunkStorageMetadata, MetadataIndex, TensorProperties
        from torch.distributed.checkpoint.planner import TensorWriteData, WriteItem, WriteItemType

        gs = self._geom[3]
        return [
            WriteItem(
                index=MetadataIndex(fqn, torch.Size(off)),
                type=WriteItemType.SHARD,
                tensor_data=TensorWriteData(
                    chunk=ChunkStorageMetadata(offsets=torch.Size(off), sizes=torch.Size(sz)),
                    properties=TensorProperties.create_from_tensor(self._local_tensor),
                    size=torch.Size(gs),
                ),
            )
            for off, sz, _ in self.boxes()
        ]

    def __create_chunk_list__(self):
        from torch.distributed.checkpoint.metadata import ChunkStorageMetadata

        return [ChunkStorageMetadata(offsets=torch.Size(off), sizes=torch.Size(sz)) for off, sz, _ in self.boxes()]

    def __get_tensor_shard__(self, index):
        want = tuple(index.offset) if index.offset is not None else None
        for off, sz, pos in self.boxes():
            if want is None or off == want:
                return self._local_tensor.narrow(0, pos, math.prod(sz)).view(sz)
        raise ValueError(f"no box at offset {want} in {self!r}")
