"""Placement types: how one tensor dimension (or the flat storage) is laid out along one mesh dim.

Pure math, no communication.  Collectives that realise a placement change live in
``vescale_b200.dtensor.redistribute`` / ``vescale_b200.comm``.

Parity: reference ``vescale/dtensor/placement_types.py:45-230`` (RaggedShard, _StridedRaggedShard),
``legacy/vescale/dtensor/placement_types.py:64-372`` (Shard, Replicate, Partial, InterleavedShard).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

__all__ = [
    "Placement",
    "Shard",
    "Replicate",
    "Partial",
    "_Partial",
    "RaggedShard",
    "_StridedRaggedShard",
    "_StridedShard",
    "InterleavedShard",
    "is_ragged_shard",
    "normalize_placements",
]


class Placement:
    """Base class.  Subclasses are frozen, hashable value types."""

    def is_shard(self, dim: Optional[int] = None) -> bool:
        return False

    def is_replicate(self) -> bool:
        return False

    def is_partial(self, reduce_op: Optional[str] = None) -> bool:
        return False

    def is_ragged_shard(self) -> bool:
        return False

    def is_interleaved_shard(self, dim: Optional[int] = None) -> bool:
        return False


def shard_size_and_offset(size: int, num_chunks: int, idx: int) -> Tuple[int, int]:
    """``torch.chunk`` semantics: ceil-sized leading chunks, possibly short/empty tail chunks."""
    full = (size + num_chunks - 1) // num_chunks if num_chunks > 0 else size
    start = min(full * idx, size)
    end = min(full * (idx + 1), size)
    return end - start, start


@dataclass(frozen=True)
class Shard(Placement):
    """Even (torch.chunk-style) sharding of tensor dim ``dim`` along a mesh dim."""

    dim: int

    def is_shard(self, dim: Optional[int] = None) -> bool:
        return dim is None or self.dim == dim

    def local_size_and_offset(self, size: int, num_chunks: int, idx: int) -> Tuple[int, int]:
        return shard_size_and_offset(size, num_chunks, idx)

    def split_tensor(
        self, tensor: torch.Tensor, num_chunks: int, *, with_padding: bool = False, contiguous: bool = True
    ) -> Tuple[List[torch.Tensor], List[int]]:
        """Split ``tensor`` on ``self.dim`` into exactly ``num_chunks`` pieces (tail pieces may be empty).

        With ``with_padding`` every piece is zero-padded to the size of the first one; the per-piece
        pad amounts are returned so the caller can un-pad after an even collective.
        """
        dim = self.dim if self.dim >= 0 else self.dim + tensor.ndim
        size = tensor.size(dim)
        full = (size + num_chunks - 1) // num_chunks
        pieces: List[torch.Tensor] = []
        pads: List[int] = []
        for i in range(num_chunks):
            n, off = shard_size_and_offset(size, num_chunks, i)
            p = tensor.narrow(dim, off, n)
            pad = full - n
            if with_padding and pad > 0:
                shape = list(p.shape)
                shape[dim] = pad
                p = torch.cat([p, p.new_zeros(shape)], dim=dim)
            elif contiguous:
                p = p.contiguous()
            pieces.append(p)
            pads.append(pad)
        return pieces, pads

    # legacy spellings (``legacy/vescale/dtensor/placement_types.py:69-170``; padding on by default there)
    def _split_tensor(self, tensor: torch.Tensor, num_chunks: int, *, with_padding: bool = True, contiguous: bool = True):
        return self.split_tensor(tensor, num_chunks, with_padding=with_padding, contiguous=contiguous)

    def _pad_tensor(self, tensor: torch.Tensor, pad_size: int) -> torch.Tensor:
        """Append ``pad_size`` zeros along the sharded dim."""
        if pad_size <= 0:
            return tensor
        dim = self.dim if self.dim >= 0 else self.dim + tensor.ndim
        shape = list(tensor.shape)
        shape[dim] = pad_size
        return torch.cat([tensor, tensor.new_zeros(shape)], dim=dim)

    def _unpad_tensor(self, tensor: torch.Tensor, pad_size: int) -> torch.Tensor:
        if pad_size <= 0:
            return tensor
        dim = self.dim if self.dim >= 0 else self.dim + tensor.ndim
        return tensor.narrow(dim, 0, tensor.size(dim) - pad_size)

    def __hash__(self) -> int:
        # not hash((dim,)): CPython hashes -1 and -2 alike, and Shard(-1) / Shard(-2) key different cache entries
        return hash((type(self).__name__, self.dim + (1 << 20)))

    def __repr__(self) -> str:
        return f"Shard(dim={self.dim})"

    def __str__(self) -> str:
        return f"S({self.dim})"


@dataclass(frozen=True)
class _StridedShard(Shard):
    """Shard whose chunks were produced *after* a later (inner) mesh dim already sharded the same
    tensor dim ``split_factor`` ways (FSDP-over-TP ordering).  Local data of mesh coordinate ``i``
    is the concatenation, over the ``split_factor`` inner pieces, of the ``i``-th sub-chunk."""

    split_factor: int = 1

    def __hash__(self) -> int:
        return hash((type(self).__name__, self.dim + (1 << 20), self.split_factor))

    def __repr__(self) -> str:
        return f"_StridedShard(dim={self.dim}, sf={self.split_factor})"

    def __str__(self) -> str:
        return f"_S({self.dim}, {self.split_factor})"


@dataclass(frozen=True)
class Replicate(Placement):
    def is_replicate(self) -> bool:
        return True

    def __repr__(self) -> str:
        return "Replicate()"

    def __str__(self) -> str:
        return "R"


_REDUCE_OPS = ("sum", "avg", "max", "min", "product", "band", "bor", "bxor")


@dataclass(frozen=True)
class Partial(Placement):
    """Pending reduction along a mesh dim.  ``reduce_op`` is a c10d op name, or ``"norm{p}"`` for a
    p-norm partial (local values are p-norms of disjoint pieces)."""

    reduce_op: str = "sum"

    def __post_init__(self):
        op = self.reduce_op
        if hasattr(op, "name"):  # c10d ReduceOp enum instance
            object.__setattr__(self, "reduce_op", str(op.name).lower())
        elif not isinstance(op, str):
            object.__setattr__(self, "reduce_op", str(op).lower().split(".")[-1])
        else:
            object.__setattr__(self, "reduce_op", op.lower())

    def is_partial(self, reduce_op: Optional[str] = None) -> bool:
        return reduce_op is None or reduce_op == self.reduce_op

    @property
    def norm_type(self) -> Optional[float]:
        if self.reduce_op.startswith("norm"):
            return float(self.reduce_op[4:])
        return None

    def __repr__(self) -> str:
        return f"Partial({self.reduce_op})"

    def __str__(self) -> str:
        return "P" if self.reduce_op == "sum" else f"P({self.reduce_op})"


_Partial = Partial


@dataclass(frozen=True)
class RaggedShard(Placement):
    """Asymmetric sharding of the *flattened contiguous storage* of the leading ``dims``.

    ``local_units[i]`` is the relative share held by mesh coordinate ``i`` (zeros allowed, so
    "whole tensor on one rank" is ``(0, .., 1, .., 0)``).  One unit is
    ``numel(tensor) // sum(local_units)`` elements; with ``dims=(0,..,k-1)`` every local piece is a
    whole number of ``prod(shape[k:])``-element rows.  Local tensors are 1-D.

    Parity: reference ``vescale/dtensor/placement_types.py:45-226``.
    """

    dims: Tuple[int, ...]
    local_units: Tuple[int, ...]

    def __post_init__(self):
        object.__setattr__(self, "dims", tuple(int(d) for d in self.dims))
        object.__setattr__(self, "local_units", tuple(int(u) for u in self.local_units))
        if any(u < 0 for u in self.local_units) or sum(self.local_units) <= 0:
            raise ValueError(f"local_units must be non-negative with a positive sum, got {self.local_units}")

    def is_ragged_shard(self) -> bool:
        return True

    @property
    def total_units(self) -> int:
        return sum(self.local_units)

    def unit_prefix(self, idx: int) -> int:
        return sum(self.local_units[:idx])

    def flat_range(self, numel: int, idx: int) -> Tuple[int, int]:
        """[start, end) in flat element index of coordinate ``idx`` for a tensor of ``numel`` elements."""
        tot = self.total_units
        if numel % tot != 0:
            raise ValueError(f"numel {numel} is not divisible by sum(local_units)={tot}")
        r = numel // tot
        s = self.unit_prefix(idx) * r
        return s, s + self.local_units[idx] * r

    def split_tensor(self, tensor: torch.Tensor, num_chunks: int) -> List[torch.Tensor]:
        if not tensor.is_contiguous():
            raise ValueError("RaggedShard expects a contiguous tensor")
        if num_chunks != len(self.local_units):
            raise ValueError("num_chunks must equal len(local_units)")
        flat = tensor.reshape(-1)
        out = []
        for i in range(num_chunks):
            s, e = self.flat_range(flat.numel(), i)
            out.append(flat.narrow(0, s, e - s))
        return out

    # reference-compatible private aliases
    _split_tensor = split_tensor

    def reconstruct_tensor_from_flat(self, flat_tensor: torch.Tensor, shape: Sequence[int]) -> torch.Tensor:
        if flat_tensor.ndim != 1:
            raise ValueError("flat_tensor must be 1-D")
        n = len(self.dims)
        if self.dims != tuple(range(n)):
            raise ValueError(f"dims must be (0, 1, ..., k-1), got {self.dims}")
        trailing = tuple(shape[n:])
        if flat_tensor.numel() % max(1, math.prod(trailing)) != 0:
            raise ValueError("flat numel is not a whole number of rows")
        return flat_tensor.view(-1, *trailing)

    def __repr__(self) -> str:
        return f"RaggedShard(dims={self.dims}, local_units={self.local_units})"

    __str__ = __repr__


@dataclass(frozen=True)
class _StridedRaggedShard(RaggedShard):
    """RaggedShard applied on a tensor whose dim 0 is *also* sharded by a later mesh dim
    (``[ _StridedRaggedShard, Shard(0) ]``): logically Shard(0) cuts first (``split_factor`` pieces),
    then each piece is ragged-sharded.  Parity: reference ``placement_types.py:228-230`` and
    ``docs/texts/raggedshard.md:60-63``."""

    split_factor: int = 1

    def __repr__(self) -> str:
        return f"_StridedRaggedShard(dims={self.dims}, local_units={self.local_units}, sf={self.split_factor})"

    __str__ = __repr__


@dataclass(frozen=True)
class InterleavedShard(Shard):
    """Shard of a dim that is itself a concatenation of ``interleaved_size`` equal sections (merged
    QKV / gate-up weights): each section is sharded evenly and coordinate ``i`` holds the ``i``-th
    piece of every section, concatenated.  Parity: ``legacy/vescale/dtensor/placement_types.py:284``."""

    interleaved_size: int = 1

    def is_interleaved_shard(self, dim: Optional[int] = None) -> bool:
        return dim is None or self.dim == dim

    def is_shard(self, dim: Optional[int] = None) -> bool:  # distinct from plain Shard in rules
        return False

    def split_tensor(self, tensor, num_chunks, *, with_padding=False, contiguous=True):
        dim = self.dim if self.dim >= 0 else self.dim + tensor.ndim
        size = tensor.size(dim)
        k = self.interleaved_size
        if size % (k * num_chunks) != 0:
            raise ValueError(f"InterleavedShard needs size % (interleaved_size*num_chunks) == 0, got {size}")
        shp = list(tensor.shape)
        view = tensor.reshape(*shp[:dim], k, num_chunks, size // (k * num_chunks), *shp[dim + 1 :])
        out = []
        for i in range(num_chunks):
            p = view.select(dim + 1, i).reshape(*shp[:dim], size // num_chunks, *shp[dim + 1 :])
            out.append(p.contiguous() if contiguous else p)
        return out, [0] * num_chunks

    def __repr__(self) -> str:
        return f"InterleavedShard(dim={self.dim}, interleaved_size={self.interleaved_size})"

    def __str__(self) -> str:
        return f"IS({self.dim},{self.interleaved_size})"


def is_ragged_shard(p: Placement) -> bool:
    return isinstance(p, RaggedShard)


def normalize_placements(placements, mesh_ndim: int, tensor_ndim: Optional[int] = None) -> Tuple[Placement, ...]:
    """None → all-Replicate; negative Shard dims normalised; length checked."""
    if placements is None:
        return tuple(Replicate() for _ in range(mesh_ndim))
    placements = tuple(placements)
    if len(placements) > mesh_ndim:
        raise ValueError(f"`placements` has {len(placements)} entries, more than the mesh has dims ({mesh_ndim}): {placements}")
    if len(placements) < mesh_ndim:
        # as in the legacy package (``dtensor/dtensor.py:60``): the trailing mesh dims are replicated, with a warning
        import warnings

        warnings.warn(f"`placements` has fewer entries ({len(placements)}) than the mesh has dims ({mesh_ndim}); appending Replicate()", UserWarning, stacklevel=3)
        placements = placements + tuple(Replicate() for _ in range(mesh_ndim - len(placements)))
    out = []
    for p in placements:
        if not isinstance(p, Placement):
            raise TypeError(f"not a Placement: {p!r}")
        if isinstance(p, Shard) and p.dim < 0:
            if tensor_ndim is None:
                raise ValueError("negative shard dim needs tensor ndim")
            kw = {k: getattr(p, k) for k in p.__dataclass_fields__}
            kw["dim"] = p.dim + tensor_ndim
            p = type(p)(**kw)
        out.append(p)
    return tuple(out)


def __getattr__(name):  # ``vescale.dtensor.placement_types`` also carries DTensorSpec / TensorMeta in the reference
    if name in ("DTensorSpec", "TensorMeta"):
        from . import spec

        return getattr(spec, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
