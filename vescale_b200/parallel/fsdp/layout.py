"""Unit buffer layout for RaggedShard FSDP.

All parameters of an FSDP unit live back to back in ONE flat buffer of ``world * shard_size`` elements; rank
``r`` owns the contiguous slice ``[r*shard_size, (r+1)*shard_size)``.  Every parameter has a *block
granularity* ``g`` (a row by default, ``block_rows`` rows for block-quantised weights, one element for 1-D
params) and the planner guarantees that (a) a parameter is contiguous in the gathered buffer and (b) rank
boundaries never cut through a block.  It does so by aligning each parameter's offset to its granularity
and choosing ``shard_size`` as a multiple of the lcm of the granularities.  Consequences:

* all-gather / reduce-scatter of the unit are single, even, zero-copy collectives over the flat buffer (no
  interleaved copy-in/copy-out as with per-parameter ``Shard(0)``);
* each parameter is exactly a ``RaggedShard(dims=(0,), local_units=blocks_per_rank)`` DTensor whose local
  tensor is a view into the owner's slice (ranks that hold none of it have 0 units).

Design source: reference ``docs/texts/raggedshard.md:67-77`` (zero-copy batched collectives, block-wise
quantisation granularity); the wrapper itself is not in the reference tree.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

from ...placement import RaggedShard

__all__ = ["ParamSlot", "UnitLayout", "row_granularity"]


def row_granularity(shape: Sequence[int], block_rows: int = 1) -> int:
    """Elements per un-cuttable block: ``block_rows`` rows of a >=2-D weight, 1 element for vectors."""
    if len(shape) <= 1:
        return 1
    row = math.prod(shape[1:])
    br = block_rows if shape[0] % block_rows == 0 else 1
    return row * br


@dataclass
class ParamSlot:
    name: str
    shape: Tuple[int, ...]
    numel: int
    granularity: int
    offset: int = 0  # element offset inside the full unit buffer

    @property
    def end(self) -> int:
        return self.offset + self.numel


class UnitLayout:
    def __init__(
        self,
        params: Sequence[Tuple[str, Sequence[int]]],
        world: int,
        *,
        align: int = 64,
        granularity_fn: Optional[Callable[[str, Sequence[int]], int]] = None,
    ):
        self.world = world
        self.align = align
        gf = granularity_fn or (lambda n, s: row_granularity(s))
        self.slots: List[ParamSlot] = []
        for name, shape in params:
            shape = tuple(int(x) for x in shape)
            numel = math.prod(shape) if shape else 1
            g = int(gf(name, shape))
            if numel % g != 0:
                raise ValueError(f"{name}: numel {numel} is not a whole number of blocks of {g}")
            self.slots.append(ParamSlot(name, shape, numel, g))
        self._plan()

    def _plan(self) -> None:
        W, A = self.world, self.align
        G = A
        for s in self.slots:
            G = math.lcm(G, s.granularity)
        # a rank boundary k*S inside param i must satisfy (k*S - off_i) % g_i == 0: off_i % g_i == 0 and S % g_i == 0
        def place() -> int:
            pos = 0
            for s in self.slots:
                a = math.lcm(s.granularity, A) if s.granularity > 1 else A
                pos = (pos + a - 1) // a * a
                s.offset = pos
                pos += s.numel
            return pos

        used = place()
        S = max(G, (((used + W - 1) // W) + G - 1) // G * G)
        self.shard_size = S
        self.total = S * W
        self.used = used
        self.block_lcm = G

    # ------------------------------------------------------------------ queries
    def slot(self, name: str) -> ParamSlot:
        for s in self.slots:
            if s.name == name:
                return s
        raise KeyError(name)

    def rank_range(self, slot: ParamSlot, rank: int) -> Tuple[int, int]:
        """[lo, hi) of the part of ``slot`` owned by ``rank``, as offsets inside that rank's shard."""
        S = self.shard_size
        lo = max(slot.offset, rank * S)
        hi = min(slot.end, (rank + 1) * S)
        if hi <= lo:
            return 0, 0
        return lo - rank * S, hi - rank * S

    def local_units(self, slot: ParamSlot) -> Tuple[int, ...]:
        units = []
        for r in range(self.world):
            lo, hi = self.rank_range(slot, r)
            n = hi - lo
            if n % slot.granularity != 0:
                raise AssertionError(f"{slot.name}: rank {r} boundary cuts a block ({n} % {slot.granularity})")
            units.append(n // slot.granularity)
        return tuple(units)

    def placement(self, slot: ParamSlot) -> RaggedShard:
        dims = (0,) if len(slot.shape) >= 1 else ()
        if len(slot.shape) == 0:
            dims = ()
        return RaggedShard(dims if len(slot.shape) > 0 else (), self.local_units(slot))

    def padding_fraction(self) -> float:
        real = sum(s.numel for s in self.slots)
        return 1.0 - real / self.total

    def segments(self, rank: int) -> List[Tuple[int, int, str]]:
        """(lo, hi, name) pieces of this rank's shard that belong to real parameters (rest is padding)."""
        out = []
        for s in self.slots:
            lo, hi = self.rank_range(s, rank)
            if hi > lo:
                out.append((lo, hi, s.name))
        return out

    def __repr__(self) -> str:
        return f"UnitLayout(world={self.world}, shard={self.shard_size}, total={self.total}, pad={self.padding_fraction():.4%}, params={len(self.slots)})"
