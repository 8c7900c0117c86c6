"""Debug tooling for the in-kernel signalling protocols (the reference has no race tooling, and NCCL's own
checks do not see flags written by our kernels).

* ``poison(t)`` fills a symmetric buffer with NaNs (floating point) or a sentinel pattern: a consumer that reads a region
  before its producer has delivered it then shows up as NaNs / the sentinel in the numerics tests instead of silently using
  stale-but-plausible data from the previous step.
* ``check_epochs(arena)`` snapshots the signal pad and validates the protocol invariants that every kernel relies on:
  epochs are monotonic per (slot, source rank) between two snapshots, and within one slot all sources are within one epoch
  of each other at a quiescent point (after a device synchronize + barrier).
* ``SymmDebug(arena)`` is a context manager that poisons the free part of every chunk on entry and checks epochs on exit.

For memory errors and intra-kernel races use ``tools/sanitize.sh`` (compute-sanitizer memcheck / racecheck / synccheck on the
single-GPU kernel tests).  Enabled in the FSDP / TP paths by ``VESCALE_B200_SYMM_DEBUG=1`` (buffers are poisoned when they
are returned to a pool).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.distributed as dist

__all__ = ["poison", "check_epochs", "SymmDebug", "debug_enabled"]

_SENTINEL = 0x5A


def debug_enabled() -> bool:
    return os.environ.get("VESCALE_B200_SYMM_DEBUG", "0") == "1"


def poison(t: torch.Tensor) -> torch.Tensor:
    if t.is_floating_point():
        t.fill_(float("nan"))
    else:
        t.view(torch.uint8).fill_(_SENTINEL)
    return t


def check_epochs(arena, previous: Optional[torch.Tensor] = None, quiescent: bool = True) -> torch.Tensor:
    """Returns the pad snapshot ([slots, world] int64 on the host); raises ``AssertionError`` on a violated invariant."""
    if arena.pad.is_cuda:
        torch.cuda.synchronize(arena.device)
    if quiescent and dist.is_initialized():
        dist.barrier(group=arena.group)
    snap = arena.pad.detach().cpu().to(torch.int64).view(-1, arena.world) & 0xFFFFFFFF
    if previous is not None:
        went_back = (snap - previous) < 0
        # epochs are compared with wrap-around in the kernels; a decrease of more than 2^31 is a wrap, not a regression
        went_back &= (previous - snap) < (1 << 31)
        assert not went_back.any(), f"signal epochs went backwards at (slot, src) {went_back.nonzero().tolist()[:8]}"
    if quiescent:
        used = snap.amax(dim=1) > 0
        spread = snap.amax(dim=1) - snap.amin(dim=1)
        bad = used & (spread > 1)
        assert not bad.any(), f"slots whose sources disagree by more than one epoch at a quiescent point: {bad.nonzero().flatten().tolist()[:8]}"
    return snap


class SymmDebug:
    def __init__(self, arena):
        self.arena = arena
        self.before: Optional[torch.Tensor] = None

    def __enter__(self):
        self.before = check_epochs(self.arena, quiescent=True)
        for c in self.arena.chunks:  # never-allocated tail of every chunk
            if c.used < c.nbytes:
                c.tensor[c.used :].fill_(_SENTINEL)
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            check_epochs(self.arena, self.before, quiescent=True)
        return False
