"""Run ordinary torch code on the global view: inside ``EmulatorInstrumentation`` the listed torch functions accept *lists*
of per-rank tensors and are mapped rank by rank, so a single-process script computes what every rank would compute and the
emulated collectives can sit in between (legacy ``emulator/emulator_instrumentation.py``)."""
from __future__ import annotations

import functools
from typing import Callable, Iterable, List, Tuple

import torch

__all__ = ["EmulatorInstrumentation", "map_over_ranks"]


def _is_rank_list(x, world: int) -> bool:
    return isinstance(x, (list, tuple)) and len(x) == world and all(isinstance(t, torch.Tensor) for t in x)


def map_over_ranks(fn: Callable, world: int) -> Callable:
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        if not any(_is_rank_list(a, world) for a in list(args) + list(kwargs.values())):
            return fn(*args, **kwargs)
        outs = []
        for r in range(world):
            a = [x[r] if _is_rank_list(x, world) else x for x in args]
            k = {n: (x[r] if _is_rank_list(x, world) else x) for n, x in kwargs.items()}
            outs.append(fn(*a, **k))
        return outs

    return wrapped


class EmulatorInstrumentation:
    """``with EmulatorInstrumentation(world, [(torch, "add"), (torch, "mm"), (torch.nn.functional, "relu")]): ...``"""

    def __init__(self, world: int, targets: Iterable[Tuple[object, str]]):
        self.world = world
        self.targets: List[Tuple[object, str]] = list(targets)
        self._saved: List[Tuple[object, str, Callable]] = []

    def __enter__(self):
        for mod, name in self.targets:
            orig = getattr(mod, name)
            self._saved.append((mod, name, orig))
            setattr(mod, name, map_over_ranks(orig, self.world))
        return self

    def __exit__(self, *exc):
        for mod, name, orig in reversed(self._saved):
            setattr(mod, name, orig)
        self._saved.clear()
        return False
