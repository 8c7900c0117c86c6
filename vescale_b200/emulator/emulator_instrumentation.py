"""Run ordinary torch code on the global view: inside ``EmulatorInstrumentation`` the listed torch functions accept *lists*
of per-rank tensors and are mapped rank by rank, so a single-process script computes what every rank would compute and the
emulated collectives can sit in between (legacy ``emulator/emulator_instrumentation.py``)."""
from __future__ import annotations

import functools
from typing import Callable, Iterable, List, Tuple

import torch

__all__ = ["EmulatorInstrumentation", "map_over_ranks", "decorate_function", "instrument", "revert_instrument", "instrument_all", "revert_instrument_all"]


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


# ---- explicit-index form: the caller says which positional arguments are per-rank lists (the reference's call shape:
# ``EmulatorInstrumentation(torch, ["mm", "nn.functional.relu"], [(0, 1), (0,)])``)
def decorate_function(func: Callable, indices) -> Callable:
    """``func`` mapped over the ranks of the list arguments at positions ``indices`` (every other argument is shared)."""
    if not callable(func):
        return func
    indices = tuple(indices)

    @functools.wraps(func)
    def per_rank(*args, **kwargs):
        world = len(args[indices[0]])
        return [func(*[a[r] if i in indices else a for i, a in enumerate(args)], **kwargs) for r in range(world)]

    return per_rank


def _resolve(obj, dotted: str):
    *path, leaf = dotted.split(".")
    for name in path:
        obj = getattr(obj, name)
    return obj, leaf


def instrument(obj, func_name: str, indices) -> Callable:
    """Replace ``obj.<dotted func_name>`` by its per-rank form; returns the original for ``revert_instrument``."""
    owner, leaf = _resolve(obj, func_name)
    orig = getattr(owner, leaf)
    setattr(owner, leaf, decorate_function(orig, indices))
    return orig


def revert_instrument(obj, func_name: str, orig: Callable) -> None:
    owner, leaf = _resolve(obj, func_name)
    setattr(owner, leaf, orig)


def instrument_all(obj, func_names, indices_list) -> dict:
    return {name: instrument(obj, name, idx) for name, idx in zip(func_names, indices_list)}


def revert_instrument_all(obj, func_names, originals: dict) -> None:
    for name in func_names:
        revert_instrument(obj, name, originals[name])


class EmulatorInstrumentation:
    """Two call forms:

    * ``EmulatorInstrumentation(world, [(torch, "add"), (torch.nn.functional, "relu")])`` — any argument that is a list of ``world``
      tensors is treated as per-rank;
    * ``EmulatorInstrumentation(torch, ["mm", "nn.functional.relu"], [(0, 1), (0,)])`` — dotted names under one object with the
      positions of the per-rank list arguments spelled out (the reference's shape)."""

    def __init__(self, world_or_obj, targets_or_names, indices_list=None):
        self._saved: List[Tuple[object, str, Callable]] = []
        if indices_list is None and isinstance(world_or_obj, int):
            self.world = world_or_obj
            self._wrap = [(mod, name, functools.partial(map_over_ranks, world=self.world)) for mod, name in targets_or_names]
        else:
            names = list(targets_or_names)
            idx = list(indices_list) if indices_list is not None else [(0,)] * len(names)
            self._wrap = []
            for name, ix in zip(names, idx):
                owner, leaf = _resolve(world_or_obj, name)
                self._wrap.append((owner, leaf, functools.partial(decorate_function, indices=ix)))

    def __enter__(self):
        for owner, name, wrap in self._wrap:
            orig = getattr(owner, name)
            self._saved.append((owner, name, orig))
            setattr(owner, name, wrap(orig))
        return self

    def __exit__(self, *exc):
        for owner, name, orig in reversed(self._saved):
            setattr(owner, name, orig)
        self._saved.clear()
        return False
