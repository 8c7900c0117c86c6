"""``@patch_method(target, "name")``: replace ``target.name`` with the decorated function, keeping the original reachable
as ``fn.__wrapped_original__`` (reference ``vescale/utils/monkey_patch.py:20-35``; it uses this to graft ``is_ragged_shard``
onto torch's ``Placement``).  ``unpatch_all()`` restores everything — the reference has no undo."""
from __future__ import annotations

from typing import Callable, List, Tuple

__all__ = ["patch_method", "unpatch_all"]

_PATCHED: List[Tuple[object, str, object, bool]] = []


def patch_method(target: object, name: str) -> Callable[[Callable], Callable]:
    def deco(fn: Callable) -> Callable:
        had = hasattr(target, name)
        orig = getattr(target, name, None)
        fn.__wrapped_original__ = orig
        setattr(target, name, fn)
        _PATCHED.append((target, name, orig, had))
        return fn

    return deco


def unpatch_all() -> None:
    while _PATCHED:
        target, name, orig, had = _PATCHED.pop()
        if had:
            setattr(target, name, orig)
        else:
            try:
                delattr(target, name)
            except AttributeError:
                pass
