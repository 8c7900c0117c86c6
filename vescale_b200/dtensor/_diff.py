"""Behaviour switches that make this package differ from stock DTensor, gathered under the reference's module name (legacy
``dtensor/_diff.py``): environment flags, the dry-run / dump decorators of the pipeline tooling, and ``DeferReshardMode``.

    VESCALE_DISABLE_REDISTRIBUTE=1   ops never reshard their inputs implicitly; a mismatch raises (every communication is user-planned)
    VESCALE_DUMMY_P2P=1              pipeline p2p functions decorated with ``dummy_p2p`` log a line instead of communicating
    VESCALE_DUMP_INSTRUCTION=1       functions decorated with ``manage_dump_file`` write what they executed to a per-stage file"""
from __future__ import annotations

import functools
import os
from typing import Callable

import torch

__all__ = ["VESCALE_DISABLE_REDISTRIBUTE", "VESCALE_DUMMY_P2P", "VESCALE_DUMP_INSTRUCTION", "global_counter", "get_counter", "set_counter", "dummy_p2p", "manage_dump_file",
           "DeferReshardMode", "EnablePartialMode"]


def _flag(name: str, default: str = "0") -> bool:
    return os.environ.get(name, default) == "1"


VESCALE_DISABLE_REDISTRIBUTE = _flag("VESCALE_DISABLE_REDISTRIBUTE")  # snapshot at import; dispatch reads the environment live
VESCALE_DUMMY_P2P = _flag("VESCALE_DUMMY_P2P")
VESCALE_DUMP_INSTRUCTION = _flag("VESCALE_DUMP_INSTRUCTION")

global_counter = 0  # running index of the dumped / logged p2p calls (one numbering per process)


def get_counter() -> int:
    return global_counter


def set_counter(value: int) -> None:
    global global_counter
    global_counter = int(value)


def dummy_p2p(func: Callable) -> Callable:
    """Dry-run decorator for p2p functions: with ``VESCALE_DUMMY_P2P=1`` the call is logged (index, function, tensor shapes) to
    ``dummy_p2p_rank{STAGE_ID}.txt`` and skipped; a received tensor, if the function would return one, is fabricated from its
    ``tensor_shape`` / ``recv_dtype`` keyword arguments.  Without the flag: the function itself."""

    @functools.wraps(func)
    def wrap(*args, **kwargs):
        if not _flag("VESCALE_DUMMY_P2P"):
            return func(*args, **kwargs)
        global global_counter
        shapes = [tuple(a.shape) for a in list(args) + list(kwargs.values()) if isinstance(a, torch.Tensor)]
        line = f"{global_counter}: {func.__name__} tensors={shapes} kw={sorted(k for k, v in kwargs.items() if not isinstance(v, torch.Tensor))}"
        global_counter += 1
        with open(f"dummy_p2p_rank{os.environ.get('STAGE_ID', os.environ.get('RANK', '0'))}.txt", "a") as f:
            f.write(line + "\n")
        shape = kwargs.get("tensor_shape")
        if func.__name__.startswith("recv") or "recv" in func.__name__.split("_"):
            return torch.zeros(tuple(shape), dtype=kwargs.get("recv_dtype") or torch.float32) if shape is not None else None
        return None

    return wrap


def manage_dump_file(func: Callable) -> Callable:
    """With ``VESCALE_DUMP_INSTRUCTION=1``: before the decorated executor runs, the file ``instruction_dump_stage{id}.txt`` is started
    afresh and, if the bound object can render its program (``dump(stage)`` / ``gen_instruction_str_list()``), the program is written
    into it — so what each stage was ABOUT to execute survives a hang.  ``stage_id`` is the first positional argument after self."""

    @functools.wraps(func)
    def wrap(self, *args, **kwargs):
        if _flag("VESCALE_DUMP_INSTRUCTION"):
            stage = args[0] if args else kwargs.get("stage_id", os.environ.get("STAGE_ID", "0"))
            text = None
            if hasattr(self, "dump"):
                try:
                    text = self.dump(stage)
                except Exception:  # noqa: BLE001
                    text = None
            if text is None and hasattr(self, "gen_instruction_str_list"):
                lst = self.gen_instruction_str_list()
                text = lst[stage] if isinstance(stage, int) and stage < len(lst) else "\n".join(lst)
            with open(f"instruction_dump_stage{stage}.txt", "w") as f:
                f.write((text or "") + "\n")
        return func(self, *args, **kwargs)

    return wrap


class DeferReshardMode:
    """``with DeferReshardMode():`` — ``Partial + Partial`` stays ``Partial`` inside (one reduction for a chain of additions).  That is
    this package's default already; the context manager exists for code written against the reference and pins the behaviour even
    if the default was turned off by an outer ``defer_resharding(False)``."""

    def __init__(self, enabled: bool = True):
        from .api import defer_resharding

        self._cm = defer_resharding(enabled)

    def __enter__(self):
        self._cm.__enter__()
        return self

    def __exit__(self, *exc):
        return self._cm.__exit__(*exc)

    @staticmethod
    def is_enabled() -> bool:
        from .rules import pointwise

        return bool(pointwise.DEFER_RESHARD[0])


EnablePartialMode = DeferReshardMode  # older spelling in the reference's docs
