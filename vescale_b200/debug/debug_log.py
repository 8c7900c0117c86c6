"""DebugLogger: log every mesh collective (with the user-code frame that triggered it) and every dispatched
DTensor op (op, input specs, output specs).  Switch on with ``VESCALE_DEBUG_MODE=1`` or
``set_vescale_debug_mode(True, rank_to_print=(0,), logger=...)``.
Parity: ``legacy/vescale/debug/debug_log.py:40-361``."""
from __future__ import annotations

import logging
import os
import traceback
from typing import Optional, Sequence

import torch.distributed as dist

from ..comm import collectives as C

__all__ = ["DebugLogger", "set_vescale_debug_mode"]


class DebugLogger:
    """Logs every mesh collective and every dispatched DTensor op (legacy ``debug/debug_log.py:40-361``)."""
    enabled = False
    ranks: Optional[Sequence[int]] = None
    logger: Optional[logging.Logger] = None
    records = []
    _installed = False

    @classmethod
    def _rank_ok(cls) -> bool:
        r = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        return cls.ranks is None or r in cls.ranks or -1 in cls.ranks

    @classmethod
    def _emit(cls, msg: str) -> None:
        cls.records.append(msg)
        if cls.logger is not None:
            cls.logger.info(msg)
        else:
            print(msg, flush=True)

    @classmethod
    def _user_frame(cls) -> str:
        for fr in reversed(traceback.extract_stack()[:-3]):
            if "vescale_b200" not in fr.filename and "torch/" not in fr.filename:
                return f"{os.path.basename(fr.filename)}:{fr.lineno} in {fr.name}"
        return "?"

    @classmethod
    def _comm_hook(cls, name, nbytes, group, kw):
        if cls.enabled and cls._rank_ok():
            cls._emit(f"[vescale_b200][comm] {name} bytes={nbytes} {kw} @ {cls._user_frame()}")

    @classmethod
    def _op_hook(cls, op, schema, out_sh):
        if cls.enabled and cls._rank_ok():
            ins = [str(s) for s in schema.tensor_specs()]
            redis = None if out_sh.redistribute_specs is None else [None if s is None else str(s) for s in out_sh.redistribute_specs]
            cls._emit(f"[vescale_b200][op] {op} in={ins} out={out_sh.output_spec} redistribute={redis}")

    @classmethod
    def install(cls):
        if cls._installed:
            return
        from ..dtensor.dispatch import dispatcher

        C.add_comm_hook(cls._comm_hook)
        dispatcher._hooks.append(cls._op_hook)
        cls._installed = True


def set_vescale_debug_mode(on: bool = True, *, rank_to_print: Optional[Sequence[int]] = None, logger: Optional[logging.Logger] = None) -> None:
    DebugLogger.enabled = bool(on)
    DebugLogger.ranks = rank_to_print
    DebugLogger.logger = logger
    if on:
        DebugLogger.install()


if os.environ.get("VESCALE_DEBUG_MODE", "") not in ("", "0"):
    set_vescale_debug_mode(True)
