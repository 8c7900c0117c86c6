"""Checkpoint logger: one named logger whose level comes from ``VESCALE_CHECKPOINT_LOGGING_LEVEL`` (name or number; default WARNING),
with timing helpers used around the save / load phases (legacy ``checkpoint/utilities/logger.py:175-251``)."""
from __future__ import annotations

import contextlib
import logging
import os
import time
from typing import Dict, Optional

__all__ = ["get_vescale_checkpoint_logger", "timed", "PHASE_SECONDS"]

_LOGGER: Optional[logging.Logger] = None
PHASE_SECONDS: Dict[str, float] = {}  # last duration of every timed phase (also what the report service publishes)


def _level_from_env() -> int:
    raw = os.environ.get("VESCALE_CHECKPOINT_LOGGING_LEVEL", "WARNING").strip()
    if raw.lstrip("-").isdigit():
        return int(raw)
    lv = logging.getLevelName(raw.upper())
    return lv if isinstance(lv, int) else logging.WARNING


def get_vescale_checkpoint_logger() -> logging.Logger:
    global _LOGGER
    if _LOGGER is None:
        lg = logging.getLogger("vescale_b200.checkpoint")
        if not lg.handlers:
            h = logging.StreamHandler()
            h.setFormatter(logging.Formatter("[%(asctime)s][%(levelname)s][checkpoint][rank " + os.environ.get("RANK", "0") + "] %(message)s"))
            lg.addHandler(h)
            lg.propagate = False
        _LOGGER = lg
    _LOGGER.setLevel(_level_from_env())
    return _LOGGER


@contextlib.contextmanager
def timed(phase: str, level: int = logging.INFO):
    """``with timed("d2h"):`` logs the phase's duration and records it in ``PHASE_SECONDS``."""
    t0 = time.perf_counter()
    try:
        yield
    finally:
        dt = time.perf_counter() - t0
        PHASE_SECONDS[phase] = dt
        get_vescale_checkpoint_logger().log(level, f"{phase}: {dt * 1e3:.1f} ms")


class VeScaleCheckpointLogger:
    """Holder form (legacy ``utilities/logger.py:229-251``): ``VeScaleCheckpointLogger().logger`` is the shared logger."""

    def __new__(cls):
        if not hasattr(cls, "_inst"):
            cls._inst = super().__new__(cls)
            cls._inst.logger = get_vescale_checkpoint_logger()
        return cls._inst
