"""The subsystem's logger: one process-wide ``logging.Logger`` named "ndtimeline" whose level comes from
``VESCALE_NDTIMELINE_LOG_LEVEL`` (default INFO; unknown names fall back to WARNING) — legacy ``ndtimeline/logger.py``."""
import logging
import os
import sys

_LOGGER = None


def get_logger() -> logging.Logger:
    global _LOGGER
    if _LOGGER is None:
        name = os.getenv("VESCALE_NDTIMELINE_LOG_LEVEL", "INFO").upper()
        level = getattr(logging, name, None)
        lg = logging.getLogger("ndtimeline")
        h = logging.StreamHandler(stream=sys.stderr)
        h.setFormatter(logging.Formatter("[%(asctime)s][%(levelname)s][%(filename)s:%(lineno)d][pid:%(process)d] - %(message)s", datefmt="%Y-%m-%d %H:%M:%S"))
        lg.addHandler(h)
        lg.setLevel(level if isinstance(level, int) else logging.WARNING)
        lg.propagate = False
        _LOGGER = lg
    return _LOGGER


class NDTimelineLogger:
    """``NDTimelineLogger()`` returns the shared logger."""

    def __new__(cls):
        return get_logger()
