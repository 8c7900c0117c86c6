"""Loader for the in-tree sm_100a extension (``vescale_b200/_C.so``, built by ``__graft_entry__.build()``
or ``python csrc/build.py``).  Ops are registered with TORCH_LIBRARY under ``torch.ops.vescale_b200``.

Policy: on a machine with a CUDA device the extension MUST load — a missing/broken build raises instead of
silently falling back to eager PyTorch.  On CPU-only machines (unit tests, planners) the pure-PyTorch
reference implementations in ``vescale_b200.ops.reference`` are used.
"""
from __future__ import annotations

import glob
import os
import threading

import torch

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LOCK = threading.Lock()
_STATE = {"loaded": False, "error": None, "path": None}


def so_path() -> str | None:
    cands = sorted(glob.glob(os.path.join(_PKG_DIR, "_C*.so")))
    return cands[0] if cands else None


def load(required: bool | None = None) -> bool:
    """Load the extension once.  ``required`` defaults to "a CUDA device is present"."""
    if _STATE["loaded"]:
        return True
    if required is None:
        required = torch.cuda.is_available() and os.environ.get("VESCALE_B200_ALLOW_FALLBACK", "0") != "1"
    with _LOCK:
        if _STATE["loaded"]:
            return True
        p = so_path()
        if p is None:
            _STATE["error"] = "vescale_b200/_C.so not found — run `python -c 'import __graft_entry__ as g; g.build()'`"
        else:
            try:
                torch.ops.load_library(p)
                _STATE.update(loaded=True, path=p, error=None)
                return True
            except Exception as e:  # noqa: BLE001
                _STATE["error"] = f"failed to load {p}: {e}"
        if required:
            raise RuntimeError(f"[vescale_b200] native sm_100a extension is required on a GPU box but unavailable: {_STATE['error']}")
        return False


def available() -> bool:
    """True iff kernels can run here: extension loaded AND a CUDA device present."""
    if not torch.cuda.is_available():
        return False
    return load()


def ops():
    load(required=True)
    return torch.ops.vescale_b200


LAUNCH_COUNTER = {"n": 0, "enabled": False, "by_op": {}}


def count_launch(name: str, n: int = 1) -> None:
    if LAUNCH_COUNTER["enabled"]:
        LAUNCH_COUNTER["n"] += n
        LAUNCH_COUNTER["by_op"][name] = LAUNCH_COUNTER["by_op"].get(name, 0) + n
