"""Which CUDA stream a communication runs on, for timing it (legacy ``ndtimeline/stream.py:25-79``).

The reference patches ``ProcessGroupNCCL`` to expose NCCL's internal streams.  Here the hot collectives are our own kernels
launched on streams we own (``FSDPState.ag_stream`` / ``rs_stream``, or simply the current stream), so the answer is known
without touching c10d internals; for c10d NCCL calls issued with ``async_op=False`` the work is ordered with the current
stream, which is what these helpers return."""
from __future__ import annotations

from typing import Dict, Optional

import torch

__all__ = ["get_nccl_coll_stream", "get_nccl_p2p_stream", "register_comm_stream"]

_STREAMS: Dict[str, "torch.cuda.Stream"] = {}


def register_comm_stream(name: str, stream) -> None:
    """Name one of our communication streams (e.g. ``register_comm_stream("fsdp-ag", state.ag_stream)``)."""
    _STREAMS[name] = stream


def _lookup(name: str) -> Optional["torch.cuda.Stream"]:
    s = _STREAMS.get(name)
    if s is None and torch.cuda.is_available():
        s = torch.cuda.current_stream()
    return s


def get_nccl_coll_stream(name: str, nccl_pg=None, nccl_tensor: Optional[torch.Tensor] = None):
    return _lookup(name)


def get_nccl_p2p_stream(name: str, nccl_pg=None, peer=None, is_batched: bool = False):
    return _lookup(name)
