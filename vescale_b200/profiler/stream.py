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
    """Stream a collective named ``name`` runs on.  Resolution order: a stream registered under ``name``; a stream registered for
    the process group (``register_comm_stream(pg.group_name, s)`` — FSDP registers its all-gather / reduce-scatter streams as
    ``"fsdp-ag" / "fsdp-rs"`` and per group); the current stream of the tensor's device (c10d collectives issued with
    ``async_op=False`` are ordered with it)."""
    s = _STREAMS.get(name)
    if s is None and nccl_pg is not None:
        s = _STREAMS.get(getattr(nccl_pg, "group_name", None) or str(id(nccl_pg)))
    if s is None and nccl_tensor is not None and nccl_tensor.is_cuda:
        s = torch.cuda.current_stream(nccl_tensor.device)
    return s if s is not None else _lookup(name)


def get_nccl_p2p_stream(name: str, nccl_pg=None, peer=None, is_batched: bool = False):
    """Stream of a pipeline p2p operation: per (group, peer) registration first (``register_comm_stream(f"p2p:{peer}", s)``), then
    the group's, then the current stream (torch's NCCL p2p kernels are enqueued behind the current stream's work)."""
    s = _STREAMS.get(name)
    if s is None and peer is not None:
        s = _STREAMS.get(f"p2p:{peer}")
    if s is None and nccl_pg is not None:
        s = _STREAMS.get(getattr(nccl_pg, "group_name", None) or str(id(nccl_pg)))
    return s if s is not None else _lookup(name)
