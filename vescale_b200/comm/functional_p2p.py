"""Functional point-to-point collectives (native piece N-C4 of SURVEY §2C: the reference adds ``funcol.send / recv / isend / irecv``
plus meta kernels to torch's ``_functional_collectives`` so pipeline stages can be expressed inside traced graphs,
``legacy/patches/patched_pytorch_v2.2.1_rc3.patch:1595-1738``).

Here they are ordinary ``torch.library`` custom ops — no patched torch:

    y   = recv(shape, dtype, src, group)          # a NEW tensor; nothing is mutated, so the op composes with tracing / export
    tok = send(x, dst, group)                     # returns a 0-element token that depends on ``x`` (keeps the send alive in a graph)

Both are differentiable: the backward of ``send`` receives the gradient of what was sent from ``dst`` and the backward of
``recv`` sends the incoming gradient back to ``src`` — a pipeline stage boundary written with these two calls back-propagates
across ranks by itself.  Meta ("fake") kernels give shapes without communicating.  ``isend`` / ``irecv`` return a waitable handle
whose ``wait()`` yields the functional result.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

__all__ = ["send", "recv", "isend", "irecv", "P2PHandle"]

_GROUPS = {}


def _gid(group) -> int:
    """Custom-op schemas take ints: groups are registered under an id."""
    if group is None:
        return 0
    k = id(group)
    _GROUPS[k] = group
    return k


def _group(gid: int):
    return None if gid == 0 else _GROUPS[gid]


def _global(rank: int, group) -> int:
    return dist.get_global_rank(group, rank) if group is not None else rank


@torch.library.custom_op("vescale_b200::p2p_send", mutates_args=())
def _send_op(x: torch.Tensor, dst: int, gid: int) -> torch.Tensor:
    g = _group(gid)
    dist.send(x.contiguous(), _global(dst, g), group=g)
    return x.new_empty(0)


@_send_op.register_fake
def _(x, dst, gid):
    return x.new_empty(0)


@torch.library.custom_op("vescale_b200::p2p_recv", mutates_args=())
def _recv_op(like: torch.Tensor, shape: Sequence[int], src: int, gid: int) -> torch.Tensor:
    g = _group(gid)
    out = like.new_empty(list(shape))
    dist.recv(out, _global(src, g), group=g)
    return out


@_recv_op.register_fake
def _(like, shape, src, gid):
    return like.new_empty(list(shape))


def _send_bwd(ctx, grad_tok):
    # the gradient of what was sent comes back from the receiver
    like = torch.empty(0, dtype=ctx.dtype, device=ctx.device)
    return _recv_op(like, ctx.shape, ctx.dst, ctx.gid), None, None


def _send_setup(ctx, inputs, output):
    x, dst, gid = inputs
    ctx.shape, ctx.dtype, ctx.device, ctx.dst, ctx.gid = list(x.shape), x.dtype, x.device, dst, gid


def _recv_bwd(ctx, grad):
    _send_op(grad.contiguous(), ctx.src, ctx.gid)
    return None, None, None, None


def _recv_setup(ctx, inputs, output):
    _, _, src, gid = inputs
    ctx.src, ctx.gid = src, gid


_send_op.register_autograd(_send_bwd, setup_context=_send_setup)
_recv_op.register_autograd(_recv_bwd, setup_context=_recv_setup)


def send(x: torch.Tensor, dst: int, group=None) -> torch.Tensor:
    """Blocking functional send of ``x`` to group rank ``dst``; returns a token tensor (0 elements) to thread through the graph."""
    return _send_op(x, int(dst), _gid(group))


def recv(shape: Sequence[int], dtype: torch.dtype, src: int, group=None, device=None, requires_grad: bool = False) -> torch.Tensor:
    """Functional receive: a new ``shape`` / ``dtype`` tensor filled by group rank ``src``.  With ``requires_grad`` the backward
    pass sends the gradient back to ``src``."""
    like = torch.empty(0, dtype=dtype, device=device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu"), requires_grad=requires_grad)
    return _recv_op(like, [int(s) for s in shape], int(src), _gid(group))


class P2PHandle:
    """Result of ``isend`` / ``irecv``: ``wait()`` returns what the blocking form would have returned."""

    def __init__(self, work, value: torch.Tensor, keep=None):
        self.work, self.value, self._keep = work, value, keep

    def wait(self) -> torch.Tensor:
        if self.work is not None:
            self.work.wait()
            self.work = None
        return self.value

    def is_completed(self) -> bool:
        return self.work is None or self.work.is_completed()


def isend(x: torch.Tensor, dst: int, group=None) -> P2PHandle:
    x = x.detach().contiguous()
    return P2PHandle(dist.isend(x, _global(dst, group), group=group), x.new_empty(0), keep=x)


def irecv(shape: Sequence[int], dtype: torch.dtype, src: int, group=None, device=None) -> P2PHandle:
    out = torch.empty([int(s) for s in shape], dtype=dtype, device=device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu"))
    return P2PHandle(dist.irecv(out, _global(src, group), group=group), out)
