"""Mesh collectives: the c10d (NCCL / Gloo) backend used for plumbing, CPU tests and as the measured
baseline.  The sm_100a symmetric-memory kernels in ``vescale_b200.comm.symm`` implement the same
operations fused with their adjacent compute; ``vescale_b200.dtensor.redistribute`` picks a backend.

Every call is reported to the active ``CommDebugMode`` / debug logger through ``_note``.

Parity: ``legacy/vescale/dtensor/_collective_utils.py:50-400`` (mesh_scatter, mesh_all_to_all(_single),
mesh_broadcast, mesh_reduce_scatter, mesh_all_gather, mesh_all_reduce), reference
``vescale/dtensor/_collective_utils.py:66-99`` (mesh_scatter_ragged) and ``placement_types.py:128-192``.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from ..utils.env import flag as _flag

__all__ = [
    "mesh_all_gather",
    "mesh_all_gather_uneven",
    "mesh_all_reduce",
    "mesh_reduce_scatter",
    "mesh_all_to_all_single",
    "mesh_all_to_all_uneven",
    "mesh_ragged_exchange",
    "mesh_broadcast",
    "mesh_scatter",
    "mesh_scatter_ragged",
    "add_comm_hook",
    "remove_comm_hook",
]

_HOOKS: List[Callable] = []


def add_comm_hook(fn: Callable) -> None:
    _HOOKS.append(fn)


def remove_comm_hook(fn: Callable) -> None:
    if fn in _HOOKS:
        _HOOKS.remove(fn)


def _note(name: str, tensor: Optional[torch.Tensor], group, **kw) -> None:
    if _HOOKS:
        nbytes = 0 if tensor is None else tensor.numel() * tensor.element_size()
        for h in list(_HOOKS):
            h(name, nbytes, group, kw)


_OPS = {
    "sum": dist.ReduceOp.SUM,
    "avg": dist.ReduceOp.AVG,
    "max": dist.ReduceOp.MAX,
    "min": dist.ReduceOp.MIN,
    "product": dist.ReduceOp.PRODUCT,
    "band": dist.ReduceOp.BAND,
    "bor": dist.ReduceOp.BOR,
    "bxor": dist.ReduceOp.BXOR,
}


def _backend(group) -> str:
    if group is _SOLO:
        return "solo"
    try:
        return dist.get_backend(group)
    except Exception:
        return "gloo"


def _group_size(group) -> int:
    return 1 if group is _SOLO else dist.get_world_size(group)


class _Solo:
    """Stand-in group of a size-1 mesh dim: every collective is the identity and needs no process group (fake / single
    process meshes built with ``_init_process_groups=False``)."""

    def __repr__(self):
        return "<solo group>"


_SOLO = _Solo()


def _group_of(mesh, mesh_dim):
    if mesh.size(mesh_dim) == 1:
        return _SOLO
    return mesh.get_group(mesh_dim)


def _group_rank(group) -> int:
    return 0 if group is _SOLO else dist.get_rank(group)


def _symm(group, tensor):
    """The symmetric-memory backend registered for this group (``enable_symmetric_collectives``), if the tensor qualifies."""
    if group is _SOLO or not tensor.is_cuda:
        return None
    from .symm_collectives import symm_backend_for

    return symm_backend_for(group, tensor)


def mesh_all_reduce(tensor: torch.Tensor, mesh, reduce_op: str = "sum", mesh_dim: int = 0, *, inplace: bool = False) -> torch.Tensor:
    group = _group_of(mesh, mesh_dim)
    n = _group_size(group)
    out = tensor if inplace else tensor.clone(memory_format=torch.contiguous_format)
    _note("all_reduce", out, group, op=reduce_op)
    if n == 1:
        return out
    if reduce_op in ("sum", "avg") and out.dtype in (torch.float32, torch.bfloat16) and out.is_contiguous():
        sc = _symm(group, out)
        if sc is not None:
            return sc.all_reduce(out, reduce_op)
    if reduce_op == "avg" and (_backend(group) != "nccl" or not out.is_floating_point()):
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out.div_(n) if out.is_floating_point() else out.div_(n, rounding_mode="floor")
    if out.dtype == torch.bool:
        tmp = out.to(torch.uint8)
        dist.all_reduce(tmp, op=_OPS[reduce_op], group=group)
        return out.copy_(tmp != 0)
    dist.all_reduce(out, op=_OPS[reduce_op], group=group)
    return out


def mesh_all_gather(tensor: torch.Tensor, mesh, mesh_dim: int = 0, gather_dim: int = 0) -> torch.Tensor:
    """Even all-gather: every rank contributes the same shape; result concatenated on ``gather_dim``."""
    group = _group_of(mesh, mesh_dim)
    n = _group_size(group)
    tensor = tensor.contiguous()
    _note("all_gather", tensor, group)
    if n == 1:
        return tensor.clone()
    out = tensor.new_empty((n * tensor.shape[0], *tensor.shape[1:])) if tensor.ndim else tensor.new_empty((n,))
    src = tensor if tensor.ndim else tensor.reshape(1)
    dist.all_gather_into_tensor(out, src, group=group)
    if tensor.ndim == 0:
        return out
    if gather_dim != 0:
        chunks = out.chunk(n, dim=0)
        out = torch.cat(chunks, dim=gather_dim)
    return out


def mesh_all_gather_uneven(tensor: torch.Tensor, sizes: Sequence[int], mesh, mesh_dim: int = 0) -> List[torch.Tensor]:
    """All-gather of 1-D pieces with per-rank lengths ``sizes`` (zeros allowed).  Returns the list.
    NCCL takes the list form directly; elsewhere pieces are padded to the max length."""
    group = _group_of(mesh, mesh_dim)
    n = _group_size(group)
    flat = tensor.reshape(-1).contiguous()
    _note("all_gather", flat, group, uneven=True)
    if n == 1:
        return [flat.clone()]
    # One even all_gather_into_tensor of max-size padded pieces on every backend: the list form of
    # dist.all_gather degrades to per-rank broadcasts for uneven sizes (and zero-size entries are fragile).
    mx = max(int(s) for s in sizes)
    if mx == 0:
        return [flat.new_empty(0) for _ in sizes]
    padded = flat.new_zeros(mx)
    padded[: flat.numel()] = flat
    buf = flat.new_empty(n * mx)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return [buf[i * mx : i * mx + int(sizes[i])] for i in range(n)]


def mesh_reduce_scatter(tensor: torch.Tensor, mesh, reduce_op: str = "sum", mesh_dim: int = 0, scatter_dim: int = 0) -> torch.Tensor:
    """Even reduce-scatter along ``scatter_dim`` (size divisible by the group size)."""
    group = _group_of(mesh, mesh_dim)
    n = _group_size(group)
    _note("reduce_scatter", tensor, group, op=reduce_op)
    if n == 1:
        return tensor.clone()
    if tensor.shape[scatter_dim] % n != 0:
        raise ValueError(f"reduce_scatter dim {scatter_dim} of size {tensor.shape[scatter_dim]} not divisible by {n}")
    if scatter_dim != 0:
        tensor = torch.cat(tensor.chunk(n, dim=scatter_dim), dim=0)
    tensor = tensor.contiguous()
    out_shape = (tensor.shape[0] // n, *tensor.shape[1:])
    if reduce_op in ("sum", "avg") and tensor.dtype in (torch.float32, torch.bfloat16) and _flag("VESCALE_B200_SYMM_RS"):
        sc = _symm(group, tensor)
        if sc is not None and (math.prod(out_shape[1:]) * tensor.element_size()) % 16 == 0:
            return sc.reduce_scatter(tensor, reduce_op)
    if _backend(group) == "nccl":
        out = tensor.new_empty(out_shape)
        op = _OPS[reduce_op]
        dist.reduce_scatter_tensor(out, tensor, op=op, group=group)
        return out
    red = mesh_all_reduce(tensor, mesh, reduce_op, mesh_dim)
    r = _group_rank(group)
    return red.narrow(0, r * out_shape[0], out_shape[0]).clone()


def _p2p_all_to_all(outs: List[torch.Tensor], ins: List[torch.Tensor], group) -> None:
    """all_to_all for backends without a native one (Gloo): ordered isend/irecv pairs."""
    n, me = _group_size(group), _group_rank(group)
    outs[me].copy_(ins[me])
    reqs = []
    for step in range(1, n):
        dst = (me + step) % n
        src = (me - step) % n
        if ins[dst].numel():
            reqs.append(dist.isend(ins[dst].contiguous(), dist.get_global_rank(group, dst), group=group))
        if outs[src].numel():
            reqs.append(dist.irecv(outs[src], dist.get_global_rank(group, src), group=group))
    for r in reqs:
        r.wait()


def mesh_all_to_all_uneven(ins: List[torch.Tensor], out_sizes: Sequence[int], mesh, mesh_dim: int = 0) -> List[torch.Tensor]:
    """List all-to-all of 1-D pieces: send ``ins[j]`` to coordinate ``j``; receive ``out_sizes[j]`` elements from it."""
    group = _group_of(mesh, mesh_dim)
    ref = ins[0]
    outs = [ref.new_empty(int(s)) for s in out_sizes]
    _note("all_to_all", torch.cat([i.reshape(-1) for i in ins]) if ins else None, group, uneven=True)
    if _group_size(group) == 1:
        outs[0].copy_(ins[0])
        return outs
    if _backend(group) == "nccl":
        dist.all_to_all(outs, [i.contiguous() for i in ins], group=group)
    else:
        _p2p_all_to_all(outs, ins, group)
    return outs


def mesh_ragged_exchange(local: torch.Tensor, src_ranges: Sequence, dst_ranges: Sequence, mesh, mesh_dim: int = 0) -> torch.Tensor:
    """Flat interval exchange: ``local`` holds flat range ``src_ranges[me]`` of a global buffer, the result holds
    ``dst_ranges[me]`` (RaggedShard -> RaggedShard; one-hot destinations = gather-to-root).  Symmetric-memory put kernel
    when enabled, else the uneven list all-to-all of the reference (``placement_types.py:152-192``)."""
    group = _group_of(mesh, mesh_dim)
    n, me = _group_size(group), mesh.get_local_rank(mesh_dim)
    local = local.contiguous().view(-1)
    sc = _symm(group, local) if n > 1 else None
    if sc is not None:
        _note("all_to_all", local, group, uneven=True, backend="symm")
        return sc.ragged_exchange(local, src_ranges, dst_ranges)
    s_lo, s_hi = src_ranges[me]
    d_lo, d_hi = dst_ranges[me]
    ins, out_sizes = [], []
    for j in range(n):
        lo, hi = max(s_lo, dst_ranges[j][0]), min(s_hi, dst_ranges[j][1])
        ins.append(local.narrow(0, lo - s_lo, hi - lo) if hi > lo else local.new_empty(0))
        out_sizes.append(max(0, min(src_ranges[j][1], d_hi) - max(src_ranges[j][0], d_lo)))
    outs = mesh_all_to_all_uneven(ins, out_sizes, mesh, mesh_dim)
    return torch.cat(outs) if outs else local.new_empty(0)


def mesh_all_to_all_single(tensor: torch.Tensor, mesh, mesh_dim: int, split_dim: int, concat_dim: int) -> torch.Tensor:
    """Even Shard(concat_dim) -> Shard(split_dim): split my tensor on ``split_dim`` into n pieces, piece j
    goes to coordinate j, received pieces are concatenated on ``concat_dim``."""
    group = _group_of(mesh, mesh_dim)
    n = _group_size(group)
    _note("all_to_all", tensor, group)
    if n == 1:
        return tensor.clone()
    if tensor.shape[split_dim] % n != 0:
        raise ValueError("all_to_all_single needs an evenly divisible split dim")
    if tensor.element_size() >= 2:
        sc = _symm(group, tensor)
        if sc is not None:
            return sc.all_to_all_permute(tensor, concat_dim, split_dim)
    pieces = [p.contiguous() for p in tensor.chunk(n, dim=split_dim)]
    if _backend(group) == "nccl":
        send = torch.stack(pieces, 0)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group)
        outs = list(recv.unbind(0))
    else:
        outs = [torch.empty_like(p) for p in pieces]
        _p2p_all_to_all(outs, pieces, group)
    return torch.cat(outs, dim=concat_dim)


def mesh_broadcast(tensor: torch.Tensor, mesh, mesh_dim: int = 0, group_src: int = 0) -> torch.Tensor:
    group = _group_of(mesh, mesh_dim)
    _note("broadcast", tensor, group)
    if _group_size(group) == 1:
        return tensor
    dist.broadcast(tensor, src=dist.get_global_rank(group, group_src), group=group)
    return tensor


def mesh_scatter(output: torch.Tensor, scatter_list: Optional[List[torch.Tensor]], mesh, mesh_dim: int = 0, group_src: int = 0) -> torch.Tensor:
    """Even scatter from coordinate ``group_src``."""
    group = _group_of(mesh, mesh_dim)
    n, me = _group_size(group), _group_rank(group)
    _note("scatter", output, group)
    if n == 1:
        output.copy_(scatter_list[0])
        return output
    src = dist.get_global_rank(group, group_src)
    if me == group_src:
        dist.scatter(output, scatter_list=[t.contiguous() for t in scatter_list], src=src, group=group)
    else:
        dist.scatter(output, scatter_list=None, src=src, group=group)
    return output


def mesh_scatter_ragged(output: torch.Tensor, scatter_list: Optional[List[torch.Tensor]], mesh, mesh_dim: int = 0, group_src: int = 0) -> torch.Tensor:
    """Uneven scatter.  The reference serialises blocking sends (``_collective_utils.py:83-94``, with a TODO);
    here all sends are posted at once and waited together."""
    group = _group_of(mesh, mesh_dim)
    n, me = _group_size(group), _group_rank(group)
    _note("scatter", output, group, uneven=True)
    if me == group_src:
        reqs = []
        for j in range(n):
            if j == me:
                output.copy_(scatter_list[j])
            elif scatter_list[j].numel():
                reqs.append(dist.isend(scatter_list[j].contiguous(), dist.get_global_rank(group, j), group=group))
        for r in reqs:
            r.wait()
    elif output.numel():
        dist.recv(output, dist.get_global_rank(group, group_src), group=group)
    return output
