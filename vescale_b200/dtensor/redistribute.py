"""Redistribute: turn a local shard laid out as ``cur_spec`` into one laid out as ``tgt_spec``.

Planner + executor.  The planner works on *logical shard chains*: for every tensor dim, the ordered
list of mesh dims that cut it (plain shards outer-to-inner in mesh order, strided shards after the
later mesh dims).  Chains of source and target are compared; the differing suffix of the source chain
is un-sharded innermost-first, the differing suffix of the target chain is sharded outermost-first.
Partial→Shard becomes a reduce-scatter and Shard(i)→Shard(j) an all-to-all when the chain allows it;
everything else decomposes into all-gather / all-reduce / local chunk.

RaggedShard (reference ``vescale/dtensor/_redistribute.py:48-127``): ragged→ragged on one mesh dim is an
uneven all-to-all by interval intersection; ragged→X first un-rags with an uneven all-gather; X→ragged
replicates that mesh dim then slices the flat storage (Partial sources are reduced first).

Parity (non-ragged): ``legacy/vescale/dtensor/redistribute.py:35-455`` transition table.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import torch

from ..comm import collectives as C
from ..layout import _shard_order, dim_intervals, get_ragged_shard, shape_and_offset_before_ragged
from ..placement import InterleavedShard, Partial, Placement, RaggedShard, Replicate, Shard, _StridedShard, shard_size_and_offset
from ..spec import DTensorSpec

__all__ = ["redistribute_local_tensor", "Redistribute", "redistribute_cost"]


# --------------------------------------------------------------------------- primitive transitions
def _gather_shard(local: torch.Tensor, mesh, i: int, d: int, full: int) -> torch.Tensor:
    """S(d)->R on mesh dim i, where the un-sharded size of dim d is ``full`` (pads uneven pieces)."""
    n = mesh.size(i)
    if n == 1:
        return local
    chunk = (full + n - 1) // n
    if full % n == 0:
        return C.mesh_all_gather(local, mesh, i, gather_dim=d)
    # uneven: pad my piece to the full chunk, gather, drop the padding
    pad = chunk - local.shape[d]
    if pad:
        shp = list(local.shape)
        shp[d] = pad
        local = torch.cat([local, local.new_zeros(shp)], dim=d)
    out = C.mesh_all_gather(local, mesh, i, gather_dim=d)
    return out.narrow(d, 0, full).contiguous()  # views downstream assume the contiguous strides of the global metadata


def _gather_interleaved(local: torch.Tensor, mesh, i: int, p: InterleavedShard) -> torch.Tensor:
    n, d, k = mesh.size(i), p.dim, p.interleaved_size
    if n == 1:
        return local
    g = C.mesh_all_gather(local, mesh, i, gather_dim=0)  # [n*L0, ...] pieces stacked on dim 0
    shp = list(local.shape)
    g = g.reshape(n, *shp)  # [n, ..., k*piece, ...]
    piece = shp[d] // k
    g = g.reshape(n, *shp[:d], k, piece, *shp[d + 1 :])  # n at 0, k at d+1, piece at d+2
    g = g.movedim(0, d + 1)  # [..., k, n, piece, ...]
    return g.reshape(*shp[:d], k * n * piece, *shp[d + 1 :]).contiguous()


def _chunk_shard(local: torch.Tensor, mesh, i: int, p: Shard, coord) -> torch.Tensor:
    """R->S(d) on mesh dim i: keep my piece of the current local dim."""
    n = mesh.size(i)
    if isinstance(p, InterleavedShard):
        pieces, _ = p.split_tensor(local, n)
        return pieces[coord[i]]
    ln, off = shard_size_and_offset(local.shape[p.dim], n, coord[i])
    return local.narrow(p.dim, off, ln).clone(memory_format=torch.contiguous_format)


def _reduce_scatter_shard(local: torch.Tensor, mesh, i: int, d: int, op: str, coord) -> torch.Tensor:
    n = mesh.size(i)
    if n == 1:
        return local
    full = local.shape[d]
    if full % n == 0:
        return C.mesh_reduce_scatter(local, mesh, op, i, scatter_dim=d)
    chunk = (full + n - 1) // n
    shp = list(local.shape)
    shp[d] = chunk * n - full
    padded = torch.cat([local, local.new_zeros(shp)], dim=d)
    out = C.mesh_reduce_scatter(padded, mesh, op, i, scatter_dim=d)
    ln, _ = shard_size_and_offset(full, n, coord[i])
    return out.narrow(d, 0, ln).contiguous()


def _reduce_partial(local: torch.Tensor, mesh, i: int, p: Partial) -> torch.Tensor:
    nt = p.norm_type
    if nt is None:
        return C.mesh_all_reduce(local, mesh, p.reduce_op, i)
    if math.isinf(nt):
        return C.mesh_all_reduce(local, mesh, "max" if nt > 0 else "min", i)
    if nt == 0:
        return C.mesh_all_reduce(local, mesh, "sum", i)
    if nt == 1:
        return C.mesh_all_reduce(local, mesh, "sum", i)
    acc = C.mesh_all_reduce(local.pow(nt), mesh, "sum", i)
    return acc.pow_(1.0 / nt)


def _replicate_to_partial(local: torch.Tensor, mesh, i: int, p: Partial, coord) -> torch.Tensor:
    if p.reduce_op == "sum":
        return local.clone() if coord[i] == 0 else torch.zeros_like(local)
    if p.reduce_op in ("avg", "max", "min", "band", "bor"):
        return local
    if p.norm_type is not None:
        return local.clone() if coord[i] == 0 else torch.zeros_like(local)
    raise NotImplementedError(f"Replicate -> {p}")


# --------------------------------------------------------------------------- ragged transitions
def _ragged_to_replicate(local: torch.Tensor, mesh, i: int, rp: RaggedShard, before_shape: Sequence[int]) -> torch.Tensor:
    numel = math.prod(before_shape)
    sizes = [rp.flat_range(numel, j)[1] - rp.flat_range(numel, j)[0] for j in range(mesh.size(i))]
    outs = C.mesh_all_gather_uneven(local, sizes, mesh, i)
    return torch.cat(outs).view(tuple(before_shape))


def _ragged_to_ragged(local: torch.Tensor, mesh, i: int, src: RaggedShard, dst: RaggedShard, numel: int, coord) -> torch.Tensor:
    """Uneven all-to-all by intersecting source and destination flat intervals
    (reference ``placement_types.py:152-192``)."""
    n = mesh.size(i)
    return C.mesh_ragged_exchange(local, [src.flat_range(numel, j) for j in range(n)], [dst.flat_range(numel, j) for j in range(n)], mesh, i)


# --------------------------------------------------------------------------- planner / executor
def _same_shard(a: Placement, b: Placement) -> bool:
    return type(a) is type(b) and a == b


def redistribute_local_tensor(local: torch.Tensor, cur: DTensorSpec, tgt: DTensorSpec, *, async_op: bool = False) -> torch.Tensor:
    if cur.mesh != tgt.mesh:
        raise NotImplementedError("cross-mesh redistribute: use vescale_b200.dtensor.cross_mesh_redistribute")
    mesh = cur.mesh
    coord = mesh.get_coordinate()
    if coord is None:
        return local
    cur_p, tgt_p = list(cur.placements), list(tgt.placements)
    if cur_p == tgt_p:
        return local
    gshape = tuple(cur.shape)
    mshape = tuple(mesh.shape)

    # ---- ragged handling -------------------------------------------------------------------------
    c_idx, c_rp = get_ragged_shard(cur_p) if any(isinstance(p, RaggedShard) for p in cur_p) else (None, None)
    t_idx, t_rp = get_ragged_shard(tgt_p) if any(isinstance(p, RaggedShard) for p in tgt_p) else (None, None)
    if c_rp is not None and t_rp is not None and c_idx == t_idx:
        rest_same = all(a == b for k, (a, b) in enumerate(zip(cur_p, tgt_p)) if k != c_idx)
        if rest_same:
            before, _ = shape_and_offset_before_ragged(gshape, mshape, tuple(cur_p), coord)
            return _ragged_to_ragged(local, mesh, c_idx, c_rp, t_rp, math.prod(before), coord)
    if c_rp is not None:
        before, _ = shape_and_offset_before_ragged(gshape, mshape, tuple(cur_p), coord)
        local = _ragged_to_replicate(local, mesh, c_idx, c_rp, before)
        cur_p[c_idx] = Replicate()
    tgt_nonrag = list(tgt_p)
    if t_rp is not None:
        tgt_nonrag[t_idx] = Replicate()

    local = _redistribute_plain(local, mesh, coord, gshape, cur_p, tgt_nonrag)

    if t_rp is not None:
        before, _ = shape_and_offset_before_ragged(gshape, mshape, tuple(tgt_p), coord)
        if tuple(local.shape) != tuple(before):
            raise RuntimeError(f"internal: shape before ragged {tuple(local.shape)} != {before}")
        lo, hi = t_rp.flat_range(math.prod(before), coord[t_idx])
        local = local.contiguous().view(-1).narrow(0, lo, hi - lo).clone()
    return local


def _redistribute_plain(local, mesh, coord, gshape, cur_p: List[Placement], tgt_p: List[Placement]) -> torch.Tensor:
    if cur_p == tgt_p:
        return local
    mshape = tuple(mesh.shape)
    ndim = len(gshape)
    to_unshard: List[List[int]] = []
    to_shard: List[List[int]] = []
    for d in range(ndim):
        cc, tc = _shard_order(cur_p, d), _shard_order(tgt_p, d)
        k = 0
        while k < len(cc) and k < len(tc) and cc[k] == tc[k] and _same_shard(cur_p[cc[k]], tgt_p[tc[k]]):
            k += 1
        to_unshard.append(cc[k:][::-1])
        to_shard.append(tc[k:])

    def full_size(d: int, without: int) -> int:
        tmp = list(cur_p)
        tmp[without] = Replicate()
        return sum(x[1] for x in dim_intervals(gshape[d], d, mshape, tuple(tmp), coord))

    # 1. partials
    for i, p in enumerate(list(cur_p)):
        if not p.is_partial():
            continue
        t = tgt_p[i]
        if t == p:
            continue
        if (
            isinstance(t, Shard)
            and not isinstance(t, (InterleavedShard,))
            and p.norm_type is None
            and to_shard[t.dim][:1] == [i]
            and not to_unshard[t.dim]
            and not any(j != i and isinstance(q, Shard) and q.dim == t.dim and j in to_shard[t.dim] for j, q in enumerate(tgt_p))
        ):
            local = _reduce_scatter_shard(local, mesh, i, t.dim, p.reduce_op, coord)
            cur_p[i] = t
            to_shard[t.dim].pop(0)
        else:
            local = _reduce_partial(local, mesh, i, p)
            cur_p[i] = Replicate()

    # 2. un-shard (innermost first per chain); Shard(d)->Shard(d2) as all-to-all when legal
    for d in range(ndim):
        for i in to_unshard[d]:
            p = cur_p[i]
            t = tgt_p[i]
            n = mesh.size(i)
            if isinstance(p, InterleavedShard):
                local = _gather_interleaved(local, mesh, i, p)
                cur_p[i] = Replicate()
                continue
            full = full_size(d, i)
            if (
                isinstance(t, Shard)
                and not isinstance(t, (InterleavedShard, _StridedShard))
                and not isinstance(p, _StridedShard)
                and t.dim != d
                and to_shard[t.dim] == [i]
                and not to_unshard[t.dim]
                and full % n == 0
                and local.shape[t.dim] % n == 0
            ):
                local = C.mesh_all_to_all_single(local, mesh, i, split_dim=t.dim, concat_dim=d)
                cur_p[i] = t
                to_shard[t.dim] = []
            else:
                local = _gather_shard(local, mesh, i, d, full)
                cur_p[i] = Replicate()

    # 3. shard (outermost first per chain)
    for d in range(ndim):
        for i in to_shard[d]:
            if not cur_p[i].is_replicate():
                raise RuntimeError(f"internal: mesh dim {i} should be replicated before sharding, got {cur_p[i]}")
            local = _chunk_shard(local, mesh, i, tgt_p[i], coord)
            cur_p[i] = tgt_p[i]

    # 4. replicate -> partial
    for i, t in enumerate(tgt_p):
        if t.is_partial() and cur_p[i].is_replicate():
            local = _replicate_to_partial(local, mesh, i, t, coord)
            cur_p[i] = t
    if cur_p != tgt_p:
        raise RuntimeError(f"redistribute could not reach {tgt_p} (stopped at {cur_p})")
    return local


def redistribute_cost(cur: DTensorSpec, tgt: DTensorSpec) -> float:
    """Relative bytes-moved cost used to rank sharding strategies
    (intra-node factor as in ``legacy/vescale/dtensor/_collective_utils.py:406-507``)."""
    if cur.placements == tgt.placements:
        return 0.0
    mesh = cur.mesh
    nbytes = math.prod(cur.shape) * cur.dtype.itemsize / max(cur.num_shards, 1)
    cost = 0.0
    for i, (c, t) in enumerate(zip(cur.placements, tgt.placements)):
        if c == t:
            continue
        n = mesh.size(i)
        if n == 1:
            continue
        if isinstance(c, (Shard, RaggedShard)) and t.is_replicate():
            nbytes *= n
            cost += nbytes * (n - 1) / n + 1.0
        elif isinstance(c, (Shard, RaggedShard)) and isinstance(t, (Shard, RaggedShard)):
            cost += nbytes * (n - 1) / n + 2.0
        elif c.is_partial() and t.is_replicate():
            cost += 2 * nbytes * (n - 1) / n + 1.0
        elif c.is_partial() and isinstance(t, (Shard, RaggedShard)):
            cost += nbytes * (n - 1) / n + 1.0
            nbytes /= n
        elif c.is_replicate() and isinstance(t, (Shard, RaggedShard)):
            nbytes /= n
            cost += 0.01
        elif c.is_replicate() and t.is_partial():
            cost += 0.01
        elif isinstance(c, (Shard, RaggedShard)) and t.is_partial():
            nbytes *= n
            cost += nbytes * (n - 1) / n + 1.0
        else:
            cost += nbytes
    return cost


class Redistribute(torch.autograd.Function):
    """Differentiable redistribute.  Backward redistributes the gradient from the target layout back to
    the source layout; Partial layouts are normalised to Replicate for gradients
    (reference ``_redistribute.py:160-203``)."""

    @staticmethod
    def forward(ctx, dt, mesh, placements, async_op: bool = False):
        from .api import DTensor

        cur = dt._spec
        ctx.cur_spec = cur
        tgt = DTensorSpec(mesh, tuple(placements), cur.tensor_meta)
        ctx.tgt_spec = tgt
        if cur.placements == tgt.placements:
            out_local = dt._local_tensor
        else:
            out_local = redistribute_local_tensor(dt._local_tensor, cur, tgt, async_op=async_op)
        return DTensor(out_local, tgt, requires_grad=dt.requires_grad)

    @staticmethod
    def backward(ctx, grad):
        from .api import DTensor

        prev, now = ctx.cur_spec, ctx.tgt_spec
        g_cur = grad._spec
        # gradient layouts never carry Partial on the way back
        tgt_pl = tuple(Replicate() if p.is_partial() else p for p in prev.placements)
        g_tgt = DTensorSpec(prev.mesh, tgt_pl, g_cur.tensor_meta)
        # the incoming gradient keeps its own layout: a Partial gradient (e.g. from a column-parallel GEMM's
        # dgrad) is reduced here — Partial -> Shard is the reduce-scatter of Megatron SP
        local = redistribute_local_tensor(grad._local_tensor, g_cur, g_tgt)
        return DTensor(local, g_tgt, requires_grad=grad.requires_grad), None, None, None


def __getattr__(name):  # ``CrossMeshRedistribute`` lives with the cross-mesh transport (import cycle otherwise)
    if name == "CrossMeshRedistribute":
        from .cross_mesh import CrossMeshRedistribute

        return CrossMeshRedistribute
    raise AttributeError(name)


def substitute_ragged_spec(spec: DTensorSpec) -> DTensorSpec:
    """The spec with every ragged placement replaced by ``Replicate`` — the layout a ragged tensor is gathered to before an op
    without a ragged rule runs (``vescale/dtensor/_redistribute.py:44-45``)."""
    from ..layout import substitute_ragged_with_replicate

    return DTensorSpec(spec.mesh, substitute_ragged_with_replicate(spec.placements), spec.tensor_meta)
