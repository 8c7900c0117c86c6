"""Custom op handlers: ops whose distributed semantics are not "run the aten op on the shards".

Parity: reference ``vescale/dtensor/_dispatch.py`` — fused_adamw_sgd_op_handler:118-132,
found_inf_reduce_handler:60-96, ragged_norm_op_handler:154-244 — and legacy
``dtensor/_dispatch_bypass.py:26-169`` (is_same_size, _local_scalar_dense, equal, nonzero, _to_copy).
"""
from __future__ import annotations

import math

import torch

from ..comm import collectives as C
from ..layout import shape_and_offset_before_ragged
from ..placement import Partial, RaggedShard, Replicate, Shard
from ..spec import DTensorSpec, TensorMeta, contiguous_stride
from .dispatch import dispatcher, register_op_handler

aten = torch.ops.aten


def _local(x):
    from .api import DTensor

    if isinstance(x, DTensor):
        return x._local_tensor
    if isinstance(x, (list, tuple)):
        return type(x)(_local(i) for i in x) if isinstance(x, tuple) else [_local(i) for i in x]
    return x


# ------------------------------------------------------------------------------- fused optimizers
_FUSED = [getattr(aten, n) for n in ("_fused_adamw_", "_fused_adam_", "_fused_sgd_", "_fused_adagrad_") if hasattr(aten, n)]


@register_op_handler(_FUSED)
def fused_optimizer_handler(op, args, kwargs):
    """Multi-tensor optimizer step: params/grads/states are identically laid out shards, so one local
    launch over the unwrapped lists is the whole step (no DTensor dispatch per tensor)."""
    from .api import DTensor

    lists = [a for a in args if isinstance(a, (list, tuple))]
    if lists:
        n = len(lists[0])
        for k in range(n):
            pls = {l[k].placements for l in lists if k < len(l) and isinstance(l[k], DTensor) and l[k].ndim > 0}
            if len(pls) > 1:
                raise RuntimeError(f"{op}: tensor {k} has mismatching placements across param/grad/state lists: {pls}")
    return op(*[_local(a) for a in args], **{k: _local(v) for k, v in kwargs.items()})


# ------------------------------------------------------------------------------- amp found-inf
@register_op_handler([aten._amp_foreach_non_finite_check_and_unscale_.default])
def found_inf_handler(op, args, kwargs):
    from .api import DTensor

    grads, found_inf, inv_scale = args[0], args[1], args[2]
    op(_local(grads), _local(found_inf), _local(inv_scale))
    meshes = {}
    for g in grads:
        if isinstance(g, DTensor):
            for i, p in enumerate(g.placements):
                if not p.is_replicate():
                    meshes[(id(g.device_mesh), i)] = (g.device_mesh, i)
    fi = _local(found_inf)
    for mesh, i in meshes.values():
        fi.copy_(C.mesh_all_reduce(fi, mesh, "max", i))
    return None


# ------------------------------------------------------------------------------- scalar extraction / equality
@register_op_handler([aten._local_scalar_dense.default])
def local_scalar_handler(op, args, kwargs):
    return op(args[0].full_tensor() if hasattr(args[0], "full_tensor") else args[0])


@register_op_handler([aten.is_same_size.default])
def same_size_handler(op, args, kwargs):
    return tuple(args[0].shape) == tuple(args[1].shape)


@register_op_handler([aten.equal.default])
def equal_handler(op, args, kwargs):
    from .api import DTensor

    a, b = args
    if not isinstance(a, DTensor) or not isinstance(b, DTensor) or tuple(a.shape) != tuple(b.shape):
        return False
    if a.placements != b.placements:
        b = b.redistribute(a.device_mesh, [Replicate() if p.is_partial() else p for p in a.placements])
        a = a.redistribute(a.device_mesh, [Replicate() if p.is_partial() else p for p in a.placements])
    ok = torch.equal(a._local_tensor, b._local_tensor)
    mesh = a.device_mesh
    if mesh.has_groups():
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=a._local_tensor.device)
        for i in range(mesh.ndim):
            flag = C.mesh_all_reduce(flag, mesh, "min", i)
        ok = bool(flag.item())
    return ok


# ------------------------------------------------------------------------------- ragged vector norm
def ragged_norm_local(local: torch.Tensor, spec: DTensorSpec, ord_: float, dims, keepdim: bool, dtype=None) -> torch.Tensor:
    """Segmented p-norm partial of a flat ragged shard over a subset of dims, without materialising the
    global tensor (the reference builds zeros(global) and copies the slice in, ``_dispatch.py:146-151``).

    The shard is ``rows x trailing`` whole rows starting at global row ``row0``.  Rows are grouped by
    the coordinates of the kept leading dims; kept trailing dims stay; the result is the p-th-power
    sum laid out in the kept-dims shape (a ``Partial(norm p)`` contribution before the root)."""
    mesh = spec.mesh
    coord = mesh.get_coordinate()
    ridx = next(i for i, p in enumerate(spec.placements) if isinstance(p, RaggedShard))
    rp: RaggedShard = spec.placements[ridx]
    before, _ = shape_and_offset_before_ragged(tuple(spec.shape), tuple(mesh.shape), spec.placements, coord)
    k = len(rp.dims)
    lead, trail = tuple(before[:k]), tuple(before[k:])
    row_elems = max(1, math.prod(trail))
    lo, hi = rp.flat_range(math.prod(before), coord[ridx])
    rows = (hi - lo) // row_elems
    row0 = lo // row_elems
    red = set(dims)
    fast = _ragged_norm_kernel(local, rows, row_elems, lead, trail, k, red, ord_, row0, before, keepdim, dtype)
    if fast is not None:
        return fast
    x = local.reshape(rows, *trail).to(dtype or (torch.float32 if local.dtype in (torch.float16, torch.bfloat16) else local.dtype))
    # power-sum elementwise
    if math.isinf(ord_):
        pw = x.abs()
    elif ord_ == 0:
        pw = (x != 0).to(x.dtype)
    elif ord_ == 1:
        pw = x.abs()
    elif ord_ == 2:
        pw = x * x
    else:
        pw = x.abs().pow(ord_)
    # reduce trailing dims that are reduced
    tr_red = [d - k + 1 for d in sorted(red) if d >= k]
    if tr_red:
        pw = pw.amax(tr_red, keepdim=True) if math.isinf(ord_) else pw.sum(tr_red, keepdim=True)
    kept_lead = [d for d in range(k) if d not in red]
    out_lead = tuple(lead[d] for d in kept_lead)
    tr_shape = tuple(pw.shape[1:])
    n_out_rows = max(1, math.prod(out_lead))
    if rows == 0:
        acc = x.new_zeros((n_out_rows, *tr_shape))
    elif len(kept_lead) == k:
        # no leading reduction: scatter rows to their global position
        acc = x.new_zeros((n_out_rows, *tr_shape))
        acc[row0 : row0 + rows] = pw
    else:
        ridx_t = torch.arange(row0, row0 + rows, device=x.device)
        # linear index over kept leading dims
        lin = torch.zeros_like(ridx_t)
        rem = ridx_t
        strides = contiguous_stride(lead)
        mult = contiguous_stride(out_lead) if out_lead else ()
        for d in range(k):
            c = rem // strides[d]
            rem = rem % strides[d]
            if d in kept_lead:
                lin = lin + c * mult[kept_lead.index(d)]
        acc = x.new_zeros((n_out_rows, *tr_shape))
        if math.isinf(ord_):
            acc.index_reduce_(0, lin, pw, "amax", include_self=True)
        else:
            acc.index_add_(0, lin, pw)
    # final shape
    if keepdim:
        shape = [1 if d in red else before[d] for d in range(len(before))]
    else:
        shape = [before[d] for d in range(len(before)) if d not in red]
    acc = acc.reshape(shape)
    if math.isinf(ord_) or ord_ in (0, 1):
        return acc
    return acc.pow(1.0 / ord_)


def _ragged_norm_kernel(local, rows, row_elems, lead, trail, k, red, ord_, row0, before, keepdim, dtype):
    """sm_100a fast path (``csrc/ragged_norm.cu``): one pass over the shard in its storage dtype when the reduced dims are exactly
    the trailing dims (per-row partials) or exactly the leading dims (per-column partials) and p is 1, 2 or inf."""
    from ..ops import _ext

    if not (local.is_cuda and local.dtype in (torch.bfloat16, torch.float32) and local.is_contiguous() and _ext.available()):
        return None
    if not (math.isinf(ord_) or ord_ in (1.0, 2.0)) or (dtype is not None and dtype != torch.float32) or not hasattr(_ext.ops(), "ragged_norm_partial"):
        return None
    nd = len(before)
    trailing, leading = set(range(k, nd)), set(range(k))
    p = 0 if math.isinf(ord_) else int(ord_)
    if red == trailing and k == 1:
        out = torch.zeros(max(1, lead[0]), dtype=torch.float32, device=local.device)
        if rows:
            _ext.count_launch("ragged_norm")
            _ext.ops().ragged_norm_partial(local, out[row0 : row0 + rows], rows, row_elems, 0, p)
        shape = [lead[0]] + ([1] * len(trail) if keepdim else [])
    elif red == leading and nd > k:
        out = torch.zeros(row_elems, dtype=torch.float32, device=local.device)
        if rows:
            _ext.count_launch("ragged_norm")
            _ext.ops().ragged_norm_partial(local, out, rows, row_elems, 1, p)
        shape = ([1] * k if keepdim else []) + list(trail)
    else:
        return None
    out = out.reshape(shape)
    return out if p != 2 else out.sqrt()


def _ragged_norm_cond(args, kwargs) -> bool:
    from .api import DTensor

    x = args[0]
    if not isinstance(x, DTensor) or not x._spec.is_ragged_shard():
        return False
    dims = args[2] if len(args) > 2 else kwargs.get("dim", None)
    if dims is None or len(dims) == 0 or len(set(d % x.ndim for d in dims)) == x.ndim:
        return False  # full reduction goes through the ordinary rule (flat local norm)
    return True


_orig_dispatch = dispatcher.dispatch


def ragged_norm_handler(op, args, kwargs):
    from .api import DTensor

    if not _ragged_norm_cond(args, kwargs):
        return None
    x: DTensor = args[0]
    ord_ = args[1] if len(args) > 1 and args[1] is not None else 2
    dims = tuple(sorted(d % x.ndim for d in (args[2] if len(args) > 2 else kwargs["dim"])))
    keepdim = bool(args[3]) if len(args) > 3 else bool(kwargs.get("keepdim", False))
    dtype = kwargs.get("dtype", None)
    spec = x._spec
    mesh = spec.mesh
    local = ragged_norm_local(x._local_tensor, spec, float(ord_), dims, keepdim, dtype)
    ridx = next(i for i, p in enumerate(spec.placements) if isinstance(p, RaggedShard))
    rp = spec.placements[ridx]
    k = len(rp.dims)
    lead_reduced = any(d < k for d in dims)
    pl = []
    for i, p in enumerate(spec.placements):
        if i == ridx:
            pl.append(Partial(f"norm{float(ord_)}"))
        elif isinstance(p, Shard):
            if p.dim in dims:
                pl.append(Partial(f"norm{float(ord_)}"))
            else:
                pl.append(Shard(p.dim if keepdim else p.dim - sum(1 for d in dims if d < p.dim)))
        else:
            pl.append(p)
    out_shape = tuple(1 if d in dims else s for d, s in enumerate(spec.shape)) if keepdim else tuple(s for d, s in enumerate(spec.shape) if d not in dims)
    ospec = DTensorSpec(mesh, tuple(pl), TensorMeta(out_shape, contiguous_stride(out_shape), local.dtype))
    return DTensor(local, ospec)


dispatcher._cond_handlers = {aten.linalg_vector_norm.default: ragged_norm_handler}


def _dispatch_with_cond(op, args, kwargs):
    h = dispatcher._cond_handlers.get(op)
    if h is not None:
        r = h(op, args, kwargs)
        if r is not None:
            return r
    return _orig_dispatch(op, args, kwargs)


dispatcher.dispatch = _dispatch_with_cond


# ------------------------------------------------------------------------------- bypass table as an object
class BypassOpDispatch:
    """The ops that never reach sharding propagation because their result is not a tensor layout question — scalar extraction,
    size comparison, equality (legacy ``_dispatch_bypass.py:26-158``: ``BypassOpDispatch.apply(op, args, kwargs)`` answers
    ``(handled, result)``).  Backed by the same functions the dispatcher's handler table holds, so registering one here is
    registering it there."""

    op_handlers = {
        aten._local_scalar_dense.default: local_scalar_handler,
        aten.is_same_size.default: same_size_handler,
        aten.equal.default: equal_handler,
    }

    @classmethod
    def register(cls, op, fn) -> None:
        cls.op_handlers[op] = fn
        register_op_handler([op], fn)

    @classmethod
    def apply(cls, op_call, args, kwargs=None):
        h = cls.op_handlers.get(op_call)
        if h is None:
            return False, None
        return True, h(op_call, args, kwargs or {})


def is_contiguous(spec: DTensorSpec) -> bool:
    """Whether the global tensor the spec describes is laid out row-major (the ragged handlers address the flat storage)."""
    expect = 1
    for size, stride in zip(reversed(spec.tensor_meta.shape), reversed(spec.tensor_meta.stride)):
        if size != 1 and stride != expect:
            return False
        expect *= size
    return True


# the new package's names for the custom handlers (``vescale/dtensor/_dispatch.py:31-250``)
found_inf_reduce_handler = found_inf_handler
fused_adamw_sgd_op_handler = fused_optimizer_handler
ragged_norm_op_handler = ragged_norm_handler
ragged_norm_kernel = _ragged_norm_kernel
from . import dispatch as _dispatch_mod  # noqa: E402

for _n in ("found_inf_reduce_handler", "fused_adamw_sgd_op_handler", "ragged_norm_op_handler", "ragged_norm_kernel", "is_contiguous", "BypassOpDispatch"):
    setattr(_dispatch_mod, _n, globals()[_n])
