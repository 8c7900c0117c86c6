"""loss_parallel: cross-entropy on logits sharded along the class (vocab) dimension without gathering them.

Inside the context ``_log_softmax`` / ``nll_loss_forward`` and their backwards get custom handlers: local
max → all-reduce(MAX), local sum-exp → all-reduce(SUM), masked local gather of the target log-prob →
Partial.  Parity: ``legacy/vescale/dtensor/loss.py:39-474`` (C20: AR(MAX), AR(SUM) of [tokens] fp32).
The fused sm_100a vocab-parallel CE kernel (``csrc/vp_ce.cu``) does local max/sum-exp in one pass.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch

from ..comm import collectives as C
from ..layout import compute_local_shape_and_global_offset
from ..placement import Partial, Replicate, Shard
from ..spec import DTensorSpec, TensorMeta, contiguous_stride
from .dispatch import dispatcher

aten = torch.ops.aten

__all__ = ["loss_parallel"]


def _class_mesh_dim(spec: DTensorSpec, dim: int) -> Optional[int]:
    for i, p in enumerate(spec.placements):
        if isinstance(p, Shard) and p.dim == dim:
            return i
    return None


def _log_softmax_handler(op, args, kwargs):
    from .api import DTensor

    x: DTensor = args[0]
    dim = args[1] % x.ndim
    half_to_float = args[2]
    md = _class_mesh_dim(x._spec, dim)
    if md is None:
        if any(p.is_shard() for p in x.placements):
            # same contract as torch / the reference (legacy ``dtensor/loss.py:105``): inside ``loss_parallel`` the logits must be
            # sharded on the class dim; anything else is almost certainly a plan mistake (the loss would gather the logits)
            raise ValueError(f"loss_parallel() only supports logits sharded on the class dimension {dim}, got placements {x.placements}")
        return _fallthrough(op, args, kwargs)
    mesh = x.device_mesh
    lx = x._local_tensor.float() if half_to_float else x._local_tensor
    lmax = lx.amax(dim, keepdim=True)
    gmax = C.mesh_all_reduce(lmax, mesh, "max", md)
    shifted = lx - gmax
    sumexp = C.mesh_all_reduce(shifted.exp().sum(dim, keepdim=True), mesh, "sum", md)
    out = shifted - sumexp.log()
    spec = DTensorSpec(mesh, x.placements, TensorMeta(tuple(x.shape), contiguous_stride(x.shape), out.dtype))
    return DTensor(out, spec)


def _log_softmax_bwd_handler(op, args, kwargs):
    from .api import DTensor

    g, out = args[0], args[1]
    dim = args[2] % out.ndim
    md = _class_mesh_dim(out._spec, dim)
    if md is None:
        return _fallthrough(op, args, kwargs)
    mesh = out.device_mesh
    if g.placements != out.placements:
        g = g.redistribute(mesh, out.placements)
    gl, ol = g._local_tensor, out._local_tensor
    s = C.mesh_all_reduce(gl.sum(dim, keepdim=True), mesh, "sum", md)
    gi = gl - ol.exp() * s
    return DTensor(gi.to(args[3]) if len(args) > 3 and isinstance(args[3], torch.dtype) else gi, out._spec.with_meta(TensorMeta(tuple(out.shape), contiguous_stride(out.shape), gi.dtype)))


def _nll_fwd_handler(op, args, kwargs):
    from .api import DTensor

    x, target = args[0], args[1]
    weight, reduction, ignore_index = args[2], args[3], args[4]
    cdim = 1 if x.ndim >= 2 else 0
    md = _class_mesh_dim(x._spec, cdim)
    if md is None or weight is not None:
        return _fallthrough(op, args, kwargs)
    mesh = x.device_mesh
    if isinstance(target, DTensor):
        target = target.redistribute(mesh, [Replicate()] * mesh.ndim)._local_tensor
    (lshape, goff) = compute_local_shape_and_global_offset(x.shape, mesh, x.placements)
    lo, n = goff[cdim], lshape[cdim]
    lx = x._local_tensor
    valid = target != ignore_index
    mine = (target >= lo) & (target < lo + n) & valid
    idx = (target - lo).clamp(0, max(n - 1, 0))
    picked = lx.gather(cdim, idx.unsqueeze(cdim)).squeeze(cdim) if x.ndim >= 2 else lx[idx]
    nll = torch.where(mine, -picked, torch.zeros_like(picked))
    nll = C.mesh_all_reduce(nll, mesh, "sum", md)
    total_w = valid.sum().to(lx.dtype)
    if reduction == 0:
        loss = nll
    elif reduction == 1:
        loss = nll.sum() / total_w
    else:
        loss = nll.sum()
    rep = tuple(Replicate() if i == md else p for i, p in enumerate(x.placements))
    rep = tuple(Replicate() if isinstance(p, Shard) and reduction != 0 else p for p in rep)
    lspec = DTensorSpec(mesh, rep, TensorMeta(tuple(loss.shape), contiguous_stride(loss.shape), loss.dtype))
    wspec = DTensorSpec(mesh, tuple(Replicate() for _ in rep), TensorMeta((), (), total_w.dtype))
    return DTensor(loss, lspec), DTensor(total_w, wspec)


def _nll_bwd_handler(op, args, kwargs):
    from .api import DTensor

    g, x, target, weight, reduction, ignore_index, total_w = args[:7]
    cdim = 1 if x.ndim >= 2 else 0
    md = _class_mesh_dim(x._spec, cdim)
    if md is None or weight is not None:
        return _fallthrough(op, args, kwargs)
    mesh = x.device_mesh
    rep = [Replicate()] * mesh.ndim
    if isinstance(target, DTensor):
        target = target.redistribute(mesh, rep)._local_tensor
    gl = g.redistribute(mesh, rep)._local_tensor if isinstance(g, DTensor) else g
    tw = total_w._local_tensor if isinstance(total_w, DTensor) else total_w
    (lshape, goff) = compute_local_shape_and_global_offset(x.shape, mesh, x.placements)
    lo, n = goff[cdim], lshape[cdim]
    valid = target != ignore_index
    mine = (target >= lo) & (target < lo + n) & valid
    idx = (target - lo).clamp(0, max(n - 1, 0))
    gi = torch.zeros(lshape, dtype=x.dtype, device=x._local_tensor.device)
    if reduction == 1:
        scale = gl / tw
    else:
        scale = gl
    val = torch.where(mine, -torch.ones_like(idx, dtype=x.dtype), torch.zeros_like(idx, dtype=x.dtype))
    val = val * (scale if scale.ndim == 0 or reduction != 0 else scale)
    gi.scatter_(cdim, idx.unsqueeze(cdim), val.unsqueeze(cdim).to(x.dtype))
    return DTensor(gi, x._spec)


_ACTIVE = [0]
_SAVED = {}


def _fallthrough(op, args, kwargs):
    h = dispatcher._custom.pop(op)
    try:
        return dispatcher.dispatch(op, args, kwargs)
    finally:
        dispatcher._custom[op] = h


_HANDLERS = {
    aten._log_softmax.default: _log_softmax_handler,
    aten._log_softmax_backward_data.default: _log_softmax_bwd_handler,
    aten.nll_loss_forward.default: _nll_fwd_handler,
    aten.nll_loss_backward.default: _nll_bwd_handler,
}


@contextlib.contextmanager
def loss_parallel():
    """``with loss_parallel(): loss = F.cross_entropy(dt_logits_sharded_on_vocab, target); loss.backward()``"""
    _ACTIVE[0] += 1
    if _ACTIVE[0] == 1:
        for op, h in _HANDLERS.items():
            dispatcher._custom[op] = h
    try:
        yield
    finally:
        _ACTIVE[0] -= 1
        if _ACTIVE[0] == 0:
            for op in _HANDLERS:
                dispatcher._custom.pop(op, None)
