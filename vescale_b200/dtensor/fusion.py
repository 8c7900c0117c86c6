"""DTensor-level fusion of a tensor-parallel collective with the matmul next to it.

The two redistribute → matmul shapes every Megatron-style plan produces are recognised *inside the op dispatcher*:

* ``mm(x: Shard(0), w_t: Shard(1))`` on a mesh dim — a sequence-/row-sharded activation entering a column-parallel weight.  The
  generic path all-gathers ``x`` (``redistribute(Shard→Replicate)``, reference ``legacy/vescale/dtensor/redistribute.py:122`` issued from
  ``legacy/vescale/dmodule/_hook.py:97``) and then multiplies; here it becomes ONE ``ag_gemm`` launch (peers' rows are pulled
  through shared memory by a copy warp while the MMA warps work on the rows that already arrived).
* ``mm(x: Shard(1), w_t: Shard(0))`` whose ``Partial`` result is about to be resharded to ``Shard`` (``mm(Partial) →
  redistribute(→Shard)``, reference ``redistribute.py:341`` issued from ``_hook.py:230``) becomes ONE ``gemm_rs`` launch.  The
  dispatcher cannot see the future, so the resharding target is announced as a *hint*: DModule output plans push it for the
  duration of the module's forward (the plan-level hint of legacy ``PlacementsInterface.defer_reshard``), and user code can say
  ``with fuse_reshard(mesh, [Shard(0)]): y = x @ w.t()``.  Producing ``Shard`` instead of ``Partial`` is always a legal answer for
  a DTensor op, so a hint can never change results, only where the reduction happens.

Back ends: ``FusedTP`` (``csrc/gemm_fused_tp.cu``, sm_100a) on CUDA; a c10d implementation of the same two calls for CPU/gloo
tests and as the measured baseline (``VESCALE_B200_FUSE_TP=c10d``).  ``VESCALE_B200_FUSE_TP=off`` disables the pattern match.
Backward runs through the ordinary DTensor rules (the hand-fused backward duals live in ``comm/fused_tp.py`` for models written
against them, ``models/llama_tp.py``).
"""
from __future__ import annotations

import contextlib
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ..placement import InterleavedShard, Partial, Placement, RaggedShard, Replicate, Shard
from ..spec import DTensorSpec, TensorMeta
from ..utils.env import flag as _flag

__all__ = ["fuse_reshard", "mm_fusion_handler", "stats", "reset_stats", "push_hint", "pop_hint"]

aten = torch.ops.aten
_TLS = threading.local()
stats: Dict[str, int] = {"ag_gemm": 0, "gemm_rs": 0, "fallback": 0}


def reset_stats() -> None:
    for k in stats:
        stats[k] = 0


def _hints() -> List[Tuple[int, object, Tuple[Placement, ...], bool]]:
    h = getattr(_TLS, "hints", None)
    if h is None:
        h = _TLS.hints = []
    return h


def push_hint(owner: int, mesh, placements: Sequence[Placement], rows_contiguous: bool = True) -> None:
    """Announce that the ``Partial`` result of a matmul on ``mesh`` will be resharded to ``placements``."""
    hs = _hints()
    hs[:] = [h for h in hs if h[0] != owner]  # a forward that raised may have left a stale entry of this owner
    hs.append((owner, mesh, tuple(placements), rows_contiguous))


def pop_hint(owner: int) -> None:
    hs = _hints()
    hs[:] = [h for h in hs if h[0] != owner]


@contextlib.contextmanager
def fuse_reshard(mesh, placements: Sequence[Placement]):
    """Within the block, a matmul whose result would be ``Partial`` on a mesh dim where ``placements`` says ``Shard(0)`` is
    computed by the fused GEMM ⊕ reduce-scatter kernel and comes out ``Shard(0)`` directly."""
    tok = id(object())
    push_hint(tok, mesh, placements)
    try:
        yield
    finally:
        pop_hint(tok)


# ------------------------------------------------------------------------------------------------- back ends
class _C10dBackend:
    """The same two calls on ordinary collectives + library GEMMs (gloo / NCCL)."""

    name = "c10d"

    def __init__(self, mesh, md: int):
        self.mesh, self.md = mesh, md
        self.group = mesh.get_group(md)
        self.world = mesh.size(md)

    def can_ag(self, x: torch.Tensor, w: torch.Tensor) -> bool:
        return True

    def can_rs(self, x: torch.Tensor, w: torch.Tensor) -> bool:
        return x.shape[0] % self.world == 0

    def ag_gemm(self, x_local: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        x_full = torch.empty(x_local.shape[0] * self.world, x_local.shape[1], dtype=x_local.dtype, device=x_local.device)
        dist.all_gather_into_tensor(x_full, x_local.contiguous(), group=self.group)
        return x_full @ w.t()

    def gemm_rs(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        part = (x @ w.t()).contiguous()
        out = torch.empty(part.shape[0] // self.world, part.shape[1], dtype=part.dtype, device=part.device)
        if dist.get_backend(self.group) == "gloo":  # gloo has no reduce_scatter
            dist.all_reduce(part, group=self.group)
            r = dist.get_rank(self.group)
            out.copy_(part[r * out.shape[0] : (r + 1) * out.shape[0]])
        else:
            dist.reduce_scatter_tensor(out, part, group=self.group)
        return out


class _FusedBackend:
    """``comm.fused_tp.FusedTP``: all-gather ⊕ GEMM and GEMM ⊕ reduce-scatter in one sm_100a kernel each."""

    name = "fused"

    def __init__(self, mesh, md: int, device):
        from ..comm.fused_tp import FusedTP

        self.tp = FusedTP(mesh, md, device)
        self.world = self.tp.world

    @staticmethod
    def _ok(t: torch.Tensor) -> bool:
        return t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous() and t.data_ptr() % 16 == 0

    def can_ag(self, x, w) -> bool:  # constraints of csrc/gemm_fused_tp.cu::ag_gemm
        return self._ok(x) and self._ok(w) and x.shape[0] % 256 == 0 and x.shape[1] % 256 == 0 and w.shape[0] % 8 == 0

    def can_rs(self, x, w) -> bool:
        return self._ok(x) and self._ok(w) and x.shape[0] % (256 * self.world) == 0 and x.shape[1] % 64 == 0 and w.shape[0] % 8 == 0

    def ag_gemm(self, x_local, w):
        return self.tp.ag_gemm(x_local, w)[0]

    def gemm_rs(self, x, w):
        return self.tp.gemm_rs(x, w)


_BACKENDS: Dict[Tuple[int, int, str], object] = {}


def _backend(mesh, md: int, sample: torch.Tensor):
    mode = _flag("VESCALE_B200_FUSE_TP")
    if mode == "off" or mesh.size(md) <= 1 or not mesh.has_groups():
        return None
    kind = "c10d" if mode == "c10d" else ("fused" if sample.is_cuda else None)
    if kind is None:
        return None
    key = (id(mesh.get_group(md)), md, kind)
    b = _BACKENDS.get(key)
    if b is None:
        try:
            if kind == "fused":
                from ..ops import _ext

                if not _ext.available():
                    return None
                b = _FusedBackend(mesh, md, sample.device)
            else:
                b = _C10dBackend(mesh, md)
        except Exception:  # noqa: BLE001 — symmetric memory unavailable: stay on the generic path
            b = False
        _BACKENDS[key] = b
    return b or None


# ------------------------------------------------------------------------------------------------- the pattern match
def _plain_shard(p, dim: int) -> bool:
    return type(p) is Shard and p.dim == dim or (isinstance(p, InterleavedShard) and p.dim == dim and p.interleaved_size == 1)


def _kmajor(w_t_local: torch.Tensor) -> Optional[torch.Tensor]:
    """``w_t_local`` is the local [K, N] operand of ``mm``; the kernels want the K-major [N, K] weight.  For ``x @ w.t()`` that is a
    free view; otherwise fusing would cost a transpose copy and the generic path is taken."""
    w = w_t_local.t()
    return w if w.is_contiguous() else None


def mm_fusion_handler(dispatcher, op, args, kwargs):
    """Custom handler of ``aten.mm``: returns a DTensor when one of the two patterns matched, ``NotImplemented`` otherwise."""
    from .api import DTensor

    a, b = args[0], args[1]
    if type(a) is not DTensor or type(b) is not DTensor or a.ndim != 2 or b.ndim != 2:
        return NotImplemented
    mesh = a._spec.mesh
    if b._spec.mesh != mesh or mesh.get_coordinate() is None:
        return NotImplemented
    pa, pb = a._spec.placements, b._spec.placements
    md = None
    for i in range(mesh.ndim):
        if pa[i].is_replicate() and pb[i].is_replicate():
            continue
        if md is not None:
            return NotImplemented  # sharded on more than one mesh dim: generic path
        md = i
    if md is None:
        return NotImplemented
    x_local, wt_local = a._local_tensor, b._local_tensor
    M, K = a.shape
    N = b.shape[1]
    W = mesh.size(md)
    if _plain_shard(pa[md], 0) and _plain_shard(pb[md], 1) and M % W == 0 and N % W == 0:
        # ---- redistribute(Shard(0) -> Replicate) -> mm  ==>  all-gather ⊕ GEMM
        be = _backend(mesh, md, x_local)
        w = _kmajor(wt_local)
        if be is not None and w is not None and be.can_ag(x_local, w):
            y = be.ag_gemm(x_local.contiguous(), w)
            stats["ag_gemm"] += 1
            pl = tuple(Shard(1) if i == md else Replicate() for i in range(mesh.ndim))
            return DTensor(y, DTensorSpec(mesh, pl, TensorMeta((M, N), (N, 1), y.dtype)), requires_grad=False)
    elif _plain_shard(pa[md], 1) and _plain_shard(pb[md], 0) and M % W == 0:
        # ---- mm -> Partial, about to be resharded to Shard(0)  ==>  GEMM ⊕ reduce-scatter (needs the hint)
        hint = next((h for h in reversed(_hints()) if h[1] == mesh), None)
        if hint is not None:
            tgt = hint[2][md] if md < len(hint[2]) else None
            rows_ok = isinstance(tgt, Shard) and not isinstance(tgt, (InterleavedShard,)) and (tgt.dim == 0 or (tgt.dim == 1 and hint[3]))
            if rows_ok:
                be = _backend(mesh, md, x_local)
                w = _kmajor(wt_local)
                if be is not None and w is not None and be.can_rs(x_local, w):
                    y = be.gemm_rs(x_local.contiguous(), w)
                    stats["gemm_rs"] += 1
                    pl = tuple(Shard(0) if i == md else Replicate() for i in range(mesh.ndim))
                    return DTensor(y, DTensorSpec(mesh, pl, TensorMeta((M, N), (N, 1), y.dtype)), requires_grad=False)
    stats["fallback"] += 1
    return NotImplemented
