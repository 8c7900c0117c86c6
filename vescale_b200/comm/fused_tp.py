"""Tensor-parallel linear layers whose collective is fused into the GEMM kernel (``csrc/gemm_fused_tp.cu``).

    tp = FusedTP(mesh, "TP")
    y_full_rows = tp.ag_linear(x_seq_shard, w_col_shard)     # all-gather(x) ⊕ GEMM      (SP -> column parallel)
    y_seq_shard = tp.linear_rs(x_full_rows, w_row_shard)     # GEMM ⊕ reduce-scatter     (row parallel -> SP)

These are what ``DTensor.redistribute(Shard(seq)→Replicate) → mm`` and ``mm(Partial) → redistribute(→Shard(seq))``
become on B200.  Backward uses the dual collectives (all-gather ↔ reduce-scatter) on
NCCL + cuBLAS for now; the forward kernels never call NCCL.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops import _ext
from ..utils.env import flag as _flag

__all__ = ["FusedTP", "PlainTP"]


class FusedTP:
    """``fused_backward=True`` (default) also runs the backward duals on the fused kernels: d(ag_linear)/dx is
    ``gemm_rs(dy, W^T)`` and d(linear_rs)/dx is ``ag_gemm(dy_local, W^T)``; the transposed weight copy is a
    few tens of microseconds per layer."""

    def __init__(self, mesh, mesh_dim=0, device=None, fused_backward: bool = True, rs_impl: Optional[str] = None):
        self.fused_backward = fused_backward
        self.rs_impl = rs_impl or _flag("VESCALE_B200_GEMM_RS")  # "staged" | "nvls"
        if self.rs_impl not in ("staged", "nvls"):
            raise ValueError(f"rs_impl must be 'staged' or 'nvls', got {self.rs_impl!r}")
        self._sc = None
        from ..parallel.fsdp.api import _COMM_CACHE
        from .symm import SymmUnitComm

        md = mesh._dim_index(mesh_dim)
        self.mesh, self.md = mesh, md
        self.group = mesh.get_group(md)
        self.world = mesh.size(md)
        self.rank = mesh.get_local_rank(md)
        dev = device or torch.device("cuda", torch.cuda.current_device())
        key = (id(self.group), dev.index)
        comm = _COMM_CACHE.get(key)
        if comm is None:
            comm = _COMM_CACHE[key] = SymmUnitComm(mesh, md, dev)
        self.comm, self.arena, self.device = comm, comm.arena, dev
        self.ops = _ext.ops()
        self._sites: Dict[Tuple, dict] = {}

    def _site(self, kind: str, *shape) -> dict:
        k = (kind, *shape)
        st = self._sites.get(k)
        if st is None:
            s0 = self.arena.new_slots(2)
            W = self.world
            st = {"epoch": 0, "flag_ptrs": [p + s0 * W * 4 for p in self.arena.pad_ptrs]}
            self._sites[k] = st
        return st

    # ------------------------------------------------------------------ all-gather ⊕ GEMM
    def ag_gemm(self, x_local: torch.Tensor, w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """x_local [M/W, K], w [Nr, K] -> (y [M, Nr], x_full [M, K])."""
        Ml, K = x_local.shape
        W = self.world
        st = self._site("ag", Ml, K, w.shape[0])
        if "x_sym" not in st:
            st["x_sym"] = self.arena.alloc(Ml * K, torch.bfloat16).view(Ml, K)
            st["x_ptrs"] = self.arena.peer_ptrs(st["x_sym"])
            st["arrive"] = torch.zeros(max(1, Ml * W // 256), dtype=torch.int32, device=self.device)
        st["x_sym"].copy_(x_local)
        x_full = torch.empty(Ml * W, K, dtype=torch.bfloat16, device=self.device)
        y = torch.empty(Ml * W, w.shape[0], dtype=torch.bfloat16, device=self.device)
        st["epoch"] += 1
        _ext.count_launch("ag_gemm")
        self.ops.ag_gemm(st["x_sym"], st["x_ptrs"], w, x_full, y, st["arrive"], st["flag_ptrs"], self.rank, st["epoch"])
        # own rows are consumed straight from x_sym; fill them in the gathered copy for backward's wgrad
        x_full[self.rank * Ml : (self.rank + 1) * Ml].copy_(x_local)
        return y, x_full

    # ------------------------------------------------------------------ GEMM ⊕ reduce-scatter
    def gemm_rs(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        """x [M, Kr], w [N, Kr] -> y [M/W, N] (sum over ranks, row-scattered)."""
        if self.rs_impl == "nvls":
            return self._gemm_rs_nvls(x, w)
        M, _ = x.shape
        N, W = w.shape[0], self.world
        st = self._site("rs", M, N)
        if "staging" not in st:
            st["staging"] = self.arena.alloc(M * N, torch.bfloat16)  # [W, M/W, N]
            st["staging_ptrs"] = self.arena.peer_ptrs(st["staging"])
            st["done"] = torch.zeros(W, dtype=torch.int32, device=self.device)
        y = torch.empty(M // W, N, dtype=torch.bfloat16, device=self.device)
        st["epoch"] += 1
        _ext.count_launch("gemm_rs")
        self.ops.gemm_rs(x, w, y, st["staging_ptrs"], st["done"], st["flag_ptrs"], self.rank, st["epoch"])
        return y

    def _gemm_rs_nvls(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        """``rs_impl="nvls"``: the GEMM writes this rank's partial [M, N] into symmetric memory and the owner of each row block
        pulls the sum through the switch (``SymmCollectives.reduce_scatter``: one ``multimem.ld_reduce`` per 16 bytes), so a
        GPU receives M*N/W reduced values instead of (W-1)/W * M*N staged partials.  The staged form wins at W=2, this one is
        meant for W >= 4 (profiles/roofline_r1.md: staged gemm_rs at 0.82-0.98x of NCCL at W=8)."""
        from ..ops import functional as F
        from .symm_collectives import SymmCollectives

        M, N = x.shape[0], w.shape[0]
        st = self._sites.get(("rs_nvls", M, N))
        if st is None:
            st = self._sites[("rs_nvls", M, N)] = {"partial": self.arena.alloc(M * N, torch.bfloat16).view(M, N)}
        if self._sc is None:
            self._sc = SymmCollectives(self.mesh, self.md, self.device)
        F.gemm_nt(x, w, out=st["partial"])
        # the reduce-scatter kernel retires only after every peer has finished reading my partial (end barrier), so the next
        # call's GEMM may overwrite it without further synchronisation
        return self._sc.reduce_scatter(st["partial"])

    # ------------------------------------------------------------------ autograd front-ends
    def ag_linear(self, x_local: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        return _AGLinear.apply(x_local, w, self)

    def linear_rs(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        return _LinearRS.apply(x, w, self)


class _AGLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, w, tp: FusedTP):
        shape = x_local.shape
        y, x_full = tp.ag_gemm(x_local.reshape(-1, shape[-1]).contiguous(), w)
        ctx.save_for_backward(x_full, w)
        ctx.tp, ctx.shape = tp, shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x_full, w = ctx.saved_tensors
        tp = ctx.tp
        dy = dy.contiguous()
        if tp.fused_backward and dy.shape[0] % (256 * tp.world) == 0 and dy.shape[1] % 64 == 0 and w.shape[1] % 8 == 0:
            dx = tp.gemm_rs(dy, w.t().contiguous())  # GEMM ⊕ reduce-scatter: [M/W, K]
        else:
            dx_full = dy @ w  # [M, K] partial over the TP group
            dx = torch.empty(dx_full.shape[0] // tp.world, dx_full.shape[1], dtype=dx_full.dtype, device=dx_full.device)
            dist.reduce_scatter_tensor(dx, dx_full, group=tp.group)
        mg = getattr(w, "main_grad", None)
        if mg is not None:
            from ..ops.functional import gemm_tn

            gemm_tn(dy, x_full, out=mg, accumulate=getattr(w, "_main_grad_initialised", False))
            w._main_grad_initialised = True
            dw = None
        else:
            dw = dy.t() @ x_full
        return dx.view(ctx.shape), dw, None


class _LinearRS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, tp: FusedTP):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        ctx.save_for_backward(x2, w)
        ctx.tp, ctx.shape = tp, x.shape
        return tp.gemm_rs(x2, w)

    @staticmethod
    def backward(ctx, dy_local):
        x2, w = ctx.saved_tensors
        tp = ctx.tp
        dy_local = dy_local.contiguous()
        if tp.fused_backward and dy_local.shape[0] % 256 == 0 and dy_local.shape[1] % 256 == 0 and w.shape[1] % 8 == 0:
            dx, dy = tp.ag_gemm(dy_local, w.t().contiguous())  # all-gather ⊕ GEMM; the gathered dy feeds the wgrad
        else:
            dy = torch.empty(dy_local.shape[0] * tp.world, dy_local.shape[1], dtype=dy_local.dtype, device=dy_local.device)
            dist.all_gather_into_tensor(dy, dy_local, group=tp.group)
            dx = dy @ w
        mg = getattr(w, "main_grad", None)
        if mg is not None:
            from ..ops.functional import gemm_tn

            gemm_tn(dy, x2, out=mg, accumulate=getattr(w, "_main_grad_initialised", False))
            w._main_grad_initialised = True
            dw = None
        else:
            dw = dy.t() @ x2
        return dx.view(ctx.shape), dw, None


class PlainTP:
    """The same two tensor-parallel linears on ordinary collectives (NCCL / gloo) + library GEMMs: the measured baseline of
    ``FusedTP`` and the CPU-testable implementation (legacy ``redistribute.py:122,341`` → ``mm`` call pattern)."""

    def __init__(self, mesh, mesh_dim=0):
        md = mesh._dim_index(mesh_dim)
        self.mesh, self.md = mesh, md
        self.group = mesh.get_group(md)
        self.world = mesh.size(md)
        self.rank = mesh.get_local_rank(md)

    def ag_linear(self, x_local: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        return _PlainAGLinear.apply(x_local, w, self)

    def linear_rs(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        return _PlainLinearRS.apply(x, w, self)


def _wgrad_into(w, dy2, x2):
    from ..ops.functional import gemm_tn

    mg = getattr(w, "main_grad", None)
    if mg is not None:
        gemm_tn(dy2, x2, out=mg, accumulate=getattr(w, "_main_grad_initialised", False))
        w._main_grad_initialised = True
        return None
    return dy2.t() @ x2


class _PlainAGLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, w, tp: PlainTP):
        from . import collectives as C

        x2 = x_local.reshape(-1, x_local.shape[-1]).contiguous()
        x_full = C.mesh_all_gather(x2, tp.mesh, tp.md, 0)
        ctx.save_for_backward(x_full, w)
        ctx.tp, ctx.shape = tp, x_local.shape
        return x_full @ w.t()

    @staticmethod
    def backward(ctx, dy):
        from . import collectives as C

        x_full, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = C.mesh_reduce_scatter(dy @ w, ctx.tp.mesh, "sum", ctx.tp.md, 0)
        return dx.view(ctx.shape), _wgrad_into(w, dy, x_full), None


class _PlainLinearRS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, tp: PlainTP):
        from . import collectives as C

        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        ctx.save_for_backward(x2, w)
        ctx.tp, ctx.shape = tp, x.shape
        return C.mesh_reduce_scatter(x2 @ w.t(), tp.mesh, "sum", tp.md, 0)

    @staticmethod
    def backward(ctx, dy_local):
        from . import collectives as C

        x2, w = ctx.saved_tensors
        dy = C.mesh_all_gather(dy_local.contiguous(), ctx.tp.mesh, ctx.tp.md, 0)
        return (dy @ w).view(ctx.shape), _wgrad_into(w, dy, x2), None
