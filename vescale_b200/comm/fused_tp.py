"""Tensor-parallel linear layers whose collective is fused into the GEMM kernel (``csrc/gemm_fused_tp.cu``).

    tp = FusedTP(mesh, "TP")
    y_full_rows = tp.ag_linear(x_seq_shard, w_col_shard)     # all-gather(x) ⊕ GEMM      (SP -> column parallel)
    y_seq_shard = tp.linear_rs(x_full_rows, w_row_shard)     # GEMM ⊕ reduce-scatter     (row parallel -> SP)

These are what ``DTensor.redistribute(Shard(seq)→Replicate) → mm`` and ``mm(Partial) → redistribute(→Shard(seq))``
become on B200.  Backward uses the dual collectives (all-gather ↔ reduce-scatter) on
NCCL + cuBLAS for now; the forward kernels never call NCCL.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist

from ..ops import _ext

__all__ = ["FusedTP"]


class FusedTP:
    def __init__(self, mesh, mesh_dim=0, device=None):
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
