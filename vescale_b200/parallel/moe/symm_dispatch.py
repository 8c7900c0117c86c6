"""MoE expert-parallel forward without host synchronisation: symmetric-memory dispatch / combine kernels and the
device-driven grouped tcgen05 GEMM (``csrc/moe_dispatch.cu``, ``grouped_gemm_nt`` in ``csrc/gemm_sm100.cu``).

Forward pipeline for T local tokens x k copies over W EP ranks (E experts, E/W per rank):

    sort copies by expert (argsort / bincount)                                   on device
    moe_exchange_counts   -> counts_all[W, E] on every rank                      1 tiny kernel, no .tolist()
    moe_plan              -> send offsets, 256-row aligned receive segments, tile->expert map
    moe_dispatch_put      -> my rows land in the peers' receive buffers (P2P stores), delivery flags
    grouped_gemm_nt (gate|up) -> swiglu -> grouped_gemm_nt (down)                ragged M per expert, device-side
    moe_signal / moe_wait -> expert outputs ready everywhere
    moe_combine_get       -> each copy's output row is pulled back from its expert's rank

The receive buffers are sized by ``capacity_factor`` x the balanced load (dropless up to that bound; an
overflow traps on the device).  The dispatched activation is the ``RaggedShard`` token DTensor of
``layer.ragged_token_placement``.  Backward reuses the same two data movers (gradients are *put* to the
experts' ranks and *pulled* back); it synchronises the per-expert row counts to the host once to run the
per-expert wgrad/dgrad GEMMs on exact row ranges.

Parity: replaces the 3x ``all_to_all_single`` + ``.tolist()`` + Python expert loop + ``index_add_`` of
``legacy/vescale/moe/_scheduler.py:162-265``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from ...ops import _ext

__all__ = ["SymmMoEDispatcher", "symm_moe_forward"]

_TILE = 256


class SymmMoEDispatcher:
    """Per-(shape) resources shared by all MoE layers of a model: symmetric receive buffers, count matrix, flags."""

    def __init__(self, ep_mesh, num_experts: int, hidden: int, ffn: int, max_tokens: int, top_k: int, capacity_factor: float = 2.0, device=None):
        from ..fsdp.api import _COMM_CACHE
        from ...comm.symm import SymmUnitComm

        self.mesh = ep_mesh
        self.group = ep_mesh.get_group(0)
        self.W = ep_mesh.size(0)
        self.rank = ep_mesh.get_local_rank(0)
        self.E, self.H, self.F, self.k = num_experts, hidden, ffn, top_k
        self.El = num_experts // self.W
        dev = device or torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        key = (id(self.group), dev.index)
        comm = _COMM_CACHE.get(key)
        if comm is None:
            comm = _COMM_CACHE[key] = SymmUnitComm(ep_mesh, 0, dev)
        self.arena = comm.arena
        self.ops = _ext.ops()
        cap = int(capacity_factor * max_tokens * top_k) + self.El * _TILE
        self.C = (cap + _TILE - 1) // _TILE * _TILE
        A = self.arena
        self.recv_x = A.alloc(self.C * hidden, torch.bfloat16).view(self.C, hidden)  # dispatched tokens (and dy in backward)
        self.out_y = A.alloc(self.C * hidden, torch.bfloat16).view(self.C, hidden)  # expert outputs (and dx in backward)
        self.counts_all = A.alloc(self.W * num_experts, torch.int32).view(self.W, num_experts)
        self.recv_ptrs, self.out_ptrs, self.counts_ptrs = A.peer_ptrs(self.recv_x), A.peer_ptrs(self.out_y), A.peer_ptrs(self.counts_all)
        s0 = A.new_slots(3)  # [counts ready | rows delivered | outputs ready], W entries each
        self.flag_ptrs = [p + s0 * self.W * 4 for p in A.pad_ptrs]
        self.my_flags = A.my_pad + s0 * self.W * 4
        self.epoch = 0
        self.done_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.recv_seg_start = torch.zeros(self.El * self.W, **i32)
        self.send_off = torch.zeros(num_experts, **i32)
        self.tile_expert = torch.full((self.C // _TILE,), -1, **i32)
        self.expert_rows = torch.zeros(self.El, **i32)
        self.total_rows = torch.zeros(1, **i32)
        # receive buffers are kept zero outside their live range (each consumer zeroes them once it has read them),
        # so the padding rows of the tile-aligned expert segments are always finite
        self.recv_x.zero_()
        self.out_y.zero_()
        self.counts_all.zero_()
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group, device_ids=[dev.index])

    # ------------------------------------------------------------------ routing metadata (device only)
    def route(self, topi: torch.Tensor):
        flat_e = topi.reshape(-1).to(torch.int32)
        order = torch.argsort(flat_e.long(), stable=True)
        counts = torch.bincount(flat_e.long(), minlength=self.E).to(torch.int32)
        row_expert = flat_e[order].contiguous()
        starts = torch.cumsum(counts, 0, dtype=torch.int64) - counts
        row_pos = (torch.arange(order.numel(), device=order.device) - starts[row_expert.long()]).to(torch.int32)
        inv = torch.empty_like(order)
        inv[order] = torch.arange(order.numel(), device=order.device)
        return order, counts, row_expert, row_pos, inv

    def exchange_and_plan(self, counts: torch.Tensor) -> None:
        self.epoch += 1
        _ext.count_launch("moe_exchange_counts")
        self.ops.moe_exchange_counts(counts, self.counts_ptrs, self.flag_ptrs, self.rank, self.epoch)
        _ext.count_launch("moe_plan")
        self.ops.moe_plan(self.counts_all, self.recv_seg_start, self.send_off, self.tile_expert, self.expert_rows, self.total_rows, self.rank, _TILE)

    def put(self, rows_sorted: torch.Tensor, row_expert, row_pos, into_out: bool = False) -> None:
        """Rows (sorted by expert) -> destination ranks' receive buffer; returns after *my* stores are issued; call
        ``wait_delivered`` before reading the local receive buffer."""
        _ext.count_launch("moe_dispatch_put")
        self.ops.moe_dispatch_put(rows_sorted, row_expert, row_pos, self.send_off, self.out_ptrs if into_out else self.recv_ptrs, self.flag_ptrs,
                                  self.done_counter, self.E, self.rank, self._put_epoch())

    def _put_epoch(self) -> int:
        self._puts = getattr(self, "_puts", 0) + 1
        return self._puts

    def wait_delivered(self) -> None:
        self.ops.moe_wait(self.my_flags, self.W, self.W, self._puts)

    def signal_outputs_ready(self) -> None:
        self._outs = getattr(self, "_outs", 0) + 1
        self.ops.moe_signal(self.flag_ptrs, 2 * self.W, self.rank, self._outs)
        self.ops.moe_wait(self.my_flags, self.W, 2 * self.W, self._outs)

    def slots(self, topi: torch.Tensor, row_pos, inv) -> Tuple[torch.Tensor, torch.Tensor]:
        flat_e = topi.reshape(-1).long()
        slot_rank = (flat_e // self.El).to(torch.int32)
        slot_row = (self.send_off[flat_e].long() + row_pos[inv].long()).to(torch.int32)
        return slot_rank, slot_row

    def pull(self, slot_rank, slot_row, from_out: bool = True, gate: Optional[torch.Tensor] = None, k: int = 1) -> torch.Tensor:
        n = slot_rank.numel() // k
        out = torch.empty(n, self.H, dtype=torch.bfloat16, device=self.device)
        _ext.count_launch("moe_combine_get")
        self.ops.moe_combine_get(out, gate, slot_rank, slot_row, self.out_ptrs if from_out else self.recv_ptrs, k)
        return out


class _SymmMoEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, topv, topi, w1, w2, disp: SymmMoEDispatcher):
        ops, El, F, H, k = disp.ops, disp.El, disp.F, disp.H, disp.k
        order, counts, row_expert, row_pos, inv = disp.route(topi)
        disp.exchange_and_plan(counts)
        xs = x2[order // k].contiguous()
        disp.put(xs, row_expert, row_pos)
        disp.wait_delivered()
        gu = torch.zeros(disp.C, 2 * F, dtype=torch.bfloat16, device=x2.device)  # skipped (empty) tiles stay zero
        _ext.count_launch("grouped_gemm_nt", 2)
        ops.grouped_gemm_nt(disp.recv_x, w1.view(El * 2 * F, H), gu, disp.tile_expert, 2 * F)
        recv_saved = disp.recv_x.clone()
        disp.recv_x.zero_()  # consumed: peers may write into it again only after my next flag (stream-ordered after this)
        act = ops.swiglu_fwd(gu)
        ops.grouped_gemm_nt(act, w2.view(El * H, F), disp.out_y, disp.tile_expert, H)
        disp.signal_outputs_ready()
        slot_rank, slot_row = disp.slots(topi, row_pos, inv)
        y_copies = disp.pull(slot_rank, slot_row, from_out=True)  # [T*k, H], one row per copy
        T = x2.shape[0]
        out = (y_copies.view(T, k, H).float() * topv.unsqueeze(-1).float()).sum(1).to(x2.dtype)
        # everyone must have pulled before the next layer overwrites the shared buffers
        disp.signal_outputs_ready()
        ctx.save_for_backward(x2, topv, topi, w1, w2, y_copies, recv_saved, disp.expert_rows.clone(), disp.send_off.clone(), disp.tile_expert.clone())
        ctx.disp = disp
        ctx.route = (order, row_expert, row_pos, inv, slot_rank, slot_row)
        return out

    @staticmethod
    def backward(ctx, dout):
        x2, topv, topi, w1, w2, y_copies, recv_x, expert_rows, send_off, tile_expert = ctx.saved_tensors
        disp: SymmMoEDispatcher = ctx.disp
        order, row_expert, row_pos, inv, slot_rank, slot_row = ctx.route
        ops, El, F, H, k, W = disp.ops, disp.El, disp.F, disp.H, disp.k, disp.W
        T = x2.shape[0]
        dout = dout.contiguous()
        dtopv = (y_copies.view(T, k, H).float() * dout.unsqueeze(1).float()).sum(-1).to(topv.dtype)
        # dy of every copy, pushed to the expert's rank (same slots as the forward dispatch)
        dy_copies = (dout.unsqueeze(1).float() * topv.unsqueeze(-1).float()).to(torch.bfloat16).view(T * k, H)
        disp.send_off.copy_(send_off)
        disp.put(dy_copies[order].contiguous(), row_expert, row_pos)
        disp.wait_delivered()
        rows = expert_rows.tolist()  # one host sync in backward: exact per-expert row ranges for wgrad
        dw1, dw2 = torch.zeros_like(w1), torch.zeros_like(w2)
        dx_recv = disp.out_y
        pos = 0
        for e, n in enumerate(rows):
            if n:
                xe, dye = recv_x[pos : pos + n], disp.recv_x[pos : pos + n]
                gu = xe @ w1[e].t()
                act = ops.swiglu_fwd(gu)
                dw2[e] = dye.t() @ act
                dgu = ops.swiglu_bwd(dye @ w2[e], gu)
                dw1[e] = dgu.t() @ xe
                dx_recv[pos : pos + n] = dgu @ w1[e]
            pos = (pos + n + _TILE - 1) // _TILE * _TILE
        disp.recv_x.zero_()
        disp.signal_outputs_ready()
        dx_copies = disp.pull(slot_rank, slot_row, from_out=True)  # gradient of every copy's input row
        dx = dx_copies.view(T, k, H).float().sum(1).to(x2.dtype)
        disp.signal_outputs_ready()
        return dx, dtopv, None, dw1, dw2, None


def symm_moe_forward(x2: torch.Tensor, topv: torch.Tensor, topi: torch.Tensor, w_gate_up: torch.Tensor, w_down: torch.Tensor, disp: SymmMoEDispatcher) -> torch.Tensor:
    """x2 [T, H] bf16, topv/topi [T, k], w_gate_up [E_local, 2F, H], w_down [E_local, H, F] -> [T, H]."""
    return _SymmMoEFn.apply(x2, topv, topi, w_gate_up, w_down, disp)
