"""Symmetric-memory substrate for one mesh dimension + the FSDP unit collectives built on it.

``SymmArena``: a growable pool of peer-mapped device memory.  Chunks are allocated with CUDA VMM and
exchanged between the ranks of the mesh dim at rendezvous (``torch.distributed._symmetric_memory`` does the
cuMem export/import and, when the fabric supports it, the NVLS multicast binding); every rank performs the
same sequence of ``alloc`` calls, so a (chunk, offset) pair names the same logical buffer on every rank and
``peer_ptrs(tensor)`` yields the addresses of that buffer in all peers.  A uint32 signal pad lives in the
first chunk.

``SymmUnitComm``: what ``FSDPUnit`` calls — pull all-gather of the unit's parameter shards and the fused
reduce-scatter kernels in ``csrc/symm_comm.cu``.  No NCCL kernel runs on these paths.

Parity: this replaces the reference's list-``all_gather`` + ``cat`` (``placement_types.py:128-150``) and
all-reduce-then-slice (``_redistribute.py:111-120``) for RaggedShard, and legacy's bucket
``_reduce_scatter_base`` / ``_all_gather_base`` + separate scale/cast/Adam passes.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops import _ext

__all__ = ["SymmArena", "SymmUnitComm", "symm_available", "get_unit_comm"]

_MAX_SLOTS = 2048  # signal slots per rank


def symm_available() -> bool:
    try:
        import torch.distributed._symmetric_memory  # noqa: F401

        return torch.cuda.is_available()
    except Exception:
        return False


class _Chunk:
    def __init__(self, tensor: torch.Tensor, handle, nbytes: int):
        self.tensor = tensor  # uint8 [nbytes], symmetric
        self.handle = handle
        self.nbytes = nbytes
        self.used = 0
        self.base = tensor.data_ptr()
        self.peer_bases: List[int] = [int(p) for p in handle.buffer_ptrs]
        mc = getattr(handle, "multicast_ptr", 0) or 0
        self.multicast_base = int(mc)


class SymmArena:
    """Growable pool of peer-mapped device memory with a signal pad (see the module docstring)."""
    def __init__(self, group, device: torch.device, chunk_bytes: int = 1 << 30):
        import torch.distributed._symmetric_memory as symm_mem

        self._symm = symm_mem
        self.group = group
        self.group_name = group.group_name
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.chunk_bytes = chunk_bytes
        self.chunks: List[_Chunk] = []
        try:
            symm_mem.enable_symm_mem_for_group(self.group_name)
        except Exception:
            pass
        # signal pad: [_MAX_SLOTS, world] uint32, zero-initialised on every rank before anyone signals
        self.pad = self.alloc(_MAX_SLOTS * self.world, torch.int32)
        self.pad.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=group, device_ids=[device.index])
        self.pad_ptrs = self.peer_ptrs(self.pad)
        self.my_pad = self.pad.data_ptr()
        self._next_slot = 0

    # ------------------------------------------------------------------ allocation
    def _new_chunk(self, min_bytes: int) -> _Chunk:
        nbytes = max(self.chunk_bytes, (min_bytes + (1 << 21) - 1) // (1 << 21) * (1 << 21))
        t = self._symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        h = self._symm.rendezvous(t, group=self.group_name)
        c = _Chunk(t, h, nbytes)
        self.chunks.append(c)
        return c

    def alloc(self, numel: int, dtype: torch.dtype, align: int = 1024) -> torch.Tensor:
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        chunk = None
        for c in self.chunks:
            off = (c.used + align - 1) // align * align
            if off + nbytes <= c.nbytes:
                chunk = c
                break
        if chunk is None:
            chunk = self._new_chunk(nbytes)
        off = (chunk.used + align - 1) // align * align
        chunk.used = off + nbytes
        t = chunk.tensor[off : off + nbytes].view(dtype)
        t._symm_chunk = chunk
        t._symm_off = off
        return t

    def _locate(self, t: torch.Tensor) -> Tuple[_Chunk, int]:
        p = t.data_ptr()
        for c in self.chunks:
            if c.base <= p < c.base + c.nbytes:
                return c, p - c.base
        raise ValueError("tensor does not live in this symmetric arena")

    def peer_ptrs(self, t: torch.Tensor) -> List[int]:
        c, off = self._locate(t)
        return [b + off for b in c.peer_bases]

    def multicast_ptr(self, t: torch.Tensor) -> int:
        c, off = self._locate(t)
        return c.multicast_base + off if c.multicast_base else 0

    def new_slots(self, n: int) -> int:
        s = self._next_slot
        self._next_slot += n
        if self._next_slot > _MAX_SLOTS:
            raise RuntimeError("out of signal slots")
        return s

    def allocated_bytes(self) -> int:
        return sum(c.nbytes for c in self.chunks)


class SymmUnitComm:
    """FSDP unit collectives over a SymmArena (one per FSDP mesh dim)."""

    symmetric = True

    def __init__(self, mesh, mesh_dim: int, device: torch.device, chunk_bytes: Optional[int] = None):
        self.mesh = mesh
        self.group = mesh.get_group(mesh_dim)
        self.world = mesh.size(mesh_dim)
        self.rank = mesh.get_local_rank(mesh_dim)
        self.device = device
        cb = chunk_bytes or int(os.environ.get("VESCALE_B200_SYMM_CHUNK_MB", "2048")) << 20
        self.arena = SymmArena(self.group, device, cb)
        self.use_multimem = os.environ.get("VESCALE_B200_MULTIMEM", "1") == "1"
        self.ops = _ext.ops()
        self._unit_slots: Dict[int, Tuple[int, int, int]] = {}
        self._fused_sites: Dict[Tuple[int, str], dict] = {}
        self._epochs: Dict[Tuple[int, str], int] = {}
        # all-gather transport: "ce" = peer cudaMemcpyAsync on the copy engines (default: zero SMs), "pull" = SM pull kernel
        self.ag_impl = os.environ.get("VESCALE_B200_AG_IMPL", "ce")
        self.ag_ctas = int(os.environ.get("VESCALE_B200_AG_CTAS", "0"))
        self.rs_ctas = int(os.environ.get("VESCALE_B200_RS_CTAS", "0"))

    # ------------------------------------------------------------------ memory
    def alloc(self, numel: int, dtype: torch.dtype) -> torch.Tensor:
        t = self.arena.alloc(numel, dtype)
        t.zero_()
        return t

    def _slots(self, unit) -> Tuple[int, int, int]:
        k = id(unit)
        if k not in self._unit_slots:
            s = self.arena.new_slots(3)
            self._unit_slots[k] = (s, s + 1, s + 2)  # ag-ready, rs-ready, rs-done
        return self._unit_slots[k]

    def _bump(self, unit, kind: str) -> int:
        k = (id(unit), kind)
        self._epochs[k] = self._epochs.get(k, 0) + 1
        return self._epochs[k]

    # ------------------------------------------------------------------ collectives (launch on the current stream)
    def all_gather(self, shard: torch.Tensor, full: torch.Tensor, unit, *, only: Optional[Tuple[int, int]] = None, skip: Optional[Tuple[int, int]] = None,
                   handshake: bool = True) -> None:
        """Pull all-gather of the unit.  ``only=(lo, hi)`` gathers just that element range of the flat unit (small parameters
        needed before a fused first GEMM); ``skip=(lo, hi)`` gathers everything else (the fused kernel gathers that range)."""
        ag_slot, _, _ = self._slots(unit)
        epoch = self._bump(unit, "ag")
        esz = shard.element_size()
        mode, lo, hi = (1, *only) if only is not None else (2, *skip) if skip is not None else (0, 0, 0)
        if mode == 2 and (lo * esz) % 16:
            raise ValueError("skip range must start on a 16-byte boundary")
        if mode == 2:
            hi = hi * esz // 16 * 16 // esz  # never skip a partially covered vector
        if self.ag_impl == "ce":
            # copy engines (peer cudaMemcpyAsync): no SM is taken from the GEMMs the gather overlaps
            _ext.count_launch("symm_all_gather_ce")
            args = (self.arena.peer_ptrs(shard), full, shard.numel() * esz, self.rank)
            pads = self.arena.pad_ptrs if handshake else []
            if mode == 0:
                self.ops.symm_all_gather_ce(*args, pads, ag_slot, epoch, 0, 0)
            elif mode == 1:
                self.ops.symm_all_gather_ce(*args, pads, ag_slot, epoch, lo * esz, hi * esz)
            else:
                total = full.numel() * esz
                # [0, lo) with the handshake (an empty range, lo == hi == 1, still exchanges the flags), then [hi, total)
                self.ops.symm_all_gather_ce(*args, pads, ag_slot, epoch, 0 if lo > 0 else 1, lo * esz if lo > 0 else 1)
                if hi * esz < total:
                    self.ops.symm_all_gather_ce(*args, [], ag_slot, epoch, hi * esz, total)
            return
        _ext.count_launch("symm_all_gather")
        self.ops.symm_all_gather(
            self.arena.peer_ptrs(shard), full, shard.numel() * esz, self.rank, self.arena.pad_ptrs if handshake else [], ag_slot, epoch, self.ag_ctas, mode,
            lo * esz // 16 * 16, hi * esz,
        )

    # ------------------------------------------------------------------ all-gather ⊕ first GEMM of the unit
    def fusable_slot(self, unit, name: str):
        """The layout slot of weight ``name`` if the fused kernel can gather it: 2-D bf16, N % 256 == 0, K % 256 == 0 and every
        rank boundary inside it on a multiple of 32 rows (``fully_shard(..., block_rows=32)``)."""
        slot = unit.layout.slot(name)
        if len(slot.shape) != 2 or unit.param_dtype != torch.bfloat16:
            return None
        N, K = slot.shape
        if N % 256 or K % 256 or slot.offset % 8:
            return None
        S = unit.layout.shard_size
        for p in range(self.world + 1):
            b = min(max(p * S, slot.offset), slot.end) - slot.offset
            if b % (32 * K):
                return None
        return slot

    def fused_first_linear(self, x: torch.Tensor, unit, slot, full: torch.Tensor, handshake: bool = True) -> torch.Tensor:
        """y = x @ W^T where W (``slot``) is gathered from the peers' parameter shards *by the GEMM kernel itself*, straight into
        its place in ``full``; the math starts as soon as the first 256 rows have arrived."""
        N, K = slot.shape
        M = x.shape[0]
        if x.dtype != torch.bfloat16 or not x.is_contiguous() or M % 256:
            raise ValueError("fused_first_linear needs a contiguous bf16 activation with a multiple of 256 rows")
        S, W = unit.layout.shard_size, self.world
        key = (id(unit), slot.name)
        st = self._fused_sites.get(key)
        if st is None:
            s0 = self.arena.new_slots(2)
            bounds, ptrs = [0], []
            base = self.arena.peer_ptrs(unit.param_shard)
            for p in range(W):
                lo, hi = min(max(p * S, slot.offset), slot.end), min(max((p + 1) * S, slot.offset), slot.end)
                bounds.append((hi - slot.offset) // K)
                ptrs.append(base[p] + (lo - p * S) * 2 if hi > lo else base[p])
            st = self._fused_sites[key] = {
                "epoch": 0,
                "flag_ptrs": [q + s0 * W * 4 for q in self.arena.pad_ptrs],
                "arrive": torch.zeros(N // 256, dtype=torch.int32, device=self.device),
                "bounds": bounds,
                "ptrs": ptrs,
            }
        st["epoch"] += 1
        w_full = full[slot.offset : slot.end].view(N, K)
        y = torch.empty(M, N, dtype=torch.bfloat16, device=x.device)
        _ext.count_launch("wag_gemm")
        self.ops.wag_gemm(x, w_full, st["ptrs"], st["bounds"], y, st["arrive"], st["flag_ptrs"], self.rank, st["epoch"], handshake)
        return y

    def wait_buffer_free(self, buf: torch.Tensor) -> None:
        """Before a symmetric gradient buffer is reused: every peer must have finished reading it."""
        tag = getattr(buf, "_symm_last_use", None)
        if tag is None:
            return
        done_slot, epoch = tag
        _ext.count_launch("symm_wait")
        self.ops.symm_wait(self.arena.my_pad, self.world, done_slot, epoch)
        buf._symm_last_use = None

    def reduce_scatter(self, full_grad: torch.Tensor, out: torch.Tensor, scale: float, unit) -> None:
        _, rs_slot, done_slot = self._slots(unit)
        epoch = self._bump(unit, "rs")
        if unit.sumsq is None:
            unit.sumsq = torch.zeros(1, dtype=torch.float32, device=self.device)
        else:
            unit.sumsq.zero_()
        mc = self.arena.multicast_ptr(full_grad) if self.use_multimem else 0
        _ext.count_launch("symm_reduce_scatter")
        self.ops.symm_reduce_scatter(
            self.arena.peer_ptrs(full_grad), out, unit.sumsq, out.numel(), self.rank, float(scale), self.arena.pad_ptrs, rs_slot, epoch, mc, self.rs_ctas
        )
        # tell the peers that I am done reading their copy of this buffer
        _ext.count_launch("symm_signal")
        self.ops.symm_signal(self.arena.pad_ptrs, self.rank, done_slot, epoch)
        full_grad._symm_last_use = (done_slot, epoch)

    def reduce_scatter_adamw(self, full_grad: torch.Tensor, unit, scale: float, hp: dict) -> None:
        """Fully fused: reduce-scatter ⊕ scale ⊕ AdamW ⊕ bf16 cast into the all-gather source shard."""
        _, rs_slot, done_slot = self._slots(unit)
        epoch = self._bump(unit, "rs")
        if unit.sumsq is None:
            unit.sumsq = torch.zeros(1, dtype=torch.float32, device=self.device)
        else:
            unit.sumsq.zero_()
        _ext.count_launch("symm_rs_adamw")
        self.ops.symm_rs_adamw(
            self.arena.peer_ptrs(full_grad), unit.master, unit.exp_avg, unit.exp_avg_sq, unit.param_shard, unit.wd_table, hp["coef"], unit.sumsq,
            self.rank, float(scale), self.arena.pad_ptrs, rs_slot, epoch, hp["lr"], hp["b1"], hp["b2"], hp["eps"], hp["wd"], hp["bc1"], hp["bc2"], self.rs_ctas,
        )
        _ext.count_launch("symm_signal")
        self.ops.symm_signal(self.arena.pad_ptrs, self.rank, done_slot, epoch)
        full_grad._symm_last_use = (done_slot, epoch)


# one arena per (process group, device): FSDP units, fused TP layers, MoE dispatch and the tensor collectives all share it
_COMM_CACHE: Dict[Tuple[int, Optional[int]], "SymmUnitComm"] = {}


def get_unit_comm(mesh, mesh_dim: int, device: torch.device) -> "SymmUnitComm":
    key = (id(mesh.get_group(mesh_dim)), device.index)
    comm = _COMM_CACHE.get(key)
    if comm is None:
        comm = _COMM_CACHE[key] = SymmUnitComm(mesh, mesh_dim, device)
    return comm
