"""Tensor-level collectives over symmetric memory (``csrc/symm_collectives.cu``) — no NCCL kernel on these paths.

    sc = SymmCollectives(mesh, "TP")
    sc.all_reduce(x)                               # NVLS two-shot / P2P two-shot / one-shot (small)   — SURVEY §2F C11/C17/C19
    y = sc.reduce_scatter(x)                       # NVLS / P2P pull of my slice, one kernel           — C5/C10/C18
    y = sc.all_to_all_permute(x, i, j)             # Shard(i) -> Shard(j), both permutes folded in     — C12
    y = sc.ragged_exchange(local, src_rng, dst_rng)  # ragged->ragged / scatter / gather-to-root puts — C2/C3/C21
    loss = sc.vocab_parallel_cross_entropy(logits_shard, target)   # one launch, no all-reduce         — C20

``enable_symmetric_collectives(mesh)`` registers an instance per mesh dim; ``vescale_b200.comm.collectives`` (and therefore
``DTensor.redistribute`` / ``loss_parallel``) then routes qualifying CUDA tensors through these kernels.

Reference call sites replaced: legacy ``dtensor/_collective_utils.py:222,354`` (all_to_all_single + two permutes, all_reduce),
reference ``placement_types.py:152-192`` (uneven list all-to-all), ``_collective_utils.py:66-99`` (serialized send/recv scatter),
legacy ``dtensor/loss.py:138,141`` and ``model/patch/vp_cross_entropy.py:47,79,84`` (two all-reduces per CE).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ..ops import _ext
from .symm import SymmArena, get_unit_comm

__all__ = ["SymmCollectives", "enable_symmetric_collectives", "disable_symmetric_collectives", "symm_backend_for"]

_ONESHOT_BYTES = 512 * 1024
_DT = {torch.float32: 0, torch.bfloat16: 1}


class SymmCollectives:
    """Tensor collectives of one mesh dim over the shared symmetric arena (see the module docstring)."""
    def __init__(self, mesh, mesh_dim=0, device: Optional[torch.device] = None):
        md = mesh._dim_index(mesh_dim)
        dev = device or torch.device("cuda", torch.cuda.current_device())
        comm = get_unit_comm(mesh, md, dev)
        self.mesh, self.md, self.device = mesh, md, dev
        self.arena: SymmArena = comm.arena
        self.world, self.rank = comm.world, comm.rank
        # NVLS two-shot wins from 4 ranks up (profiles/coll_bench_w8_r1.json: 1.2-1.6x NCCL at >= 2 MB); between two GPUs the
        # P2P two-shot is faster than going through the switch (coll_bench_w2_r1.json)
        self.use_multimem = comm.use_multimem and self.world > 2
        self.ops = _ext.ops()
        self.slot = self.arena.new_slots(2)
        self.epoch = 0
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self._staging: Dict[int, torch.Tensor] = {}
        self._ce: Optional[dict] = None
        self.num_ctas = 0

    # ------------------------------------------------------------------ staging
    def reserve(self, nbytes: int) -> None:
        """Allocate the symmetric staging block up front (a rendezvous is collective: do it outside the step)."""
        self._stage(nbytes)

    def _stage(self, nbytes: int) -> torch.Tensor:
        bucket = max(1 << 16, 1 << (max(1, nbytes) - 1).bit_length())
        buf = self._staging.get(bucket)
        if buf is None:
            buf = self._staging[bucket] = self.arena.alloc(bucket, torch.uint8)
        return buf

    def _next(self) -> int:
        self.epoch += 1
        return self.epoch

    def empty(self, shape, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
        """A tensor in symmetric memory (same call sequence on every rank).  Collectives on such tensors are zero-copy: no
        staging pass in, none out."""
        shape = tuple(shape) if not isinstance(shape, int) else (shape,)
        return self.arena.alloc(math.prod(shape), dtype).view(shape)

    def _is_symmetric(self, x: torch.Tensor) -> bool:
        try:
            self.arena._locate(x)
            return True
        except ValueError:
            return False

    # ------------------------------------------------------------------ all-reduce
    def all_reduce(self, x: torch.Tensor, op: str = "sum") -> torch.Tensor:
        """Sum / avg of a contiguous bf16 or fp32 CUDA tensor over the mesh dim; returns the reduced tensor (``x`` itself, reduced
        in place, except for small tensors that already live in symmetric memory: those are reduced one-shot into a new tensor)."""
        if x.dtype not in _DT or not x.is_contiguous() or op not in ("sum", "avg"):
            raise ValueError("symmetric all_reduce handles contiguous bf16/fp32 sum/avg")
        if x.numel() == 0:
            return x
        nbytes = x.numel() * x.element_size()
        if nbytes % 16 == 0 and x.data_ptr() % 16 == 0 and self._is_symmetric(x):  # zero-copy
            scale = 1.0 / self.world if op == "avg" else 1.0
            ptrs = self.arena.peer_ptrs(x)
            _ext.count_launch("symm_all_reduce")
            if nbytes <= _ONESHOT_BYTES:
                out = torch.empty_like(x)
                self.ops.symm_all_reduce(ptrs, 0, out, x.numel(), _DT[x.dtype], scale, self.rank, self.arena.pad_ptrs, self.slot, self._next(), self.counter, self.num_ctas)
                return out
            mc = self.arena.multicast_ptr(x) if self.use_multimem else 0
            self.ops.symm_all_reduce(ptrs, mc, None, x.numel(), _DT[x.dtype], scale, self.rank, self.arena.pad_ptrs, self.slot, self._next(), self.counter, self.num_ctas)
            return x
        padded = (nbytes + 15) // 16 * 16
        st = self._stage(padded)
        flat = x.view(-1).view(torch.uint8)
        st[:nbytes].copy_(flat)
        scale = 1.0 / self.world if op == "avg" else 1.0
        numel_p = padded // x.element_size()
        ptrs = self.arena.peer_ptrs(st)
        _ext.count_launch("symm_all_reduce")
        if padded <= _ONESHOT_BYTES:
            out = flat if padded == nbytes else torch.empty(padded, dtype=torch.uint8, device=x.device)
            self.ops.symm_all_reduce(ptrs, 0, out, numel_p, _DT[x.dtype], scale, self.rank, self.arena.pad_ptrs, self.slot, self._next(), self.counter, self.num_ctas)
            if out is not flat:
                flat.copy_(out[:nbytes])
        else:
            mc = self.arena.multicast_ptr(st) if self.use_multimem else 0
            self.ops.symm_all_reduce(ptrs, mc, None, numel_p, _DT[x.dtype], scale, self.rank, self.arena.pad_ptrs, self.slot, self._next(), self.counter, self.num_ctas)
            flat.copy_(st[:nbytes])
        return x

    # ------------------------------------------------------------------ reduce-scatter
    def reduce_scatter(self, x: torch.Tensor, op: str = "sum", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``x`` [W * rows, ...] (contiguous bf16 / fp32) holds this rank's partial values; returns rows ``[rank * rows,
        (rank + 1) * rows)`` of the sum (or mean) over the mesh dim as a private tensor.  One kernel: with NVLS the switch adds
        (``multimem.ld_reduce``), otherwise every peer's copy of my slice is pulled over NVLink.  ``x`` in symmetric memory
        (``empty``) is read in place; other tensors are staged once.  ``out`` may be a column slice of a wider 2-D tensor."""
        if x.dtype not in _DT or not x.is_contiguous() or op not in ("sum", "avg"):
            raise ValueError("symmetric reduce_scatter handles contiguous bf16/fp32 sum/avg")
        W = self.world
        if x.ndim == 0 or x.shape[0] % W:
            raise ValueError("reduce_scatter needs a leading dim divisible by the group size")
        rows = x.shape[0] // W
        row = math.prod(x.shape[1:])
        esz = x.element_size()
        if out is None:
            out = torch.empty((rows, *x.shape[1:]), dtype=x.dtype, device=x.device)
        if out.dtype != x.dtype or tuple(out.shape) != (rows, *x.shape[1:]):
            raise ValueError("reduce_scatter: out must be [rows, ...] of x's dtype")
        if rows == 0 or row == 0:
            return out
        if out.is_contiguous():
            ostride = row
        elif out.ndim == 2 and out.stride(1) == 1:
            ostride = out.stride(0)
        else:
            raise ValueError("reduce_scatter: out must be contiguous or a column slice of a 2-D tensor")
        if (row * esz) % 16 or (ostride * esz) % 16 or out.data_ptr() % 16:
            raise ValueError("reduce_scatter: rows, the output row stride and the output base must be multiples of 16 bytes")
        if x.data_ptr() % 16 == 0 and self._is_symmetric(x):
            src = x
        else:
            nbytes = x.numel() * esz
            src = self._stage(nbytes)[:nbytes]
            src.copy_(x.view(-1).view(torch.uint8))
        scale = 1.0 / W if op == "avg" else 1.0
        mc = self.arena.multicast_ptr(src) if self.use_multimem else 0
        _ext.count_launch("symm_reduce_scatter_t")
        self.ops.symm_reduce_scatter_t(self.arena.peer_ptrs(src), mc, out, rows * row, row, ostride, _DT[x.dtype], scale, self.rank, self.arena.pad_ptrs, self.slot,
                                       self._next(), self.counter, self.num_ctas)
        return out

    # ------------------------------------------------------------------ Shard(i) -> Shard(j)
    def all_to_all_permute(self, x: torch.Tensor, src_shard_dim: int, dst_shard_dim: int) -> torch.Tensor:
        """``x`` is my block of a tensor sharded on ``src_shard_dim``; returns my block of the same tensor sharded on
        ``dst_shard_dim`` (which must divide evenly).  One put kernel; the chunk/stack/cat permutes never materialise."""
        i, j, W = src_shard_dim % x.ndim, dst_shard_dim % x.ndim, self.world
        if i == j:
            return x.clone()
        x = x.contiguous()
        s = list(x.shape)
        if s[j] % W:
            raise ValueError("all_to_all_permute needs an evenly divisible destination dim")
        esz = x.element_size()
        I, J = s[i], s[j]
        Jl = J // W
        if i < j:
            a, b, c = math.prod(s[:i]), math.prod(s[i + 1 : j]), math.prod(s[j + 1 :])
            ss = [I * b * J * c, b * J * c, J * c, c, 1]
            ds = [I * W * b * Jl * c, b * Jl * c, Jl * c, c, 1]
            sps, drs = Jl * c, I * b * Jl * c
        else:
            a, b, c = math.prod(s[:j]), math.prod(s[j + 1 : i]), math.prod(s[i + 1 :])
            ss = [J * b * I * c, c, I * c, b * I * c, 1]
            ds = [Jl * b * I * W * c, c, I * W * c, b * I * W * c, 1]
            sps, drs = Jl * b * I * c, I * c
        n = [a, I, b, Jl, c]
        out_shape = list(s)
        out_shape[i], out_shape[j] = I * W, Jl
        if esz < 2:
            raise ValueError("all_to_all_permute: element size below 2 bytes is not supported")
        vec = next(v for v in (16, 8, 4, 2) if (c * esz) % v == 0)

        def to_vec(v: int) -> int:  # elements -> vec units; exact because every stride carries the factor c
            return v * esz // vec

        n[4] = to_vec(c)
        ss = [to_vec(v) for v in ss[:4]] + [1]
        ds = [to_vec(v) for v in ds[:4]] + [1]
        nbytes = x.numel() * esz
        st = self._stage(nbytes)
        _ext.count_launch("symm_a2a_permute")
        self.ops.symm_a2a_permute(x, self.arena.peer_ptrs(st), n, ss, ds, to_vec(sps), to_vec(drs), vec, self.rank, self.arena.pad_ptrs, self.slot,
                                  self._next(), self.counter, self.num_ctas)
        return st[:nbytes].view(x.dtype).view(out_shape).clone()

    # ------------------------------------------------------------------ flat interval exchange
    def ragged_exchange(self, local: torch.Tensor, src_ranges: Sequence[Tuple[int, int]], dst_ranges: Sequence[Tuple[int, int]]) -> torch.Tensor:
        """``local`` holds flat elements ``src_ranges[rank]`` of a global buffer; returns flat elements ``dst_ranges[rank]``.
        Every rank puts the intersections of its source interval with each destination interval straight into the owner's
        symmetric block (ragged->ragged redistribute; one-hot ``dst_ranges`` = gather-to-root for Muon; one-hot
        ``src_ranges`` = scatter from a source rank)."""
        W, me = self.world, self.rank
        esz = local.element_size()
        local = local.contiguous().view(-1)
        s_lo, s_hi = src_ranges[me]
        segs: List[List[int]] = []
        for p in range(W):
            d_lo, d_hi = dst_ranges[p]
            lo, hi = max(s_lo, d_lo), min(s_hi, d_hi)
            if hi > lo:
                segs.append([lo - s_lo, p, lo - d_lo, hi - lo])
        max_out = max(hi - lo for lo, hi in dst_ranges)
        my_out = dst_ranges[me][1] - dst_ranges[me][0]
        st = self._stage(max(16, max_out * esz))
        # widest vector that divides every offset and length (base pointers are >= 1 KiB aligned)
        vec = 16
        for sg in segs:
            for v in (sg[0] * esz, sg[2] * esz, sg[3] * esz):
                while v % vec:
                    vec //= 2
        if local.data_ptr() % 16:
            vec = min(vec, esz)
        vec = max(vec, 1)
        if vec == 8:
            vec = 4
        tab = torch.tensor([[sg[0] * esz // vec, sg[1], sg[2] * esz // vec, sg[3] * esz // vec] for sg in segs] or [[0, me, 0, 0]], dtype=torch.int64)
        tab = tab.to(local.device, non_blocking=True)
        total = int(sum(sg[3] for sg in segs) * esz // vec)
        src = local if local.numel() else torch.empty(16, dtype=local.dtype, device=local.device)
        _ext.count_launch("symm_put_segments")
        self.ops.symm_put_segments(src, self.arena.peer_ptrs(st), tab, vec, max(total, 1), self.rank, self.arena.pad_ptrs, self.slot, self._next(), self.counter,
                                   self.num_ctas)
        return st[: my_out * esz].view(local.dtype).clone()

    # ------------------------------------------------------------------ vocab-parallel cross entropy
    def _ce_block(self, T: int) -> dict:
        if self._ce is None or self._ce["max_rows"] < T:
            max_rows = max(8192, 1 << (T - 1).bit_length())
            max_ctas = 2 * torch.cuda.get_device_properties(self.device).multi_processor_count
            floats = 2 * self.world * max_rows * 4 + self.world * max_ctas
            blk = self.arena.alloc(floats, torch.float32)
            blk.zero_()
            torch.cuda.current_stream().synchronize()
            import torch.distributed as dist

            dist.barrier(group=self.arena.group, device_ids=[self.device.index])
            self._ce = {"blk": blk, "ptrs": self.arena.peer_ptrs(blk), "max_rows": max_rows, "max_ctas": max_ctas, "epoch": 0}
        return self._ce

    def vocab_ce_fwd_bwd_(self, logits: torch.Tensor, target: torch.Tensor, n_valid: torch.Tensor, vocab_start: int, ignore_index: int = -100) -> torch.Tensor:
        """Per-row loss; ``logits`` (this rank's [T, V/W] bf16 slice) is overwritten with d(mean loss)/d(logits)."""
        ce = self._ce_block(logits.shape[0])
        ce["epoch"] += 1
        _ext.count_launch("symm_vocab_ce")
        return self.ops.symm_vocab_ce_(logits, target, n_valid, int(vocab_start), int(ignore_index), ce["ptrs"], self.rank, ce["epoch"], ce["max_rows"], ce["max_ctas"])

    def vocab_parallel_cross_entropy(self, logits_shard: torch.Tensor, target: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
        """Mean cross entropy over a vocabulary sharded evenly on this mesh dim.  NOTE: consumes ``logits_shard`` (contiguous
        bf16), which the forward kernel overwrites with the gradient, like ``ops.functional.cross_entropy``."""
        if not (logits_shard.is_contiguous() and logits_shard.dtype == torch.bfloat16):
            raise ValueError("vocab_parallel_cross_entropy needs a contiguous bf16 logits shard")
        return _VocabCE.apply(logits_shard, target, self, ignore_index)


class _VocabCE(torch.autograd.Function):
    """Same contract as ``ops.functional._CrossEntropy``: consumes the logits buffer (overwritten with the gradient)."""

    @staticmethod
    def forward(ctx, logits, target, sc: SymmCollectives, ignore_index):
        V = logits.shape[-1]
        l2 = logits.detach().view(-1, V)
        t = target.reshape(-1).contiguous()
        n_valid = (t != ignore_index).sum().to(torch.float32).clamp_(min=1.0).reshape(1)
        loss_rows = sc.vocab_ce_fwd_bwd_(l2, t, n_valid, sc.rank * V, ignore_index)
        ctx.grad_buf = l2
        ctx.shape = logits.shape
        return loss_rows.sum() / n_valid[0]

    @staticmethod
    def backward(ctx, g):
        grad = ctx.grad_buf
        if not getattr(ctx, "scaled", False):
            grad.mul_(g.to(grad.dtype))
            ctx.scaled = True
        return grad.view(ctx.shape), None, None, None


# --------------------------------------------------------------------------- registry consulted by comm.collectives
_REGISTRY: Dict[int, SymmCollectives] = {}


def enable_symmetric_collectives(mesh, mesh_dims: Optional[Sequence] = None, reserve_bytes: int = 0) -> List[SymmCollectives]:
    """Route DTensor redistribute / loss-parallel collectives on these mesh dims through the symmetric-memory kernels."""
    dims = range(mesh.ndim) if mesh_dims is None else [mesh._dim_index(d) for d in mesh_dims]
    out = []
    for d in dims:
        if mesh.size(d) == 1:
            continue
        g = mesh.get_group(d)
        sc = _REGISTRY.get(id(g))
        if sc is None:
            sc = _REGISTRY[id(g)] = SymmCollectives(mesh, d)
        if reserve_bytes:
            sc.reserve(reserve_bytes)
        out.append(sc)
    return out


def disable_symmetric_collectives() -> None:
    _REGISTRY.clear()


def symm_backend_for(group, tensor: Optional[torch.Tensor] = None) -> Optional[SymmCollectives]:
    sc = _REGISTRY.get(id(group))
    if sc is None or (tensor is not None and not tensor.is_cuda):
        return None
    return sc
