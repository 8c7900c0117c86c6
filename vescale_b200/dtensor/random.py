"""Random number generation for sharded tensors.

Two mechanisms:

* ``sharded_random_fill`` — *single-device-equivalent* init: the values a rank writes into its shard are
  exactly the values a single device would have produced for those global positions.  On CUDA this is the
  counter-based Philox4x32-10 kernel in ``csrc/philox_shard.cu`` (element g uses counter g/4, lane g%4 —
  no per-element ``curand_init`` skip-ahead as in the reference's patched aten kernels,
  ``legacy/patches/patched_pytorch_v2.2.1_rc3.patch:299-448``); on CPU the same function is evaluated in
  pure torch integer ops so CPU and GPU agree bit-for-bit.
* ``rng_region`` — for aten random ops running on shards (dropout, ``normal_``...): an offset-based tracker
  gives every distinct shard its own Philox offset window and replicas the same one
  (``legacy/vescale/dtensor/random.py:167`` OffsetBasedRNGTracker).

``manual_seed(seed, mesh)`` sets both (``legacy/vescale/dtensor/random.py:62``).
"""
from __future__ import annotations

import contextlib
import math
from typing import Optional

import torch
import torch.distributed as dist

from ..layout import compute_local_shape_and_global_offset, local_boxes
from ..placement import RaggedShard, Shard
from ..spec import DTensorSpec

__all__ = ["manual_seed", "rng_region", "sharded_random_fill", "philox_uniform_reference", "get_rng_state", "set_rng_state", "RNGStateTracker",
           "OffsetBasedRNGTracker", "ThreadBasedRNGTracker", "TensorParallelRNGTracker", "get_rng_tracker", "set_rng_tracker"]

_STATE = {"seed": 0, "offset": 0, "initialised": False}


def manual_seed(seed: int, device_mesh=None, *, check: bool = False) -> None:
    """Seed the sharded RNG.  All ranks must pass the same seed (checked with an object all-gather when
    ``check``; legacy always checked, ``random.py:90``)."""
    if check and dist.is_initialized():
        seeds = [None] * dist.get_world_size()
        dist.all_gather_object(seeds, int(seed))
        if len(set(seeds)) != 1:
            raise RuntimeError(f"manual_seed must be called with the same seed on every rank, got {seeds}")
    _STATE.update(seed=int(seed), offset=0, initialised=True)
    torch.manual_seed(int(seed))


def get_rng_state():
    return dict(_STATE)


def set_rng_state(st) -> None:
    _STATE.update(st)


# ------------------------------------------------------------------------------- Philox4x32-10 in torch ints
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def _mulhilo(a: int, b: torch.Tensor):
    # b, a < 2^32: the 64-bit product overflows signed int64, so multiply by the two 16-bit halves of ``a``
    a_lo, a_hi = a & 0xFFFF, a >> 16
    bl = b.to(torch.int64)
    p_lo = bl * a_lo  # < 2^48
    p_hi = bl * a_hi  # < 2^48
    lo = (p_lo + ((p_hi & 0xFFFF) << 16)) & _MASK
    carry = (p_lo + ((p_hi & 0xFFFF) << 16)) >> 32
    hi = ((p_hi >> 16) + carry) & _MASK
    return hi, lo


def philox4x32(counter_lo: torch.Tensor, counter_hi: torch.Tensor, key0: int, key1: int, sub0: int = 0, sub1: int = 0):
    """Vectorised Philox4x32-10.  Counter = (c0=counter_lo, c1=counter_hi, c2=sub0, c3=sub1); returns 4 x uint32 (as int64)."""
    c0 = counter_lo.to(torch.int64) & _MASK
    c1 = counter_hi.to(torch.int64) & _MASK
    c2 = torch.full_like(c0, sub0 & _MASK)
    c3 = torch.full_like(c0, sub1 & _MASK)
    k0, k1 = key0 & _MASK, key1 & _MASK
    for _ in range(10):
        hi0, lo0 = _mulhilo(_M0, c0)
        hi1, lo1 = _mulhilo(_M1, c2)
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def philox_uniform_reference(global_index: torch.Tensor, seed: int, offset: int) -> torch.Tensor:
    """U[0,1) float32 for each *global linear element index* (int64 tensor).  Element g uses Philox counter
    (g//4 + offset) and output lane g%4.  This is the specification the CUDA kernel implements."""
    g = global_index.to(torch.int64)
    ctr = g // 4 + int(offset)
    lane = g % 4
    r = philox4x32(ctr & _MASK, (ctr >> 32) & _MASK, seed & _MASK, (seed >> 32) & _MASK)
    bits = torch.where(lane == 0, r[0], torch.where(lane == 1, r[1], torch.where(lane == 2, r[2], r[3])))
    # 24 high bits -> [0,1)
    return (bits >> 8).to(torch.float32) * (1.0 / 16777216.0)


def philox_normal_reference(global_index: torch.Tensor, seed: int, offset: int) -> torch.Tensor:
    """N(0,1) via Box-Muller on two independent uniforms drawn from counters (g, sub=0) and (g, sub=1)."""
    g = global_index.to(torch.int64)
    ctr = g // 4 + int(offset)
    lane = g % 4

    def pick(r):
        return torch.where(lane == 0, r[0], torch.where(lane == 1, r[1], torch.where(lane == 2, r[2], r[3])))

    r0 = philox4x32(ctr & _MASK, (ctr >> 32) & _MASK, seed & _MASK, (seed >> 32) & _MASK, 0)
    r1 = philox4x32(ctr & _MASK, (ctr >> 32) & _MASK, seed & _MASK, (seed >> 32) & _MASK, 1)
    u1 = ((pick(r0) >> 8).to(torch.float32) + 1.0) * (1.0 / 16777216.0)  # (0,1]
    u2 = (pick(r1) >> 8).to(torch.float32) * (1.0 / 16777216.0)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(6.283185307179586 * u2)


def _global_linear_indices(spec: DTensorSpec, device) -> torch.Tensor:
    """Global linear (row-major) index of every local element, in local storage order."""
    shape = tuple(spec.shape)
    gst = [1] * len(shape)
    for i in range(len(shape) - 2, -1, -1):
        gst[i] = gst[i + 1] * shape[i + 1]
    parts = []
    for off, sz, _ in local_boxes(shape, spec.mesh, spec.placements):
        idx = torch.zeros(sz, dtype=torch.int64, device=device) if len(sz) else torch.zeros((), dtype=torch.int64, device=device)
        for d, (o, n) in enumerate(zip(off, sz)):
            view = [1] * len(sz)
            view[d] = n
            idx = idx + (torch.arange(o, o + n, device=device, dtype=torch.int64) * gst[d]).view(view)
        parts.append(idx.reshape(-1))
    if not parts:
        return torch.zeros(0, dtype=torch.int64, device=device)
    return torch.cat(parts) if len(parts) > 1 else parts[0]


def sharded_random_fill(local: torch.Tensor, spec: DTensorSpec, kind: str = "uniform", *, low: float = 0.0, high: float = 1.0, mean: float = 0.0, std: float = 1.0) -> torch.Tensor:
    """Fill ``local`` (the shard described by ``spec``) with the single-device-equivalent random stream and
    advance the global offset by ceil(global_numel/4)."""
    seed, offset = _STATE["seed"], _STATE["offset"]
    numel = math.prod(spec.shape)
    _STATE["offset"] = offset + (numel + 3) // 4
    if local.numel() == 0:
        return local
    boxes = local_boxes(tuple(spec.shape), spec.mesh, spec.placements)
    if local.is_cuda:
        from ..ops import philox as _ph

        if _ph.available():
            _ph.philox_fill_boxes(local, tuple(spec.shape), boxes, seed, offset, kind, low, high, mean, std, ragged=spec.is_ragged_shard())
            return local
    gi = _global_linear_indices(spec, local.device)
    if kind == "uniform":
        vals = philox_uniform_reference(gi, seed, offset) * (high - low) + low
    else:
        vals = philox_normal_reference(gi, seed, offset) * std + mean
    if spec.is_ragged_shard() or len(boxes) == 1:
        local.view(-1).copy_(vals.to(local.dtype)) if local.is_contiguous() else local.copy_(vals.view(local.shape).to(local.dtype))
    else:
        # multi-box (interleaved) local storage: boxes tile the local tensor in order along the split dims
        pos = 0
        for off, sz, loc in boxes:
            n = math.prod(sz)
            t = local
            for d, (o, m) in enumerate(zip(loc, sz)):
                t = t.narrow(d, o, m)
            t.copy_(vals[pos : pos + n].view(sz).to(local.dtype))
            pos += n
    return local


# ------------------------------------------------------------------------------- offset tracker for aten random ops
def _shard_linear_index(spec: DTensorSpec) -> int:
    """Index of my shard among all distinct shards (replicas share an index)."""
    coord = spec.mesh.get_coordinate()
    idx, mult = 0, 1
    for i in reversed(range(spec.mesh.ndim)):
        p = spec.placements[i]
        if isinstance(p, (Shard, RaggedShard)):
            idx += coord[i] * mult
            mult *= spec.mesh.size(i)
    return idx


@contextlib.contextmanager
def rng_region(spec: DTensorSpec):
    """Run an aten random op on a shard: distinct shards draw from disjoint offset windows, replicas from
    the same one; afterwards every rank's generator sits at the same post-op offset."""
    local_shape, _ = compute_local_shape_and_global_offset(spec.shape, spec.mesh, spec.placements)
    local_numel = math.prod(local_shape)
    window = ((local_numel + 3) // 4) * 4
    shard = _shard_linear_index(spec)
    nshards = spec.num_shards
    dev = spec.mesh.device_type
    if dev == "cuda" and torch.cuda.is_available():
        gen = torch.cuda.default_generators[torch.cuda.current_device()]
        base = gen.get_offset()
        gen.set_offset(base + shard * window)
        try:
            yield
        finally:
            gen.set_offset(base + nshards * window)
    else:
        base_seed = _STATE["seed"] * 1000003 + _STATE["offset"]
        with torch.random.fork_rng(devices=[]):
            torch.manual_seed((base_seed + shard * 7919) & 0x7FFFFFFFFFFF)
            try:
                yield
            finally:
                pass
        _STATE["offset"] += (nshards * window) // 4


# ------------------------------------------------------------------------------- tracker objects (legacy API shape)
class RNGStateTracker:
    """How aten random ops (dropout, ``uniform_``, ``normal_``, ``rand_like`` ...) behave on shards.  ``run`` returns
    ``NotImplemented`` for ops the tracker does not special-case; the dispatcher then runs the op inside ``region``."""

    single_device_equivalent = False

    def manual_seed(self, seed: int, device_mesh=None) -> None:
        manual_seed(seed, device_mesh)

    def region(self, spec: DTensorSpec):
        return rng_region(spec)

    def run(self, op, local_args, local_kwargs, spec: DTensorSpec):
        return NotImplemented


class OffsetBasedRNGTracker(RNGStateTracker):
    """Same seed everywhere; each distinct shard draws from its own Philox offset window, replicas share one
    (legacy ``dtensor/random.py:167``; torch DTensor's scheme).  Results depend on the sharding."""


class ThreadBasedRNGTracker(RNGStateTracker):
    """Single-device-equivalent randomness: the value at a global position does not depend on how the tensor is sharded
    (legacy ``dtensor/random.py:340-518`` — there by packing the shard geometry into the CUDA generator state for patched
    aten kernels; here the counter-based Philox of ``sharded_random_fill`` is evaluated per global index)."""

    single_device_equivalent = True

    def run(self, op, local_args, local_kwargs, spec: DTensorSpec):
        aten = torch.ops.aten
        x = local_args[0]
        if op is aten.native_dropout.default:
            p, train = float(local_args[1]), local_args[2]
            if not train or p == 0.0:
                return x.clone(), torch.ones_like(x, dtype=torch.bool)
            if x.is_cuda and x.is_contiguous() and x.numel() and p < 1.0:
                from ..ops import philox as _ph

                if _ph.dropout_available():  # one fused pass: Philox + mask + scale (csrc/philox_shard.cu)
                    seed, offset = _STATE["seed"], _STATE["offset"]
                    _STATE["offset"] = offset + (math.prod(spec.shape) + 3) // 4
                    boxes = local_boxes(tuple(spec.shape), spec.mesh, spec.placements)
                    return _ph.philox_dropout_boxes(x, tuple(spec.shape), boxes, seed, offset, p, ragged=spec.is_ragged_shard())
            u = sharded_random_fill(torch.empty(x.shape, dtype=torch.float32, device=x.device), spec, "uniform")
            mask = u >= p
            return x * mask.to(x.dtype) * (1.0 / (1.0 - p)), mask
        if op is aten.uniform_.default:
            lo = float(local_args[1]) if len(local_args) > 1 else float(local_kwargs.get("from", 0.0))
            hi = float(local_args[2]) if len(local_args) > 2 else float(local_kwargs.get("to", 1.0))
            return sharded_random_fill(x, spec, "uniform", low=lo, high=hi)
        if op is aten.normal_.default:
            mean = float(local_args[1]) if len(local_args) > 1 else float(local_kwargs.get("mean", 0.0))
            std = float(local_args[2]) if len(local_args) > 2 else float(local_kwargs.get("std", 1.0))
            return sharded_random_fill(x, spec, "normal", mean=mean, std=std)
        if op is aten.rand_like.default:
            return sharded_random_fill(torch.empty_like(x), spec, "uniform")
        if op is aten.randn_like.default:
            return sharded_random_fill(torch.empty_like(x), spec, "normal")
        if op is aten.bernoulli_.float:
            p = float(local_args[1]) if len(local_args) > 1 else float(local_kwargs.get("p", 0.5))
            u = sharded_random_fill(torch.empty(x.shape, dtype=torch.float32, device=x.device), spec, "uniform")
            return x.copy_((u < p).to(x.dtype))
        if op is aten.bernoulli.default:
            u = sharded_random_fill(torch.empty(x.shape, dtype=torch.float32, device=x.device), spec, "uniform")
            return (u < x.float()).to(x.dtype)
        return NotImplemented


class TensorParallelRNGTracker(RNGStateTracker):
    """Megatron-style: ranks of the tensor-parallel mesh dim draw from *different* streams (dropout inside a TP region),
    every other mesh dim shares one (legacy ``dtensor/random.py:521``)."""

    def __init__(self, tp_mesh_dim=-1, offset: int = 2718):
        self.tp_mesh_dim, self.offset = tp_mesh_dim, offset

    @contextlib.contextmanager
    def region(self, spec: DTensorSpec):
        d = self.tp_mesh_dim if self.tp_mesh_dim >= 0 else spec.mesh.ndim + self.tp_mesh_dim
        tp_rank = spec.mesh.get_coordinate()[d]
        devices = [torch.cuda.current_device()] if (spec.mesh.device_type == "cuda" and torch.cuda.is_available()) else []
        with torch.random.fork_rng(devices=devices):
            torch.manual_seed((_STATE["seed"] + self.offset + tp_rank + 1000003 * _STATE["offset"]) & 0x7FFFFFFFFFFF)
            yield
        _STATE["offset"] += 1


_TRACKER = {"t": None}


def get_rng_tracker() -> RNGStateTracker:
    if _TRACKER["t"] is None:
        import os

        # the legacy default: VESCALE_SINGLE_DEVICE_RAND=1 selects the thread-based (single-device-equivalent) tracker
        _TRACKER["t"] = ThreadBasedRNGTracker() if os.environ.get("VESCALE_SINGLE_DEVICE_RAND", "0") == "1" else OffsetBasedRNGTracker()
    return _TRACKER["t"]


def set_rng_tracker(tracker: Optional[RNGStateTracker]) -> None:
    _TRACKER["t"] = tracker


def init_vescale_rng_tracker(device_type: str = "cuda") -> RNGStateTracker:
    """Create and install the process-wide tracker: thread-based (single-device-equivalent sampling) when
    ``VESCALE_SINGLE_DEVICE_RAND=1``, offset-based otherwise (legacy ``dtensor/random.py:30``)."""
    set_rng_tracker(None)
    return get_rng_tracker()


def is_rng_supported_mesh(device_mesh) -> bool:
    """DTensor random ops work on every mesh here: the Philox stream is evaluated per *global* element index, on the GPU by
    ``csrc/philox_shard.cu`` and on the CPU by the reference implementation in this module (the reference supports CUDA
    meshes only, ``dtensor/random.py:37``)."""
    return device_mesh is not None and device_mesh.device_type in ("cuda", "cpu", "meta")
