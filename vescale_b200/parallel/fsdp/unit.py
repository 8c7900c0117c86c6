"""FSDPUnit: the flat buffers of one FSDP unit and its collectives.

One unit = the parameters of one wrapped module (one decoder layer, the embedding, the head...).  Buffers
(all 1-D; W = world, S = layout.shard_size):

    param_shard  [S]    compute dtype (bf16)  — my slice of the gathered buffer; peers read it over NVLink
    full_param   [W*S]  compute dtype         — gathered parameters; module weights are views into it
    full_grad    [W*S]  compute dtype         — wgrad GEMMs write here directly (``weight.main_grad`` views)
    grad_shard   [S]    reduce dtype (fp32)   — reduce-scattered, pre-scaled gradient of my slice
    master/exp_avg/exp_avg_sq [S] fp32        — optimizer state (owned by the optimizer)

Two communication backends:
  * ``nccl``: ``all_gather_into_tensor`` / ``reduce_scatter_tensor`` over the flat buffers (zero-copy thanks
    to the RaggedShard layout) + separate cast/scale — the NCCL-only baseline.
  * ``symm``: sm_100a kernels over symmetric memory (``vescale_b200.comm.symm``): pull all-gather of peer
    shards, and reduce-scatter ⊕ scale ⊕ cast ⊕ grad-norm (⊕ AdamW) reading the 8 peer gradient slices
    directly over NVLink.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from ...dtensor.api import DTensor
from ...mesh import DeviceMesh
from ...placement import RaggedShard
from ...spec import DTensorSpec, TensorMeta, contiguous_stride
from .layout import ParamSlot, UnitLayout

__all__ = ["FSDPUnit", "MixedPrecisionPolicy"]


class MixedPrecisionPolicy:
    """dtypes of the gathered compute parameters, the reduced gradients and the sharded master weights."""
    def __init__(self, param_dtype: Optional[torch.dtype] = torch.bfloat16, reduce_dtype: Optional[torch.dtype] = torch.float32, master_dtype: torch.dtype = torch.float32):
        self.param_dtype = param_dtype
        self.reduce_dtype = reduce_dtype
        self.master_dtype = master_dtype


class _NullStream:
    def wait_stream(self, s):
        pass

    def wait_event(self, e):
        pass

    def synchronize(self):
        pass


class _NullEvent:
    def record(self, s=None):
        pass

    def wait(self, s=None):
        pass

    def synchronize(self):
        pass

    def query(self):
        return True


def make_event(device: torch.device):
    return torch.cuda.Event() if device.type == "cuda" else _NullEvent()


class FSDPUnit:
    """One FSDP unit = the parameters of one wrapped module in a flat ``UnitLayout`` buffer: fp32 master shard (exposed as
    RaggedShard DTensor parameters), bf16 shard in symmetric memory (all-gather source), gathered compute views, flat gradient
    buffer with ``main_grad`` views."""
    def __init__(
        self,
        module: nn.Module,
        named_params: List[Tuple[str, nn.Parameter]],
        mesh: DeviceMesh,
        mesh_dim: int,
        mp: MixedPrecisionPolicy,
        *,
        name: str = "",
        comm=None,
        block_rows: int = 1,
        granularity_fn=None,
    ):
        self.module = module
        self.name = name
        self.mesh = mesh
        self.mesh_dim = mesh_dim
        self.group = mesh.get_group(mesh_dim) if mesh.has_groups() else None
        self.world = mesh.size(mesh_dim)
        self.rank = mesh.get_local_rank(mesh_dim)
        self.mp = mp
        self.comm = comm
        self.params: List[nn.Parameter] = [p for _, p in named_params]
        self.param_names: List[str] = [n for n, _ in named_params]
        dev = self.params[0].device
        self.device = dev
        orig_dtype = self.params[0].dtype
        self.param_dtype = mp.param_dtype or orig_dtype
        self.reduce_dtype = mp.reduce_dtype or self.param_dtype
        self.master_dtype = mp.master_dtype
        from .layout import row_granularity

        gf = granularity_fn or (lambda n, s: row_granularity(s, block_rows))
        self.layout = UnitLayout([(n, tuple(p.shape)) for n, p in named_params], self.world, granularity_fn=gf)
        S, W = self.layout.shard_size, self.world
        self.S = S

        # ---- master shard initialised from the (replicated) module parameters
        self.master = torch.zeros(S, dtype=self.master_dtype, device=dev)
        for slot, p in zip(self.layout.slots, self.params):
            lo, hi = self.layout.rank_range(slot, self.rank)
            if hi > lo:
                g_lo = self.rank * S + lo - slot.offset
                self.master[lo:hi].copy_(p.detach().reshape(-1)[g_lo : g_lo + (hi - lo)])
        self.param_shard = self._alloc_shard(self.param_dtype, symmetric=True)
        self.param_shard.copy_(self.master)
        self.bf16_fresh = True
        self.full_param: Optional[torch.Tensor] = None
        self.full_grad: Optional[torch.Tensor] = None
        self.grad_shard: Optional[torch.Tensor] = None
        self.grad_ready = False  # grad_shard holds this step's reduced gradient
        self.sumsq: Optional[torch.Tensor] = None  # device scalar: sum of squares of my reduced grad shard
        self.ag_event = None
        self.rs_event = None
        self.unsharded = False
        self.pending_grad_params = 0

        # ---- sharded DTensor view of the master weights (what ``module.parameters()`` exposes outside fwd/bwd)
        self.sharded_params: List[nn.Parameter] = []
        for slot in self.layout.slots:
            lo, hi = self.layout.rank_range(slot, self.rank)
            local = self.master[lo:hi]
            spec = DTensorSpec(mesh if mesh.ndim == 1 else mesh, self._placements(slot), TensorMeta(slot.shape, contiguous_stride(slot.shape), self.master_dtype))
            sp = nn.Parameter(DTensor(local, spec), requires_grad=True)
            sp._fsdp_unit = self
            sp._fsdp_slot = slot
            self.sharded_params.append(sp)
        # the original parameters become the *unsharded compute parameters*: storage freed until all-gather
        for p in self.params:
            p.data = torch.empty(0, dtype=self.param_dtype, device=dev)
            p._fsdp_unit = self

    # ------------------------------------------------------------------ helpers
    def _placements(self, slot: ParamSlot):
        from ...placement import Replicate

        pl = [Replicate() for _ in range(self.mesh.ndim)]
        pl[self.mesh_dim] = self.layout.placement(slot)
        return tuple(pl)

    def _alloc_shard(self, dtype, symmetric: bool = False) -> torch.Tensor:
        if symmetric and self.comm is not None and getattr(self.comm, "symmetric", False):
            return self.comm.alloc(self.S, dtype)
        return torch.zeros(self.S, dtype=dtype, device=self.device)

    def _alloc_full(self, dtype, symmetric: bool = False) -> torch.Tensor:
        n = self.S * self.world
        if symmetric and self.comm is not None and getattr(self.comm, "symmetric", False):
            return self.comm.alloc(n, dtype)
        return torch.empty(n, dtype=dtype, device=self.device)

    def nbytes_full(self) -> int:
        return self.S * self.world * torch.empty((), dtype=self.param_dtype).element_size()

    # ------------------------------------------------------------------ parameter (un)sharding
    def refresh_param_shard(self) -> bool:
        """bf16 shard <- master (needed when a foreign optimizer updated the fp32 DTensor params)."""
        if not self.bf16_fresh:
            self.param_shard.copy_(self.master)
            self.bf16_fresh = True
            return True
        return False

    def all_gather(self, full: torch.Tensor, handshake: bool = True) -> None:
        """Launch the all-gather of the unit into ``full`` on the current stream."""
        handshake = self.refresh_param_shard() or handshake
        if self.world == 1:
            if full.data_ptr() != self.param_shard.data_ptr():
                full.copy_(self.param_shard)
            return
        if self.comm is not None:
            self.comm.all_gather(self.param_shard, full, self, handshake=handshake)
        else:
            dist.all_gather_into_tensor(full, self.param_shard, group=self.group)

    def use_full(self, full: torch.Tensor) -> None:
        """Point the module's compute parameters at views of the gathered buffer."""
        self.full_param = full
        for slot, p in zip(self.layout.slots, self.params):
            p.data = full[slot.offset : slot.end].view(slot.shape)
        self.unsharded = True

    def release_full(self) -> Optional[torch.Tensor]:
        full, self.full_param = self.full_param, None
        for p in self.params:
            p.data = torch.empty(0, dtype=self.param_dtype, device=self.device)
        self.unsharded = False
        return full

    # ------------------------------------------------------------------ gradients
    def attach_grad_buffer(self, full_grad: torch.Tensor, accumulate: bool) -> None:
        self.full_grad = full_grad
        if not accumulate:
            # padding between / after parameters is never written by a wgrad kernel: keep it zero
            pos = 0
            for slot in self.layout.slots:
                if slot.offset > pos:
                    full_grad[pos : slot.offset].zero_()
                pos = slot.end
            if pos < full_grad.numel():
                full_grad[pos:].zero_()
        for slot, p in zip(self.layout.slots, self.params):
            p.main_grad = full_grad[slot.offset : slot.end].view(slot.shape)
            p._main_grad_initialised = accumulate
        self.pending_grad_params = len(self.params)

    def collect_autograd_grads(self) -> None:
        """Parameters used by generic (non-``vescale_b200.ops``) modules get ordinary ``.grad``s: fold them
        into the flat buffer.  Parameters that never received a gradient are zero-filled."""
        for p in self.params:
            mg = getattr(p, "main_grad", None)
            if mg is None:
                continue
            if p.grad is not None:
                if getattr(p, "_main_grad_initialised", False):
                    mg.add_(p.grad.to(mg.dtype))
                else:
                    mg.copy_(p.grad)
                    p._main_grad_initialised = True
                p.grad = None
            if not getattr(p, "_main_grad_initialised", False):
                mg.zero_()
                p._main_grad_initialised = True

    def detach_grad_buffer(self) -> Optional[torch.Tensor]:
        fg, self.full_grad = self.full_grad, None
        for p in self.params:
            p.main_grad = None
            p._main_grad_initialised = False
        return fg

    def reduce_scatter(self, scale: float) -> None:
        """full_grad -> grad_shard (= scale * sum over ranks of my slice), plus ``sumsq`` of the result.
        Launches on the current stream."""
        fg = self.full_grad
        W, S = self.world, self.S
        if W == 1:
            # the bf16 buffer *is* the reduced gradient; no fp32 copy is kept (180 GB budget at 8B params)
            self.grad_shard = fg
            self.grad_scale_pending = scale
        elif self.comm is not None:
            if self.grad_shard is None or self.grad_shard.dtype != self.reduce_dtype or self.grad_shard.data_ptr() == 0:
                self.grad_shard = torch.empty(S, dtype=self.reduce_dtype, device=self.device)
            self.comm.reduce_scatter(fg, self.grad_shard, scale, self)
            self.grad_scale_pending = 1.0
        else:
            if self.grad_shard is None or self.grad_shard.dtype != self.reduce_dtype or self.grad_shard.numel() != S:
                self.grad_shard = torch.empty(S, dtype=self.reduce_dtype, device=self.device)
            backend = dist.get_backend(self.group)
            if backend == "nccl":
                # the cross-rank sum runs in ``reduce_dtype`` (MixedPrecisionPolicy default fp32, like FSDP2 and the legacy
                # fp32 grad buffer ``ddp/grad_buffer.py:132-147``), as the gloo path and the symmetric-memory kernel do
                red = fg if fg.dtype == self.reduce_dtype else fg.to(self.reduce_dtype)
                dist.reduce_scatter_tensor(self.grad_shard, red, op=dist.ReduceOp.SUM, group=self.group)
                self.grad_shard.mul_(scale)
            else:
                red = fg.to(self.reduce_dtype)
                dist.all_reduce(red, group=self.group)
                self.grad_shard.copy_(red[self.rank * S : (self.rank + 1) * S]).mul_(scale)
            self.grad_scale_pending = 1.0
        self.grad_ready = True

    def expose_sharded_grads(self) -> None:
        """Make ``sharded_param.grad`` DTensors over ``grad_shard`` (for foreign optimizers / clip_grad_norm_)."""
        gs = self.grad_shard
        if gs is None:
            return
        if gs.dtype != self.master_dtype or getattr(self, "grad_scale_pending", 1.0) != 1.0 or gs.numel() != self.S:
            src = gs[self.rank * self.S : (self.rank + 1) * self.S] if gs.numel() != self.S else gs
            gs = src.to(self.master_dtype) * getattr(self, "grad_scale_pending", 1.0)
            self.grad_shard_exposed = gs
        for slot, sp in zip(self.layout.slots, self.sharded_params):
            lo, hi = self.layout.rank_range(slot, self.rank)
            sp.grad = DTensor(gs[lo:hi], sp.data._spec)

    def zero_grad(self) -> None:
        self.grad_ready = False
        for sp in self.sharded_params:
            sp.grad = None

    def __repr__(self) -> str:
        return f"FSDPUnit({self.name}, {self.layout})"
