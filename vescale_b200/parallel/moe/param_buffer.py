"""``MoELayerParamBuffer``: per-layer expert parameter buffers sharded over the data-parallel replicas of each EP rank.

The reference packs every MoE layer's expert weights into per-DP-mesh flat buffers, all-gathers layer i+1 while layer i
computes, reduce-scatters expert gradients from a tensor hook in backward, and can re-allocate experts between ranks at run
time while migrating optimizer state (``legacy/vescale/moe/_moe_param_buffer.py:50-403``, ``_scheduler.py:125-134,217-222``).

Here that machinery IS the RaggedShard FSDP engine (``parallel/fsdp``): each layer's :class:`GroupedExperts` becomes one FSDP
unit over the DP mesh dim, all units share one :class:`FSDPState`, so

* the unit's flat buffer is the layer's expert parameter buffer (zero-copy views, rank boundaries on block granularity);
* ``prefetch`` layers ahead are all-gathered on the communication stream while the current layer computes
  (copy engines / symmetric-memory kernels on CUDA, c10d on CPU) — ``_moe_param_buffer.py:384-393,436-446``;
* gradients are reduce-scattered (⊕ 1/dp scale ⊕ fp32 cast ⊕ sum of squares) from the unit's post-backward hook —
  ``:395-403``;
* :meth:`refresh_buffer` moves experts between EP ranks together with their fp32 master weights and AdamW moments —
  ``refresh_buffer:183-337``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from ..fsdp import MixedPrecisionPolicy, fully_shard
from ..fsdp.api import FSDPState
from ..fsdp.unit import FSDPUnit
from .api import _expert_move_plan
from .layer import MoELayer

__all__ = ["MoELayerParamBuffer"]


class MoELayerParamBuffer:
    def __init__(self, layers: Sequence[MoELayer], dp_mesh, *, mesh_dim=0, prefetch: int = 1, mp_policy: Optional[MixedPrecisionPolicy] = None,
                 comm_backend: str = "auto", reshard_after_forward: bool = True):
        self.layers: List[MoELayer] = list(layers)
        self.dp_mesh = dp_mesh
        state: Optional[FSDPState] = None
        for l in self.layers:
            fully_shard(l.experts, dp_mesh, mesh_dim=mesh_dim, mp_policy=mp_policy, reshard_after_forward=reshard_after_forward, prefetch=prefetch,
                        comm_backend=comm_backend, state=state)
            state = l.experts._fsdp_state
        self.state = state
        self.units: List[FSDPUnit] = [l.experts._fsdp_unit for l in self.layers]

    # ------------------------------------------------------------------ explicit control (the hooks do this on their own)
    def all_gather(self, index: int) -> None:
        """Kick the all-gather of layer ``index``'s expert buffer on the communication stream."""
        self.state.lazy_init()
        self.state.launch_all_gather(self.units[index])

    def wait(self, index: int) -> None:
        self.state.wait_all_gather(self.units[index])

    def buffer_bytes(self) -> List[int]:
        return [u.nbytes_full() for u in self.units]

    # ------------------------------------------------------------------ dynamic re-allocation
    @torch.no_grad()
    def refresh_buffer(self, index: int, new_slot_of_expert, optimizer=None) -> None:
        """Re-allocate the experts of layer ``index`` (``new_slot_of_expert[e]`` = global slot rank * E/W + local index) and
        migrate everything that belongs to an expert with it: the fp32 master weights, the bf16 all-gather source, and — when
        ``optimizer`` is the ``FSDPAdamW`` driving these units — its two moment buffers.  Collective over the DP group (gather /
        re-shard of the unit) and the EP group (one all-to-all per tensor)."""
        layer, u = self.layers[index], self.units[index]
        new, move = _expert_move_plan(layer, new_slot_of_expert)
        flats = [u.master]
        if optimizer is not None and hasattr(u, "exp_avg"):
            flats += [u.exp_avg, u.exp_avg_sq]
        S, W, r = u.S, u.world, u.rank
        for flat in flats:
            # gather the DP shards of the unit, permute expert slots across EP ranks, keep my shard again
            full = torch.empty(S * W, dtype=flat.dtype, device=flat.device)
            if W > 1:
                dist.all_gather_into_tensor(full, flat.contiguous(), group=u.group)
            else:
                full.copy_(flat)
            for slot in u.layout.slots:
                move(full[slot.offset : slot.end].view(slot.shape))
            flat.copy_(full[r * S : (r + 1) * S])
        u.bf16_fresh = False  # the all-gather source is rebuilt from the master shard before the next gather
        layer.slot_of_expert.copy_(torch.tensor(new, device=layer.slot_of_expert.device))
        self.state.invalidate_params()
