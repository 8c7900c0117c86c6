"""Expert parallelism for UNMODIFIED third-party MoE blocks (HuggingFace Mixtral-style).

The reference makes a stock ``MixtralSparseMoeBlock`` expert-parallel without touching its source: every expert's ``forward`` is
replaced by a recorder that returns a placeholder, ``Tensor.index_add_`` is intercepted, and once all experts have queued their
work one batched dispatch runs (``legacy/vescale/moe/_moe_tensor.py:42-99``, ``_scheduler.py:224-277``).  The same effect — the
block's class, router and call signature stay as they are, the per-expert Python loop becomes one dispatch → grouped GEMM →
combine — is obtained here by swapping the *expert container*, not by patching ``torch.Tensor``:

* **fused container** (transformers >= 5: ``block.experts`` is one module holding ``gate_up_proj [E, 2I, H]`` and ``down_proj
  [E, H, I]``, called as ``experts(hidden, top_k_index, top_k_weights)``): replaced by :class:`EPExperts` with the same signature;
* **per-expert ModuleList** (transformers 4.x / custom blocks: ``block.experts[i]`` has ``w1/w3/w2`` or ``gate_proj/up_proj/
  down_proj`` and the block's ``forward`` loops over experts with ``index_add_``): the block instance gets a bound ``forward``
  that runs the block's own ``gate``, the standard softmax → top-k → renormalise routing, and the batched expert path; it
  returns ``(hidden_states, router_logits)`` like the 4.x block.

Either way only the experts this EP rank hosts keep their weights (``ExpertsAllocator``), tokens travel by the dispatch back
ends of :class:`MoELayer` (NCCL all-to-all, or the device-side symmetric-memory dispatcher), and expert parameters are tagged
``_is_expert_param`` so data-parallel wrappers leave them out of the EP-wide gradient average.
"""
from __future__ import annotations

import types
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from .layer import MoEConfig, MoELayer

__all__ = ["EPExperts", "hijack_moe_block", "is_hijackable"]


def _expert_linears(mod: nn.Module) -> Optional[Tuple[nn.Linear, nn.Linear, nn.Linear]]:
    """(gate, up, down) projections of one SwiGLU expert MLP under the two common naming schemes."""
    for names in (("w1", "w3", "w2"), ("gate_proj", "up_proj", "down_proj")):
        if all(isinstance(getattr(mod, n, None), nn.Linear) for n in names):
            return tuple(getattr(mod, n) for n in names)
    return None


def _kind(block: nn.Module) -> Optional[str]:
    ex = getattr(block, "experts", None)
    if ex is None or getattr(block, "gate", None) is None:
        return None
    if isinstance(getattr(ex, "gate_up_proj", None), torch.Tensor) and isinstance(getattr(ex, "down_proj", None), torch.Tensor):
        return "fused"
    if isinstance(ex, nn.ModuleList) and len(ex) > 0 and all(_expert_linears(e) is not None for e in ex):
        return "list"
    return None


def is_hijackable(block: nn.Module) -> bool:
    return _kind(block) is not None


class EPExperts(nn.Module):
    """Drop-in for a fused expert container: ``forward(hidden_states [T, H], top_k_index [T, k], top_k_weights [T, k])``."""

    def __init__(self, layer: MoELayer):
        super().__init__()
        self.layer = layer
        self.num_experts = layer.cfg.num_experts

    @property
    def gate_up_proj(self):  # the local experts' weights under the container's original attribute names
        return self.layer.experts.w_gate_up

    @property
    def down_proj(self):
        return self.layer.experts.w_down

    def forward(self, hidden_states: torch.Tensor, top_k_index: torch.Tensor, top_k_weights: torch.Tensor) -> torch.Tensor:
        return self.layer.experts_forward(hidden_states, top_k_weights, top_k_index).to(hidden_states.dtype)


def _build_layer(E: int, H: int, I: int, k: int, dtype, device, ep_group, comm_backend: str) -> MoELayer:
    W = dist.get_world_size(ep_group) if ep_group is not None else 1
    cfg = MoEConfig(H, I, E, k, ep_size=W, dtype=dtype, comm_backend=comm_backend)
    layer = MoELayer(cfg, ep_group, device=device)
    layer.router = nn.Identity()  # routing stays with the hijacked block's own gate
    return layer


def hijack_moe_block(block: nn.Module, ep_group=None, placement: Optional[List[List[int]]] = None, top_k: Optional[int] = None,
                     comm_backend: str = "nccl") -> nn.Module:
    """Make ``block`` expert-parallel over ``ep_group`` in place (see the module docstring).  ``placement[e]`` lists the EP
    ranks hosting expert ``e`` (default: contiguous blocks of E / W experts per rank)."""
    kind = _kind(block)
    if kind is None:
        raise TypeError(f"{type(block).__name__}: expected a block with `gate` and `experts` (fused container or ModuleList of SwiGLU MLPs)")
    W = dist.get_world_size(ep_group) if ep_group is not None else 1
    rank = dist.get_rank(ep_group) if ep_group is not None else 0
    ex = block.experts
    if kind == "fused":
        E, two_i, H = ex.gate_up_proj.shape
        I = two_i // 2
        dtype, device = ex.gate_up_proj.dtype, ex.gate_up_proj.device
    else:
        E = len(ex)
        g0, _, d0 = _expert_linears(ex[0])
        I, H = g0.weight.shape
        dtype, device = g0.weight.dtype, g0.weight.device
    k = int(top_k or getattr(block, "top_k", None) or getattr(getattr(block, "gate", None), "top_k", None) or 2)
    if E % W:
        raise ValueError(f"{E} experts cannot be spread evenly over {W} EP ranks")
    per = E // W
    placement = placement or [[e // per] for e in range(E)]
    mine = [e for e in range(E) if rank in placement[e]]
    if len(mine) != per:
        raise ValueError("the expert placement must give every EP rank E / W experts")
    layer = _build_layer(E, H, I, k, dtype, device, ep_group, comm_backend)
    # the routing table maps an expert id to the global slot (rank * per + local index) that hosts it
    slot = [0] * E
    counters = [0] * W
    for e in range(E):
        r = placement[e][0]
        slot[e] = r * per + counters[r]
        counters[r] += 1
    layer.slot_of_expert.copy_(torch.tensor(slot, device=layer.slot_of_expert.device))
    with torch.no_grad():
        for le, e in enumerate(mine):
            if kind == "fused":
                layer.experts.w_gate_up[le].copy_(ex.gate_up_proj[e])
                layer.experts.w_down[le].copy_(ex.down_proj[e])
            else:
                g, u, d = _expert_linears(ex[e])
                if any(l.bias is not None for l in (g, u, d)):
                    raise NotImplementedError("expert MLPs with biases are not supported by the grouped expert GEMM")
                layer.experts.w_gate_up[le].copy_(torch.cat([g.weight, u.weight], 0))
                layer.experts.w_down[le].copy_(d.weight)
    if kind == "fused":
        block.experts = EPExperts(layer)
    else:
        block.experts = nn.ModuleList()  # weights of the experts hosted elsewhere are released
        block.ep_experts = EPExperts(layer)

        def forward(self, hidden_states: torch.Tensor):
            shape = hidden_states.shape
            x = hidden_states.reshape(-1, shape[-1])
            router_logits = self.gate(x)
            w = F.softmax(router_logits, dim=1, dtype=torch.float)
            w, idx = torch.topk(w, k, dim=-1)
            w = w / w.sum(dim=-1, keepdim=True)
            out = self.ep_experts(x, idx, w.to(x.dtype))
            return out.reshape(shape), router_logits

        block.forward = types.MethodType(forward, block)
    block._vb_moe_layer = layer
    return block
