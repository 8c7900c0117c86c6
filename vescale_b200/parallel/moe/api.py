"""``parallelize_experts``: turn the expert modules of an existing model into an expert-parallel MoE.

    parallelize_experts(model, experts_expr=r"layers\\.\\d+\\.moe", ep_mesh=mesh["EP"], config={...})

The reference intercepts each expert's ``forward`` and ``Tensor.index_add_`` to batch all experts' work into one
dispatch (``legacy/vescale/moe/_moe_tensor.py:42-99``); here matching ``MoELayer``s (or HF-style blocks exposing
``gate`` + ``experts``) are rebuilt as :class:`MoELayer` over the EP group, their expert weights re-laid to the
owning rank (``ExpertsAllocator``), and expert parameters are excluded from data-parallel gradient reduction
across EP ranks (``param_to_ignore`` in DDP).  ``TokenDispatcher`` picks (expert, replica) per token copy.

Parity: ``legacy/vescale/moe/api.py:29-49``, ``experts_allocator.py:63``, ``token_dispatcher.py:30``, ``moe_optimizer.py``.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from .layer import MoELayer

__all__ = ["parallelize_experts", "ExpertsAllocator", "BasicExpertsAllocator", "TokenDispatcher", "BasicTokenDispatcher", "MoEOptimizer", "is_moe", "reallocate_experts", "balanced_allocation"]


class ExpertsAllocator:
    """Where the experts of each MoE layer live.  Two levels of contract:

    * the reference's (``legacy/vescale/moe/experts_allocator.py:26-62``): ``allocate_experts(layer_id, iter)`` returns one
      ``[DP, TP]`` mesh per expert (a DeviceMesh, a tensor or nested lists of EP ranks: ``DP`` replicas, each sharded ``TP`` ways)
      or ``None`` = keep the current allocation; ``collect_performance(perf, iter)`` receives the loads of every layer run.  Layers
      built from such an allocator are :class:`~.scheduler.ScheduledMoELayer` s driven by a :class:`~.scheduler.MoEScheduler`;
    * the simple one: ``allocate()`` returns, per expert, the list of ranks hosting it (whole experts, plain EP).

    The default ``allocate_experts`` lifts ``allocate()`` (each hosting rank = one unsharded replica) the first time a layer asks."""

    def __init__(self, num_experts: Optional[int] = None, ep_size: Optional[int] = None, model_config=None, env_config=None):
        self.num_experts, self.ep_size = num_experts, ep_size
        self.model_config, self.env_config = model_config, env_config
        self._visited = set()

    def allocate(self) -> List[List[int]]:
        raise NotImplementedError

    def collect_performance(self, perf, iter=-1) -> None:
        pass

    def allocate_experts(self, layer_id, iter=-1):
        if layer_id in self._visited:
            return None
        self._visited.add(layer_id)
        return [[[r] for r in ranks] for ranks in self.allocate()]

    def allocate_experts_internal(self, layer_id, iter=-1) -> Optional[Dict]:
        """``{"experts_alloc", "dp_size", "tp_size"}`` for the dispatcher, or ``None`` when nothing changes."""
        from .scheduler import ExpertsAllocation

        got = self.allocate_experts(layer_id, iter)
        if got is None:
            return None
        alloc = got if isinstance(got, ExpertsAllocation) else ExpertsAllocation(got, self.ep_size or (dist.get_world_size() if dist.is_initialized() else 1))
        return alloc.info()


class BasicExpertsAllocator(ExpertsAllocator):
    """Plain expert parallelism: expert ``e`` whole on rank ``e // (E / W)``."""

    def allocate(self) -> List[List[int]]:
        per = max(1, self.num_experts // self.ep_size)
        return [[min(e // per, self.ep_size - 1)] for e in range(self.num_experts)]


class TokenDispatcher:
    """Chooses the replica every routed token copy goes to (reference ``token_dispatcher.py:26-70``): the scheduler calls
    ``set_experts_alloc(info)`` when the allocation changes, ``assign_task(layer_id, token_id, expert_id, hidden_state, token_weight)``
    with the routed copies of the layer, then ``dispatch_token(layer_id) -> (expert_id, replica_id)``; ``collect_performance`` as
    for the allocator."""

    def set_experts_alloc(self, experts_alloc_info: Dict) -> None:
        self.experts_alloc = experts_alloc_info["experts_alloc"]
        self.num_replicate = experts_alloc_info["dp_size"]

    def assign_task(self, layer_id, token_id, expert_id, hidden_state, token_weight) -> None:
        self.expert_id, self.token_id = expert_id, token_id

    def collect_performance(self, perf, iter=-1) -> None:
        pass

    def dispatch_token(self, layer_id: int):
        raise NotImplementedError


class BasicTokenDispatcher(TokenDispatcher):
    """Uniformly random replica among the ones hosting the expert (legacy ``token_dispatcher.py:46-70``)."""

    def dispatch_token(self, layer_id: int):
        n = self.num_replicate.to(self.expert_id.device)[self.expert_id]
        return self.expert_id, torch.randint_like(n, 65535) % n


def is_moe(module: nn.Module) -> bool:
    return isinstance(module, MoELayer)


_TAG_EXPERTS_PARALLELIZED = "_vb_experts_parallelized"


def is_experts_parallized(module: nn.Module) -> bool:
    """True once ``parallelize_experts`` has run on ``module`` (legacy ``moe/_experts.py:39``; spelling kept)."""
    return bool(getattr(module, _TAG_EXPERTS_PARALLELIZED, False))


def parallelize_experts(module: nn.Module, experts_expr: str = r".*moe.*", ep_mesh=None, experts_allocator: Optional[ExpertsAllocator] = None, token_dispatcher: Optional[TokenDispatcher] = None, config: Optional[Dict] = None) -> nn.Module:
    rx = re.compile(experts_expr)
    group = ep_mesh.get_group(0) if ep_mesh is not None and ep_mesh.has_groups() else None
    W = dist.get_world_size(group) if group is not None else 1
    rank = dist.get_rank(group) if group is not None else 0
    from .hijack import hijack_moe_block, is_hijackable

    scheduled: List[nn.Module] = []
    for fqn, sub in list(module.named_modules()):
        if not rx.fullmatch(fqn):
            continue
        if not isinstance(sub, MoELayer) and is_hijackable(sub) and not hasattr(sub, "_vb_moe_layer"):
            # an unmodified third-party block (HF Mixtral style): its expert container is swapped for the EP path (hijack.py)
            ex = sub.experts
            E = ex.gate_up_proj.shape[0] if hasattr(ex, "gate_up_proj") else len(ex)
            alloc = (experts_allocator or BasicExpertsAllocator(E, W)).allocate()
            hijack_moe_block(sub, group, placement=alloc, top_k=(config or {}).get("top_k"), comm_backend=(config or {}).get("comm_backend", "nccl"))
            continue
        if not isinstance(sub, MoELayer) or sub.ep_size == W:
            continue
        cfg = sub.cfg
        meshes = _reference_shaped_allocation(experts_allocator, len(scheduled), W)
        if meshes is not None:
            # per-expert DP x TP meshes (replicated / sharded experts): the layer becomes a ScheduledMoELayer
            from .scheduler import ExpertsAllocation, ScheduledMoELayer

            alloc_obj = meshes if isinstance(meshes, ExpertsAllocation) else ExpertsAllocation(meshes, W)
            new = ScheduledMoELayer(cfg, alloc_obj, group, device=sub.router.weight.device, layer_id=len(scheduled))
            new.router.weight.data.copy_(sub.router.weight.data)
            new.load_full_experts(sub.experts.w_gate_up.data, sub.experts.w_down.data)
            _replace_module(module, fqn, new)
            scheduled.append(new)
            continue
        alloc = (experts_allocator or BasicExpertsAllocator(cfg.num_experts, W)).allocate()
        new = MoELayer(cfg, group, device=sub.router.weight.device)
        new.router.weight.data.copy_(sub.router.weight.data)
        mine = [e for e in range(cfg.num_experts) if rank in alloc[e]]
        assert len(mine) == new.num_local, "allocator must give every rank E/W experts"
        with torch.no_grad():
            for le, e in enumerate(mine):
                new.experts.w_gate_up[le].copy_(sub.experts.w_gate_up[e])
                new.experts.w_down[le].copy_(sub.experts.w_down[e])
        _replace_module(module, fqn, new)
        for p in new.experts.parameters():
            p._is_expert_param = True
    module._ep_group = group
    if scheduled:
        from .scheduler import MoEScheduler

        module._moe_scheduler = MoEScheduler(experts_allocator, token_dispatcher, config, optimizer=(config or {}).get("optimizer")).register(module)
    setattr(module, _TAG_EXPERTS_PARALLELIZED, True)
    return module


def _replace_module(root: nn.Module, fqn: str, new: nn.Module) -> None:
    parent = root
    parts = fqn.split(".")
    for p in parts[:-1]:
        parent = getattr(parent, p)
    setattr(parent, parts[-1], new)


def _reference_shaped_allocation(allocator, layer_id: int, world: int):
    """The per-expert meshes of a reference-shaped allocator (one that overrides ``allocate_experts``), or ``None`` for the simple
    ``allocate()`` kind / no allocator."""
    if allocator is None:
        return None
    fn = getattr(type(allocator), "allocate_experts", None)
    if fn is None or fn is ExpertsAllocator.allocate_experts:
        return None
    return allocator.allocate_experts(layer_id, 0)


class MoEOptimizer:
    """Optimizer wrapper for EP models: expert grads are averaged over the *data-parallel replicas of that
    expert* only (not over EP ranks); the grad-norm sums expert shards across EP (legacy ``moe_optimizer.py:40-107``)."""

    def __init__(self, optimizer: torch.optim.Optimizer, model: nn.Module, dp_group=None, ep_group=None, clip_grad: float = 0.0):
        self.optimizer, self.model, self.dp_group, self.ep_group, self.clip_grad = optimizer, model, dp_group, ep_group, clip_grad

    def step(self):
        dense = [p for p in self.model.parameters() if p.grad is not None and not getattr(p, "_is_expert_param", False)]
        expert = [p for p in self.model.parameters() if p.grad is not None and getattr(p, "_is_expert_param", False)]
        world_group = dist.group.WORLD if dist.is_initialized() else None
        if world_group is not None and dist.get_world_size() > 1:
            for p in dense:
                dist.all_reduce(p.grad, group=world_group)
                p.grad.div_(dist.get_world_size())
            if self.dp_group is not None and dist.get_world_size(self.dp_group) > 1:
                for p in expert:
                    dist.all_reduce(p.grad, group=self.dp_group)
            ep = dist.get_world_size(self.ep_group) if self.ep_group is not None else 1
            for p in expert:  # the loss is averaged over all ranks' tokens; each expert saw tokens from every EP rank
                p.grad.div_(dist.get_world_size())
        if self.clip_grad > 0:
            from ...optim.clip_grads import get_grad_norm_fp32

            nd = get_grad_norm_fp32([p.grad for p in dense], 2.0)
            ne = get_grad_norm_fp32([p.grad for p in expert], 2.0, [self.ep_group] if self.ep_group is not None else [])
            total = (nd**2 + ne**2).sqrt()
            coef = torch.clamp(self.clip_grad / (total + 1e-6), max=1.0)
            torch._foreach_mul_([p.grad for p in dense + expert], coef)
        self.optimizer.step()

    def zero_grad(self, set_to_none=True):
        self.optimizer.zero_grad(set_to_none)


def balanced_allocation(load_per_expert, ep_size: int) -> List[int]:
    """Greedy load balancing: experts in decreasing load order go to the least-loaded rank that still has a free slot.
    Returns ``slot_of_expert`` (global slot = rank * E/W + local index)."""
    loads = [float(x) for x in load_per_expert]
    E = len(loads)
    per = E // ep_size
    rank_load, rank_free = [0.0] * ep_size, [per] * ep_size
    slot = [0] * E
    for e in sorted(range(E), key=lambda i: -loads[i]):
        r = min((r for r in range(ep_size) if rank_free[r] > 0), key=lambda r: rank_load[r])
        slot[e] = r * per + (per - rank_free[r])
        rank_free[r] -= 1
        rank_load[r] += loads[e]
    return slot


def _expert_move_plan(layer: MoELayer, new_slot_of_expert):
    """(new slot list, ``move(t)``): ``move`` permutes a ``[E_local, ...]`` tensor of per-slot data across the EP group in place
    according to the re-allocation (one all-to-all)."""
    from ...comm.collectives import _p2p_all_to_all

    E, W, El, me, group = layer.cfg.num_experts, layer.ep_size, layer.num_local, layer.ep_rank, layer.ep_group
    new = [int(x) for x in (new_slot_of_expert.tolist() if isinstance(new_slot_of_expert, torch.Tensor) else new_slot_of_expert)]
    if sorted(new) != list(range(E)):
        raise ValueError("new_slot_of_expert must be a permutation of range(num_experts)")
    old = [int(x) for x in layer.slot_of_expert.tolist()]
    expert_at_old = {s: e for e, s in enumerate(old)}
    expert_at_new = {s: e for e, s in enumerate(new)}
    # my outgoing experts per destination rank (ordered by their new local index), incoming per source rank likewise
    send = [[] for _ in range(W)]
    for i in range(El):
        e = expert_at_old[me * El + i]
        send[new[e] // El].append((new[e] % El, i))
    recv = [[] for _ in range(W)]
    for j in range(El):
        e = expert_at_new[me * El + j]
        recv[old[e] // El].append(j)
    for lst in send:
        lst.sort()

    def move(t: torch.Tensor) -> None:  # t: [El, ...] stacked per local slot
        ins = [torch.stack([t[i] for _, i in send[d]]) if send[d] else t.new_empty((0, *t.shape[1:])) for d in range(W)]
        outs = [t.new_empty((len(recv[s]), *t.shape[1:])) for s in range(W)]
        if W == 1:
            outs[0].copy_(ins[0])
        elif dist.get_backend(group) == "nccl":
            dist.all_to_all(outs, ins, group=group)
        else:
            _p2p_all_to_all(outs, ins, group)
        for s in range(W):
            for k, j in enumerate(sorted(recv[s])):
                t[j].copy_(outs[s][k])

    return new, move


@torch.no_grad()
def reallocate_experts(layer: MoELayer, new_slot_of_expert, optimizer: Optional[torch.optim.Optimizer] = None) -> None:
    """Move experts between EP ranks while training: weights *and* the optimizer state of every moved expert migrate to the
    new owner, and the layer's routing table is updated (legacy ``MoELayerParamBuffer.refresh_buffer``,
    ``_moe_param_buffer.py:183-337``).  ``new_slot_of_expert[e]`` is the global slot (rank * E/W + local index) expert ``e``
    moves to; it must be a permutation of ``range(E)``.  One all-to-all per tensor; collective over the EP group."""
    new, move = _expert_move_plan(layer, new_slot_of_expert)
    for p in (layer.experts.w_gate_up, layer.experts.w_down):
        move(p.data)
        if optimizer is not None:
            st = optimizer.state.get(p, {})
            for v in st.values():
                if isinstance(v, torch.Tensor) and v.shape == p.shape:
                    move(v)
        if getattr(p, "main_grad", None) is not None and p.main_grad.shape == p.shape:
            move(p.main_grad)
    layer.slot_of_expert.copy_(torch.tensor(new, device=layer.slot_of_expert.device))
