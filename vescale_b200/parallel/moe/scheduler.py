"""MoE scheduling with per-expert DP x TP allocations: hot experts replicated on several ranks, experts sharded over several
ranks, allocations that change while training.

The reference describes where an expert lives by ONE DEVICE MESH PER EXPERT — shape ``[DP, TP]``: ``DP`` replicas, each sharded
``TP`` ways — produced by a user-supplied ``ExpertsAllocator``; a ``TokenDispatcher`` picks the replica for every routed token
copy, and a scheduler batches the experts' work of one layer into dispatch -> local experts -> combine and re-lays the parameter
buffers when the allocator returns a new allocation (``legacy/vescale/moe/_scheduler.py:78-277``, ``experts_allocator.py:26-62``,
``token_dispatcher.py:26-70``, ``_moe_param_buffer.py:183-337``).

Here the same contract runs on the dispatch / grouped-GEMM / combine machinery of :mod:`.layer`, through one observation: a SwiGLU
expert is a sum over chunks of its intermediate dimension,

    expert(x) = sum_t ( silu(x Wg_t^T) * (x Wu_t^T) ) Wd_t^T ,

so shard ``t`` of replica ``r`` of expert ``e`` is just another (smaller) expert — a VIRTUAL SLOT on rank ``mesh_e[r, t]`` — and a
token routed to ``e`` is sent to the ``TP`` slots of the replica the dispatcher chose; the combine's scatter-add sums the partial
outputs.  No second code path for tensor-parallel experts, no extra collective.

* :class:`ExpertsAllocation` — the per-expert meshes, the slot table derived from them;
* :class:`ScheduledMoELayer` — router + the local slots' stacked weights; ``forward`` = route, pick replicas, dispatch, grouped
  FFN, combine; ``sync_replica_grads`` sums weight gradients over the replicas of each shard (one all-reduce of only the replicated
  shards); ``reallocate`` moves weights and optimizer state to a new allocation (p2p, only what changes host);
* :class:`MoEScheduler` — drives a model's layers from a reference-shaped allocator / dispatcher pair: asks the allocator for a new
  allocation before each layer's forward (``None`` = keep), feeds the dispatcher, reports per-expert / per-rank loads back
  (``collect_performance``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from .layer import GroupedExperts, MoEConfig, TopKRouter, _gloo_a2a_counts, all_to_all_uneven

__all__ = ["ExpertsAllocation", "ScheduledMoELayer", "MoEScheduler", "MoETask"]


def _mesh_tensor(m) -> torch.Tensor:
    """``[DP, TP]`` tensor of EP-group ranks from a DeviceMesh (ours or torch's), a tensor or nested lists."""
    t = getattr(m, "mesh", m)
    t = torch.as_tensor(t, dtype=torch.long).cpu()
    if t.dim() == 1:
        t = t.view(-1, 1)  # a list of ranks = that many replicas, unsharded
    if t.dim() != 2:
        raise ValueError(f"an expert's allocation is a [DP, TP] mesh, got shape {tuple(t.shape)}")
    return t


class ExpertsAllocation:
    """Where every expert lives: ``meshes[e]`` is a ``[DP_e, TP]`` tensor of EP ranks (``TP`` uniform over the layer so that the
    local slots stack into one grouped GEMM; ``DP_e`` free per expert).  Derived: ``slots_of[e, r, t]`` = global slot id
    (``rank * L + local index``, ``-1`` where ``r >= DP_e``), ``hosted[rank]`` = [(e, r, t)] in local-slot order, ``L`` = slots per
    rank (the maximum over ranks; ranks with fewer keep unused slots)."""

    def __init__(self, meshes: Sequence, world: int):
        self.meshes = [_mesh_tensor(m) for m in meshes]
        self.world = int(world)
        tps = {int(m.shape[1]) for m in self.meshes}
        if len(tps) != 1:
            raise ValueError(f"the experts of one layer must use the same TP degree (got {sorted(tps)}): their shards are stacked into one grouped GEMM")
        self.tp = tps.pop()
        self.num_experts = len(self.meshes)
        self.dp_size = torch.tensor([int(m.shape[0]) for m in self.meshes], dtype=torch.long)
        self.hosted: List[List[Tuple[int, int, int]]] = [[] for _ in range(self.world)]
        for e, m in enumerate(self.meshes):
            if int(m.min()) < 0 or int(m.max()) >= self.world:
                raise ValueError(f"expert {e}: rank outside the EP group of size {self.world}")
            for r in range(m.shape[0]):
                for t in range(m.shape[1]):
                    self.hosted[int(m[r, t])].append((e, r, t))
        self.slots_per_rank = max(1, max(len(h) for h in self.hosted))
        L, rmax = self.slots_per_rank, int(self.dp_size.max())
        self.slots_of = torch.full((self.num_experts, rmax, self.tp), -1, dtype=torch.long)
        for rank, h in enumerate(self.hosted):
            for i, (e, r, t) in enumerate(h):
                self.slots_of[e, r, t] = rank * L + i

    @classmethod
    def even(cls, num_experts: int, world: int) -> "ExpertsAllocation":
        """Plain expert parallelism: expert ``e`` whole on rank ``e // (E / W)``."""
        per = max(1, num_experts // world)
        return cls([[[min(e // per, world - 1)]] for e in range(num_experts)], world)

    def hosts_of_shard(self, e: int, t: int) -> List[int]:
        return [int(x) for x in self.meshes[e][:, t]]

    def replicated_shards(self) -> List[Tuple[int, int]]:
        return [(e, t) for e in range(self.num_experts) if self.meshes[e].shape[0] > 1 for t in range(self.tp)]

    def info(self, device_type: str = "cpu") -> Dict:
        """The dictionary the reference hands to ``TokenDispatcher.set_experts_alloc`` (``experts_allocator.py:41-60``)."""
        return {"experts_alloc": self.meshes, "dp_size": self.dp_size.to(device_type), "tp_size": torch.full((self.num_experts,), self.tp, dtype=torch.long, device=device_type)}

    def __eq__(self, other) -> bool:
        return isinstance(other, ExpertsAllocation) and self.world == other.world and len(self.meshes) == len(other.meshes) and all(torch.equal(a, b) for a, b in zip(self.meshes, other.meshes))


class MoETask:
    """What the dispatcher is told about one layer's routed token copies (reference ``MoETask``, ``_scheduler.py:31-40``)."""

    __slots__ = ("layer_id", "token_id", "expert_id", "hidden_state", "token_weight")

    def __init__(self, layer_id, token_id, expert_id, hidden_state, token_weight):
        self.layer_id, self.token_id, self.expert_id, self.hidden_state, self.token_weight = layer_id, token_id, expert_id, hidden_state, token_weight


class ScheduledMoELayer(nn.Module):
    """An MoE layer whose experts live where an :class:`ExpertsAllocation` says (see the module docstring)."""

    def __init__(self, cfg: MoEConfig, allocation: ExpertsAllocation, ep_group=None, device=None, layer_id: int = 0):
        super().__init__()
        self.cfg, self.ep_group, self.layer_id = cfg, ep_group, layer_id
        self.ep_size = dist.get_world_size(ep_group) if dist.is_initialized() else 1
        self.ep_rank = dist.get_rank(ep_group) if dist.is_initialized() else 0
        if allocation.world != self.ep_size:
            raise ValueError(f"allocation is for {allocation.world} ranks, the EP group has {self.ep_size}")
        if cfg.ffn_size % allocation.tp:
            raise ValueError(f"ffn_size {cfg.ffn_size} is not divisible by the experts' TP degree {allocation.tp}")
        self.alloc = allocation
        self.router = TopKRouter(cfg, device)
        self.experts = self._new_bank(allocation, device)
        self.dispatcher = None  # reference-shaped TokenDispatcher (assign_task / dispatch_token); None = uniform random replica
        self.last_tokens_per_expert: Optional[torch.Tensor] = None
        self.last_tokens_per_rank: Optional[List[int]] = None
        self._gen = None

    # ------------------------------------------------------------------ construction / weights
    def _shard_cfg(self, allocation: ExpertsAllocation) -> MoEConfig:
        c = self.cfg
        return MoEConfig(c.hidden_size, c.ffn_size // allocation.tp, c.num_experts, c.top_k, c.ep_size, c.dtype, c.init_std, c.comm_backend, c.aux_loss_coef)

    def _new_bank(self, allocation: ExpertsAllocation, device) -> GroupedExperts:
        bank = GroupedExperts(self._shard_cfg(allocation), allocation.slots_per_rank, device)
        with torch.no_grad():
            bank.w_gate_up.zero_()
            bank.w_down.zero_()
        return bank

    @staticmethod
    def shard_of(w_gate_up_e: torch.Tensor, w_down_e: torch.Tensor, t: int, tp: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Shard ``t`` of ``tp`` of one expert: the ``t``-th chunk of the gate rows, of the up rows (``w_gate_up`` is gate|up
        stacked) and of the down columns."""
        two_i = w_gate_up_e.shape[0]
        i = two_i // 2
        c = i // tp
        gu = torch.cat([w_gate_up_e[t * c : (t + 1) * c], w_gate_up_e[i + t * c : i + (t + 1) * c]], 0)
        return gu, w_down_e[:, t * c : (t + 1) * c]

    @torch.no_grad()
    def load_full_experts(self, w_gate_up: torch.Tensor, w_down: torch.Tensor) -> None:
        """Fill the local slots from the full expert weights ``[E, 2I, H]`` / ``[E, H, I]`` (every rank passes the same tensors)."""
        for i, (e, _r, t) in enumerate(self.alloc.hosted[self.ep_rank]):
            gu, dn = self.shard_of(w_gate_up[e], w_down[e], t, self.alloc.tp)
            self.experts.w_gate_up[i].copy_(gu)
            self.experts.w_down[i].copy_(dn)

    # ------------------------------------------------------------------ forward
    def _pick_replicas(self, eid: torch.Tensor, token_id: torch.Tensor, x2: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        dp = self.alloc.dp_size.to(eid.device)[eid]
        if self.dispatcher is not None:
            self.dispatcher.assign_task(self.layer_id, token_id=token_id, expert_id=eid, hidden_state=x2, token_weight=weight)
            got_e, rid = self.dispatcher.dispatch_token(self.layer_id)
            if got_e is not eid and not torch.equal(got_e, eid):
                raise ValueError("TokenDispatcher.dispatch_token must return the expert ids it was given (it only chooses the replica)")
            return rid.to(eid.device).long() % dp
        if int(self.alloc.dp_size.max()) == 1:
            return torch.zeros_like(eid)
        return torch.randint(0, 65535, eid.shape, device=eid.device, generator=self._gen) % dp  # legacy BasicTokenDispatcher

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        topv, topi, probs = self.router(x2)
        return self.experts_forward(x2, topv, topi, probs).view(shape)

    def experts_forward(self, x2: torch.Tensor, topv: torch.Tensor, topi: torch.Tensor, probs: Optional[torch.Tensor] = None) -> torch.Tensor:
        T, k, E, W, L, TP = x2.shape[0], topi.shape[-1], self.cfg.num_experts, self.ep_size, self.alloc.slots_per_rank, self.alloc.tp
        dev = x2.device
        if self.cfg.aux_loss_coef > 0 and probs is not None:
            frac = torch.zeros(E, device=dev).index_add_(0, topi.reshape(-1), torch.ones(T * k, device=dev)) / (T * k)
            self.last_aux_loss = self.cfg.aux_loss_coef * E * (frac * probs.mean(0)).sum()
        eid = topi.reshape(-1)
        token_id = torch.arange(T, device=dev).repeat_interleave(k)
        wflat = topv.reshape(-1)
        rid = self._pick_replicas(eid, token_id, x2, wflat)
        slots = self.alloc.slots_of.to(dev)[eid, rid]  # [T*k, TP]: every shard of the chosen replica gets the token
        flat = slots.reshape(-1)
        token_of_copy = token_id.repeat_interleave(TP)
        w_of_copy = wflat.repeat_interleave(TP)
        order = torch.argsort(flat, stable=True)
        token_of = token_of_copy[order]
        xs = x2[token_of]
        counts = torch.bincount(flat, minlength=W * L)
        self.last_tokens_per_expert = torch.bincount(eid, minlength=E)
        if W > 1:
            send_counts = counts.view(W, L)
            recv_counts = torch.empty_like(send_counts)
            if dist.get_backend(self.ep_group) == "nccl":
                dist.all_to_all_single(recv_counts, send_counts.contiguous(), group=self.ep_group)
            else:
                _gloo_a2a_counts(recv_counts, send_counts, self.ep_group)
            in_splits = send_counts.sum(1).tolist()
            rc = recv_counts.cpu()
            out_splits = rc.sum(1).tolist()
            recv = all_to_all_uneven(xs, in_splits, out_splits, self.ep_group)
            sid = torch.repeat_interleave(torch.arange(L, device=dev).repeat(W), recv_counts.reshape(-1))
            perm2 = torch.argsort(sid, stable=True)
            y = self.experts(recv[perm2], rc.sum(0).tolist())
            inv2 = torch.empty_like(perm2)
            inv2[perm2] = torch.arange(perm2.numel(), device=dev)
            back = all_to_all_uneven(y[inv2], out_splits, in_splits, self.ep_group)
            self.last_tokens_per_rank = out_splits
        else:
            back = self.experts(xs, counts.tolist())
            self.last_tokens_per_rank = [int(counts.sum())]
        out = torch.zeros_like(x2).index_add_(0, token_of, back * w_of_copy[order].to(back.dtype).unsqueeze(-1))
        return out

    # ------------------------------------------------------------------ replicas: gradient sync
    def _grad_of(self, p: torch.Tensor) -> Optional[torch.Tensor]:
        mg = getattr(p, "main_grad", None)
        return mg if mg is not None else p.grad

    @torch.no_grad()
    def sync_replica_grads(self) -> None:
        """Sum the weight gradients of every replicated shard over its replicas (each replica saw different tokens of the same
        batch).  One all-reduce per weight over the EP group of ONLY the replicated shards — ranks that do not host a shard
        contribute zeros — so nothing moves when no expert is replicated and no per-expert process groups are needed."""
        rep = self.alloc.replicated_shards()
        if not rep or self.ep_size == 1:
            return
        index = {et: i for i, et in enumerate(rep)}
        mine = [(i, index[(e, t)]) for i, (e, _r, t) in enumerate(self.alloc.hosted[self.ep_rank]) if (e, t) in index]
        for p in (self.experts.w_gate_up, self.experts.w_down):
            g = self._grad_of(p)
            buf = torch.zeros((len(rep), *p.shape[1:]), dtype=torch.float32, device=p.device)
            if g is not None:
                for li, bi in mine:
                    buf[bi] += g[li].float()
            dist.all_reduce(buf, group=self.ep_group)
            if g is None and mine:
                p.grad = g = torch.zeros_like(p)
            for li, bi in mine:
                g[li].copy_(buf[bi].to(g.dtype))

    # ------------------------------------------------------------------ re-allocation
    @torch.no_grad()
    def reallocate(self, new_alloc: ExpertsAllocation, optimizer: Optional[torch.optim.Optimizer] = None) -> None:
        """Adopt ``new_alloc``: every shard that gets a new host (or an extra replica) is copied from a current host — weights and,
        when ``optimizer`` is given, every parameter-shaped optimizer state (Adam moments) — then the local banks are rebuilt and
        the optimizer re-pointed at them.  Only shards that change host travel (batched p2p).  Collective over the EP group.
        Reference: ``MoELayerParamBuffer.refresh_buffer`` (``_moe_param_buffer.py:183-337``)."""
        old = self.alloc
        if new_alloc == old:
            return
        if new_alloc.tp != old.tp:
            raise NotImplementedError("re-allocation keeps the experts' TP degree (change it by rebuilding the layer from full weights)")
        if new_alloc.world != old.world or new_alloc.num_experts != old.num_experts:
            raise ValueError("re-allocation must keep the EP group and the expert count")
        me, dev = self.ep_rank, self.experts.w_gate_up.device
        old_params = (self.experts.w_gate_up, self.experts.w_down)
        # per parameter: the tensors that travel together (weight + parameter-shaped optimizer states)
        packs: List[List[Tuple[str, torch.Tensor]]] = []
        for p in old_params:
            pack = [("param", p.data)]
            if optimizer is not None:
                for name, v in optimizer.state.get(p, {}).items():
                    if isinstance(v, torch.Tensor) and v.shape == p.shape:
                        pack.append((name, v))
            packs.append(pack)
        new_bank = self._new_bank(new_alloc, dev)
        new_params = (new_bank.w_gate_up, new_bank.w_down)
        new_packs = [[(name, torch.zeros_like(np_) if name != "param" else np_.data) for name, _ in pack] for pack, np_ in zip(packs, new_params)]
        old_index = {}  # (e, t) -> (rank, local index) of the first current host
        for rank, h in enumerate(old.hosted):
            for i, (e, _r, t) in enumerate(h):
                old_index.setdefault((e, t), (rank, i))
        ops, keep = [], []
        for dst, h in enumerate(new_alloc.hosted):
            for j, (e, _r, t) in enumerate(h):
                # prefer a copy the destination already holds
                local = next((i for i, (e2, _r2, t2) in enumerate(old.hosted[dst]) if (e2, t2) == (e, t)), None)
                src, i = (dst, local) if local is not None else old_index[(e, t)]
                if src == dst:
                    if dst == me:
                        for pack, npack in zip(packs, new_packs):
                            for (_n, a), (_n2, b) in zip(pack, npack):
                                b[j].copy_(a[i])
                    continue
                for pack, npack in zip(packs, new_packs):
                    for (_n, a), (_n2, b) in zip(pack, npack):
                        if me == src:
                            buf = a[i].contiguous()
                            keep.append(buf)
                            ops.append(dist.P2POp(dist.isend, buf, self._global_rank(dst), self.ep_group))
                        elif me == dst:
                            ops.append(dist.P2POp(dist.irecv, b[j], self._global_rank(src), self.ep_group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        # swap the banks in; re-point the optimizer (state keyed by the new Parameter objects)
        for p_old, p_new, npack in zip(old_params, new_params, new_packs):
            for attr in ("_is_expert_param",):
                setattr(p_new, attr, True)
            if optimizer is not None:
                st_old = optimizer.state.pop(p_old, None)
                if st_old is not None:
                    st_new = {k: v for k, v in st_old.items() if not (isinstance(v, torch.Tensor) and v.shape == p_old.shape)}
                    for name, t in npack:
                        if name != "param":
                            st_new[name] = t
                    optimizer.state[p_new] = st_new
                for grp in optimizer.param_groups:
                    grp["params"] = [p_new if q is p_old else q for q in grp["params"]]
        self.experts = new_bank
        self.alloc = new_alloc
        if self.dispatcher is not None and hasattr(self.dispatcher, "set_experts_alloc"):
            self.dispatcher.set_experts_alloc(new_alloc.info(dev.type))

    def _global_rank(self, group_rank: int) -> int:
        return dist.get_global_rank(self.ep_group, group_rank) if self.ep_group is not None else group_rank


class MoEScheduler:
    """Drives the :class:`ScheduledMoELayer` s of a model from a reference-shaped allocator / dispatcher pair
    (``legacy/vescale/moe/_scheduler.py:78-277``):

    * before a layer runs, ``experts_allocator.allocate_experts(layer_id, iter)`` may return a new list of per-expert
      ``[DP, TP]`` meshes (``None`` = keep the current one) — the layer's weights and optimizer state migrate (``reallocate``);
    * the ``token_dispatcher`` is told the allocation (``set_experts_alloc``), the routed copies (``assign_task``) and chooses
      replicas (``dispatch_token``);
    * after the layer ran, both get ``collect_performance({"layer_id", "tokens_per_expert", "tokens_per_rank"}, iter)``.

    ``step_end()`` advances the iteration counter and sums the replica gradients of every layer (call it after backward)."""

    def __init__(self, experts_allocator=None, token_dispatcher=None, config: Optional[Dict] = None, optimizer: Optional[torch.optim.Optimizer] = None):
        self.allocator, self.dispatcher, self.config, self.optimizer = experts_allocator, token_dispatcher, dict(config or {}), optimizer
        self.layers: List[ScheduledMoELayer] = []
        self.iter = 0
        self._hooks = []

    def register(self, model: nn.Module) -> "MoEScheduler":
        for m in model.modules():
            if isinstance(m, ScheduledMoELayer) and m not in self.layers:
                m.layer_id = len(self.layers)
                m.dispatcher = self.dispatcher
                self.layers.append(m)
                self._hooks.append(m.register_forward_pre_hook(self._pre))
                self._hooks.append(m.register_forward_hook(self._post))
                if self.dispatcher is not None and hasattr(self.dispatcher, "set_experts_alloc"):
                    self.dispatcher.set_experts_alloc(m.alloc.info(m.experts.w_gate_up.device.type))
        return self

    def _new_allocation(self, layer: ScheduledMoELayer) -> Optional[ExpertsAllocation]:
        if self.allocator is None:
            return None
        fn = getattr(self.allocator, "allocate_experts", None)
        got = fn(layer.layer_id, self.iter) if fn is not None else None
        if got is None:
            return None
        return got if isinstance(got, ExpertsAllocation) else ExpertsAllocation(got, layer.ep_size)

    def _pre(self, layer: ScheduledMoELayer, args):
        new = self._new_allocation(layer)
        if new is not None and new != layer.alloc:
            layer.reallocate(new, self.optimizer)
        if self.dispatcher is not None and hasattr(self.dispatcher, "set_experts_alloc"):
            self.dispatcher.set_experts_alloc(layer.alloc.info(layer.experts.w_gate_up.device.type))

    def _post(self, layer: ScheduledMoELayer, args, out):
        perf = {"layer_id": layer.layer_id, "tokens_per_expert": layer.last_tokens_per_expert, "tokens_per_rank": layer.last_tokens_per_rank}
        for who in (self.allocator, self.dispatcher):
            fn = getattr(who, "collect_performance", None) if who is not None else None
            if fn is not None:
                fn(perf, self.iter)

    def step_end(self) -> None:
        for m in self.layers:
            m.sync_replica_grads()
        self.iter += 1

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
