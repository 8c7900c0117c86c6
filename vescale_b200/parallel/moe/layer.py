"""Expert-parallel MoE layer: top-k routing, token dispatch / combine over the EP group, grouped expert FFN.

Dataflow for ``T`` local tokens, ``E`` experts spread over ``W`` EP ranks (``E/W`` experts each):

    router (fp32) → top-k → sort token copies by expert (device-side; counts by ``bincount``)
    counts exchange  [W, E/W] ints                         (C13: 64 B)
    dispatch  all-to-all of token rows, uneven             (C14)
    grouped SwiGLU FFN over the local experts' ragged row groups (tcgen05 GEMMs per expert group)
    combine   all-to-all back, × gate weight, scatter-add  (C15)

Backends: ``"nccl"`` = ``all_to_all_single`` with host-side split sizes (what legacy does, incl. the
``.tolist()`` sync, ``legacy/vescale/moe/_scheduler.py:162-215``); ``"symm"`` = sm_100a put/get kernels over
symmetric memory with device-side counts (``csrc/moe_dispatch.cu``) — no host sync, rows land directly in the
destination expert's ragged buffer.  The dispatched activation is a ``RaggedShard`` over the token dim
(``ragged_token_placement``), as ``docs/texts/raggedshard.md:97-99`` suggests.

Parity: ``legacy/vescale/moe/`` (api, _scheduler, token_dispatcher, experts_allocator, _moe_param_buffer).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from ... import ops as O
from ...placement import RaggedShard

__all__ = ["MoEConfig", "MoELayer", "TopKRouter", "all_to_all_uneven", "ragged_token_placement", "GroupedExperts"]


class _AllToAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, in_splits, out_splits, group):
        ctx.in_splits, ctx.out_splits, ctx.group = in_splits, out_splits, group
        out = x.new_empty((sum(out_splits), *x.shape[1:]))
        if group is None or dist.get_world_size(group) == 1:
            return x.clone()
        if dist.get_backend(group) == "nccl":
            dist.all_to_all_single(out, x.contiguous(), out_splits, in_splits, group=group)
        else:
            from ...comm.collectives import _p2p_all_to_all

            ins = list(x.contiguous().split(in_splits, 0))
            outs = list(out.split(out_splits, 0))
            _p2p_all_to_all(outs, ins, group)
        return out

    @staticmethod
    def backward(ctx, g):
        return _AllToAll.apply(g.contiguous(), ctx.out_splits, ctx.in_splits, ctx.group), None, None, None


def all_to_all_uneven(x: torch.Tensor, in_splits: List[int], out_splits: List[int], group) -> torch.Tensor:
    return _AllToAll.apply(x, in_splits, out_splits, group)


def ragged_token_placement(tokens_per_rank: Sequence[int]) -> RaggedShard:
    """The dispatched [sum(tokens), H] activation as a DTensor placement: uneven rows per EP rank."""
    g = math.gcd(*[int(t) for t in tokens_per_rank]) or 1
    return RaggedShard((0,), tuple(int(t) // g for t in tokens_per_rank))


class MoEConfig:
    """Sizes and routing hyper-parameters of one MoE layer (Mixtral: 8 experts, top-2)."""
    def __init__(self, hidden_size: int, ffn_size: int, num_experts: int = 8, top_k: int = 2, ep_size: int = 1, dtype=torch.bfloat16, init_std: float = 0.02, comm_backend: str = "nccl", aux_loss_coef: float = 0.0):
        self.hidden_size, self.ffn_size, self.num_experts, self.top_k = hidden_size, ffn_size, num_experts, top_k
        self.ep_size, self.dtype, self.init_std, self.comm_backend, self.aux_loss_coef = ep_size, dtype, init_std, comm_backend, aux_loss_coef


class TopKRouter(nn.Module):
    """Softmax router returning (top-k weights renormalised, top-k expert ids, full probabilities)."""
    def __init__(self, cfg: MoEConfig, device=None):
        super().__init__()
        self.cfg = cfg
        self.weight = nn.Parameter(torch.empty(cfg.num_experts, cfg.hidden_size, dtype=torch.float32, device=device))

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        logits = F.linear(x.float(), self.weight)
        probs = torch.softmax(logits, dim=-1)
        topv, topi = torch.topk(probs, self.cfg.top_k, dim=-1)
        topv = topv / topv.sum(-1, keepdim=True)
        return topv, topi, probs


class GroupedExperts(nn.Module):
    """``E_local`` SwiGLU experts with stacked weights; rows arrive grouped by expert (ragged M per expert)."""

    def __init__(self, cfg: MoEConfig, num_local: int, device=None):
        super().__init__()
        self.cfg, self.num_local = cfg, num_local
        self.w_gate_up = nn.Parameter(torch.empty(num_local, 2 * cfg.ffn_size, cfg.hidden_size, dtype=cfg.dtype, device=device))
        self.w_down = nn.Parameter(torch.empty(num_local, cfg.hidden_size, cfg.ffn_size, dtype=cfg.dtype, device=device))
        # expert weights differ per EP rank: data-parallel wrappers must not average them across the EP group
        self.w_gate_up._is_expert_param = True
        self.w_down._is_expert_param = True

    def forward(self, x: torch.Tensor, rows_per_expert: List[int]) -> torch.Tensor:
        if sum(rows_per_expert) == 0:
            # no token for any local expert.  The (empty) result must still depend on BOTH the input rows and the weights: the rows
            # came out of the dispatch all-to-all, and a graph that does not reach it makes this rank skip that collective's backward
            # while its peers run theirs (hang on NCCL, size mismatch on gloo); the weights must get (zero) gradients for the optimizer.
            return x[:, : self.cfg.hidden_size] * 1.0 + 0.0 * (self.w_gate_up.sum() + self.w_down.sum()).to(x.dtype)
        return _GroupedSwiGLUFFN.apply(x, self.w_gate_up, self.w_down, tuple(int(n) for n in rows_per_expert))


class _GroupedSwiGLUFFN(torch.autograd.Function):
    """All local experts' SwiGLU FFNs over rows grouped by expert, as ONE autograd node that saves the *stacked parameters
    themselves* (not per-expert slices of them).  Under FSDP the parameters' storage is re-pointed at the freshly gathered unit
    buffer before backward (``reshard_after_forward``); a slice view taken in forward would still alias the recycled buffer of
    the forward pass, so backward re-slices the live parameter.  Weight gradients go straight into ``main_grad`` (a view of the
    unit's flat gradient buffer) when the wrapper provides one; the gate|up activation is kept, its SwiGLU product recomputed."""

    @staticmethod
    def forward(ctx, x, w_gate_up, w_down, rows):
        Fn = O.functional
        outs, gus, pos = [], [], 0
        for e, n in enumerate(rows):
            if n == 0:
                continue
            xe = x[pos : pos + n].contiguous()
            pos += n
            gu = Fn.gemm_nt(xe, w_gate_up[e])
            gus.append(gu)
            outs.append(Fn.gemm_nt(Fn._swiglu_fwd(gu), w_down[e]))
        ctx.save_for_backward(x, torch.cat(gus, 0), w_gate_up, w_down)
        ctx.rows = rows
        return torch.cat(outs, 0)

    @staticmethod
    def backward(ctx, dy):
        Fn = O.functional
        x, gu_all, w_gate_up, w_down = ctx.saved_tensors
        dy = dy.contiguous()
        grads = []
        for w in (w_gate_up, w_down):
            mg = getattr(w, "main_grad", None)
            if mg is not None:
                if not getattr(w, "_main_grad_initialised", False):
                    mg.zero_()  # experts that received no token keep a zero gradient
                    w._main_grad_initialised = True
                grads.append(mg)
            else:
                grads.append(torch.zeros_like(w))
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        pos = 0
        for e, n in enumerate(ctx.rows):
            if n == 0:
                continue
            sl = slice(pos, pos + n)
            pos += n
            ge, dye, xe = gu_all[sl], dy[sl], x[sl].contiguous()
            act = Fn._swiglu_fwd(ge)
            grads[1][e].add_(Fn.gemm_tn(dye, act).to(grads[1].dtype))
            dact = Fn.gemm_nn(dye, w_down[e])
            if Fn._use_kernels(dact) and dact.dtype == torch.bfloat16:
                dgu = Fn._ext.ops().swiglu_bwd(dact, ge.contiguous())
            else:
                dgu = Fn._swiglu_bwd_ref(dact, ge)
            grads[0][e].add_(Fn.gemm_tn(dgu, xe).to(grads[0].dtype))
            if dx is not None:
                dx[sl] = Fn.gemm_nn(dgu, w_gate_up[e])
        for w in (w_gate_up, w_down):
            hook = getattr(w, "_post_main_grad_hook", None)
            if hook is not None and getattr(w, "main_grad", None) is not None:
                hook(w)
        return dx, (None if getattr(w_gate_up, "main_grad", None) is not None else grads[0]), (None if getattr(w_down, "main_grad", None) is not None else grads[1]), None


class MoELayer(nn.Module):
    """Expert-parallel MoE layer: route → dispatch (NCCL all-to-all with host-side split sizes, or the device-side symmetric
    dispatcher) → grouped expert GEMMs → combine.  Parity: legacy ``moe/_scheduler.py:162-277``."""
    def __init__(self, cfg: MoEConfig, ep_group=None, device=None):
        super().__init__()
        self.cfg = cfg
        self.ep_group = ep_group
        self.ep_size = dist.get_world_size(ep_group) if ep_group is not None else 1
        self.ep_rank = dist.get_rank(ep_group) if ep_group is not None else 0
        assert cfg.num_experts % self.ep_size == 0
        self.num_local = cfg.num_experts // self.ep_size
        self.router = TopKRouter(cfg, device)
        self.experts = GroupedExperts(cfg, self.num_local, device)
        self.last_aux_loss: Optional[torch.Tensor] = None
        self.last_tokens_per_rank: Optional[List[int]] = None
        # global slot (= ep_rank * num_local + local index) that currently hosts each expert; identity until
        # ``reallocate_experts`` moves experts between ranks (dynamic load balancing)
        self.register_buffer("slot_of_expert", torch.arange(cfg.num_experts, device=device), persistent=True)
        self.symm_dispatcher = None  # set by ``use_symmetric_dispatch`` (sm_100a kernels, no host sync in forward)

    def reset_buffers(self) -> None:
        """(Re)initialise non-parameter state after a meta-device materialisation: experts start at their home slots."""
        self.slot_of_expert.copy_(torch.arange(self.cfg.num_experts, device=self.slot_of_expert.device))

    def use_symmetric_dispatch(self, dispatcher) -> None:
        self.symm_dispatcher = dispatcher

    def reset_parameters(self, generator=None):
        with torch.no_grad():
            self.router.weight.normal_(0, self.cfg.init_std, generator=generator)
            # every rank draws all experts so that expert e has the same weights whatever the EP size
            for e in range(self.cfg.num_experts):
                gu = torch.empty_like(self.experts.w_gate_up[0]).normal_(0, self.cfg.init_std, generator=generator)
                dn = torch.empty_like(self.experts.w_down[0]).normal_(0, self.cfg.init_std, generator=generator)
                le = e - self.ep_rank * self.num_local
                if 0 <= le < self.num_local:
                    self.experts.w_gate_up[le].copy_(gu)
                    self.experts.w_down[le].copy_(dn)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        topv, topi, probs = self.router(x2)
        return self.experts_forward(x2, topv, topi, probs).view(shape)

    def experts_forward(self, x2: torch.Tensor, topv: torch.Tensor, topi: torch.Tensor, probs: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Everything after routing: dispatch the ``[T, H]`` token rows to the ranks hosting their ``top_k`` experts, run the
        grouped expert FFN, combine with the gate weights.  Also the entry point of hijacked third-party MoE blocks, whose own
        router produced ``topv`` / ``topi`` (``hijack.py``)."""
        shape = x2.shape
        T, k, E, W, El = x2.shape[0], topi.shape[-1], self.cfg.num_experts, self.ep_size, self.num_local
        slots = self.slot_of_expert[topi]  # where each chosen expert lives now
        if self.symm_dispatcher is not None and x2.is_cuda and W > 1:
            from .symm_dispatch import symm_moe_forward

            out = symm_moe_forward(x2.contiguous(), topv, slots, self.experts.w_gate_up, self.experts.w_down, self.symm_dispatcher)
            return out.view(shape)
        x = x2
        if self.cfg.aux_loss_coef > 0 and probs is not None:
            frac = torch.zeros(E, device=x.device).index_add_(0, topi.reshape(-1), torch.ones(T * k, device=x.device)) / (T * k)
            self.last_aux_loss = self.cfg.aux_loss_coef * E * (frac * probs.mean(0)).sum()
        flat_e = slots.reshape(-1)
        order = torch.argsort(flat_e, stable=True)
        token_of = order // k
        xs = x2[token_of]
        counts = torch.bincount(flat_e, minlength=E)
        if W > 1:
            send_counts = counts.view(W, El)
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts.contiguous(), group=self.ep_group) if dist.get_backend(self.ep_group) == "nccl" else _gloo_a2a_counts(recv_counts, send_counts, self.ep_group)
            in_splits = send_counts.sum(1).tolist()  # host sync: the NCCL baseline needs split sizes on the host
            rc = recv_counts.cpu()
            out_splits = rc.sum(1).tolist()
            recv = all_to_all_uneven(xs, in_splits, out_splits, self.ep_group)
            # received rows are ordered (src rank, local expert): regroup by local expert
            eid = torch.repeat_interleave(torch.arange(El, device=x.device).repeat(W), recv_counts.reshape(-1))
            perm2 = torch.argsort(eid, stable=True)
            rows = rc.sum(0).tolist()
            y = self.experts(recv[perm2], rows)
            inv2 = torch.empty_like(perm2)
            inv2[perm2] = torch.arange(perm2.numel(), device=x.device)
            back = all_to_all_uneven(y[inv2], out_splits, in_splits, self.ep_group)
            self.last_tokens_per_rank = out_splits
        else:
            back = self.experts(xs, counts.tolist())
        w = topv.reshape(-1)[order].to(back.dtype).unsqueeze(-1)
        out = torch.zeros_like(x2).index_add_(0, token_of, back * w)
        return out.view(shape)


def _gloo_a2a_counts(recv, send, group):
    from ...comm.collectives import _p2p_all_to_all

    outs = list(recv.unbind(0))
    _p2p_all_to_all(outs, list(send.unbind(0)), group)


def global_all_to_all_single(tensor: torch.Tensor, input_split_sizes: Optional[List[int]] = None, output_split_sizes: Optional[List[int]] = None, async_op: bool = False, group=None) -> torch.Tensor:
    """Differentiable ``all_to_all_single`` along dim 0 (legacy ``moe/_utils.py:26-67``): even split when no sizes are given.  The
    gradient is the all-to-all with the split lists swapped."""
    n = dist.get_world_size(group) if dist.is_initialized() else 1
    if input_split_sizes is None:
        if tensor.shape[0] % n:
            raise ValueError(f"an even all-to-all needs dim 0 ({tensor.shape[0]}) divisible by the group size ({n})")
        input_split_sizes = [tensor.shape[0] // n] * n
    if output_split_sizes is None:
        output_split_sizes = list(input_split_sizes)
    return _AllToAll.apply(tensor, list(input_split_sizes), list(output_split_sizes), group)


_AllToAllSingle = _AllToAll  # the reference's name for the autograd function
__all__ += ["global_all_to_all_single"]
