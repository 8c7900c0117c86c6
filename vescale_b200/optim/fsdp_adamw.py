"""FSDPAdamW: AdamW over the flat unit shards of a ``fully_shard``-ed model.

One fused launch per unit (``csrc/adamw.cu``): reads the reduce-scattered gradient shard (fp32, or bf16 at
world size 1), applies the global-norm clip coefficient *from a device scalar* (no host sync), updates the
fp32 master / exp_avg / exp_avg_sq shards and writes the bf16 parameter shard that the next all-gather reads
— i.e. "copy grads→main, clip, Adam, copy main→model params" of the reference's DistributedOptimizer
(``legacy/vescale/optim/distributed_optimizer.py:1142-1261``) in a single pass over HBM.  Weight decay is
per-parameter through a small segment table (norm weights and other 1-D params are not decayed).

When ``max_grad_norm is None`` and the symmetric-memory backend is active, ``fused_reduce=True`` folds the
update into the reduce-scatter kernel itself (``rs_adamw``): the reduced gradient never touches HBM.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ..ops import _ext
from ..profiler import ndtimeit, predefined
from ..parallel.fsdp.api import FSDPState, get_fsdp_state
from ..parallel.fsdp.unit import FSDPUnit

__all__ = ["FSDPAdamW"]


class FSDPAdamW:
    """See the module docstring.  ``tp_group`` / ``replicate_group`` add the 2-D cases (FSDP × TP, HSDP)."""
    def __init__(
        self,
        model: torch.nn.Module,
        lr: float = 1e-3,
        betas: Tuple[float, float] = (0.9, 0.95),
        eps: float = 1e-8,
        weight_decay: float = 0.1,
        max_grad_norm: Optional[float] = 1.0,
        no_decay: Optional[Callable[[str, Sequence[int]], bool]] = None,
        state_dtype: torch.dtype = torch.float32,
        fused_reduce: bool = False,
        tp_group=None,
        tp_sharded: Optional[Callable[[str], bool]] = None,
        replicate_group=None,
    ):
        st = None
        for m in model.modules():
            st = get_fsdp_state(m)
            if st is not None:
                break
        if st is None:
            raise ValueError("model is not wrapped with fully_shard")
        self.state: FSDPState = st
        self.units: List[FSDPUnit] = st.units
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        self.no_decay = no_decay or (lambda name, shape: len(shape) <= 1)
        dev = st.device
        for u in self.units:
            u.exp_avg = torch.zeros(u.S, dtype=state_dtype, device=dev)
            u.exp_avg_sq = torch.zeros(u.S, dtype=state_dtype, device=dev)
            segs = []
            for slot in u.layout.slots:
                lo, hi = u.layout.rank_range(slot, u.rank)
                if hi > lo:
                    segs.append((lo, hi, 0.0 if self.no_decay(slot.name, slot.shape) else 1.0))
            u.wd_segments = segs
            if dev.type == "cuda":
                # [n, 3] int64/float table consumed by the kernel: (lo, hi, decay_flag)
                u.wd_table = torch.tensor([[s[0], s[1], int(s[2])] for s in segs] or [[0, 0, 0]], dtype=torch.int64, device=dev)
        # ---- 2-D (FSDP x TP): parameters that are replicated over the TP group (norms, embedding, lm-head under sequence
        # parallelism) carry partial gradients; they are summed over ``tp_group`` once per step, and the global grad norm
        # counts TP-sharded parameters over all TP ranks but replicated ones once (legacy ``_grad_sync.py:98-101``,
        # ``clip_grads.py:60-110``).
        self.tp_group = tp_group if (tp_group is not None and dist.get_world_size(tp_group) > 1) else None
        if self.tp_group is not None:
            if fused_reduce:
                raise ValueError("fused_reduce cannot be combined with tp_group (replicated gradients need a TP all-reduce first)")
            from ..models.llama_tp import TP_SHARDED_PARAMS

            is_sharded = tp_sharded or (lambda name: name.rsplit(".", 1)[-1] in TP_SHARDED_PARAMS)
            for u in self.units:
                u.tp_replicated_segments = []
                for slot in u.layout.slots:
                    lo, hi = u.layout.rank_range(slot, u.rank)
                    if hi > lo and not is_sharded(slot.name):
                        u.tp_replicated_segments.append((lo, hi))
                real = sum(hi - lo for lo, hi, _ in u.layout.segments(u.rank))
                u.tp_all_replicated = sum(hi - lo for lo, hi in u.tp_replicated_segments) == real
        # ---- HSDP (replicate x shard): the model is sharded over the FSDP mesh dim and replicated over ``replicate_group``
        # (each replica sees different data); reduce-scattered gradient shards are averaged over the replicas once per step.
        self.replicate_group = replicate_group if (replicate_group is not None and dist.get_world_size(replicate_group) > 1) else None
        if self.replicate_group is not None and fused_reduce:
            raise ValueError("fused_reduce cannot be combined with replicate_group (gradients need the replica average first)")
        self._norm_buf = torch.zeros(1, dtype=torch.float32, device=dev)
        self._norm_rep = torch.zeros(1, dtype=torch.float32, device=dev)
        self._coef = torch.ones(1, dtype=torch.float32, device=dev)
        self.last_grad_norm: Optional[torch.Tensor] = None
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay}]
        # fully fused mode: the update happens inside the reduce-scatter kernel during backward.  It needs the
        # clip coefficient before the global norm of *this* step exists, so it is only legal without clipping.
        self.fused_reduce = bool(fused_reduce)
        if self.fused_reduce:
            if max_grad_norm is not None:
                raise ValueError("fused_reduce=True folds AdamW into the reduce-scatter kernel and cannot apply a global-norm clip; pass max_grad_norm=None")
            if st.comm is None or not getattr(st.comm, "symmetric", False):
                raise ValueError("fused_reduce=True needs the symmetric-memory comm backend")
            st.fused_optimizer = self

    def fused_update(self, u: FSDPUnit, scale: float) -> None:
        """Called from post-backward on the reduce-scatter stream (see FSDPState.post_backward)."""
        step = self.step_count + 1
        b1, b2 = self.betas
        hp = dict(coef=self._coef, lr=float(self.param_groups[0]["lr"]), b1=float(b1), b2=float(b2), eps=float(self.eps), wd=float(self.weight_decay),
                  bc1=float(1.0 - b1**step), bc2=float(1.0 - b2**step))
        u.comm.reduce_scatter_adamw(u.full_grad, u, scale, hp)
        u.grad_ready = False
        u.bf16_fresh = True
        u._fused_done = True

    # ------------------------------------------------------------------ 2-D: TP-replicated gradients
    def _grad_view(self, u: FSDPUnit) -> torch.Tensor:
        g = u.grad_shard
        return g[u.rank * u.S : (u.rank + 1) * u.S] if g.numel() != u.S else g

    def _sync_tp_replicated_grads(self) -> None:
        small: List[torch.Tensor] = []
        for u in self.units:
            if not u.grad_ready or not u.tp_replicated_segments:
                continue
            g = self._grad_view(u)
            if u.tp_all_replicated:
                dist.all_reduce(g, group=self.tp_group)
            else:
                small.extend(g[lo:hi] for lo, hi in u.tp_replicated_segments)
        if small:  # norm weights of every block: one flat bucket
            flat = torch.cat([t.reshape(-1) for t in small])
            dist.all_reduce(flat, group=self.tp_group)
            off = 0
            for t in small:
                t.copy_(flat[off : off + t.numel()])
                off += t.numel()

    def _sumsq_into(self, g: torch.Tensor, out: torch.Tensor, sc: float) -> None:
        if g.numel() == 0:
            return
        if g.is_cuda and _ext.available() and g.data_ptr() % 16 == 0:
            _ext.count_launch("sumsq")
            _ext.ops().sumsq_accumulate(g, out, float(sc))
        else:
            out += (g.float() * sc).pow(2).sum()

    def _global_grad_norm_2d(self) -> torch.Tensor:
        tot, rep = self._norm_buf.zero_(), self._norm_rep.zero_()
        for u in self.units:
            if not u.grad_ready:
                continue
            g, sc = self._grad_view(u), getattr(u, "grad_scale_pending", 1.0)
            self._sumsq_into(g, tot, sc)
            for lo, hi in u.tp_replicated_segments:
                self._sumsq_into(g[lo:hi], rep, sc)
        tot -= rep  # TP-sharded part only
        dist.all_reduce(tot, group=self.tp_group)
        tot += rep
        if self.units and self.units[0].world > 1:
            dist.all_reduce(tot, group=self.units[0].group)
        return tot.clamp_(min=0).sqrt()

    # ------------------------------------------------------------------ grad norm (device-side)
    def _global_grad_norm(self) -> torch.Tensor:
        if self.tp_group is not None:
            return self._global_grad_norm_2d()
        tot = self._norm_buf.zero_()
        for u in self.units:
            if not u.grad_ready:
                continue
            if u.sumsq is not None:  # produced by the fused reduce-scatter kernel
                tot += u.sumsq
                continue
            g = u.grad_shard
            sc = getattr(u, "grad_scale_pending", 1.0)
            if g.is_cuda and _ext.available():
                _ext.count_launch("sumsq")
                _ext.ops().sumsq_accumulate(g, tot, float(sc))
            else:
                tot += (g.float() * sc).pow(2).sum()
        if self.state.mesh.has_groups() and self.units and self.units[0].world > 1:
            dist.all_reduce(tot, group=self.units[0].group)
        return tot.sqrt()

    @torch.no_grad()
    def step(self) -> Optional[torch.Tensor]:
        with ndtimeit(predefined.OPTIMIZER_STEP):
            return self._step()

    def _step(self) -> Optional[torch.Tensor]:
        st = self.state
        st.wait_grads()
        self.step_count += 1
        if self.fused_reduce:
            # every unit was already updated by its reduce-scatter kernel; only book-keeping is left
            tot = self._norm_buf.zero_()
            for u in self.units:
                if u.sumsq is not None:
                    tot += u.sumsq
                u._fused_done = False
            if self.units and self.units[0].world > 1:
                dist.all_reduce(tot, group=self.units[0].group)
            self.last_grad_norm = tot.sqrt()
            st.invalidate_params()
            st.iteration += 1
            return self.last_grad_norm
        b1, b2 = self.betas
        lr = self.param_groups[0]["lr"]
        bc1 = 1.0 - b1**self.step_count
        bc2 = 1.0 - b2**self.step_count
        norm = None
        if self.replicate_group is not None:
            n_rep = dist.get_world_size(self.replicate_group)
            for u in self.units:
                if u.grad_ready:
                    g = self._grad_view(u)
                    dist.all_reduce(g, group=self.replicate_group)
                    g.div_(n_rep)
                    u.sumsq = None  # the reduce-scatter kernel's partial norm predates the replica average
        if self.tp_group is not None:
            self._sync_tp_replicated_grads()
        if self.max_grad_norm is not None:
            norm = self._global_grad_norm()
            torch.clamp(self.max_grad_norm / (norm + 1e-6), max=1.0, out=self._coef)
            self.last_grad_norm = norm
        else:
            self._coef.fill_(1.0)
        for u in self.units:
            if not u.grad_ready:
                continue
            g = u.grad_shard
            sc = getattr(u, "grad_scale_pending", 1.0)
            if g.is_cuda and _ext.available():
                _ext.count_launch("adamw")
                _ext.ops().fused_adamw_(
                    u.master, u.exp_avg, u.exp_avg_sq, g, u.param_shard, u.wd_table, self._coef,
                    float(lr), float(b1), float(b2), float(self.eps), float(self.weight_decay), float(bc1), float(bc2), float(sc),
                )
            else:
                self._step_reference(u, g, sc, lr, b1, b2, bc1, bc2)
            u.bf16_fresh = True
            u.grad_ready = False
            u.sumsq = None
        st.invalidate_params()
        st.iteration += 1
        return norm

    def _step_reference(self, u: FSDPUnit, g, sc, lr, b1, b2, bc1, bc2) -> None:
        gf = g.float() * (self._coef * sc)
        u.exp_avg.mul_(b1).add_(gf, alpha=1 - b1)
        u.exp_avg_sq.mul_(b2).addcmul_(gf, gf, value=1 - b2)
        denom = (u.exp_avg_sq / bc2).sqrt_().add_(self.eps)
        upd = (u.exp_avg / bc1) / denom
        for lo, hi, dec in u.wd_segments:
            if dec:
                u.master[lo:hi].mul_(1 - lr * self.weight_decay)
        # padding has zero grad and zero state: update is 0 there
        u.master.add_(upd, alpha=-lr)
        u.param_shard.copy_(u.master)

    def zero_grad(self, set_to_none: bool = True) -> None:
        for u in self.units:
            u.zero_grad()

    # ------------------------------------------------------------------ checkpointing (DTensor views → DCP)
    def state_dict(self) -> dict:
        """Optimizer state as RaggedShard DTensors keyed like the parameters, so DCP can reshard it
        (reference ``OptimizerStateSpec`` resharding, ``distributed_optimizer.py:51-93,748-880``)."""
        from ..dtensor.api import DTensor

        out = {"step": self.step_count, "state": {}}
        for u in self.units:
            for slot, sp in zip(u.layout.slots, u.sharded_params):
                lo, hi = u.layout.rank_range(slot, u.rank)
                key = f"{u.name}.{u._index}.{slot.name}"
                out["state"][key] = {
                    "exp_avg": DTensor(u.exp_avg[lo:hi], sp.data._spec),
                    "exp_avg_sq": DTensor(u.exp_avg_sq[lo:hi], sp.data._spec),
                }
        return out

    def load_state_dict(self, sd: dict) -> None:
        self.step_count = int(sd.get("step", 0))
        for u in self.units:
            for slot in u.layout.slots:
                lo, hi = u.layout.rank_range(slot, u.rank)
                key = f"{u.name}.{u._index}.{slot.name}"
                ent = sd["state"].get(key)
                if ent is None:
                    continue
                for nm, buf in (("exp_avg", u.exp_avg), ("exp_avg_sq", u.exp_avg_sq)):
                    t = ent[nm]
                    t = t._local_tensor if hasattr(t, "_local_tensor") else t
                    buf[lo:hi].copy_(t.reshape(-1))
