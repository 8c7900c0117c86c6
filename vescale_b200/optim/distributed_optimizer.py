"""DistributedOptimizer (ZeRO-2+): optimizer state and fp32 master weights sharded over the DP group.

Works on top of ``DistributedDataParallel(use_distributed_optimizer=True)``:

* every gradient bucket is reduce-scattered; DP rank r owns the r-th equal slice of each bucket (slice
  boundaries ignore parameter edges, ``legacy/vescale/optim/distributed_optimizer.py:454-517``);
* fp32 *main* shards mirror the owned slices; the wrapped optimizer's param groups are re-pointed at views of
  the main shards (one view per parameter piece, so per-group hyper-parameters survive);
* ``step``: main_grad ← owned grad slice (fp32), global-norm clip, inner ``optimizer.step()``, model-dtype copy
  into the owned slice of the flat *parameter buffer*, then one ``all_gather_into_tensor`` per bucket — eagerly,
  or deferred to forward pre-hooks so the gather of bucket i+1 overlaps the forward that uses bucket i
  (``overlap_param_gather``, ``:995-1076``); module parameters are views into the parameter buffer, so the
  gather is zero-copy;
* ``state_dict`` exposes each state tensor with its ``OptimizerStateSpec`` (global shape / local shape /
  global offset of the owned piece) for resharding checkpoints (``:51-93,748-880``).

On B200 the same step for FSDP-wrapped models is the fused kernel path (``FSDPAdamW``); this class is the general
wrapper for arbitrary torch optimizers and DDP models.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from ..dtensor.api import DTensor
from ..parallel.ddp import DistributedDataParallel, GradBuffer, _local
from .base_optimizer import OptimizerBase
from .clip_grads import get_grad_norm_fp32

__all__ = ["DistributedOptimizer", "OptimizerStateSpec", "Range", "convert_dict_with_sharded", "convert_dict_sharded_to_tensor", "initialize_optimizer_state"]


@dataclass
class OptimizerStateSpec:
    global_shape: Tuple[int, ...]
    local_shape: Tuple[int, ...]
    global_offset: Tuple[int, ...]
    local_tensor: torch.Tensor
    dp_ranks_ranges: Optional[Dict[int, Tuple[int, int]]] = None


class Range:
    """Half-open index range ``[start, end)`` of a shard inside a flat buffer (legacy ``distributed_optimizer.py:26-48``)."""

    __slots__ = ("start", "end", "size")

    def __init__(self, start: int, end: int):
        if end < start:
            raise ValueError(f"empty-negative range [{start}, {end})")
        self.start, self.end, self.size = int(start), int(end), int(end) - int(start)

    def normalize(self, start: int = 0) -> "Range":
        """The same length re-based at ``start`` (world offset -> offset inside a bucket / a parameter)."""
        return Range(start, start + self.size)

    def intersect(self, other: "Range") -> Optional["Range"]:
        lo, hi = max(self.start, other.start), min(self.end, other.end)
        return Range(lo, hi) if hi > lo else None

    def __len__(self) -> int:
        return self.size

    def __eq__(self, other) -> bool:
        return isinstance(other, Range) and (self.start, self.end) == (other.start, other.end)

    def __hash__(self) -> int:
        return hash((self.start, self.end))

    def __repr__(self) -> str:
        return f"Range({self.start},{self.end} [{self.size}])"

    __str__ = __repr__


def convert_dict_with_sharded(param_state: dict, global_shape: Tuple[int, ...], local_shape: Tuple[int, ...], global_offset: Tuple[int, ...], dp_ranks_ranges: Optional[Dict[int, "Range"]] = None) -> dict:
    """One parameter's optimizer state for saving: every tensor entry (``exp_avg`` ...) is wrapped in an ``OptimizerStateSpec`` that
    says where the flat piece sits in the global tensor; scalars (``step``) pass through.  Without ``dp_ranks_ranges`` the piece must
    be the whole model-parallel shard."""
    import math

    out = {}
    for k, v in param_state.items():
        if isinstance(v, torch.Tensor) and v.dim() >= 1:
            if not dp_ranks_ranges and math.prod(local_shape) != v.numel():
                raise ValueError(f"state {k!r}: {v.numel()} elements do not fill the local shard {tuple(local_shape)} of global {tuple(global_shape)} at {tuple(global_offset)}")
            out[k] = OptimizerStateSpec(tuple(global_shape), tuple(local_shape), tuple(global_offset), v, dp_ranks_ranges)
        else:
            out[k] = v
    return out


def convert_dict_sharded_to_tensor(param_state: dict, range_1d: Optional["Range"] = None) -> dict:
    """The inverse after loading: specs back to flat tensors, cut to this rank's ``range_1d`` of the shard when the state is spread
    over several DP ranks.  In place; returns the dict."""
    for k, v in param_state.items():
        if isinstance(v, OptimizerStateSpec):
            flat = v.local_tensor.flatten()
            param_state[k] = flat[range_1d.start:range_1d.end] if range_1d is not None else flat
    return param_state


class _BucketShard:
    def __init__(self, gb: GradBuffer, bucket, dtype, rank: int, dp: int):
        self.gb = gb
        self.bucket = bucket
        n = bucket.data.numel() // dp
        self.lo = bucket.offset + rank * n  # offsets in the dtype's flat buffer
        self.hi = self.lo + n
        self.n = n


class DistributedOptimizer(OptimizerBase):
    """ZeRO-2+ wrapper: every DDP bucket is sharded evenly over the DP group, fp32 main-parameter shards are updated by the inner
    optimizer and all-gathered back into the (aliased) parameter buffer, optionally overlapped with the next forward; the state
    dict is expressed as ``OptimizerStateSpec``s so it reshards on load.  Parity: legacy ``optim/distributed_optimizer.py:131-1296``."""
    def __init__(
        self,
        optimizer: torch.optim.Optimizer,
        models: Sequence[DistributedDataParallel],
        *,
        clip_grad: float = 0.0,
        overlap_param_gather: bool = False,
        grad_to_fp32: bool = True,
        extra_norm_groups: Sequence = (),
    ):
        self.optimizer = optimizer
        self.models = list(models) if isinstance(models, (list, tuple)) else [models]
        self.clip_grad = clip_grad
        self.overlap_param_gather = overlap_param_gather
        self.extra_norm_groups = list(extra_norm_groups)
        m0 = self.models[0]
        assert all(isinstance(m, DistributedDataParallel) and m.use_distributed_optimizer for m in self.models), "models must be DDP(use_distributed_optimizer=True)"
        self.group = m0.group
        self.dp = m0.dp_size
        self.rank = dist.get_rank(self.group) if self.group is not None and self.dp > 1 else 0

        # ---- flat parameter buffers mirroring the grad buffers; parameters become views
        self.param_buffers: Dict[Tuple[int, torch.dtype], torch.Tensor] = {}
        self.shards: List[_BucketShard] = []
        self.main_shards: Dict[int, torch.Tensor] = {}  # id(_BucketShard) -> fp32 flat
        self.piece_of: Dict[int, List[Tuple[_BucketShard, int, int, int]]] = {}  # id(param) -> [(shard, p_lo, p_hi, shard_off)]
        for mi, m in enumerate(self.models):
            for dt, gb in m.grad_buffers.items():
                pdt = _local(gb.params[0]).dtype
                pbuf = torch.empty(gb.numel, dtype=pdt, device=gb.data.device)
                self.param_buffers[(mi, dt)] = pbuf
                for p in gb.params:
                    s, e, _ = gb.param_index[id(p)]
                    lp = _local(p)
                    pbuf[s:e].copy_(lp.detach().reshape(-1))
                    view = pbuf[s:e].view(lp.shape)
                    # a DModule parameter IS a DTensor (wrapper subclass): its local shard is re-pointed on the parameter object
                    # itself — ``p.data`` of a wrapper subclass is a fresh alias object, assigning to it would be lost and the
                    # optimizer would update a buffer the model never reads
                    if isinstance(p, DTensor):
                        p._local_tensor = view
                    elif isinstance(p.data, DTensor):
                        p.data._local_tensor = view
                    else:
                        p.data = view
                for b in gb.buckets:
                    sh = _BucketShard(gb, b, dt, self.rank, self.dp)
                    sh.pbuf = pbuf
                    self.shards.append(sh)
                    self.main_shards[id(sh)] = pbuf[sh.lo : sh.hi].float().clone()
                    for p in b.params:
                        s, e, _ = gb.param_index[id(p)]
                        lo, hi = max(s, sh.lo), min(e, sh.hi)
                        if hi > lo:
                            self.piece_of.setdefault(id(p), []).append((sh, lo - s, hi - s, lo - sh.lo))
        # ---- re-point the inner optimizer at fp32 main pieces
        self._opt_param_ids = {id(p) for g in self.optimizer.param_groups for p in g["params"]}
        self.main_params: Dict[int, List[nn.Parameter]] = {}
        for g in self.optimizer.param_groups:
            new = []
            for p in g["params"]:
                for sh, p_lo, p_hi, off in self.piece_of.get(id(p), []):
                    mp = nn.Parameter(self.main_shards[id(sh)][off : off + (p_hi - p_lo)], requires_grad=True)
                    mp._orig_param, mp._piece = p, (p_lo, p_hi)
                    mp._shard, mp._off = sh, off
                    self.main_params.setdefault(id(p), []).append(mp)
                    new.append(mp)
            g["params"] = new
        self.optimizer.state.clear()
        self._pending_gathers: List[Tuple[_BucketShard, object]] = []
        self._hooks = []
        if overlap_param_gather:
            self._install_forward_hooks()

    @property
    def param_groups(self):
        return self.optimizer.param_groups

    # ------------------------------------------------------------------ step
    def _main_grads(self) -> List[torch.Tensor]:
        gs = []
        for plist in self.main_params.values():
            for mp in plist:
                sh, off = mp._shard, mp._off
                g = sh.gb.data[sh.lo + off : sh.lo + off + mp.numel()]
                mp.grad = g.float()
                gs.append(mp.grad)
        return gs

    @torch.no_grad()
    def step(self, closure=None):
        for m in self.models:
            m.finish_grad_sync()
        grads = self._main_grads()
        norm = None
        if self.clip_grad and self.clip_grad > 0:
            # model-parallel aware: each piece carries the spec of the parameter it belongs to (TP-sharded pieces are summed over
            # the TP mesh, TP-replicated ones counted once), then the DP group joins the ZeRO shards
            from .clip_grads import _spec_of

            specs = [_spec_of(mp._orig_param) for plist in self.main_params.values() for mp in plist]
            norm = get_grad_norm_fp32(grads, 2.0, ([self.group] if self.dp > 1 else []) + self.extra_norm_groups, specs)
            coef = torch.clamp(self.clip_grad / (norm + 1e-6), max=1.0)
            if grads:
                torch._foreach_mul_(grads, coef)
        self.optimizer.step(closure)
        # main -> model dtype into the owned slice of the parameter buffer, then gather
        for sh in self.shards:
            sh.pbuf[sh.lo : sh.hi].copy_(self.main_shards[id(sh)])
        if self.dp > 1:
            for sh in self.shards:
                full = sh.pbuf[sh.bucket.offset : sh.bucket.offset + sh.bucket.data.numel()]
                mine = sh.pbuf[sh.lo : sh.hi]
                if self.overlap_param_gather:
                    self._pending_gathers.append((sh, dist.all_gather_into_tensor(full, mine, group=self.group, async_op=True)))
                else:
                    dist.all_gather_into_tensor(full, mine, group=self.group)
        return norm

    def _install_forward_hooks(self):
        """Wait for the parameter gathers right before the first module that needs them runs."""

        def pre(mod, args):
            self.finish_param_gather()

        for m in self.models:
            self._hooks.append(m.module.register_forward_pre_hook(pre))

    def finish_param_gather(self):
        for _, h in self._pending_gathers:
            if h is not None:
                h.wait()
        self._pending_gathers.clear()

    def zero_grad(self, set_to_none: bool = True):
        self.optimizer.zero_grad(set_to_none)
        for m in self.models:
            m.zero_grad_buffer()

    # ------------------------------------------------------------------ resharding checkpoint view
    def _dp_mesh(self):
        if getattr(self, "_mesh", None) is None:
            from ..mesh import DeviceMesh

            dev = next(iter(self.param_buffers.values())).device.type
            if self.group is not None and self.dp > 1:
                ranks = dist.get_process_group_ranks(self.group)
                self._mesh = DeviceMesh(dev, list(ranks), mesh_dim_names=("DP",), _init_process_groups=False, _dim_groups=[self.group])
            else:
                self._mesh = DeviceMesh(dev, [dist.get_rank() if dist.is_initialized() else 0], mesh_dim_names=("DP",), _init_process_groups=False)
        return self._mesh

    def _ensure_state(self, mp) -> dict:
        """Optimizer state of a main-parameter piece, created the way the inner optimizer would on its first step (so that a
        fresh optimizer can be loaded into)."""
        st = self.optimizer.state.setdefault(mp, {})
        if not st:
            name = type(self.optimizer).__name__.lower()
            if "adam" in name:
                st["step"] = torch.zeros((), dtype=torch.float32, device=mp.device)
                st["exp_avg"] = torch.zeros_like(mp.data)
                st["exp_avg_sq"] = torch.zeros_like(mp.data)
            elif "sgd" in name and any(g.get("momentum", 0) for g in self.optimizer.param_groups):
                st["momentum_buffer"] = torch.zeros_like(mp.data)
        return st

    def checkpoint_state(self) -> dict:
        """Flat ``{"<param fqn>.<key>": DTensor}`` view for ``vescale_b200.checkpoint``: every parameter's main weights and
        optimizer moments are 1-D ``RaggedShard`` DTensors over the DP group whose ``local_units`` are the element counts each DP
        rank owns (bucket shards ignore parameter edges, so the counts are uneven and often zero).  Saved this way the state
        reloads under any other DP size or bucket size (legacy ``OptimizerStateSpec`` resharding,
        ``optim/distributed_optimizer.py:51-93,748-880``)."""
        from ..placement import RaggedShard
        from ..spec import DTensorSpec, TensorMeta

        mesh = self._dp_mesh()
        names = {}
        for m in self.models:
            names.update(m.param_names)
        out: Dict[str, object] = {}
        steps = {}
        for mi, m in enumerate(self.models):
            for dt, gb in m.grad_buffers.items():
                for b in gb.buckets:
                    n = b.data.numel() // self.dp
                    for p in b.params:
                        s, e, _ = gb.param_index[id(p)]
                        numel = e - s
                        units = tuple(max(0, min(e, b.offset + (r + 1) * n) - max(s, b.offset + r * n)) for r in range(self.dp))
                        spec_of = lambda dtype: DTensorSpec(mesh, (RaggedShard((0,), units),), TensorMeta((numel,), (1,), dtype))  # noqa: E731
                        mine = [mp for mp in self.main_params.get(id(p), [])]
                        fq = names.get(id(p), str(id(p)))
                        geom = self._model_parallel_geometry(p)
                        if geom is None:
                            wrap = lambda t, lo, hi: DTensor(t, spec_of(t.dtype))  # noqa: E731
                        else:
                            # the parameter is itself sharded over a model-parallel mesh: this rank's range of the LOCAL shard is
                            # saved as boxes of the GLOBAL tensor (checkpoint/flat_piece.py) — every (DP, TP) rank writes disjoint
                            # boxes under one key, reloadable under another DP size, bucket size or TP degree
                            from ..checkpoint.flat_piece import FlatPiece

                            wrap = lambda t, lo, hi: FlatPiece(t.reshape(-1), geom[0], lo, hi, geom[1], geom[2])  # noqa: E731
                        if mine:
                            mp = mine[0]
                            st = self._ensure_state(mp)
                            lo, hi = mp._piece
                            out[f"{fq}.main"] = wrap(mp.data, lo, hi)
                            for k, v in st.items():
                                if torch.is_tensor(v) and v.numel() == mp.numel() and v.dim() >= 1:
                                    out[f"{fq}.{k}"] = wrap(v, lo, hi)
                                elif k == "step":
                                    steps[fq] = float(v)
                        elif id(p) in self._opt_param_ids:
                            # this rank owns no element of the parameter: empty local shards keep the key set identical on all ranks
                            ref_dtype = torch.float32
                            out[f"{fq}.main"] = wrap(torch.empty(0, dtype=ref_dtype, device=b.data.device), 0, 0)
                            for k in self._state_keys():
                                out[f"{fq}.{k}"] = wrap(torch.empty(0, dtype=ref_dtype, device=b.data.device), 0, 0)
        out["__steps__"] = steps
        return out

    @staticmethod
    def _model_parallel_geometry(p):
        """``(local shape, global shape, global offset of the local shard)`` of a parameter that is a DTensor sharded over a
        model-parallel mesh, ``None`` for plain / fully replicated parameters (their flat ranges are saved as 1-D RaggedShard
        DTensors over the DP group).  Layouts whose local shard is not one box of the global tensor (``InterleavedShard``) are not
        expressible this way and raise."""
        if not isinstance(p, DTensor) or all(pl.is_replicate() for pl in p.placements):
            return None
        from ..layout import local_boxes

        boxes = list(local_boxes(p.shape, p.device_mesh, p.placements))
        local_shape = tuple(p._local_tensor.shape)
        if len(boxes) != 1 or tuple(boxes[0][1]) != local_shape:
            raise NotImplementedError(f"DistributedOptimizer.checkpoint_state: parameter with placements {p.placements} has a local shard that is not one box of the global tensor")
        return local_shape, tuple(p.shape), tuple(boxes[0][0])

    def _state_keys(self):
        name = type(self.optimizer).__name__.lower()
        if "adam" in name:
            return ("exp_avg", "exp_avg_sq")
        if "sgd" in name and any(g.get("momentum", 0) for g in self.optimizer.param_groups):
            return ("momentum_buffer",)
        return ()

    def load_checkpoint_state(self, flat: dict) -> None:
        """After ``checkpoint.load`` filled the DTensors of ``checkpoint_state()`` in place: restore step counters and push
        the loaded main weights into the model's parameter buffer (owned slice, then all-gather)."""
        steps = flat.get("__steps__", {}) or {}
        names = {}
        for m in self.models:
            names.update(m.param_names)
        step_val = max(steps.values()) if steps else None
        for pid, plist in self.main_params.items():
            for mp in plist:
                st = self.optimizer.state.get(mp, {})
                if "step" in st and (names.get(pid) in steps or step_val is not None):
                    st["step"].fill_(steps.get(names.get(pid), step_val))
        for sh in self.shards:
            sh.pbuf[sh.lo : sh.hi].copy_(self.main_shards[id(sh)])
        if self.dp > 1:
            for sh in self.shards:
                full = sh.pbuf[sh.bucket.offset : sh.bucket.offset + sh.bucket.data.numel()]
                dist.all_gather_into_tensor(full, sh.pbuf[sh.lo : sh.hi], group=self.group)

    # ------------------------------------------------------------------ checkpoint
    def state_dict(self) -> dict:
        """{param_name: {state_key: OptimizerStateSpec}} + the inner optimizer's param_groups (without tensors)."""
        out = {"param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.optimizer.param_groups], "state": {}}
        names = {}
        for m in self.models:
            names.update(m.param_names)
        for pid, plist in self.main_params.items():
            for mp in plist:
                p = mp._orig_param
                numel = _local(p).numel()
                st = self.optimizer.state.get(mp, {})
                ent = {"main": OptimizerStateSpec((numel,), (mp.numel(),), (mp._piece[0],), mp.data)}
                for k, v in st.items():
                    if torch.is_tensor(v) and v.numel() == mp.numel():
                        ent[k] = OptimizerStateSpec((numel,), (mp.numel(),), (mp._piece[0],), v)
                    else:
                        ent[k] = v
                out["state"].setdefault(names.get(pid, str(pid)), []).append(ent)
        return out

    def load_state_dict(self, sd: dict) -> None:
        names = {}
        for m in self.models:
            names.update(m.param_names)
        for g, saved in zip(self.optimizer.param_groups, sd.get("param_groups", [])):
            g.update(saved)
        for pid, plist in self.main_params.items():
            ents = sd["state"].get(names.get(pid, str(pid)), [])
            for mp in plist:
                for ent in ents:
                    spec = ent["main"]
                    s_lo = spec.global_offset[0]
                    s_hi = s_lo + spec.local_shape[0]
                    lo, hi = max(s_lo, mp._piece[0]), min(s_hi, mp._piece[1])
                    if hi <= lo:
                        continue
                    for k, v in ent.items():
                        src = v.local_tensor if isinstance(v, OptimizerStateSpec) else None
                        if src is None:
                            self.optimizer.state.setdefault(mp, {})[k] = v
                            continue
                        if k == "main":
                            dst = mp.data
                        else:
                            dst = self.optimizer.state.setdefault(mp, {}).setdefault(k, torch.zeros_like(mp.data))
                        dst[lo - mp._piece[0] : hi - mp._piece[0]].copy_(src[lo - s_lo : hi - s_lo])
        for sh in self.shards:
            sh.pbuf[sh.lo : sh.hi].copy_(self.main_shards[id(sh)])


def initialize_optimizer_state(optimizer: "DistributedOptimizer") -> None:
    """Create the inner optimizer's per-parameter state (``exp_avg`` ...) without taking a step, so that a freshly built optimizer
    has something to load a checkpoint INTO (legacy ``distributed_optimizer.py:1289-1296`` / ``checkpoint_helper.py``)."""
    for g in optimizer.optimizer.param_groups:
        for mp in g["params"]:
            optimizer._ensure_state(mp)
