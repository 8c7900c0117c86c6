"""Sharding propagation: rule registry, tensor-meta inference on the meta device, LRU cache.

Parity: ``legacy/vescale/dtensor/sharding_prop.py:54-395`` and reference
``vescale/dtensor/_sharding_prop.py:36-251`` (RaggedShard-aware shape-argument adjustment).
Independent of torch's private DTensor internals by construction.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Any, Callable, Dict, List, Optional

import torch

from ..placement import Placement, RaggedShard, Replicate
from ..spec import DTensorSpec, TensorMeta
from .op_schema import OpSchema, OutputSharding, RuleResult

__all__ = ["ShardingPropagator", "register_rule", "get_rule", "propagator"]

_RULES: Dict[Any, Callable[[OpSchema], RuleResult]] = {}


class DynamicReplicate:
    """Output spec of ops whose output *shape* depends on the data (``unique``, ``nonzero``, ``masked_select`` ...): no meta
    kernel exists, so the op runs on replicated inputs and the spec is built from the actual local result (identical on every
    rank because the inputs are).  Legacy handles ``nonzero`` / ``_unique2`` with bypass handlers (``_dispatch_bypass.py``)."""

    def __init__(self, mesh):
        self.mesh = mesh

IN_META_PROPAGATION = [0]  # re-entrancy guard consulted by DModule's factory mode


def register_rule(ops, fn: Optional[Callable] = None):
    """``@register_rule(aten.mm.default)`` or ``register_rule([ops...], fn)``.  Accepts OpOverload,
    OpOverloadPacket (registers every overload) or lists thereof."""
    if not isinstance(ops, (list, tuple)):
        ops = [ops]

    def deco(f):
        for op in ops:
            if isinstance(op, torch._ops.OpOverloadPacket):
                for name in op.overloads():
                    _RULES[getattr(op, name)] = f
            else:
                _RULES[op] = f
        return f

    return deco(fn) if fn is not None else deco


def get_rule(op):
    return _RULES.get(op)


def _meta_of(spec: DTensorSpec) -> torch.Tensor:
    tm = spec.tensor_meta
    return torch.empty_strided(tm.shape, tm.stride, dtype=tm.dtype, device="meta")


def _to_meta(x):
    if isinstance(x, DTensorSpec):
        return _meta_of(x)
    if isinstance(x, (list, tuple)):
        return type(x)(_to_meta(i) for i in x)
    if isinstance(x, torch.device):
        return torch.device("meta")
    return x


def _collect_meta(out):
    if isinstance(out, torch.Tensor):
        return TensorMeta(tuple(out.shape), tuple(out.stride()), out.dtype)
    if isinstance(out, (list, tuple)):
        return tuple(_collect_meta(o) for o in out)
    return None


class ShardingPropagator:
    def __init__(self, cache_size: int = 16384):
        self._cache: "OrderedDict[OpSchema, OutputSharding]" = OrderedDict()
        self._cache_size = cache_size
        self.hits = 0
        self.misses = 0

    # ------------------------------------------------------------------ meta
    def propagate_tensor_meta(self, schema: OpSchema):
        args = tuple(_to_meta(a) for a in schema.args_schema)
        kwargs = {k: _to_meta(v) for k, v in schema.kwargs_schema.items()}
        if "device" in kwargs and kwargs["device"] is not None:
            kwargs["device"] = torch.device("meta")
        IN_META_PROPAGATION[0] += 1
        try:
            with torch.no_grad():
                out = schema.op(*args, **kwargs)
        except Exception as e:  # noqa: BLE001
            self.last_meta_error = f"{type(e).__name__}: {e}"
            return None
        finally:
            IN_META_PROPAGATION[0] -= 1
        return _collect_meta(out)

    # ------------------------------------------------------------------ propagate
    def propagate(self, schema: OpSchema) -> OutputSharding:
        try:
            hit = self._cache.get(schema)
        except TypeError:  # unhashable static arg
            return self._propagate(schema)
        if hit is not None:
            self.hits += 1
            self._cache.move_to_end(schema)
            return hit
        self.misses += 1
        out = self._propagate(schema)
        self._cache[schema] = out
        if len(self._cache) > self._cache_size:
            self._cache.popitem(last=False)
        return out

    def cache_info(self):
        return {"hits": self.hits, "misses": self.misses, "size": len(self._cache)}

    def _propagate(self, schema: OpSchema) -> OutputSharding:
        rule = _RULES.get(schema.op)
        in_specs = schema.tensor_specs()
        mesh = schema.mesh
        if rule is None:
            res = _replicate_fallback(schema)
        else:
            res = rule(schema)
        metas = self.propagate_tensor_meta(schema)

        def mk(pl, meta):
            if pl is None or meta is None:
                return None
            if isinstance(pl, DTensorSpec):
                return pl if pl.tensor_meta is not None else pl.with_meta(meta)
            return DTensorSpec(mesh, tuple(pl), meta)

        out = res.out
        if out is None:
            out_spec = None
        elif isinstance(out, DTensorSpec):
            out_spec = mk(out, metas if isinstance(metas, TensorMeta) else None)
        elif isinstance(out, (list, tuple)) and (len(out) == 0 or not isinstance(out[0], Placement)):
            # multi-output
            ms = metas if isinstance(metas, tuple) else (None,) * len(out)
            out_spec = tuple(mk(o, m if isinstance(m, TensorMeta) else None) for o, m in zip(out, ms))
        else:
            if isinstance(metas, tuple):  # op returns a list of tensors all sharing one placement
                out_spec = tuple(mk(out, m) for m in metas)
            else:
                if metas is None and all(isinstance(p, Replicate) for p in out) and res.ins is not None and all(
                    w is not None and all(isinstance(p, Replicate) for p in w) for w in res.ins
                ):
                    redis_dyn = [None if tuple(w) == h.placements else h.with_placements(tuple(w)) for w, h in zip(res.ins, in_specs)]
                    return OutputSharding(DynamicReplicate(mesh), redis_dyn if any(r is not None for r in redis_dyn) else None, res.local_args, res.local_kwargs, res.post, res.pre)
                if metas is None:
                    raise RuntimeError(
                        f"cannot infer output metadata of {schema.op} on the meta device ({getattr(self, 'last_meta_error', '?')}); "
                        f"inputs: {[(tuple(s.shape), s.placements) for s in in_specs]}"
                    )
                out_spec = mk(out, metas)

        redis = None
        if res.ins is not None:
            if len(res.ins) != len(in_specs):
                raise RuntimeError(f"rule for {schema.op} returned {len(res.ins)} input placements for {len(in_specs)} tensor inputs")
            redis_l: List[Optional[DTensorSpec]] = []
            need = False
            for want, have in zip(res.ins, in_specs):
                if want is None or tuple(want) == have.placements:
                    redis_l.append(None)
                else:
                    redis_l.append(have.with_placements(tuple(want)))
                    need = True
            redis = redis_l if need else None
        return OutputSharding(out_spec, redis, res.local_args, res.local_kwargs, res.post, res.pre)


def _replicate_fallback(schema: OpSchema) -> RuleResult:
    """No rule registered: run the op on fully replicated inputs (correct for any op, costs gathers).
    Ops with tensor outputs return Replicate; this mirrors legacy's "default to replicate" strategies."""
    if os.environ.get("VESCALE_STRICT_RULES", "0") == "1":
        raise NotImplementedError(f"no sharding rule registered for {schema.op}")
    in_specs = schema.tensor_specs()
    nd = schema.mesh.ndim
    rep = tuple(Replicate() for _ in range(nd))
    return RuleResult(out=rep, ins=[rep for _ in in_specs])


propagator = ShardingPropagator()
