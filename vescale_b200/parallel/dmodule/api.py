"""DModule: plan-driven tensor/sequence parallelism for arbitrary ``nn.Module``s (eager SPMD on DTensor).

    parallelize_module(model, mesh["TP"], {
        "parameter": {r"layers\\.\\d+\\.attn\\.wqkv\\.weight": [Shard(0)], r".*\\.wo\\.weight": [Shard(1)]},
        "forward":   {r"layers\\.\\d+\\.input": [[Shard(1)]], r"layers\\.\\d+\\.attn\\.input": [[Replicate()]],
                      r"layers\\.\\d+\\.attn\\.output": [[Shard(1)]]},
    })

* ``parameter`` plan: fqn-regex → placements; matching parameters become DTensor parameters (sharded from the
  replicated value, or re-interpreted as already-local shards with ``is_model_sharded``).  Unmatched
  parameters are replicated DTensors, so every op inside the module is a DTensor op.
* ``forward`` plan: ``<module fqn>.input`` / ``.output`` (+ ``.weight``-style names for parameters used in the
  forward) → one placement list per positional tensor.  Plain tensors are wrapped (``from_local``), DTensors
  are redistributed — this is where SP all-gathers / reduce-scatters are issued (Megatron SP = ``Shard(1)``
  activations between blocks).
* ``PlacementsInterface`` adds per-entry options (``async_op``, ``defer_reshard``, ``run_check``, ``grad``).
* gradients that come out ``Partial`` (norm weights under SP, replicated params fed by sharded activations)
  are collected and all-reduced in flat buckets by ``finish_grad_sync`` (the optimizer wrappers call it).
* ``factory=True``: ``torch.zeros/ones/empty/full/arange/randn`` called inside forward build DTensors.

Parity: ``legacy/vescale/dmodule/api.py:33-293``, ``_dmodule.py:43-666``, ``_hook.py:76-273``,
``_grad_sync.py:60-126``, ``_factory.py:57-117``, ``placements_interface.py``.
"""
from __future__ import annotations

import contextlib
import dataclasses
import functools
import inspect
import re
import warnings
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from ...comm import collectives as C
from ...dtensor.api import DTensor, distribute_tensor
from ...mesh import DeviceMesh
from ...placement import Partial, Placement, Replicate, Shard, normalize_placements

__all__ = ["parallelize_module", "is_dmodule", "PlacementsInterface", "DModule"]


@dataclass
class PlacementsInterface:
    placements: Optional[Sequence[Placement]]
    async_op: bool = True
    defer_reshard: bool = False
    run_check: bool = False
    support_uneven: bool = True
    grad: Optional[Sequence[Placement]] = None

    @classmethod
    def from_placements(cls, p) -> "PlacementsInterface":
        if isinstance(p, cls):
            return p
        return cls(None if p is None else list(p))


def _as_pi_list(entry) -> List[Optional[PlacementsInterface]]:
    if entry is None:
        return []
    if isinstance(entry, PlacementsInterface) or (entry and isinstance(entry[0], Placement)):
        entry = [entry]
    return [None if e is None else PlacementsInterface.from_placements(e) for e in entry]


def _convert(x, pi: Optional[PlacementsInterface], mesh: DeviceMesh, allow_defer: bool = False):
    if pi is None or pi.placements is None or not isinstance(x, torch.Tensor):
        return x
    pl = normalize_placements(pi.placements, mesh.ndim, x.ndim)
    if isinstance(x, DTensor):
        if x.placements == pl:
            return x
        if allow_defer and pi.defer_reshard:
            x._deferred_placements = tuple(pl)  # the sum / difference this output enters pays the reshard (dispatch.py)
            return x
        return x.redistribute(mesh, pl, async_op=pi.async_op)
    return DTensor.from_local(x, mesh, pl, run_check=pi.run_check)


def _convert_nested(x, spec, mesh: DeviceMesh):
    """``spec`` mirrors the structure of ``x``: a placement list / ``PlacementsInterface`` for a tensor, a dict for a dict
    argument, a list of placement lists for a list / tuple argument."""
    if spec is None:
        return x
    if isinstance(spec, dict):
        if not isinstance(x, dict):
            return x
        return type(x)({k: _convert_nested(v, spec[k], mesh) if k in spec else v for k, v in x.items()})
    if isinstance(spec, (list, tuple)) and spec and not isinstance(spec[0], Placement) and isinstance(x, (list, tuple)):
        return type(x)(_convert_nested(v, spec[i] if i < len(spec) else None, mesh) for i, v in enumerate(x))
    return _convert(x, PlacementsInterface.from_placements(spec), mesh)


class DModule:
    """Mixin state attached to a parallelized module (``module._dmodule``)."""

    def __init__(self, module: nn.Module, mesh: DeviceMesh, plan: Dict[str, Dict]):
        self.module = module
        self.mesh = mesh
        self.param_plan = {re.compile(k): v for k, v in (plan.get("parameter") or {}).items()}
        self.fwd_plan = {re.compile(k): v for k, v in (plan.get("forward") or {}).items()}
        self.handles = []
        self.factory = False

    # ------------------------------------------------------------------ parameters
    def init_parameters(self, is_model_sharded: bool = False) -> None:
        mesh = self.mesh
        for mod_name, mod in list(self.module.named_modules()):
            for pname, p in list(mod._parameters.items()):
                if p is None or isinstance(p.data, DTensor) or isinstance(p, DTensor):
                    continue
                fqn = f"{mod_name}.{pname}" if mod_name else pname
                pl = None
                for rx, v in self.param_plan.items():
                    if rx.fullmatch(fqn):
                        pl = v.placements if isinstance(v, PlacementsInterface) else v
                        break
                pl = normalize_placements(pl, mesh.ndim, p.ndim)
                if p.device.type == "meta":
                    from ...initialize import materialize_dparameter

                    dt = materialize_dparameter(p, mesh, pl)
                elif is_model_sharded:
                    dt = DTensor.from_local(p.data, mesh, pl)
                else:
                    dt = distribute_tensor(p.data, mesh, pl)
                mod._parameters[pname] = nn.Parameter(dt, requires_grad=p.requires_grad)
            for bname, b in list(mod._buffers.items()):
                if b is None or isinstance(b, DTensor):
                    continue
                fqn = f"{mod_name}.{bname}" if mod_name else bname
                pl = None
                for rx, v in self.param_plan.items():
                    if rx.fullmatch(fqn):
                        pl = v.placements if isinstance(v, PlacementsInterface) else v
                        break
                if b.device.type != "meta":
                    mod._buffers[bname] = distribute_tensor(b, mesh, normalize_placements(pl, mesh.ndim, b.ndim))

    # ------------------------------------------------------------------ forward hooks
    def init_forward(self) -> None:
        mesh = self.mesh
        names = dict(self.module.named_modules())
        for rx, entry in self.fwd_plan.items():
            pat = rx.pattern
            for kind in ("input", "output"):
                if pat == kind:
                    mod_rx = re.compile("")
                elif pat.endswith("\\." + kind):
                    mod_rx = re.compile(pat[: -len(kind) - 2])
                elif pat.endswith("." + kind):
                    mod_rx = re.compile(pat[: -len(kind) - 1])
                else:
                    continue
                for fqn, mod in names.items():
                    if not mod_rx.fullmatch(fqn):
                        continue
                    if kind == "input":
                        self.handles.append(mod.register_forward_pre_hook(self._make_pre(entry, mesh), with_kwargs=True))
                    else:
                        hint = self._reshard_hint(entry, mesh)
                        if hint is not None:
                            # the output plan reshards this module's result: a row-parallel matmul inside may produce the target
                            # layout directly (GEMM ⊕ reduce-scatter, dtensor/fusion.py) instead of Partial + reduce-scatter
                            self.handles.append(mod.register_forward_pre_hook(self._make_hint_push(hint, mesh)))
                        self.handles.append(mod.register_forward_hook(self._make_post(entry, mesh, pop_hint=hint is not None)))

    @staticmethod
    def _make_pre(entry, mesh):
        """Input hook.  The call is bound to ``forward``'s signature first (so a wrong call raises ``TypeError`` before any
        conversion and defaults are visible), then a sequence plan is laid over the bound arguments in order — positional
        ones, then keyword ones — and a dict plan is matched by parameter name (a ``*args`` parameter takes a list of
        placements, a ``**kwargs`` parameter is looked through, container arguments take a nested dict / list).  A plan
        naming more arguments than the call has warns and the surplus is ignored (legacy ``dmodule/_hook.py:96-170``)."""
        is_dict = isinstance(entry, dict)
        pis = None if is_dict else _as_pi_list(entry)

        def pre(mod, args, kwargs):
            sig = inspect.signature(mod.forward)
            bound = sig.bind(*args, **kwargs)
            bound.apply_defaults()
            if not is_dict:
                pos, kw = bound.args, bound.kwargs
                n = len(pos) + len(kw)
                if len(pis) > n:
                    warnings.warn(f"forward plan lists {len(pis)} placements but the call has {n} arguments; the rest are ignored")
                full = list(pis[:n]) + [None] * (n - len(pis))
                return (
                    tuple(_convert(x, pi, mesh) for x, pi in zip(pos, full)),
                    {k: _convert(v, pi, mesh) for (k, v), pi in zip(kw.items(), full[len(pos):])},
                )
            var_pos = next((q.name for q in sig.parameters.values() if q.kind is q.VAR_POSITIONAL), None)
            var_kw = next((q.name for q in sig.parameters.values() if q.kind is q.VAR_KEYWORD), None)
            known = set(bound.arguments) - {var_kw}
            if var_kw is not None:
                known |= set(bound.arguments.get(var_kw, {}))
            unknown = set(entry) - known
            if unknown:
                warnings.warn(f"forward plan names arguments the call does not have: {sorted(map(str, unknown))}")
            for name, val in list(bound.arguments.items()):
                if name == var_kw:
                    bound.arguments[name] = {k: _convert_nested(v, entry[k], mesh) if k in entry else v for k, v in val.items()}
                elif name not in entry:
                    continue
                elif name == var_pos:
                    sub = _as_pi_list(entry[name])
                    if len(sub) > len(val):
                        warnings.warn(f"forward plan lists {len(sub)} placements for *{name} but {len(val)} were passed; the rest are ignored")
                    bound.arguments[name] = tuple(_convert(v, sub[i] if i < len(sub) else None, mesh) for i, v in enumerate(val))
                else:
                    bound.arguments[name] = _convert_nested(val, entry[name], mesh)
            return bound.args, bound.kwargs

        return pre

    @staticmethod
    def _reshard_hint(entry, mesh):
        """Placements of the (single / first) tensor output if the plan shards it, else None."""
        if isinstance(entry, dict):
            return None
        pis = _as_pi_list(entry)
        if not pis or pis[0] is None or pis[0].placements is None:
            return None
        pl = list(pis[0].placements)
        return pl if any(isinstance(p, Shard) for p in pl) else None

    @staticmethod
    def _make_hint_push(placements, mesh):
        from ...dtensor.fusion import push_hint

        def push(mod, args):
            # a Shard(1) target on a (B, S, H) output is a contiguous row shard of the token matrix only when B == 1
            batch1 = all(a.shape[0] == 1 for a in args if isinstance(a, torch.Tensor) and a.ndim == 3)
            push_hint(id(mod), mesh, placements, rows_contiguous=batch1)

        return push

    @staticmethod
    def _make_post(entry, mesh, pop_hint: bool = False):
        """Output hook: a sequence plan over a tensor / tuple / list output, a dict plan (by key / field name) over a dict,
        dict-like (``ModelOutput``) or dataclass output (legacy ``dmodule/_hook.py:213-256``)."""
        is_dict = isinstance(entry, dict)
        pis = None if is_dict else _as_pi_list(entry)

        def post(mod, args, output):
            if pop_hint:
                from ...dtensor.fusion import pop_hint as _pop

                _pop(id(mod))
            if is_dict:
                if dataclasses.is_dataclass(output) and not isinstance(output, type) and not isinstance(output, dict):
                    vals = {f.name: getattr(output, f.name) for f in dataclasses.fields(output)}
                    return type(output)(**{k: _convert_nested(v, entry[k], mesh) if k in entry else v for k, v in vals.items()})
                if isinstance(output, dict):
                    conv = {k: _convert_nested(v, entry[k], mesh) if k in entry else v for k, v in output.items()}
                    try:
                        return type(output)(**conv)
                    except TypeError:
                        return type(output)(conv)
                raise TypeError(f"a dict output plan needs a dict or dataclass output, got {type(output).__name__}")
            if isinstance(output, (tuple, list)):
                if len(output) != len(pis):
                    raise AssertionError(f"output plan has {len(pis)} entries but the module returned {len(output)} values")
                conv = [_convert(o, pi, mesh, allow_defer=True) for o, pi in zip(output, pis)]
                return type(output)(*conv) if hasattr(output, "_fields") else type(output)(conv)
            if isinstance(output, dict) or dataclasses.is_dataclass(output):
                raise TypeError("a sequence output plan cannot be applied to a dict / dataclass output; key it by name")
            return _convert(output, pis[0] if pis else None, mesh, allow_defer=True)

        return post

    # ------------------------------------------------------------------ gradient sync
    def partial_grad_params(self) -> List[nn.Parameter]:
        out = []
        for p in self.module.parameters():
            g = p.grad
            if isinstance(g, DTensor) and any(pl.is_partial() for pl in g.placements):
                out.append(p)
        return out

    def finish_grad_sync(self, bucket_bytes: int = 40 * 2**20) -> int:
        """All-reduce every ``Partial`` gradient on its mesh dims, flattened into buckets per (mesh dim, dtype, reduce op)
        (``_grad_sync.sync_gradients``; legacy ``_grad_sync.py:60-126``: 40 MB flat buckets).  Returns the number of collectives issued."""
        from ._grad_sync import sync_gradients

        params = list(self.partial_grad_params())
        reduced, n_coll = sync_gradients([p.grad for p in params], self.mesh, bucket_bytes)
        for p, g in zip(params, reduced):
            p.grad = g
        return n_coll


def is_dmodule(module: nn.Module) -> bool:
    return hasattr(module, "_dmodule")


def parallelize_module(
    module: nn.Module,
    device_mesh: DeviceMesh,
    sharding_plan: Optional[Dict[str, Dict]] = None,
    *,
    is_model_sharded: bool = False,
    factory: Union[bool, Dict] = False,
) -> nn.Module:
    plan = sharding_plan or {}
    dm = DModule(module, device_mesh, plan)
    module._dmodule = dm
    dm.init_parameters(is_model_sharded)
    dm.init_forward()
    if factory:
        from ._factory import wrap_factory_mode

        dm.factory = wrap_factory_mode(module, device_mesh, factory) > 0
    module.finish_grad_sync = dm.finish_grad_sync
    module.list_partial_grads = dm.partial_grad_params
    module.get_fqn = lambda: ""
    return module
