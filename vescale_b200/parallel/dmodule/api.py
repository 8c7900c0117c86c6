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
from ._hook import (  # noqa: F401
    PlacementsInterface, PostHookGrad, PostHookOutput, PostHookWeight, PreHookInput, PreHookWeight, _as_pi_list, _convert, _convert_nested,
)

__all__ = ["parallelize_module", "is_dmodule", "PlacementsInterface", "DModule"]


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
                        self.handles.append(mod.register_forward_pre_hook(PreHookInput.get_hook(mesh, entry), with_kwargs=True))
                    else:
                        hint = self._reshard_hint(entry, mesh)
                        if hint is not None:
                            # the output plan reshards this module's result: a row-parallel matmul inside may produce the target
                            # layout directly (GEMM ⊕ reduce-scatter, dtensor/fusion.py) instead of Partial + reduce-scatter
                            self.handles.append(mod.register_forward_pre_hook(PostHookOutput.get_hint_push(mesh, hint)))
                        self.handles.append(mod.register_forward_hook(PostHookOutput.get_hook(mesh, entry, pop_hint=hint is not None)))
        self._init_weight_plans(names)

    def _init_weight_plans(self, names: Dict[str, nn.Module]) -> None:
        """Forward-plan keys that name a PARAMETER (``r"fc1\\.weight": [Replicate()]`` or ``PlacementsInterface(..., grad=[...])``;
        legacy ``_dmodule.py:330-348``): the owning module computes with the parameter in that layout (``PreHookWeight`` /
        ``PostHookWeight`` swap a differentiable redistributed view in and out), and ``grad=`` re-labels the gradient that
        reaches the parameter (``PostHookGrad``)."""
        mesh = self.mesh
        per_module: Dict[str, Dict[str, PlacementsInterface]] = {}
        for rx, entry in self.fwd_plan.items():
            pat = rx.pattern
            if pat in ("input", "output") or pat.endswith((".input", ".output")):
                continue
            for mod_name, mod in names.items():
                for pname, p in mod._parameters.items():
                    if p is None:
                        continue
                    fqn = f"{mod_name}.{pname}" if mod_name else pname
                    if rx.fullmatch(fqn):
                        pi = entry if isinstance(entry, PlacementsInterface) else PlacementsInterface.from_placements(entry)
                        per_module.setdefault(mod_name, {})[pname] = pi
        for mod_name, pis in per_module.items():
            mod = names[mod_name]
            if any(pi.placements for pi in pis.values()):
                self.handles.append(mod.register_forward_pre_hook(PreHookWeight.get_hook(mesh, pis)))
                self.handles.append(mod.register_forward_hook(PostHookWeight.get_hook(mesh, pis), always_call=True))
            for pname, pi in pis.items():
                if pi.grad:
                    self.handles.append(mod._parameters[pname].register_hook(PostHookGrad.get_hook(mesh, pi.grad)))

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
