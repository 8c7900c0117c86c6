"""The hooks a ``DModule`` installs (parity: ``legacy/vescale/dmodule/_hook.py:72-273``).

* :class:`PreHookInput` / :class:`PostHookOutput` — lay a forward plan over a module's arguments / results (tensors become
  DTensors, DTensors are redistributed; the output hook can defer the reshard into the consumer, and tells a row-parallel matmul
  inside the module which layout its result is about to be given so GEMM (+) reduce-scatter runs as one kernel).
* :class:`PreHookWeight` / :class:`PostHookWeight` — a forward plan on a PARAMETER (``"fc1.weight": [Replicate()]``): the module
  computes with the parameter in that layout.  The reference re-registers a new ``nn.Parameter`` and leaves the "restore" hook
  unimplemented (``_hook.py:178-210``), which detaches the optimizer's parameter; here the owning ``nn.Parameter`` never moves: for
  the duration of ``forward`` the module's slot holds a differentiable redistributed view of it (backward of the all-gather is the
  reduce-scatter into the parameter's own layout), and the post hook puts the parameter back.
* :class:`PostHookGrad` — ``PlacementsInterface(grad=[...])`` on a weight plan: the gradient that reaches the parameter is
  re-labelled with those placements (no communication: the caller states what the local values already are, e.g. ``Replicate``
  for a gradient a custom kernel has already reduced).
"""
from __future__ import annotations

import dataclasses
import inspect
import warnings
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from ...dtensor.api import DTensor
from ...mesh import DeviceMesh
from ...placement import Placement, Shard, normalize_placements
from ...spec import DTensorSpec

__all__ = ["PlacementsInterface", "PreHookInput", "PreHookWeight", "PostHookWeight", "PostHookOutput", "PostHookGrad", "get_sig"]


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


def get_sig(module: nn.Module) -> inspect.Signature:
    """Signature of the module's ``forward`` (what input plans are bound against)."""
    return inspect.signature(module.forward)


class PreHookInput:
    """Input hook.  The call is bound to ``forward``'s signature first (so a wrong call raises ``TypeError`` before any
    conversion and defaults are visible), then a sequence plan is laid over the bound arguments in order — positional
    ones, then keyword ones — and a dict plan is matched by parameter name (a ``*args`` parameter takes a list of
    placements, a ``**kwargs`` parameter is looked through, container arguments take a nested dict / list).  A plan
    naming more arguments than the call has warns and the surplus is ignored (legacy ``dmodule/_hook.py:96-170``)."""

    @staticmethod
    def get_hook(device_mesh, input_pis):
        entry, mesh = input_pis, device_mesh
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

class PostHookOutput:
    @staticmethod
    def get_hint_push(device_mesh, placements):
        mesh = device_mesh
        from ...dtensor.fusion import push_hint

        def push(mod, args):
            # a Shard(1) target on a (B, S, H) output is a contiguous row shard of the token matrix only when B == 1
            batch1 = all(a.shape[0] == 1 for a in args if isinstance(a, torch.Tensor) and a.ndim == 3)
            push_hint(id(mod), mesh, placements, rows_contiguous=batch1)

        return push

    @staticmethod
    def get_hook(device_mesh, output_pis, pop_hint: bool = False):
        """Output hook: a sequence plan over a tensor / tuple / list output, a dict plan (by key / field name) over a dict,
        dict-like (``ModelOutput``) or dataclass output (legacy ``dmodule/_hook.py:213-256``)."""
        entry, mesh = output_pis, device_mesh
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

_SWAPPED = "_vescale_weight_plan_saved"


class PreHookWeight:
    """``weight_pis``: parameter name (relative to the hooked module, dotted for nested ones) -> ``PlacementsInterface``."""

    @staticmethod
    def _hook(module: nn.Module, input: Any, device_mesh: DeviceMesh, weight_pis: Dict[str, Optional[PlacementsInterface]]):
        saved = module.__dict__.setdefault(_SWAPPED, [])
        if saved:  # re-entrant forward (activation checkpointing recompute inside forward): already swapped
            return
        for fqn, pi in weight_pis.items():
            if pi is None or not pi.placements:
                continue
            path, _, name = fqn.rpartition(".")
            owner = module.get_submodule(path)
            param = owner._parameters.get(name)
            if param is None:
                raise AttributeError(f"forward plan names parameter '{fqn}', which {type(module).__name__} does not have")
            if not isinstance(param, DTensor) and not isinstance(param.data, DTensor):
                raise RuntimeError(f"forward plan on '{fqn}': only a DTensor parameter can be redistributed in forward")
            want = normalize_placements(pi.placements, device_mesh.ndim, param.ndim)
            if tuple(param.placements) == tuple(want):
                continue
            view = param.redistribute(device_mesh, want, async_op=pi.async_op)  # differentiable: grads land in param's layout
            saved.append((owner, name, param))
            owner._parameters[name] = view  # written to the dict: register_parameter would insist on a leaf nn.Parameter

    @staticmethod
    def get_hook(device_mesh: DeviceMesh, weight_pis):
        return lambda module, input: PreHookWeight._hook(module, input, device_mesh, weight_pis)


class PostHookWeight:
    @staticmethod
    def _hook(module: nn.Module, input: Any, output: Any, device_mesh: DeviceMesh, weight_pis=None):
        for owner, name, param in module.__dict__.get(_SWAPPED, ()):
            owner._parameters[name] = param
        module.__dict__[_SWAPPED] = []
        return None

    @staticmethod
    def get_hook(device_mesh: DeviceMesh, weight_pis=None):
        return lambda module, input, output: PostHookWeight._hook(module, input, output, device_mesh, weight_pis)


class PostHookGrad:
    @staticmethod
    def _hook(grad: Any, device_mesh: DeviceMesh, grad_placements: Optional[Sequence[Placement]]):
        if not grad_placements or grad is None:
            return grad
        if not isinstance(grad, DTensor):
            raise ValueError("a gradient-placement hook belongs on a DTensor parameter with a DTensor gradient")
        want = tuple(normalize_placements(grad_placements, device_mesh.ndim, grad.ndim))
        if want == tuple(grad.placements):
            return grad
        sp = grad._spec
        from ...layout import compute_local_shape

        if tuple(compute_local_shape(tuple(sp.shape), sp.mesh, want)) != tuple(grad._local_tensor.shape):
            raise ValueError(
                f"gradient placements {want} do not describe a local gradient of shape {tuple(grad._local_tensor.shape)} "
                f"(global {tuple(sp.shape)}, currently {tuple(grad.placements)}): a re-label never moves data"
            )
        return DTensor(grad._local_tensor, DTensorSpec(sp.mesh, want, sp.tensor_meta), requires_grad=grad.requires_grad)

    @staticmethod
    def get_hook(device_mesh: DeviceMesh, grad_placements: Optional[Sequence[Placement]]):
        return lambda grad: PostHookGrad._hook(grad, device_mesh, grad_placements)
