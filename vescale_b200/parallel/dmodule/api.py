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
import re
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


def _convert(x, pi: Optional[PlacementsInterface], mesh: DeviceMesh):
    if pi is None or pi.placements is None or not isinstance(x, torch.Tensor):
        return x
    pl = normalize_placements(pi.placements, mesh.ndim, x.ndim)
    if isinstance(x, DTensor):
        if x.placements == pl:
            return x
        return x.redistribute(mesh, pl, async_op=pi.async_op)
    return DTensor.from_local(x, mesh, pl, run_check=pi.run_check)


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
                        self.handles.append(mod.register_forward_hook(self._make_post(entry, mesh)))

    @staticmethod
    def _make_pre(entry, mesh):
        if isinstance(entry, dict):
            kw_pis = {k: PlacementsInterface.from_placements(v) for k, v in entry.items()}

            def pre_kw(mod, args, kwargs):
                return args, {k: _convert(v, kw_pis.get(k), mesh) for k, v in kwargs.items()}

            return pre_kw
        pis = _as_pi_list(entry)

        def pre(mod, args, kwargs):
            new = tuple(_convert(a, pis[i] if i < len(pis) else None, mesh) for i, a in enumerate(args))
            return new, kwargs

        return pre

    @staticmethod
    def _make_post(entry, mesh):
        pis = _as_pi_list(entry)

        def post(mod, args, output):
            if isinstance(output, (tuple, list)):
                return type(output)(_convert(o, pis[i] if i < len(pis) else None, mesh) for i, o in enumerate(output))
            return _convert(output, pis[0] if pis else None, mesh)

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
        """All-reduce every ``Partial`` gradient on its mesh dims, flattened into buckets per (mesh dim, dtype)
        (legacy ``_grad_sync.py:60-126``: 40 MB flat buckets).  Returns the number of collectives issued."""
        groups: Dict[Tuple[int, torch.dtype, str], List[Tuple[nn.Parameter, torch.Tensor]]] = {}
        for p in self.partial_grad_params():
            g: DTensor = p.grad
            for i, pl in enumerate(g.placements):
                if pl.is_partial():
                    groups.setdefault((i, g.dtype, pl.reduce_op), []).append((p, g._local_tensor))
        n_coll = 0
        for (md, dtype, op), items in groups.items():
            bucket, size = [], 0
            def flush():
                nonlocal n_coll, bucket, size
                if not bucket:
                    return
                flat = torch.cat([t.reshape(-1) for _, t in bucket])
                red = C.mesh_all_reduce(flat, self.mesh, op, md, inplace=True)
                off = 0
                for _, t in bucket:
                    t.copy_(red[off : off + t.numel()].view_as(t))
                    off += t.numel()
                n_coll += 1
                bucket, size = [], 0
            for p, t in items:
                bucket.append((p, t))
                size += t.numel() * t.element_size()
                if size >= bucket_bytes:
                    flush()
            flush()
            for p, _ in items:
                g = p.grad
                pl = tuple(Replicate() if (i == md and q.is_partial()) else q for i, q in enumerate(g.placements))
                p.grad = DTensor(g._local_tensor, g._spec.with_placements(pl))
        return n_coll


def is_dmodule(module: nn.Module) -> bool:
    return hasattr(module, "_dmodule")


_FACTORY_FNS = ("zeros", "ones", "empty", "full", "randn", "arange")


@contextlib.contextmanager
def _factory_mode(mesh: DeviceMesh):
    """Inside forward, plain factory calls build replicated DTensors (legacy ``_factory.py:57-117``)."""
    saved = {n: getattr(torch, n) for n in _FACTORY_FNS}

    from ...dtensor import sharding_prop as _sp

    def wrap(fn):
        def f(*a, **kw):
            t = fn(*a, **kw)
            # never inside the dispatcher's own meta-shape inference, never for meta tensors
            if _sp.IN_META_PROPAGATION[0] or not isinstance(t, torch.Tensor) or isinstance(t, DTensor) or t.is_meta:
                return t
            return DTensor.from_local(t, mesh, [Replicate()] * mesh.ndim)

        return f

    try:
        for n, fn in saved.items():
            setattr(torch, n, wrap(fn))
        yield
    finally:
        for n, fn in saved.items():
            setattr(torch, n, fn)


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
        orig_forward = module.forward

        def fwd(*a, **kw):
            with _factory_mode(device_mesh):
                return orig_forward(*a, **kw)

        module.forward = fwd
        dm.factory = True
    module.finish_grad_sync = dm.finish_grad_sync
    module.list_partial_grads = dm.partial_grad_params
    module.get_fqn = lambda: ""
    return module
