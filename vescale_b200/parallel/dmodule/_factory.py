"""DTensor factories inside ``forward``: while a factory region is ON, ``torch.zeros / ones / empty / full / randn / rand /
arange`` build DTensors of the given GLOBAL shape on the region's mesh with the placements configured per factory (Replicate when
not configured); an OFF region restores plain tensors.  Regions nest freely and are independent of any ``TorchDispatchMode``
active in between (legacy ``dmodule/_factory.py:57-117``, which intercepts the aten ops in a dispatch mode; here a
``TorchFunctionMode`` sees the Python-level call with its original arguments and consults a per-thread region stack, so nothing sits
on the dispatcher's hot path and nothing at all is installed when no region is open).

``parallelize_module(..., factory=...)`` accepts ``True`` (root module ON, everything replicated), or a dict keyed by module
CLASS whose values are ``True`` / ``False`` / ``{torch.zeros: [Shard(0)], ...}``; every submodule of a listed class opens the
matching region around its ``forward``."""
from __future__ import annotations

import functools
from typing import Any, Dict, Optional, Union

import torch
import torch.nn as nn
from torch.overrides import TorchFunctionMode

from ...dtensor import api as _dapi
from ...mesh import DeviceMesh
from ...placement import Replicate

__all__ = ["FactoryDispatchModeOn", "FactoryDispatchModeOff", "FACTORY_NAMES", "wrap_factory_mode"]

FACTORY_NAMES = ("zeros", "ones", "empty", "full", "randn", "rand", "arange")
_DFACTORY = {n: getattr(_dapi, n) for n in FACTORY_NAMES}
_BY_FUNC = {getattr(torch, n): n for n in FACTORY_NAMES}


_stack = _dapi._factory_region_stack


def _factory_name(key) -> str:
    """``torch.zeros`` / ``"zeros"`` / ``aten.zeros.default`` → ``"zeros"``."""
    name = key if isinstance(key, str) else (getattr(key, "__name__", None) or str(key))
    name = name.replace("aten::", "").replace("aten.", "").split(".")[0]
    if name not in FACTORY_NAMES:
        raise ValueError(f"{key!r} is not one of the supported factories {FACTORY_NAMES}")
    return name


def _provide_args(device_mesh: DeviceMesh, factory_pis: Optional[Dict[Any, Any]]) -> Dict[str, tuple]:
    """Normalise ``{factory: placements | PlacementsInterface}`` to ``{name: placements tuple}``."""
    out: Dict[str, tuple] = {}
    for k, v in (factory_pis or {}).items():
        pl = getattr(v, "placements", v)
        out[_factory_name(k)] = tuple(pl) if pl is not None else tuple(Replicate() for _ in range(device_mesh.ndim))
    return out


def _build(name: str, region, args, kwargs):
    """The DTensor twin of one intercepted factory call (``size`` is the GLOBAL shape)."""
    mesh, pis = region
    kw = {k: v for k, v in kwargs.items() if k in ("dtype", "layout", "requires_grad")}
    if name == "full" and "fill_value" in kwargs:
        args = (*args, kwargs["fill_value"])
    if name in ("zeros", "ones", "empty", "randn", "rand") and "size" in kwargs:
        args = (kwargs["size"],)
    return _DFACTORY[name](*args, device_mesh=mesh, placements=pis.get(name), **kw)


class _FactoryFunctionMode(TorchFunctionMode):
    """Sees every ``torch.*`` call made while an ON region is open — also through references to the builtins taken earlier
    (``f = torch.zeros; ...; f(shape)``) — and reroutes the factories according to the innermost region."""

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = _BY_FUNC.get(func)
        if name is None:
            return func(*args, **kwargs)
        st = _stack()
        region = st[-1] if st else None
        if region is None:
            return func(*args, **kwargs)
        from ...dtensor import sharding_prop as _sp

        dev = kwargs.get("device")
        if _sp.IN_META_PROPAGATION[0] or (dev is not None and torch.device(dev).type == "meta") or kwargs.get("out") is not None:
            return func(*args, **kwargs)
        return _build(name, region, args, kwargs)


class FactoryDispatchModeOn:
    """``with FactoryDispatchModeOn(mesh, {factory: placements}):`` — factories build DTensors on ``mesh``."""

    def __init__(self, device_mesh: DeviceMesh, aten_dfactory_pi: Optional[Dict[Any, Any]] = None):
        self.mesh = device_mesh
        self.pis = _provide_args(device_mesh, aten_dfactory_pi)
        self._mode = _FactoryFunctionMode()

    def __enter__(self):
        _stack().append((self.mesh, self.pis))
        self._mode.__enter__()
        return self

    def __exit__(self, *exc):
        self._mode.__exit__(*exc)
        _stack().pop()
        return False


class FactoryDispatchModeOff:
    """``with FactoryDispatchModeOff():`` — plain tensors again, whatever region encloses this one."""

    def __enter__(self):
        _stack().append(None)
        return self

    def __exit__(self, *exc):
        _stack().pop()
        return False


def _region_for(setting, mesh: DeviceMesh):
    if setting is False or setting is None:
        return FactoryDispatchModeOff
    pis = {} if setting is True else _provide_args(mesh, setting)
    return lambda: FactoryDispatchModeOn(mesh, pis)


def _wrapped_forward(inner, region):
    @functools.wraps(inner)
    def forward(*args, **kwargs):
        with region():
            return inner(*args, **kwargs)

    return forward


def wrap_factory_mode(root: nn.Module, mesh: DeviceMesh, factory: Union[bool, Dict[type, Any]]) -> int:
    """Open the configured region around ``forward`` of the root (``factory=True``) or of every submodule whose class is a key of
    ``factory``.  Returns the number of modules wrapped."""
    if not factory:
        return 0
    table = {type(root): True} if factory is True else dict(factory)
    n = 0
    for mod in root.modules():
        if type(mod) not in table or getattr(mod, "_vb_factory_wrapped", False):
            continue
        region = _region_for(table[type(mod)], mesh)
        mod.forward = _wrapped_forward(mod.forward, region)
        mod._vb_factory_wrapped = True
        n += 1
    return n
