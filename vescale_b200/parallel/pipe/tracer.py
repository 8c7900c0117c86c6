"""fx tracers for pipeline partitioning (legacy ``pipe/tracer.py``: ``ModelTracer``, ``HFModelTracer``, ``hf_symbolic_trace``,
``register_partition_module``, ``get_concrete_args``).

The partition *units* of a model (its transformer blocks, embeddings, heads — whatever the split method cuts between) must stay
opaque ``call_module`` nodes in the traced graph: a stage boundary can only fall between two nodes, and a unit whose inside was
inlined can no longer be kept together.  ``ModelTracer`` therefore treats as leaves (a) every module class registered with
``register_partition_module`` and (b) every module whose qualified name is in ``partition_units`` — and, like torch.fx, every
``torch.nn`` builtin.  ``HFModelTracer`` is the same policy on top of ``transformers.utils.fx.HFTracer`` (which knows how to
feed HuggingFace models meta inputs for their data-dependent control flow); ``hf_symbolic_trace`` picks dummy inputs from the
model's signature.  ``PipeParser`` uses these for ``TracerType.FX`` and falls back to the export tracers (``stage.py`` /
``trace.py``) for models fx cannot trace."""
from __future__ import annotations

import inspect
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Set, Type

import torch
import torch.fx as fx
import torch.nn as nn

__all__ = ["ModelTracer", "HFModelTracer", "hf_symbolic_trace", "register_partition_module", "unregister_partition_module", "registered_partition_modules",
           "get_concrete_args", "trace_model"]

_PARTITION_CLASSES: Set[Type[nn.Module]] = set()


def register_partition_module(cls: Type[nn.Module], *more: Type[nn.Module]) -> Type[nn.Module]:
    """Keep instances of ``cls`` un-inlined in traced graphs.  Usable as a class decorator."""
    for c in (cls,) + more:
        if not (isinstance(c, type) and issubclass(c, nn.Module)):
            raise TypeError(f"register_partition_module expects nn.Module classes, got {c!r}")
        _PARTITION_CLASSES.add(c)
    return cls


def unregister_partition_module(cls: Type[nn.Module]) -> None:
    _PARTITION_CLASSES.discard(cls)


def registered_partition_modules() -> Set[Type[nn.Module]]:
    return set(_PARTITION_CLASSES)


class _LeafPolicy:
    partition_units: Set[str]

    def _is_partition_leaf(self, m: nn.Module, qualname: str) -> bool:
        return qualname in self.partition_units or any(isinstance(m, c) for c in _PARTITION_CLASSES)


class ModelTracer(_LeafPolicy, fx.Tracer):
    """``ModelTracer(partition_units=["layers.0", "layers.1", ...]).trace(model)`` — units and registered classes stay leaves."""

    def __init__(self, partition_units: Optional[Iterable[str]] = None, **kw):
        super().__init__(**kw)
        self.partition_units = set(partition_units or ())

    def is_leaf_module(self, m: nn.Module, module_qualified_name: str) -> bool:
        return self._is_partition_leaf(m, module_qualified_name) or super().is_leaf_module(m, module_qualified_name)


def get_concrete_args(model: nn.Module, input_names: Sequence[str]) -> Dict[str, Any]:
    """Everything in ``model.forward``'s signature that is NOT a traced input, bound to its default — what fx needs to specialise
    optional arguments (``attention_mask=None``, ``use_cache=None``, ...) away instead of turning them into placeholders."""
    sig = inspect.signature(model.forward)
    unknown = [n for n in input_names if n not in sig.parameters]
    if unknown:
        raise ValueError(f"{type(model).__name__}.forward has no argument(s) {unknown}; it takes {list(sig.parameters)}")
    out = {}
    for p in sig.parameters.values():
        if p.name in input_names or p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
            continue
        out[p.name] = None if p.default is inspect.Parameter.empty else p.default
    return out


def _hf_tracer_base():
    try:
        from transformers.utils.fx import HFTracer

        return HFTracer
    except Exception:  # transformers without fx support: same leaf policy on the plain tracer
        return fx.Tracer


class HFModelTracer(_LeafPolicy, _hf_tracer_base()):  # type: ignore[misc]
    def __init__(self, partition_units: Optional[Iterable[str]] = None, **kw):
        super().__init__(**kw)
        self.partition_units = set(partition_units or ())

    def is_leaf_module(self, m: nn.Module, module_qualified_name: str) -> bool:
        return self._is_partition_leaf(m, module_qualified_name) or super().is_leaf_module(m, module_qualified_name)


def hf_symbolic_trace(model: nn.Module, input_names: Optional[Sequence[str]] = None, partition_units: Optional[Iterable[str]] = None, tracer_cls: Type = HFModelTracer) -> fx.GraphModule:
    """fx graph of a HuggingFace model with ``input_names`` (default: the model's ``dummy_inputs`` keys, else ``["input_ids"]``) as
    placeholders and every other forward argument fixed to its default."""
    if input_names is None:
        input_names = list(getattr(model, "dummy_inputs", {"input_ids": None}).keys())
    concrete = get_concrete_args(model, input_names)
    tracer = tracer_cls(partition_units=partition_units)
    try:
        graph = tracer.trace(model, concrete_args=concrete)
    except Exception as fx_error:  # noqa: BLE001
        # Recent transformers / Python releases break HF's fx path (it rewrites forward's code object); the export tracer captures
        # the same model at aten level.  Units cannot be kept opaque there — stage cutting on that graph is trace.py's job.
        gm = _export_graph_module(model, input_names)
        gm.fx_error = fx_error
        return _decorate(gm, model)
    return _decorate(fx.GraphModule(model, graph), model)


def _decorate(gm: fx.GraphModule, model: nn.Module) -> fx.GraphModule:
    gm.config = getattr(model, "config", None)
    gm.class_for_deserialization = type(model)
    gm.device = next((p.device for p in model.parameters()), torch.device("cpu"))
    return gm


def _export_graph_module(model: nn.Module, input_names: Sequence[str]) -> fx.GraphModule:
    dummy = dict(getattr(model, "dummy_inputs", {}) or {})
    dev = next((p.device for p in model.parameters()), torch.device("cpu"))
    kwargs = {}
    for n in input_names:
        v = dummy.get(n)
        kwargs[n] = v.to(dev) if isinstance(v, torch.Tensor) else torch.ones(1, 8, dtype=torch.long, device=dev)
    fixed = {"use_cache": False} if "use_cache" in inspect.signature(model.forward).parameters else {}  # a KV-cache output is no pytree

    class _Positional(nn.Module):  # the traced inputs become positional arguments, in ``input_names`` order
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, *args):
            return self.m(**dict(zip(input_names, args)), **fixed)

    return torch.export.export(_Positional(model), tuple(kwargs[n] for n in input_names), strict=False).module()


def trace_model(model: nn.Module, partition_units: Optional[Iterable[str]] = None, hf: Optional[bool] = None, input_names: Optional[Sequence[str]] = None) -> fx.GraphModule:
    """One entry point: HuggingFace models (anything with a ``config`` and ``dummy_inputs``) go through ``hf_symbolic_trace``,
    everything else through ``ModelTracer``."""
    if hf is None:
        hf = hasattr(model, "config") and hasattr(model, "dummy_inputs")
    if hf:
        return hf_symbolic_trace(model, input_names, partition_units)
    tracer = ModelTracer(partition_units)
    return fx.GraphModule(model, tracer.trace(model))
