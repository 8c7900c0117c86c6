"""ndtimeline regions for torch's own FSDP2 (``torch.distributed.fsdp.fully_shard``).

This framework's FSDP engine emits ``UNSHARD_AG`` / ``GRAD_RS`` itself (``parallel/fsdp/api.py``).  A job that still runs part of
its model under torch's ``fully_shard`` gets the same two rows in the same timeline after :func:`patch_fsdp`: the parameter
group's ``unshard`` (all-gather issue, on its all-gather stream) and ``post_backward`` (reduce-scatter issue, on its
reduce-scatter stream) are wrapped in timed regions tagged with the group's module name.  The public reference ships this entry
point as a stub (``legacy/vescale/ndtimeline/fsdp_patch.py:24-28``: ``patch_fsdp`` does nothing, ``is_fsdp_patched`` is False).

Idempotent; :func:`unpatch_fsdp` restores the originals.  A torch whose FSDP2 internals moved raises ``RuntimeError`` rather than
patching nothing silently.
"""
from __future__ import annotations

import functools
from typing import Callable, Dict

import torch

from . import predefined
from .timer import ndtimeit

__all__ = ["patch_fsdp", "unpatch_fsdp", "is_fsdp_patched"]

_ORIGINALS: Dict[str, Callable] = {}
_TARGETS = {"unshard": (predefined.UNSHARD_AG, "all_gather_stream"), "post_backward": (predefined.GRAD_RS, "reduce_scatter_stream")}


def _param_group_cls():
    try:
        from torch.distributed.fsdp._fully_shard._fsdp_param_group import FSDPParamGroup
    except ImportError:  # torch < 2.6 kept it under _composable
        try:
            from torch.distributed._composable.fsdp._fsdp_param_group import FSDPParamGroup
        except ImportError as e:
            raise RuntimeError("this torch has no FSDP2 parameter group to patch") from e
    return FSDPParamGroup


def _timed(orig: Callable, metric: str, stream_attr: str) -> Callable:
    @functools.wraps(orig)
    def wrapper(self, *args, **kwargs):
        stream = None
        if torch.cuda.is_available():
            stream = getattr(getattr(self, "comm_ctx", None), stream_attr, None)
        unit = getattr(self, "_module_fqn", None) or type(getattr(self, "modules", [self])[0]).__name__
        with ndtimeit(metric, stream=stream, unit=str(unit)):
            return orig(self, *args, **kwargs)

    wrapper._ndtimeline_patched = True
    return wrapper


def patch_fsdp() -> None:
    cls = _param_group_cls()
    missing = [n for n in _TARGETS if not callable(getattr(cls, n, None))]
    if missing:
        raise RuntimeError(f"FSDPParamGroup has no {missing}: torch's FSDP2 internals changed, nothing was patched")
    for name, (metric, stream_attr) in _TARGETS.items():
        cur = getattr(cls, name)
        if getattr(cur, "_ndtimeline_patched", False):
            continue
        _ORIGINALS[name] = cur
        setattr(cls, name, _timed(cur, metric, stream_attr))


def unpatch_fsdp() -> None:
    cls = _param_group_cls()
    for name, orig in list(_ORIGINALS.items()):
        setattr(cls, name, orig)
        del _ORIGINALS[name]


def is_fsdp_patched() -> bool:
    cls = _param_group_cls()
    return all(getattr(getattr(cls, n, None), "_ndtimeline_patched", False) for n in _TARGETS)
