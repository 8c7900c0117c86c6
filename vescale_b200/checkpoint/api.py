from __future__ import annotations

import atexit
import os
import threading
from concurrent.futures import Future
from typing import Any, Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from ..dtensor.api import DTensor
from ..spec import DTensorSpec
from .pinned_pool import PinnedPool

__all__ = ["save", "load", "VeScaleCheckpointer", "wait_for_async"]

_POOL = PinnedPool()
_PENDING: List[Future] = []
_ASYNC_PG = {"pg": None}


def _flatten(prefix: str, obj, out: Dict[str, Any]) -> None:
    if isinstance(obj, dict):
        for k, v in obj.items():
            _flatten(f"{prefix}.{k}" if prefix else str(k), v, out)
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            _flatten(f"{prefix}.{i}", v, out)
    else:
        out[prefix] = obj


def _state_of(obj) -> Dict[str, Any]:
    if isinstance(obj, nn.Module):
        return dict(obj.state_dict())
    if hasattr(obj, "state_dict"):
        flat: Dict[str, Any] = {}
        _flatten("", obj.state_dict(), flat)
        return flat
    if isinstance(obj, dict):
        flat = {}
        _flatten("", obj, flat)
        return flat
    raise TypeError(f"cannot checkpoint {type(obj)}")


def _stage_to_host(sd: Dict[str, Any]) -> Dict[str, Any]:
    """Device shards → pinned host copies (same DTensor specs), so the file write can run in the background."""
    out = {}
    for k, v in sd.items():
        if isinstance(v, DTensor):
            out[k] = DTensor(_POOL.stage(v._local_tensor), v._spec)
        elif isinstance(v, torch.Tensor):
            out[k] = _POOL.stage(v)
        else:
            out[k] = v
    _POOL.synchronize()
    return out


def _release(sd: Dict[str, Any]) -> None:
    for v in sd.values():
        t = v._local_tensor if isinstance(v, DTensor) else v
        if isinstance(t, torch.Tensor) and hasattr(t, "_pool_base"):
            _POOL.release(t)


def _async_group():
    if _ASYNC_PG["pg"] is None and dist.is_initialized():
        _ASYNC_PG["pg"] = dist.new_group(backend="gloo")
    return _ASYNC_PG["pg"]


class VeScaleCheckpointer:
    @classmethod
    def save(cls, path: str, checkpoint_state: Dict[str, Any], async_checkpoint: bool = False) -> Optional[List[Future]]:
        import torch.distributed.checkpoint as dcp

        futures = []
        for key, obj in checkpoint_state.items():
            sub = os.path.join(path, key)
            if not dist.is_initialized() or dist.get_rank() == 0:
                os.makedirs(sub, exist_ok=True)
            sd = _state_of(obj)
            tensors = {k: v for k, v in sd.items() if isinstance(v, torch.Tensor)}
            extras = {k: v for k, v in sd.items() if not isinstance(v, torch.Tensor)}
            if extras:
                tensors["__extras__"] = extras  # small python state (step counters, hyper-parameters) via DCP bytes
            if async_checkpoint:
                host = _stage_to_host(tensors)
                pg = _async_group()
                fut: Future = Future()

                def work(host=host, sub=sub, pg=pg, fut=fut):
                    try:
                        dcp.save(host, checkpoint_id=sub, process_group=pg)
                        fut.set_result(sub)
                    except Exception as e:  # noqa: BLE001
                        fut.set_exception(e)
                    finally:
                        _release(host)

                wait_for_async()  # one save in flight at a time: collectives of two saves must not interleave
                t = threading.Thread(target=work, daemon=True)
                t.start()
                fut._thread = t
                _PENDING.append(fut)
                futures.append(fut)
            else:
                dcp.save(tensors, checkpoint_id=sub)
        return futures or None

    @classmethod
    def load(cls, path: str, checkpoint_state: Dict[str, Any], broadcast_checkpoint: bool = False) -> None:
        import torch.distributed.checkpoint as dcp

        for key, obj in checkpoint_state.items():
            sub = os.path.join(path, key)
            sd = _state_of(obj)
            tensors = {k: v for k, v in sd.items() if isinstance(v, torch.Tensor)}
            extras_keys = [k for k, v in sd.items() if not isinstance(v, torch.Tensor)]
            req = dict(tensors)
            if extras_keys:
                req["__extras__"] = {k: sd[k] for k in extras_keys}
            dcp.load(req, checkpoint_id=sub)  # in place: DTensor/ tensor storages are filled with the resharded data
            if isinstance(obj, nn.Module):
                pass  # state_dict tensors alias the module's parameters/buffers
            elif hasattr(obj, "load_state_dict") and not isinstance(obj, dict):
                full = obj.state_dict()
                _assign(full, req)
                obj.load_state_dict(full)


def _assign(nested, flat: Dict[str, Any], prefix: str = "") -> None:
    extras = flat.get("__extras__", {})
    if isinstance(nested, dict):
        for k in list(nested.keys()):
            key = f"{prefix}.{k}" if prefix else str(k)
            v = nested[k]
            if isinstance(v, (dict, list, tuple)):
                _assign(v, flat, key)
            elif key in flat and isinstance(flat[key], torch.Tensor):
                nested[k] = flat[key]
            elif key in extras:
                nested[k] = extras[key]
    elif isinstance(nested, list):
        for i in range(len(nested)):
            key = f"{prefix}.{i}"
            if isinstance(nested[i], (dict, list, tuple)):
                _assign(nested[i], flat, key)
            elif key in flat:
                nested[i] = flat[key]


def wait_for_async() -> None:
    while _PENDING:
        f = _PENDING.pop(0)
        f.result()
        t = getattr(f, "_thread", None)
        if t is not None:
            t.join()


atexit.register(lambda: [f.result() for f in list(_PENDING) if not f.done()] if _PENDING else None)


def save(path: str, checkpoint_state: Dict[str, Any], async_checkpoint: bool = False):
    return VeScaleCheckpointer.save(path, checkpoint_state, async_checkpoint)


def load(path: str, checkpoint_state: Dict[str, Any], broadcast_checkpoint: bool = False):
    return VeScaleCheckpointer.load(path, checkpoint_state, broadcast_checkpoint)
