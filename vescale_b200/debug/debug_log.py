"""DebugLogger: one switch that makes a distributed run explain itself.

With the mode on, every mesh collective is logged with the user-code line that caused it (the *injection point*: the innermost
frame that is neither this framework nor torch) and every dispatched DTensor op is logged with the ``nn.Module`` whose
``forward`` (or the autograd backward) it ran under, its operand specs, the output spec and the redistributions the sharding
rule asked for.  Lines are also aggregated, so ``DebugLogger.summary()`` answers "which module line costs how many bytes of
which collective" after a step.

Switch on with ``VESCALE_DEBUG_MODE=1`` (read at import and again by ``update_vescale_debug_mode_from_env()``) or
``set_vescale_debug_mode(True, rank_to_print=(0,), logger=...)``.  ``rank_to_print`` takes an int, a sequence, or ``-1`` for
every rank; ``None`` means rank 0.

The hooks live in the two choke points of this framework, ``comm.collectives._note`` (every mesh collective, whichever back
end runs it) and ``dtensor.dispatch.dispatcher._hooks`` (every op after sharding propagation), so nothing is patched and the
cost with the mode off is one empty-list check.

Parity: ``legacy/vescale/debug/debug_log.py:40-361`` (``DebugLogger``, ``_CommunicationLogger._trace_to_coll_inject_point``,
``log_communication_decorator``, ``_OperatorLogger.trace_to_forward / ops_info_printer / dt_spec_debug_formatter``).
"""
from __future__ import annotations

import functools
import logging
import os
import sys
from collections import OrderedDict
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist

from ..comm import collectives as C

__all__ = ["DebugLogger", "set_vescale_debug_mode", "update_vescale_debug_mode_from_env"]

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_TORCH_DIR = os.path.dirname(os.path.abspath(torch.__file__))


def _is_user_file(filename: str) -> bool:
    return not (filename.startswith(_PKG_DIR) or filename.startswith(_TORCH_DIR) or filename.startswith("<"))


class DebugLogger:
    """Central, rank-filtered logging of collectives and DTensor ops (legacy ``debug/debug_log.py:40``)."""

    IS_DEBUG_MODE = False
    enabled = False  # alias of IS_DEBUG_MODE kept in step by set_vescale_debug_mode
    ranks: Optional[Sequence[int]] = None
    logger: Optional[logging.Logger] = None
    records: List[str] = []
    max_records = 100_000
    _installed = False
    _comm_stats: "OrderedDict[Tuple[str, str], List[int]]" = OrderedDict()  # (collective, site) -> [calls, bytes]
    _op_stats: "OrderedDict[Tuple[str, str], int]" = OrderedDict()  # (module, op) -> calls

    # ------------------------------------------------------------------ plumbing
    @classmethod
    def _rank(cls) -> int:
        return dist.get_rank() if dist.is_available() and dist.is_initialized() else int(os.environ.get("RANK", "0"))

    @classmethod
    def _rank_ok(cls) -> bool:
        if cls.ranks is None:
            return cls._rank() == 0
        return -1 in cls.ranks or cls._rank() in cls.ranks

    @classmethod
    def log(cls, *args: Any, **kwargs: Any) -> None:
        """Print-compatible sink: goes to the configured ``logging.Logger`` (level from ``level=``, default INFO) or to stdout."""
        level = kwargs.pop("level", logging.INFO)
        msg = kwargs.pop("sep", " ").join(str(a) for a in args)
        if len(cls.records) < cls.max_records:
            cls.records.append(msg)
        if cls.logger is not None:
            cls.logger.log(level, msg)
        else:
            print(msg, flush=True, **kwargs)

    _emit = log  # round-1 spelling

    # ------------------------------------------------------------------ stack inspection
    @staticmethod
    def trace_to_inject_point() -> str:
        """Innermost user frame (file:line in function): the code that asked for the collective."""
        f = sys._getframe(1)
        while f is not None:
            if _is_user_file(f.f_code.co_filename):
                return f"{os.path.basename(f.f_code.co_filename)}:{f.f_lineno} in {f.f_code.co_name}"
            f = f.f_back
        return "?"

    _user_frame = trace_to_inject_point

    @staticmethod
    def trace_to_module() -> Tuple[str, str, str]:
        """``(module class, phase, file:line)`` of the innermost ``nn.Module.forward`` on the stack; phase is ``backward`` when
        the op runs from the autograd engine (no Python forward frame between the op and the engine's entry)."""
        f = sys._getframe(1)
        site = "?"
        seen_site = False
        while f is not None:
            code = f.f_code
            if not seen_site and _is_user_file(code.co_filename):
                site, seen_site = f"{os.path.basename(code.co_filename)}:{f.f_lineno}", True
            if code.co_name == "forward":
                mod = f.f_locals.get("self", None)
                if isinstance(mod, torch.nn.Module):
                    return type(mod).__name__, "forward", f"{os.path.basename(code.co_filename)}:{f.f_lineno}"
            if code.co_name in ("backward", "_engine_run_backward") and code.co_filename.startswith(_TORCH_DIR):
                return "<autograd>", "backward", site
            f = f.f_back
        return "<no module>", "eager", site

    # ------------------------------------------------------------------ formatting
    @staticmethod
    def dt_spec_debug_formatter(spec) -> str:
        """Compact spec: ``f32[8, 16] (Shard(0), Replicate) @ mesh(2, 2)``."""
        if spec is None:
            return "None"
        if isinstance(spec, (list, tuple)):
            return "[" + ", ".join(DebugLogger.dt_spec_debug_formatter(s) for s in spec) + "]"
        meta = getattr(spec, "tensor_meta", None)
        shape = list(meta.shape) if meta is not None else list(getattr(spec, "shape", []))
        dt = str(getattr(meta, "dtype", "")).replace("torch.", "")
        mesh = getattr(spec, "mesh", None)
        mshape = tuple(mesh.shape) if mesh is not None and hasattr(mesh, "shape") else "?"
        return f"{dt}{shape} {tuple(spec.placements)} @ mesh{mshape}"

    # ------------------------------------------------------------------ hooks
    @classmethod
    def _comm_hook(cls, name, nbytes, group, kw):
        if not (cls.IS_DEBUG_MODE and cls._rank_ok()):
            return
        site = cls.trace_to_inject_point()
        st = cls._comm_stats.setdefault((name, site), [0, 0])
        st[0] += 1
        st[1] += int(nbytes)
        gs = dist.get_world_size(group) if group is not None and dist.is_initialized() and not isinstance(group, int) else "?"
        cls.log(f"[rank{cls._rank()}][comm] {name} bytes={nbytes} group_size={gs} {kw if kw else ''} @ {site}")

    @classmethod
    def log_communication(cls, func: Union[str, Callable], *args, **kwargs) -> None:
        """Manual form (legacy ``DebugLogger.log_communication(func, *args)``): log a collective that does not go through
        ``comm.collectives`` — tensors in ``args`` are summarised as dtype/shape."""
        if not (cls.IS_DEBUG_MODE and cls._rank_ok()):
            return
        name = func if isinstance(func, str) else getattr(func, "__name__", repr(func))
        brief = [f"{str(a.dtype).replace('torch.', '')}{list(a.shape)}" if isinstance(a, torch.Tensor) else repr(a) for a in args]
        nbytes = sum(a.numel() * a.element_size() for a in args if isinstance(a, torch.Tensor))
        site = cls.trace_to_inject_point()
        st = cls._comm_stats.setdefault((name, site), [0, 0])
        st[0] += 1
        st[1] += nbytes
        cls.log(f"[rank{cls._rank()}][comm] {name}({', '.join(brief)}{', ' if kwargs and brief else ''}{', '.join(f'{k}={v}' for k, v in kwargs.items())}) @ {site}")

    @classmethod
    def log_communication_decorator(cls) -> Callable:
        """``@DebugLogger.log_communication_decorator()`` on a user-defined collective wrapper (legacy ``:192``)."""

        def decorator(func):
            @functools.wraps(func)
            def wrapper(*args, **kwargs):
                cls.log_communication(func, *args, **kwargs)
                return func(*args, **kwargs)

            return wrapper

        return decorator

    @classmethod
    def _op_hook(cls, op, schema, out_sh):
        if not (cls.IS_DEBUG_MODE and cls._rank_ok()):
            return
        mod, phase, site = cls.trace_to_module()
        key = (f"{mod}.{phase}", str(op))
        cls._op_stats[key] = cls._op_stats.get(key, 0) + 1
        ins = [cls.dt_spec_debug_formatter(s) for s in schema.tensor_specs()]
        redis = None
        if out_sh.redistribute_specs is not None:
            redis = [None if s is None else str(tuple(s.placements)) for s in out_sh.redistribute_specs]
            if all(r is None for r in redis):
                redis = None
        cls.log(
            f"[rank{cls._rank()}][op] {mod} {phase}() at {site}: {op}\n"
            f"      in  = {ins}\n"
            f"      out = {cls.dt_spec_debug_formatter(out_sh.output_spec)}" + (f"\n      redistribute inputs -> {redis}" if redis else "")
        )

    log_op = _op_hook

    # ------------------------------------------------------------------ aggregate view
    @classmethod
    def summary(cls, reset: bool = False) -> str:
        """Table of collectives by call site (calls, bytes) and of ops by module since the mode was switched on."""
        lines = [f"[rank{cls._rank()}] communication by call site:"]
        for (name, site), (n, b) in sorted(cls._comm_stats.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"  {name:<28s} {n:>6d} calls {b / 2**20:>10.2f} MiB  @ {site}")
        lines.append(f"[rank{cls._rank()}] DTensor ops by module:")
        for (mod, op), n in sorted(cls._op_stats.items(), key=lambda kv: -kv[1]):
            lines.append(f"  {mod:<36s} {op:<40s} {n:>6d}")
        if reset:
            cls.reset()
        return "\n".join(lines)

    @classmethod
    def reset(cls) -> None:
        cls.records = []
        cls._comm_stats = OrderedDict()
        cls._op_stats = OrderedDict()

    # ------------------------------------------------------------------ switches
    @classmethod
    def install(cls):
        if cls._installed:
            return
        from ..dtensor.dispatch import dispatcher

        C.add_comm_hook(cls._comm_hook)
        dispatcher._hooks.append(cls._op_hook)
        cls._installed = True

    @classmethod
    def uninstall(cls):
        if not cls._installed:
            return
        from ..dtensor.dispatch import dispatcher

        C.remove_comm_hook(cls._comm_hook)
        # bound classmethods compare equal, not identical
        dispatcher._hooks[:] = [h for h in dispatcher._hooks if h != cls._op_hook]
        cls._installed = False

    @staticmethod
    def _parse_env() -> Tuple[bool, Optional[Tuple[int, ...]]]:
        """``VESCALE_DEBUG_MODE=1`` (rank 0), ``=1:0,3`` (ranks 0 and 3), ``=1:-1`` (every rank), empty / ``0`` = off."""
        raw = os.environ.get("VESCALE_DEBUG_MODE", "")
        flag, _, ranks = raw.partition(":")
        on = flag not in ("", "0")
        return on, (tuple(int(r) for r in ranks.split(",") if r.strip()) if ranks.strip() else None)

    @classmethod
    def update_vescale_debug_mode_from_env(cls) -> bool:
        on, ranks = cls._parse_env()
        if on != cls.IS_DEBUG_MODE or (on and ranks is not None and ranks != cls.ranks):
            raw = os.environ.get("VESCALE_DEBUG_MODE", "")
            set_vescale_debug_mode(on, rank_to_print=ranks if ranks is not None else cls.ranks, logger=cls.logger)
            os.environ["VESCALE_DEBUG_MODE"] = raw
        return on

    @classmethod
    def set_vescale_debug_mode(cls, on: bool = True, *, rank_to_print=None, logger: Optional[logging.Logger] = None) -> None:
        set_vescale_debug_mode(on, rank_to_print=rank_to_print, logger=logger)


def set_vescale_debug_mode(on: bool = True, *, rank_to_print: Union[None, int, Sequence[int]] = None, logger: Optional[logging.Logger] = None) -> None:
    if rank_to_print is not None and not isinstance(rank_to_print, int):
        if not (isinstance(rank_to_print, Sequence) and all(isinstance(i, int) for i in rank_to_print)):
            raise TypeError(f"expect rank_to_print to be an int or a tuple / list of int, got {type(rank_to_print)}")
    os.environ["VESCALE_DEBUG_MODE"] = str(int(bool(on)))
    DebugLogger.IS_DEBUG_MODE = DebugLogger.enabled = bool(on)
    DebugLogger.ranks = (rank_to_print,) if isinstance(rank_to_print, int) else (None if rank_to_print is None else tuple(rank_to_print))
    DebugLogger.logger = logger
    if on:
        DebugLogger.install()
    else:
        DebugLogger.uninstall()


def update_vescale_debug_mode_from_env() -> bool:
    return DebugLogger.update_vescale_debug_mode_from_env()


if DebugLogger._parse_env()[0]:
    DebugLogger.update_vescale_debug_mode_from_env()
