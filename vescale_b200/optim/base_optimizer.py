"""BasicOptimizer: wraps any torch optimizer for DModule / DDP models — before ``step`` it finishes the
module's gradient synchronisation (Partial grads of TP/SP, DDP buckets) and copies ``main_grad`` into
``param.grad``.  Parity: ``legacy/vescale/optim/base_optimizer.py:116-206``."""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch
import torch.nn as nn

from ..dtensor.api import DTensor

__all__ = ["BasicOptimizer", "GradOptimizerHookBase", "BasicOptimizerHook", "OptimizerBase"]


class GradOptimizerHookBase:
    """Hooks run before / after ``optimizer.step`` (legacy ``optim/base_optimizer.py:30-43``)."""
    @staticmethod
    def step_pre_hook(optim, *a, **kw):
        raise NotImplementedError

    @staticmethod
    def step_post_hook(optim, *a, **kw):
        raise NotImplementedError


class BasicOptimizerHook(GradOptimizerHookBase):
    """Before ``optimizer.step``: parameters whose gradient lives only in a DDP/FSDP ``main_grad`` view get a ``.grad`` over
    that same storage (wrapped as a DTensor with the parameter's placements when the parameter is one), so any
    ``torch.optim`` optimizer can consume it (legacy ``optim/base_optimizer.py:45-70``)."""

    @staticmethod
    def step_pre_hook(optim, *a, **kw):
        for group in optim.param_groups:
            for p in group["params"]:
                mg = getattr(p, "main_grad", None)
                if p.grad is not None or mg is None:
                    continue
                data = p.data if isinstance(p.data, DTensor) else (p if isinstance(p, DTensor) else None)
                p.grad = DTensor(mg.to(p.dtype), data._spec) if data is not None else mg.to(p.dtype)

    @staticmethod
    def step_post_hook(optim, *a, **kw):
        return None


class OptimizerBase:
    """What every optimizer wrapper of this package provides (legacy ``base_optimizer.py:26-113``): a ``step`` that first finishes
    gradient synchronisation, ``zero_grad`` that also clears gradient buffers, ``state_dict`` / ``load_state_dict``, and the
    ``param_groups`` of the wrapped optimizer.  ``BasicOptimizer`` and ``DistributedOptimizer`` are its two implementations."""

    optimizer: torch.optim.Optimizer

    def step(self, closure=None):  # pragma: no cover - interface
        raise NotImplementedError

    def zero_grad(self, set_to_none: bool = True):  # pragma: no cover - interface
        raise NotImplementedError

    def state_dict(self):  # pragma: no cover - interface
        raise NotImplementedError

    def load_state_dict(self, state_dict):  # pragma: no cover - interface
        raise NotImplementedError

    def get_loss_scale(self) -> float:
        """bf16 / fp32 training does not scale the loss."""
        return 1.0

    # ``optimizer.state`` / ``optimizer.param_groups`` read and written through the wrapper (learning-rate schedulers do both)
    @property
    def state(self):
        return self.optimizer.state

    @state.setter
    def state(self, value):
        self.optimizer.state = value


class BasicOptimizer(OptimizerBase):
    """Thin wrapper that finishes DModule / DDP gradient synchronisation, exposes ``main_grad`` as ``.grad``, clips, and steps the
    inner ``torch.optim`` optimizer (legacy ``optim/base_optimizer.py:116-206``)."""
    def __init__(self, optimizer: torch.optim.Optimizer, models: Union[nn.Module, Sequence[nn.Module]], grad_hook: Optional[GradOptimizerHookBase] = None, clip_grad: float = 0.0):
        self.optimizer = optimizer
        self.models = [models] if isinstance(models, nn.Module) else list(models)
        self.clip_grad = clip_grad
        self.grad_hook = grad_hook

    @property
    def param_groups(self):
        return self.optimizer.param_groups

    def _sync(self):
        from ..parallel.ddp import DistributedDataParallel

        for m in self.models:
            if isinstance(m, DistributedDataParallel):
                m.finish_grad_sync()
                for gb in m.grad_buffers.values():
                    for p in gb.params:
                        g = p.main_grad
                        if isinstance(p.data, DTensor) or isinstance(p, DTensor):
                            p.grad = DTensor(g.to(p.dtype), p._spec if isinstance(p, DTensor) else p.data._spec)
                        else:
                            p.grad = g.to(p.dtype)
                inner = m.module
            else:
                inner = m
            for sub in inner.modules():
                fn = getattr(sub, "_dmodule", None)
                if fn is not None:
                    fn.finish_grad_sync()

    def step(self, closure=None):
        if self.grad_hook is not None:
            self.grad_hook.step_pre_hook(self.optimizer)
        self._sync()
        norm = None
        if self.clip_grad and self.clip_grad > 0:
            from .clip_grads import clip_grad_norm_fp32

            params = [p for g in self.optimizer.param_groups for p in g["params"] if p.grad is not None]
            norm = clip_grad_norm_fp32(params, self.clip_grad)
        out = self.optimizer.step(closure)
        if self.grad_hook is not None:
            self.grad_hook.step_post_hook(self.optimizer)
        return norm if norm is not None else out

    def zero_grad(self, set_to_none: bool = True):
        from ..parallel.ddp import DistributedDataParallel

        self.optimizer.zero_grad(set_to_none)
        for m in self.models:
            if isinstance(m, DistributedDataParallel):
                m.zero_grad_buffer()

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, sd):
        self.optimizer.load_state_dict(sd)
