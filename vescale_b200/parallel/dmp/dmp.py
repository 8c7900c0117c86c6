"""auto_parallelize_module: derive a DModule sharding plan from a policy and apply it.
Parity: ``legacy/vescale/dmp/dmp.py:61-242``."""
from __future__ import annotations

from typing import Dict, Optional

import torch.nn as nn

from ..dmodule import parallelize_module
from .registry import get_policy

__all__ = ["auto_parallelize_module", "set_plan_overriding_policy", "get_plan_overriding_policy"]

_OVERRIDE = {"policy": "PARAM_FIRST"}


def set_plan_overriding_policy(module=None, policy: str = "PARAM_FIRST") -> None:
    _OVERRIDE["policy"] = policy


def get_plan_overriding_policy(module=None) -> str:
    return _OVERRIDE["policy"]


def generate_plan(module: nn.Module, policy: str = "MEGATRON") -> Dict[str, Dict]:
    pol = get_policy(policy)
    plan = {"parameter": {}, "forward": {}}
    for fqn, sub in module.named_modules():
        if not fqn:
            continue
        r = pol.provide(fqn, sub, module)
        if r is None:
            continue
        plan["parameter"].update(r.get("parameter", {}))
        plan["forward"].update(r.get("forward", {}))
    return plan


def auto_parallelize_module(module: nn.Module, device_mesh, policy: str = "MEGATRON", *, plan_override: Optional[Dict[str, Dict]] = None, **kw) -> nn.Module:
    plan = generate_plan(module, policy)
    if plan_override:
        plan["parameter"].update(plan_override.get("parameter", {}))
        plan["forward"].update(plan_override.get("forward", {}))
    module._auto_plan = plan
    return parallelize_module(module, device_mesh, plan, **kw)
