"""auto_parallelize_module: derive a DModule sharding plan from a policy and apply it (legacy ``dmp/dmp.py:37-242``).

Two sources of plan entries, in increasing priority:

1. the POLICY (``"MEGATRON"``): providers look at every sub-module and contribute parameter placements and forward resharding
   entries (``policies/megatron.py``; class-level providers registered through ``REGISTRY`` take precedence over the built-in ones);
2. plans the user PINNED on a module with ``set_plan_overriding_policy(module, param_sharding_plan, fwd_resharding_plan)`` — keys
   relative to that module.  A pinned plan replaces everything the policy says about that module's whole subtree (for the kind of
   plan that was pinned: parameters, forward, or both).

``PlanGenerator(model, policy).generate()`` returns the merged root plans and both ingredients; ``auto_parallelize_module`` applies
the merged plan with ``parallelize_module``."""
from __future__ import annotations

import copy
import os
import re
from typing import Dict, List, Optional, Tuple, Union

import torch.nn as nn

from ..dmodule import parallelize_module
from .registry import REGISTRY, _POLICIES, get_policy

__all__ = ["auto_parallelize_module", "set_plan_overriding_policy", "get_plan_overriding_policy", "PlanGenerator", "generate_plan"]

_PARAM_ATTR, _FWD_ATTR = "_vescale_param_plan_overriding_policy", "_vescale_fwd_plan_overriding_policy"
_DEBUG = os.environ.get("VESCALE_DEBUG_MODE", "0") == "1"


def set_plan_overriding_policy(module: nn.Module, param_sharding_plan: Optional[Dict] = None, fwd_resharding_plan: Optional[Dict] = None) -> None:
    """Pin plans on ``module`` (keys relative to it, regular expressions as in ``parallelize_module``); they cover all of its
    sub-modules and override the policy there.  Pinning inside an already pinned subtree is rejected: which one wins would be a guess."""
    if not isinstance(module, nn.Module):
        raise TypeError("set_plan_overriding_policy(module, param_sharding_plan=None, fwd_resharding_plan=None): module must be an nn.Module")
    for name, sub in module.named_modules():
        if hasattr(sub, _PARAM_ATTR) or hasattr(sub, _FWD_ATTR):
            raise NotImplementedError(f"nested set_plan_overriding_policy (already pinned at {name or '<this module>'!r}) is not supported")
    if param_sharding_plan is not None:
        setattr(module, _PARAM_ATTR, param_sharding_plan)
    if fwd_resharding_plan is not None:
        setattr(module, _FWD_ATTR, fwd_resharding_plan)


def get_plan_overriding_policy(module: nn.Module) -> Tuple[Optional[Dict], Optional[Dict]]:
    """The plans pinned on exactly this module: ``(param_sharding_plan, fwd_resharding_plan)``."""
    return getattr(module, _PARAM_ATTR, None), getattr(module, _FWD_ATTR, None)


def _prefix(fqn: str, key: str) -> str:
    """Root-relative plan key.  Keys are regular expressions matched against fully qualified names; module paths are written with
    plain dots (``blocks.0.fc1.weight`` — a dot matches a dot), which keeps generated plans readable and comparable."""
    return key if not fqn else fqn + "." + key


def _plain(key: str) -> str:
    return key.replace(r"\.", ".")


class PlanGenerator:
    _registry = REGISTRY

    def __init__(self, model: nn.Module, policy: Union[str, None, Dict[str, str]] = "MEGATRON"):
        self.model = model
        if not isinstance(policy, str):
            raise NotImplementedError("policy: the name of a registered policy (per-module policy dicts are not supported)")
        self.policy = policy.upper()
        if self.policy not in _POLICIES and not self._registry.has_policy(self.policy):
            raise ValueError(f"policy {self.policy!r} is not registered; registered policies: {sorted(self._registry.get_all_policies())}")

    # -- the two ingredients --------------------------------------------------------------------------------------------------------------------
    def pinned_plans(self) -> Tuple[Dict, Dict, List[str], List[str]]:
        """Root-relative pinned plans plus the fqns whose subtrees they claim (parameters / forward)."""
        root_p: Dict = {}
        root_f: Dict = {}
        claim_p: List[str] = []
        claim_f: List[str] = []
        for fqn, mod in self.model.named_modules():
            p, f = get_plan_overriding_policy(mod)
            if p is not None:
                claim_p.append(fqn)
                root_p.update({_prefix(fqn, k): v for k, v in p.items()})
            if f is not None:
                claim_f.append(fqn)
                root_f.update({_prefix(fqn, k): v for k, v in f.items()})
        return root_p, root_f, claim_p, claim_f

    def policy_plans(self) -> Tuple[Dict, Dict]:
        pol = get_policy(self.policy)
        root_p: Dict = {}
        root_f: Dict = {}
        for fqn, sub in self.model.named_modules():
            r = pol.provide(fqn, sub, self.model)
            if r is None:
                continue
            if _DEBUG:
                print(f"[dmp] {fqn or '<root>'} : {type(sub).__name__} --{self.policy}--> {r}")
            root_p.update({_plain(k): v for k, v in r.get("parameter", {}).items()})
            root_f.update({_plain(k): v for k, v in r.get("forward", {}).items()})
        return root_p, root_f

    @staticmethod
    def _override(claims: List[str], high: Dict, low: Dict) -> Dict:
        """``low`` without the entries that live under a claimed subtree, then ``high`` on top."""
        out = copy.copy(low)
        for fqn in claims:
            pre = fqn + "." if fqn else ""
            for k in [k for k in out if k.startswith(pre)]:
                del out[k]
        out.update(high)
        return out

    def generate(self) -> Tuple[Dict, Dict, Dict, Dict, Dict, Dict]:
        """``(param plan, forward plan, pinned param plan, pinned forward plan, policy param plan, policy forward plan)``, all relative
        to the model root."""
        pin_p, pin_f, claim_p, claim_f = self.pinned_plans()
        pol_p, pol_f = self.policy_plans()
        return self._override(claim_p, pin_p, pol_p), self._override(claim_f, pin_f, pol_f), pin_p, pin_f, pol_p, pol_f


def generate_plan(module: nn.Module, policy: str = "MEGATRON") -> Dict[str, Dict]:
    p, f, *_ = PlanGenerator(module, policy).generate()
    return {"parameter": p, "forward": f}


def auto_parallelize_module(module: nn.Module, device_mesh=None, policy: str = "MEGATRON", plan_only: bool = False, *, plan_override: Optional[Dict[str, Dict]] = None,
                            plan_to_save: Optional[Dict] = None, **kw):
    """``device_mesh=None``: the mesh of the enclosing ``with DeviceMesh(...):`` block.  ``plan_only``: generate, do not parallelize —
    returns ``(module, device_mesh, {"parameter": ..., "forward": ...})``.  ``plan_override``: extra root-relative entries laid over
    the generated plan (a one-call alternative to pinning).  ``plan_to_save``: a dict that receives the final plans under
    ``"param_sharding_plan"`` / ``"fwd_resharding_plan"``."""
    plan = generate_plan(module, policy)
    if plan_override:
        plan["parameter"].update(plan_override.get("parameter", {}))
        plan["forward"].update(plan_override.get("forward", {}))
    if plan_to_save is not None:
        plan_to_save["param_sharding_plan"], plan_to_save["fwd_resharding_plan"] = dict(plan["parameter"]), dict(plan["forward"])
    if plan_only:
        return module, device_mesh, plan
    module._auto_plan = plan
    return parallelize_module(module, device_mesh, plan, **kw)
