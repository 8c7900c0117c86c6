"""Policy / provider registry for automatic plan generation (legacy ``dmp/registry.py``)."""
from typing import Callable, Dict, List

_POLICIES: Dict[str, "Policy"] = {}


class Policy:
    def __init__(self, name: str):
        self.name = name
        self.providers: List[Callable] = []

    def provide(self, fqn, module, root):
        for p in self.providers:
            r = p(fqn, module, root)
            if r is not None:
                return r
        return None


def register_policy(name: str) -> Policy:
    return _POLICIES.setdefault(name.upper(), Policy(name.upper()))


def get_policy(name: str) -> Policy:
    return _POLICIES[name.upper()]


def register_provider(policy: str):
    def deco(fn):
        register_policy(policy).providers.append(fn)
        return fn

    return deco


class ModulePolicyRegistry:
    """Class-name keyed registration, the reference's extension point (legacy ``dmp/policies/registry.py:25-110``):

        register = REGISTRY.provide_register_for_policy("MEGATRON")

        @register(["MyMLP", "FFN"])
        def provider(fqn, module):
            return {"fc1.weight": [Shard(0)]}, {"fc1.input": [[Replicate()]]}     # keys relative to the module

    A provider registered here takes precedence over the built-in ones of the same policy; lookup is by exact upper-cased class
    name first, then by registered name contained in the class name."""

    def __init__(self):
        self.mpp: Dict[str, Dict[str, Callable]] = {}
        self._bridged = set()

    def provide_register_for_policy(self, policy_name: str) -> Callable:
        policy_name = policy_name.upper()

        def register(module_cls_name):
            names = [module_cls_name] if isinstance(module_cls_name, str) else list(module_cls_name)

            def deco(fn: Callable) -> Callable:
                for n in names:
                    slot = self.mpp.setdefault(n.upper(), {})
                    if policy_name in slot:
                        raise ValueError(f"policy {policy_name} already has a provider for module class {n}")
                    slot[policy_name] = fn
                self._bridge(policy_name)
                return fn

            return deco

        return register

    def _bridge(self, policy_name: str) -> None:
        if policy_name in self._bridged:
            return
        self._bridged.add(policy_name)
        import re

        def provider(fqn, module, root):
            fn = self.get_policy_provider(type(module).__name__, policy_name) or self.get_policy_provider_if_module_contains_registered_name(
                type(module).__name__, policy_name
            )
            if fn is None:
                return None
            param_plan, fwd_plan = fn(fqn, module)
            pre = re.escape(fqn) + r"\." if fqn else ""
            return {"parameter": {pre + re.escape(k): v for k, v in param_plan.items()}, "forward": {pre + re.escape(k): v for k, v in fwd_plan.items()}}

        register_policy(policy_name).providers.insert(0, provider)

    def get_policy_provider(self, module_cls_name: str, policy_name: str = None):
        slot = self.mpp.get(module_cls_name.upper())
        if slot is None:
            return None
        return slot if policy_name is None else slot.get(policy_name.upper())

    def get_policy_provider_if_module_contains_registered_name(self, module_cls_name: str, policy_name: str = None):
        up = module_cls_name.upper()
        for n, slot in self.mpp.items():
            if n in up and (policy_name is None or policy_name.upper() in slot):
                return slot if policy_name is None else slot[policy_name.upper()]
        return None

    def has_module(self, module_cls_name: str) -> bool:
        return module_cls_name.upper() in self.mpp

    def get_all_modules(self):
        return set(self.mpp)

    def get_all_policies(self):
        return {p for slot in self.mpp.values() for p in slot} | set(_POLICIES)

    def has_policy(self, policy_name: str) -> bool:
        return policy_name.upper() in self.get_all_policies()

    def __repr__(self) -> str:
        rows = [f"{m} : {sorted(slot)}" for m, slot in self.mpp.items()]
        return "\n".join(["== module class : policies ==", *rows, f"== built-in policies : {sorted(_POLICIES)} =="])


Registry = ModulePolicyRegistry  # the reference's class name
REGISTRY = ModulePolicyRegistry()
