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
