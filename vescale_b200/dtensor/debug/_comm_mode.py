"""CommDebugMode: count the collectives issued inside a region, per kind and per nn.Module, split into
forward and backward.  Parity: reference ``vescale/dtensor/debug/_comm_mode.py:20-103``."""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Optional

import torch

from ...comm import collectives as C

__all__ = ["CommDebugMode"]


class CommDebugMode:
    def __init__(self, module: Optional[torch.nn.Module] = None):
        self.comm_counts: Dict[str, int] = defaultdict(int)
        self.comm_bytes: Dict[str, int] = defaultdict(int)
        self.module_counts: Dict[str, Dict[str, int]] = defaultdict(lambda: defaultdict(int))
        self._module = module
        self._stack = []
        self._handles = []
        self._in_backward = False

    def _hook(self, name, nbytes, group, kw):
        self.comm_counts[name] += 1
        self.comm_bytes[name] += nbytes
        phase = "backward" if self._in_backward or torch._C._current_graph_task_id() != -1 else "forward"
        self.comm_counts[f"{phase}.{name}"] += 1
        if self._stack:
            self.module_counts[self._stack[-1]][f"{phase}.{name}"] += 1

    def __enter__(self):
        C.add_comm_hook(self._hook)
        if self._module is not None:
            for fqn, m in self._module.named_modules():
                fqn = fqn or type(m).__name__
                def _push(mod, a, _n=fqn):
                    self._stack.append(_n)

                def _pop(mod, a, o):
                    if self._stack:
                        self._stack.pop()

                self._handles.append(m.register_forward_pre_hook(_push))
                self._handles.append(m.register_forward_hook(_pop))
        return self

    def __exit__(self, *exc):
        C.remove_comm_hook(self._hook)
        for h in self._handles:
            h.remove()
        self._handles.clear()

    def get_total_counts(self) -> int:
        return sum(v for k, v in self.comm_counts.items() if "." not in k)

    def get_comm_counts(self) -> Dict[str, int]:
        return {k: v for k, v in self.comm_counts.items() if "." not in k}

    def get_phase_counts(self, phase: str) -> Dict[str, int]:
        return {k.split(".", 1)[1]: v for k, v in self.comm_counts.items() if k.startswith(phase + ".")}

    def generate_comm_debug_tracing_table(self) -> str:
        lines = ["collective            count        bytes"]
        for k, v in sorted(self.get_comm_counts().items()):
            lines.append(f"{k:<20} {v:>6} {self.comm_bytes[k]:>12}")
        for mod, d in self.module_counts.items():
            lines.append(f"  [{mod}] " + ", ".join(f"{k}={v}" for k, v in sorted(d.items())))
        return "\n".join(lines)
