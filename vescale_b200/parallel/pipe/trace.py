"""Graph-level pipeline tracer for models that are NOT a chain of units (HuggingFace transformers, anything with values that
skip over stages: rotary tables, attention masks, residual streams of other branches).

The reference traces such models with ``torch.fx`` / its HF tracer / dynamo export and splits the traced graph
(``legacy/vescale/pipe/tracer.py:81-709``, ``pipe_parser.py:46-652``).  ``transformers.utils.fx`` no longer exists in current
transformers; the capture here is ``torch.export`` (non-strict), which handles HF models as they are:

1. ``torch.export.export(model, example_inputs)`` -> one aten-level graph whose nodes remember the module they came from
   (``nn_module_stack``);
2. every node is assigned to a virtual stage by the module FQN it belongs to — stage boundaries are ``plan.split_points`` (FQN of the
   last module of each stage) or, by default, an even split of the model's longest ``ModuleList`` (the decoder layers); stage
   numbers never decrease along the graph, so helper ops between layers stay with the earlier stage;
3. ``torch.fx.passes.split_module`` cuts the graph; a liveness pass then turns the cut into a CHAIN: stage ``s`` receives every
   value that is produced before ``s`` and needed at or after ``s`` (the pipeline engine only ever ships a stage's outputs to
   the next stage), runs its sub-graph, and forwards the values later stages still need.

``trace_and_split(model, plan)`` returns one ``nn.Module`` per virtual stage; parameters keep their tensors (the sub-graphs hold
``get_attr`` references to the captured parameters), so optimizers and checkpoints see the original storage.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

__all__ = ["trace_and_split", "ChainStage"]


def _longest_module_list(model: nn.Module) -> Optional[Tuple[str, int]]:
    best = None
    for n, m in model.named_modules():
        if isinstance(m, (nn.ModuleList, nn.Sequential)) and (best is None or len(m) > best[1]):
            best = (n, len(m))
    return best


def _stage_prefixes(model: nn.Module, plan) -> List[Tuple[str, int]]:
    """[(module fqn prefix, stage)] — a node whose module stack contains ``prefix`` belongs to (at least) ``stage``."""
    n_vs = plan.num_stages * plan.virtual_chunks
    if plan.split_points:
        if len(plan.split_points) != n_vs - 1:
            raise ValueError(f"graph tracer: {n_vs} stages need {n_vs - 1} split points, got {len(plan.split_points)}")
        names = [n for n, _ in model.named_modules()]
        out = []
        cuts = {p: s for s, p in enumerate(plan.split_points)}
        stage = 0
        # modules in definition order: everything up to and including split point k is stage <= k
        for n in names:
            if not n:
                continue
            out.append((n, stage))
            if n in cuts:
                stage = cuts[n] + 1
        return out
    ml = _longest_module_list(model)
    if ml is None or ml[1] < n_vs:
        raise ValueError("graph tracer: give plan.split_points (the model has no ModuleList long enough to split evenly)")
    name, L = ml
    base, rem = divmod(L, n_vs)
    out, k = [], 0
    for s in range(n_vs):
        for _ in range(base + (1 if s < rem else 0)):
            out.append((f"{name}.{k}", s))
            k += 1
    return out


class ChainStage(nn.Module):
    """One virtual stage of a graph split: ``forward(*live_in) -> live_out`` (see the module docstring).  Parameters / buffers the
    sub-graph takes as arguments are owned by the stage (``self.weights``, the ORIGINAL tensors of the traced model)."""

    def __init__(self, sub: nn.Module, arg_src: Sequence[Tuple[str, object]], n_out: int, out_spec: Sequence[Tuple[str, int]], single: bool):
        super().__init__()
        self.sub = sub
        self.arg_src = []  # per sub-graph argument: ("live", position in live_in) | ("w", key in self.weights / self._bufs)
        self.weights = nn.ParameterDict()
        self.names: Dict[str, str] = {}  # key -> original fqn
        for kind, v in arg_src:
            if kind == "live":
                self.arg_src.append(("live", v))
            else:
                fqn, t = v
                key = fqn.replace(".", "__")
                if isinstance(t, nn.Parameter):
                    self.weights[key] = t
                else:
                    self.register_buffer(key, t, persistent=False)
                self.names[key] = fqn
                self.arg_src.append(("w", key))
        self.n_out = n_out
        self.out_spec = list(out_spec)  # live_out[k] = ("in", i) pass-through of live_in[i] | ("out", j) output j of the sub-graph
        self.single = single

    def forward(self, *live_in):
        res = self.sub(*[live_in[v] if kind == "live" else (self.weights[v] if v in self.weights else getattr(self, v)) for kind, v in self.arg_src])
        outs = (res,) if not isinstance(res, (tuple, list)) else tuple(res)
        live_out = tuple(live_in[i] if kind == "in" else outs[i] for kind, i in self.out_spec)
        return live_out[0] if (self.single and len(live_out) == 1) else live_out


def trace_and_split(model: nn.Module, plan) -> List[nn.Module]:
    from torch.fx.passes.split_module import split_module

    ex = getattr(plan, "example_inputs", None)
    if ex is None:
        raise ValueError("the graph tracer needs plan.example_inputs (a tuple of example arguments of the whole model)")
    ep = torch.export.export(model, tuple(ex), strict=False)
    gm = ep.module()
    prefixes = _stage_prefixes(model, plan)
    by_len = sorted(prefixes, key=lambda t: -len(t[0]))
    cur = [0]

    def part(node) -> int:
        st = node.meta.get("nn_module_stack")
        if st:
            fqns = [v[0] for v in st.values()]
            for pre, s in by_len:
                if any(f == pre or f.startswith(pre + ".") for f in fqns):
                    cur[0] = max(cur[0], s)
                    break
        return cur[0]

    split = split_module(gm, None, part)
    # ---- liveness over the top-level graph: values = placeholders and (stage, output index) pairs
    calls = [n for n in split.graph.nodes if n.op == "call_module"]
    stage_of = {n: i for i, n in enumerate(calls)}
    n_st = len(calls)
    value_of: Dict[torch.fx.Node, Tuple] = {}
    placeholders = [n for n in split.graph.nodes if n.op == "placeholder"]
    for i, n in enumerate(placeholders):
        value_of[n] = ("ph", i)
    n_out = [1] * n_st
    for n in split.graph.nodes:
        if n.op == "call_module":
            users_getitem = [u for u in n.users if u.op == "call_function" and getattr(u.target, "__name__", "") == "getitem"]
            if users_getitem and len(users_getitem) == len(n.users):
                n_out[stage_of[n]] = max(int(u.args[1]) for u in users_getitem) + 1
                for u in users_getitem:
                    value_of[u] = ("st", stage_of[n], int(u.args[1]))
            else:
                value_of[n] = ("st", stage_of[n], 0)
    out_node = next(n for n in split.graph.nodes if n.op == "output")
    flat_out: List[torch.fx.Node] = []

    def walk(a):
        if isinstance(a, torch.fx.Node):
            flat_out.append(a)
        elif isinstance(a, (tuple, list)):
            for x in a:
                walk(x)
        elif isinstance(a, dict):
            for x in a.values():
                walk(x)

    walk(out_node.args)
    need_at: Dict[Tuple, int] = {}  # value -> last stage that needs it as an input (n_st = needed by the model output)
    args_of: List[List[Tuple]] = []
    orig = dict(model.named_parameters(remove_duplicate=False))
    orig.update(dict(model.named_buffers(remove_duplicate=False)))

    def attr(target: str):
        if target in orig:
            return orig[target]
        obj = split
        for a in target.split("."):
            obj = getattr(obj, a)
        return obj

    for c in calls:
        vals = []
        for a in c.args:
            if a.op == "get_attr":
                vals.append(("w", str(a.target), attr(str(a.target))))
            else:
                vals.append(value_of[a])
        args_of.append(vals)
        for v in vals:
            if v[0] != "w":
                need_at[v] = max(need_at.get(v, -1), stage_of[c])
    final_vals = [value_of[n] for n in flat_out]
    for v in final_vals:
        need_at[v] = n_st

    def produced_at(v) -> int:
        return -1 if v[0] == "ph" else v[1]

    stages: List[nn.Module] = []
    live_in: List[Tuple] = [("ph", i) for i in range(len(placeholders))]
    for s, c in enumerate(calls):
        sub = getattr(split, c.target)
        arg_src = [("w", (v[1], v[2])) if v[0] == "w" else ("live", live_in.index(v)) for v in args_of[s]]
        if s + 1 < n_st:
            nxt = [v for v in live_in if need_at.get(v, -1) > s]
            nxt += [("st", s, k) for k in range(n_out[s]) if need_at.get(("st", s, k), -1) > s]
        else:
            nxt = list(final_vals)
        spec = [("in", live_in.index(v)) if produced_at(v) < s else ("out", v[2]) for v in nxt]
        stages.append(ChainStage(sub, arg_src, n_out[s], spec, single=(s + 1 == n_st)))
        live_in = nxt
    return stages
