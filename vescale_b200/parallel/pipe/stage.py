"""Model splitting and per-rank stage modules.

``construct_pipeline_stage(model, plan, mesh)`` splits the model into ``num_stages * virtual_chunks`` virtual
stages and returns a ``PipeModule`` that holds only this rank's chunks.

* STRUCTURAL tracer: the model is a chain of *units* — ``model.pipeline_units()`` if defined, else its direct
  children in registration order; stage i is the ``nn.Sequential`` of its units.  Split by MANUAL split points,
  UNIFORM unit count, or PARAMETERS (balanced parameter count).
* FX tracer: ``torch.fx.symbolic_trace`` + ``split_module`` with call_module nodes assigned to stages by the same
  unit→stage map (handles skip connections across stages as extra stage inputs/outputs).
* tied parameters listed in ``plan.shared_modules`` get a process group spanning their owning ranks;
  ``sync_shared_params`` all-reduces their values or gradients (legacy ``pipe_stage.py:200-247,311-500``).

Parity: ``legacy/vescale/pipe/pipe_parser.py:46-652`` (split methods), ``pipe_stage.py:64-563``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from .plan import PipelineParallelPlan, PipelineScheduleType, PipelineSplitMethodType, TracerType
from .schedule import stage_placement

__all__ = ["PipeModule", "construct_pipeline_stage", "split_units", "PipeParser"]


def _units(model: nn.Module) -> List[Tuple[str, nn.Module]]:
    if hasattr(model, "pipeline_units"):
        return list(model.pipeline_units())
    out = []
    for n, m in model.named_children():
        if isinstance(m, (nn.ModuleList, nn.Sequential)):
            out += [(f"{n}.{k}", c) for k, c in m.named_children()]
        else:
            out.append((n, m))
    return out


def split_units(units: Sequence[Tuple[str, nn.Module]], plan: PipelineParallelPlan) -> List[List[int]]:
    """Indices of the units of every virtual stage."""
    n_vs = plan.num_stages * plan.virtual_chunks
    n = len(units)
    if n < n_vs:
        raise ValueError(f"{n} units cannot fill {n_vs} virtual stages")
    if plan.split_method == PipelineSplitMethodType.MANUAL:
        names = [u[0] for u in units]
        cuts = [names.index(p) + 1 for p in (plan.split_points or [])]
        if len(cuts) != n_vs - 1:
            raise ValueError(f"MANUAL split needs {n_vs - 1} split points, got {len(cuts)}")
        bounds = [0] + cuts + [n]
    elif plan.split_method == PipelineSplitMethodType.PARAMETERS:
        sizes = [sum(p.numel() for p in u[1].parameters()) + 1 for u in units]
        total = sum(sizes)
        bounds, acc, tgt = [0], 0, total / n_vs
        for i, s in enumerate(sizes):
            acc += s
            left_units, left_stages = n - (i + 1), n_vs - len(bounds)
            if len(bounds) < n_vs and (acc >= tgt * len(bounds) or left_units == left_stages) and left_units >= left_stages:
                bounds.append(i + 1)
        bounds.append(n)
    else:  # UNIFORM / AUTO
        base, rem = divmod(n, n_vs)
        bounds = [0]
        for s in range(n_vs):
            bounds.append(bounds[-1] + base + (1 if s < rem else 0))
    return [list(range(bounds[s], bounds[s + 1])) for s in range(n_vs)]


class _StageSeq(nn.Module):
    def __init__(self, mods: Sequence[Tuple[str, nn.Module]]):
        super().__init__()
        self.names = [n for n, _ in mods]
        self.mods = nn.ModuleList([m for _, m in mods])

    def forward(self, *xs):
        for m in self.mods:
            xs = m(*xs) if isinstance(xs, tuple) else m(xs)
            if not isinstance(xs, tuple):
                xs = (xs,)
        return xs if len(xs) > 1 else xs[0]


class PipeParser:
    """Splits a model into virtual-stage modules (all of them; the caller keeps its own)."""

    def parse(self, model: nn.Module, plan: PipelineParallelPlan) -> List[nn.Module]:
        if plan.tracer_type == TracerType.GRAPH:
            from .trace import trace_and_split

            return trace_and_split(model, plan)
        units = _units(model)
        groups = split_units(units, plan)
        if plan.tracer_type == TracerType.FX:
            return self._parse_fx(model, units, groups)
        if plan.tracer_type == TracerType.EXPORT:
            return self._parse_export(model, units, groups, plan)
        return [_StageSeq([units[i] for i in g]) for g in groups]

    def _parse_export(self, model, units, groups, plan) -> List[nn.Module]:
        """Dynamo-export tracer (legacy ``pipe/tracer.py`` "dynamo export" mode): ``torch.export.export`` captures an aten-level
        graph with the parameters lifted; ``unflatten`` restores one graph module per original sub-module (same FQNs, same
        parameter names), and the stages are cut on the unit boundaries of that captured hierarchy.  Needs
        ``plan.example_inputs`` (a tuple of example arguments for the whole model)."""
        ex = getattr(plan, "example_inputs", None)
        if ex is None:
            raise ValueError("TracerType.EXPORT needs plan.example_inputs")
        ep = torch.export.export(model, tuple(ex))
        um = torch.export.unflatten(ep)
        captured = dict(um.named_children())
        missing = [n for n, _ in units if n.split(".")[0] not in captured]
        if missing:
            raise RuntimeError(f"export tracer: units {missing} are not top-level sub-modules of the captured graph")
        return [_StageSeq([(units[i][0], captured[units[i][0]]) for i in g]) for g in groups]

    def _parse_fx(self, model, units, groups) -> List[nn.Module]:
        import torch.fx as fx
        from torch.fx.passes.split_module import split_module

        unit_stage = {}
        for s, g in enumerate(groups):
            for i in g:
                unit_stage[units[i][0]] = s
        from .tracer import trace_model

        gm = trace_model(model, partition_units=[n for n, _ in units])  # units stay opaque call_module nodes (tracer.py)
        cur = [0]

        def part(node):
            if node.op == "call_module":
                t = str(node.target)
                for name, s in unit_stage.items():
                    if t == name or t.startswith(name + "."):
                        cur[0] = max(cur[0], s)
                        break
            return cur[0]

        split = split_module(gm, model, part)
        return [getattr(split, f"submod_{s}") for s in range(len(groups)) if hasattr(split, f"submod_{s}")]


class PipeModule(nn.Module):
    """This rank's virtual stages.  ``forward(*inputs, chunk_id=c)`` runs local chunk c."""

    def __init__(self, stage_modules: Dict[int, nn.Module], vstages: Dict[int, int], plan: PipelineParallelPlan, pp_rank: int, pp_group=None):
        super().__init__()
        self.stage_modules = nn.ModuleDict({str(c): m for c, m in stage_modules.items()})
        self.vstage_of_chunk = dict(vstages)
        self.plan = plan
        self.pp_rank = pp_rank
        self.pp_group = pp_group
        self.shared_groups: List[Tuple[List[nn.Parameter], object]] = []

    def chunk(self, c: int) -> nn.Module:
        return self.stage_modules[str(c)]

    def forward(self, *inputs, chunk_id: int = 0):
        return self.chunk(chunk_id)(*inputs)

    @property
    def num_chunks(self) -> int:
        return len(self.stage_modules)

    def sync_shared_params(self, share_params: bool = True) -> None:
        """All-reduce (average values / sum gradients) of parameters tied across stages, e.g. embeddings."""
        for params, group in self.shared_groups:
            for p in params:
                if share_params:
                    dist.all_reduce(p.data, group=group)
                    p.data.div_(dist.get_world_size(group))
                elif p.grad is not None:
                    dist.all_reduce(p.grad, group=group)


def parse_model_graph(parser: "PipeParser", model: nn.Module, plan: PipelineParallelPlan):
    """The model at split granularity: the ordered ``(fqn, module)`` units stages are cut between (legacy
    ``pipe_parser.py:579`` returns the traced fx graph; here the hierarchy is the graph, tracers only change how the stage
    modules are materialised)."""
    return _units(model)


def split_pipeline_point(model: nn.Module, plan: PipelineParallelPlan):
    """Resolve the plan's split method into split-point FQNs (the last unit of every virtual stage but the last), store them
    in ``plan.split_points`` and return ``(split_points, units, parser)`` (legacy ``pipe_parser.py:612``)."""
    parser = PipeParser()
    units = parse_model_graph(parser, model, plan)
    groups = split_units(units, plan)
    points = [units[g[-1]][0] for g in groups[:-1]]
    plan.split_points = points
    return points, units, parser


def construct_pipeline_split_graph(model: nn.Module, plan: PipelineParallelPlan, update_split_points: bool = False) -> List[nn.Module]:
    """All virtual-stage modules in order (legacy ``pipe_parser.py:632``)."""
    if update_split_points:
        method = plan.split_method
        split_pipeline_point(model, plan)
        plan.split_method = method
    return PipeParser().parse(model, plan)


def build_stage_module_and_dependency(stages: Sequence[nn.Module], num_stages: int, virtual_chunks: int, stage_id: int,
                                      schedule_type: PipelineScheduleType = PipelineScheduleType.SIMPLE_1F1B):
    """``(modules of pipeline rank stage_id keyed by local chunk, StageDeps, p2p input mapping)`` (legacy
    ``pipe_stage.py:368``).  The mapping lists, per virtual stage, where its inputs come from: all outputs of the previous
    virtual stage, in order (stage 0 reads the micro-batch)."""
    from .schedule import StageDeps

    place = stage_placement(num_stages, virtual_chunks, schedule_type)
    mine = {c: stages[v] for v, (r, c) in enumerate(place) if r == stage_id}
    deps = StageDeps(len(place))
    mapping = {v: ([] if deps.prev(v) is None else [(deps.prev(v), None)]) for v in range(len(place))}
    return mine, deps, mapping


def _pp_coords(device_mesh, pp_rank, pp_group):
    if device_mesh is not None and pp_rank is None:
        names = device_mesh.mesh_dim_names or ()
        d = names.index("PP") if "PP" in names else 0
        pp_rank = device_mesh.get_local_rank(d)
        pp_group = device_mesh.get_group(d) if device_mesh.has_groups() else None
    return pp_rank or 0, pp_group


def construct_stage_modules(model: nn.Module, plan: PipelineParallelPlan, device_mesh=None, update_split_points: bool = False, *, pp_rank: Optional[int] = None):
    """Ingredients of a PipeModule for this rank: ``(stage modules, StageDeps, p2p input mapping)`` (legacy
    ``pipe_stage.py:249``)."""
    pp_rank, _ = _pp_coords(device_mesh, pp_rank, None)
    stages = construct_pipeline_split_graph(model, plan, update_split_points)
    return build_stage_module_and_dependency(stages, plan.num_stages, plan.virtual_chunks, pp_rank, plan.schedule_type)


def build_shared_module_group(pipe_module: "PipeModule", stages: Sequence[nn.Module], num_stages: int, virtual_chunks: int,
                              shared_module_path_groups, device_mesh=None, *, pp_group=None, units=None):
    """One process group per tie (``shared_module_path_groups``: lists of module / parameter FQNs whose values and gradients are
    kept in sync across stages, e.g. input embedding and LM head) over the pipeline ranks that own a member; appended to
    ``pipe_module.shared_groups`` (legacy ``pipe_stage.py:311``).  Collective over the PP group."""
    _, grp = _pp_coords(device_mesh, None, pp_group) if device_mesh is not None else (0, pp_group)
    pp_group = grp or pipe_module.pp_group
    if not shared_module_path_groups or pp_group is None:
        return pipe_module.shared_groups
    place = stage_placement(num_stages, virtual_chunks, pipe_module.plan.schedule_type)
    mine = {int(c): m for c, m in pipe_module.stage_modules.items()}
    for tie in shared_module_path_groups:
        owners = sorted({place[v][0] for v, st in enumerate(stages) for t in tie if any(t in n or n in t for n in _stage_param_fqns(st, units))})
        ranks = [dist.get_global_rank(pp_group, r) for r in owners] if len(owners) > 1 else None
        g = dist.new_group(ranks) if ranks else None
        if g is not None and pipe_module.pp_rank in owners:
            ps = [p for c, m in mine.items() for n, p in m.named_parameters() if any(_match_tie(n, m, t) for t in tie)]
            pipe_module.shared_groups.append((ps, g))
    return pipe_module.shared_groups


def construct_pipeline_stage(model: nn.Module, plan: PipelineParallelPlan, device_mesh=None, *, pp_rank: Optional[int] = None, pp_group=None, update_split_points: bool = False) -> PipeModule:
    """Raw model -> this rank's ``PipeModule`` (legacy ``pipe_stage.py:285``)."""
    pp_rank, pp_group = _pp_coords(device_mesh, pp_rank, pp_group)
    stages = construct_pipeline_split_graph(model, plan, update_split_points)
    mine, _, _ = build_stage_module_and_dependency(stages, plan.num_stages, plan.virtual_chunks, pp_rank, plan.schedule_type)
    place = stage_placement(plan.num_stages, plan.virtual_chunks, plan.schedule_type)
    vmap = {c: v for v, (r, c) in enumerate(place) if r == pp_rank}
    pm = PipeModule(mine, vmap, plan, pp_rank, pp_group)
    build_shared_module_group(pm, stages, plan.num_stages, plan.virtual_chunks, plan.shared_modules, pp_group=pp_group, units=_units(model))
    return pm


def _stage_param_fqns(stage: nn.Module, units) -> List[str]:
    if isinstance(stage, _StageSeq):
        return [f"{un}.{pn}" for un, m in zip(stage.names, stage.mods) for pn, _ in m.named_parameters()]
    return [n for n, _ in stage.named_parameters()]


def _match_tie(local_name: str, stage: nn.Module, tie_fqn: str) -> bool:
    if isinstance(stage, _StageSeq):
        parts = local_name.split(".", 2)  # mods.<i>.<rest>
        if len(parts) == 3 and parts[0] == "mods":
            return f"{stage.names[int(parts[1])]}.{parts[2]}" == tie_fqn
    return local_name == tie_fqn
