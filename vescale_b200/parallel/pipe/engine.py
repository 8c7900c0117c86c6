"""PipeEngine: run one mini-batch through the pipeline schedule.

    engine = PipeEngine(pipe_module, mesh, loss_fn, plan)
    loss, outputs = engine(minibatch_inputs, minibatch_labels)      # list of micro-batches or a tensor to chunk

The executor interprets this rank's instruction row (``schedule.build_schedule``): F / B / W with p2p receives
right before and sends right after each op (``p2p.P2PContext``).  Backward of a chunk is
``torch.autograd.backward`` on the stage outputs; when the schedule splits weight gradients (zero-bubble), B
takes only the input gradients (``autograd.grad(..., inputs=stage_inputs)``) and W later produces the
parameter gradients from the retained graph.  ``VESCALE_DUMP_INSTRUCTION=1`` writes the per-rank instruction
files (``legacy/vescale/dtensor/_diff.py:25-71``).

Parity: ``legacy/vescale/engine/pipe.py:33-237``, ``pipe/pipe_emmiter.py:132-343`` (ScheduleEngine).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from ...profiler import ndtimeit, predefined
from .p2p import P2PContext
from .plan import ModeType, PipelineParallelPlan, PipelineScheduleType
from .schedule import INSTRUCTION_REGISTRY, Instr, build_schedule, stage_placement, validate_pipeline_schedule
from .stage import PipeModule

__all__ = ["PipeEngine", "ScheduleEngine", "PipelineEmitter"]


def _as_tuple(x):
    if isinstance(x, tuple):
        return x
    if isinstance(x, list):
        return tuple(x)
    return (x,)


class ScheduleEngine:
    """Interprets this rank's row of the global instruction schedule: receives right before, sends right after each F / B / W
    instruction (legacy ``pipe/pipe_emmiter.py:132-343``)."""
    def __init__(self, module: PipeModule, plan: PipelineParallelPlan, pp_rank: int, pp_group, loss_fn: Optional[Callable], device):
        self.module, self.plan, self.rank, self.group, self.loss_fn, self.device = module, plan, pp_rank, pp_group, loss_fn, device
        self.P, self.V = plan.num_stages, plan.virtual_chunks
        self.NV = self.P * self.V
        self.place = stage_placement(self.P, self.V, plan.schedule_type)
        self.split_w = plan.schedule_type in (PipelineScheduleType.ZERO_BUBBLE, PipelineScheduleType.ZERO_BUBBLE_V)
        self._sched_cache: Dict[int, List[List[Instr]]] = {}
        self._shape_cache: Dict = {}

    def schedule(self, M: int) -> List[List[Instr]]:
        if M not in self._sched_cache:
            self._sched_cache[M] = build_schedule(self.plan, M)
            if os.environ.get("VESCALE_DUMP_INSTRUCTION", "0") == "1":
                with open(f"pipe_instructions_rank{self.rank}.txt", "w") as f:
                    f.write("\n".join(repr(i) for i in self._sched_cache[M][self.rank]))
        return self._sched_cache[M]

    def _pair_ops(self, rows: List[List[Instr]]) -> Dict[int, List[Tuple[str, Tuple[str, int, int]]]]:
        """For every peer: all messages between us, both directions, in *pair order* — sorted by the simulated time at which the
        producing instruction ends (ties: lower sender rank first).  Both ends compute the same list (see ``p2p.py``)."""
        msgs: Dict[int, List] = {}
        for r, row in enumerate(rows):
            for ins in row:
                if ins.kind == "F" and ins.vstage + 1 < self.NV:
                    dst = self.place[ins.vstage + 1][0]
                elif ins.kind == "B" and ins.vstage > 0:
                    dst = self.place[ins.vstage - 1][0]
                else:
                    continue
                if dst == r or self.rank not in (r, dst):
                    continue
                peer = dst if r == self.rank else r
                msgs.setdefault(peer, []).append((ins.end, r, ins.kind, ins.microbatch, ins.vstage))
        out = {}
        for peer, lst in msgs.items():
            lst.sort()
            out[peer] = [("S" if src == self.rank else "R", (kind, m, v)) for _, src, kind, m, v in lst]
        return out

    def execute(self, inputs: Sequence, labels: Optional[Sequence], forward_only: bool = False):
        M = len(inputs)
        rows = self.schedule(M)
        p2p = P2PContext(self.group, self.rank, self._pair_ops(rows), self.device, self.plan.p2p_tensor_dtype, self.plan.reuse_p2p_tensor_shape,
                         prefetch=getattr(self.plan, "overlap_p2p_comm", True), batch=getattr(self.plan, "batch_p2p_comm", True))
        p2p.shapes = self._shape_cache
        self.last_p2p = p2p
        acts: Dict[Tuple[int, int], Tuple] = {}  # (m, v) -> (stage_inputs, stage_outputs)
        local_fwd: Dict[Tuple[int, int], Tuple] = {}
        local_bwd: Dict[Tuple[int, int], Tuple] = {}
        pending_w: Dict[Tuple[int, int], Tuple] = {}
        losses: List[Optional[torch.Tensor]] = [None] * M
        outputs: List = [None] * M
        timing = self._timing
        for ins in rows[self.rank]:
            h = INSTRUCTION_REGISTRY.get(ins.kind)
            if h is not None and h(self, ins) is not None:
                continue
            m, v, c = ins.microbatch, ins.vstage, ins.chunk
            if timing is not None:
                self._tick(timing, ins.kind)
            if ins.kind == "F":
                if v == 0:
                    xs = _as_tuple(inputs[m])
                elif self.place[v - 1][0] == self.rank:
                    xs = local_fwd.pop((m, v - 1))
                else:
                    xs = p2p.recv(("F", m, v - 1), self.place[v - 1][0])
                xs = tuple(x.detach().requires_grad_(x.is_floating_point() and not forward_only) if isinstance(x, torch.Tensor) else x for x in xs)
                with torch.set_grad_enabled(not forward_only), ndtimeit(predefined.FORWARD_COMPUTE, microbatch=m, vstage=v):
                    out = self.module(*xs, chunk_id=c)
                    outs = _as_tuple(out)
                    if v == self.NV - 1:
                        outputs[m] = out
                        if self.loss_fn is not None and labels is not None:
                            loss = self.loss_fn(out, labels[m]) / M
                            losses[m] = loss.detach()
                            outs = (loss,)
                acts[(m, v)] = (xs, outs)
                if v + 1 < self.NV:
                    if self.place[v + 1][0] == self.rank:
                        local_fwd[(m, v)] = tuple(o.detach() for o in outs)
                    else:
                        p2p.send(("F", m, v), outs, self.place[v + 1][0])
            elif ins.kind == "B":
                xs, outs = acts[(m, v)]
                if v == self.NV - 1:
                    gouts = tuple(torch.ones_like(o) for o in outs)
                elif self.place[v + 1][0] == self.rank:
                    gouts = local_bwd.pop((m, v + 1))
                else:
                    gouts = p2p.recv(("B", m, v + 1), self.place[v + 1][0])
                pairs = [(o, g) for o, g in zip(outs, gouts) if isinstance(o, torch.Tensor) and o.requires_grad]
                o_t, g_t = [p[0] for p in pairs], [p[1] for p in pairs]
                grad_inputs = [x for x in xs if isinstance(x, torch.Tensor) and x.requires_grad]
                with ndtimeit(predefined.BACKWARD_COMPUTE, microbatch=m, vstage=v):
                    if self.split_w:
                        gi = torch.autograd.grad(o_t, grad_inputs, g_t, retain_graph=True, allow_unused=True) if grad_inputs else ()
                        pending_w[(m, v)] = (o_t, g_t)
                    else:
                        torch.autograd.backward(o_t, g_t)
                        gi = tuple(x.grad for x in grad_inputs)
                        del acts[(m, v)]
                p2p.drain_sends(keep=2 * self.V)  # retire old sends so their buffers are freed while the pipeline keeps running
                if v > 0:
                    gi = tuple(g if g is not None else torch.zeros_like(x) for g, x in zip(gi, grad_inputs))
                    if self.place[v - 1][0] == self.rank:
                        local_bwd[(m, v)] = gi
                    else:
                        p2p.send(("B", m, v), gi, self.place[v - 1][0])
            elif ins.kind == "W":
                o_t, g_t = pending_w.pop((m, v))
                params = [p for p in self.module.chunk(c).parameters() if p.requires_grad]
                with ndtimeit(predefined.BACKWARD_COMPUTE, microbatch=m, vstage=v, part="weight-grad"):
                    gs = torch.autograd.grad(o_t, params, g_t, allow_unused=True)
                for p, g in zip(params, gs):
                    if g is not None:
                        p.grad = g if p.grad is None else p.grad + g
                del acts[(m, v)]
        if timing is not None:
            self._tick(timing, None)
        p2p.drain()
        return losses, outputs

    # ---- cost calibration (feeds the cost-driven schedule search, auto_schedule.py)
    _timing: Optional[Dict] = None

    def _tick(self, timing: Dict, kind: Optional[str]) -> None:
        """Close the running instruction's interval and open one for ``kind``.  Intervals include the p2p waits in front of the
        compute of an instruction on purpose: a schedule is better when what it measures shrinks."""
        import time

        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        now = time.perf_counter()
        if timing.get("open") is not None:
            k0, t0 = timing["open"]
            timing.setdefault(k0, []).append(now - t0)
        timing["open"] = None if kind is None else (kind, now)


class PipelineEmitter:
    """Schedule type -> the generator that writes its instruction programs (legacy ``pipe/pipe_emmiter.py:43-129``).  1F1B, interleaved
    1F1B and ZB-V have explicit programs (``_schedules/{pipedream_flush,looping_bfs,zero_bubble_v}.py``); GPipe and ZB-H1 expose the
    list scheduler's rows.  ``meshes``: one entry per pipeline stage."""

    def __init__(self, deps, meshes: Sequence, schedule, batches: int, tensor_shape=None, dtype: Optional[torch.dtype] = None, num_chunks: int = 1, input_shapes=None,
                 input_shapes_unpad=None, forward_only: bool = False, overlap_p2p_comm: bool = False, batch_p2p_comm: bool = True, param_sync_overlap: bool = False,
                 grad_sync_overlap: bool = False, **kwargs):
        from . import _schedules as S

        schedule = PipelineScheduleType(schedule) if not isinstance(schedule, PipelineScheduleType) else schedule
        self.deps, self.meshes, self.batches, self.num_chunks, self.forward_only = deps, meshes, batches, num_chunks, forward_only
        self.num_stage = len(meshes)
        self.overlap_p2p_comm, self.batch_p2p_comm, self.param_sync_overlap, self.grad_sync_overlap = overlap_p2p_comm, batch_p2p_comm, param_sync_overlap, grad_sync_overlap
        common = dict(deps=deps, meshes=meshes, batches=batches, default_shape=tensor_shape, default_dtype=dtype)
        if schedule == PipelineScheduleType.SIMPLE_1F1B:
            self.instruction_generator = S.OneFOneBInstrcutionGenerator(forward_only=forward_only, batch_p2p_comm=batch_p2p_comm, overlap_p2p_comm=overlap_p2p_comm, **common)
        elif schedule == PipelineScheduleType.INTERLEAVED_1F1B:
            self.instruction_generator = S.InterleavedOneFOneBInstructionGenerator(forward_only=forward_only, num_chunk=max(2, num_chunks), batch_shape_lists=input_shapes,
                                                                                   batch_p2p_comm=batch_p2p_comm, overlap_p2p_comm=overlap_p2p_comm, **common)
        elif schedule in (PipelineScheduleType.ZERO_BUBBLE_V,):
            self.instruction_generator = S.ZeroBubbleVInstrcutionGenerator(**common, **kwargs)
        elif schedule == PipelineScheduleType.ZERO_BUBBLE:
            self.instruction_generator = S.ZeroBubbleInstructionGenerator(forward_only=forward_only, num_chunk=num_chunks, **common)
        elif schedule == PipelineScheduleType.GPIPE:
            self.instruction_generator = S.GPipeInstructionGenerator(forward_only=forward_only, num_chunk=num_chunks, **common)
        else:
            raise NotImplementedError(f"unsupported schedule type {schedule}")
        self.instruction_list = self.gen_instruction()

    def gen_instruction(self):
        return self.instruction_generator.gen_instruction()

    def get_instruction_list(self, stage: int):
        return self.instruction_generator.get_instruction_list(stage)


class PipeEngine:
    """``engine(minibatch, labels) -> (loss, outputs)``: runs one pipeline-parallel mini-batch under the plan's schedule (legacy
    ``engine/pipe.py:33-237``)."""
    def __init__(self, module: PipeModule, global_mesh=None, loss_fn: Optional[Callable] = None, plan: Optional[PipelineParallelPlan] = None, *, pp_group=None, pp_rank: Optional[int] = None, device=None):
        self.module = module
        self.plan = plan or module.plan
        validate_pipeline_schedule(self.plan)
        self.loss_fn = loss_fn
        if global_mesh is not None and pp_rank is None:
            names = global_mesh.mesh_dim_names or ()
            d = names.index("PP") if "PP" in names else 0
            pp_rank = global_mesh.get_local_rank(d)
            pp_group = global_mesh.get_group(d) if global_mesh.has_groups() else None
            device = device or (torch.device("cuda", torch.cuda.current_device()) if global_mesh.device_type == "cuda" else torch.device("cpu"))
        self.pp_rank = pp_rank if pp_rank is not None else module.pp_rank
        self.pp_group = pp_group if pp_group is not None else module.pp_group
        self.device = device or torch.device("cpu")
        self.schedule_engine = ScheduleEngine(module, self.plan, self.pp_rank, self.pp_group, loss_fn, self.device)
        os.environ["STAGE_ID"] = str(self.pp_rank)
        if self.plan.reuse_p2p_tensor_shape:
            os.environ["REUSE_COMM_SHAPE"] = "1"

    @property
    def is_last_rank(self) -> bool:
        place = self.schedule_engine.place
        return place[-1][0] == self.pp_rank

    def forward_backward(self, minibatch, labels=None, forward_only: bool = False, num_microbatches: Optional[int] = None):
        if isinstance(minibatch, torch.Tensor):
            n = num_microbatches or self.plan.num_stages
            minibatch = list(minibatch.chunk(n))
            labels = list(labels.chunk(n)) if isinstance(labels, torch.Tensor) else labels
        if self.plan.mode == ModeType.GRAPH_EAGER:
            losses, outputs = self._run_graph_program(minibatch, labels, forward_only or self.plan.forward_only)
        else:
            losses, outputs = self.schedule_engine.execute(minibatch, labels, forward_only or self.plan.forward_only)
        loss = None
        if any(l is not None for l in losses):
            loss = torch.stack([l for l in losses if l is not None]).sum()
        return loss, outputs

    __call__ = forward_backward

    # ---- graph mode (plan.mode = GRAPH_EAGER): this rank's emitted fx program, p2p inside the graph (graph_emitter.py)
    def graph_program(self, minibatch, with_labels: bool = True):
        """The emitted program for ``len(minibatch)`` micro-batches (built once per micro-batch count and label use)."""
        from .graph_emitter import PPCollectiveOpEmitter, infer_stage_meta

        key = (len(minibatch), bool(with_labels))
        cache = self.__dict__.setdefault("_graph_programs", {})
        if key not in cache:
            ex = minibatch[0] if isinstance(minibatch[0], (tuple, list)) else (minibatch[0],)
            if "_graph_metas" not in self.__dict__:
                self._graph_metas = infer_stage_meta(self.module, self.plan, self.pp_rank, self.pp_group, ex, self.device)
            em = PPCollectiveOpEmitter(self.module, self.plan, self.pp_rank, self.pp_group, self.loss_fn)
            cache[key] = em.emit(len(minibatch), self._graph_metas, self.device, n_inputs=len(ex), with_labels=with_labels)
        return cache[key]

    def _run_graph_program(self, minibatch, labels, forward_only: bool):
        prog = self.graph_program(minibatch, with_labels=labels is not None and self.loss_fn is not None)
        losses, outs = prog.run(minibatch if prog.takes_inputs else None, labels if prog.takes_labels else None, forward_only=forward_only)
        M = len(minibatch)
        return (losses + [None] * (M - len(losses))), (outs + [None] * (M - len(outs)))

    def calibrate(self, minibatch, labels=None, num_microbatches: Optional[int] = None, comm: Optional[float] = None, warmup: int = 1, search: bool = True):
        """Measure F / B / W on this pipeline (median per instruction kind, MAX over the pipeline ranks, normalised so that F = 1),
        store them in ``plan.costs`` and — with ``search`` — switch the plan to the cost-driven schedule search for the following
        mini-batches (``plan.auto_schedule``; legacy ``zero_bubble_v.py:198-600`` takes the same four costs as constructor arguments
        and leaves measuring them to the user).  ``comm`` (same unit: multiples of one F) is taken as given; gradients produced by
        the calibration mini-batches are discarded.  Returns the cost dict."""
        import statistics

        import torch.distributed as dist

        se = self.schedule_engine
        for _ in range(max(0, warmup)):
            self.forward_backward(minibatch, labels, num_microbatches=num_microbatches)
        se._timing = {}
        try:
            self.forward_backward(minibatch, labels, num_microbatches=num_microbatches)
            t = se._timing
        finally:
            se._timing = None
        self.zero_grad()
        med = torch.tensor([statistics.median(t[k]) if t.get(k) else 0.0 for k in ("F", "B", "W")], dtype=torch.float64)
        if self.pp_group is not None and dist.is_initialized():
            med = med.to(self.device) if dist.get_backend(self.pp_group) == "nccl" else med
            dist.all_reduce(med, op=dist.ReduceOp.MAX, group=self.pp_group)
            med = med.cpu()
        f = float(med[0]) or 1.0
        costs = {"F": 1.0, "B": float(med[1]) / f, "W": float(med[2]) / f if se.split_w else 0.0, "comm": float(comm if comm is not None else self.plan.costs.get("comm", 0.0))}
        if not se.split_w:
            costs["B"], costs["W"] = costs["B"], 0.0  # un-split backward: all of it is measured as B
        self.plan.costs = costs
        self.plan.auto_schedule = bool(search)
        se._sched_cache.clear()
        self.measured_seconds = {"F": float(med[0]), "B": float(med[1]), "W": float(med[2])}
        return costs

    def sync_shared_params(self, share_params: bool = True):
        self.module.sync_shared_params(share_params)

    def parameters(self):
        return self.module.parameters()

    def zero_grad(self, set_to_none: bool = True):
        for p in self.module.parameters():
            p.grad = None
