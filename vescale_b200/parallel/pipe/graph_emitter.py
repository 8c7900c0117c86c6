"""Graph ("compile") mode of the pipeline: one ``torch.fx`` program per rank with the p2p communication INSIDE the graph.

The reference's compile mode inserts functional send / recv nodes into the forward and backward fx graphs of a stage so that a
whole pipeline rank becomes one traceable graph (``legacy/vescale/pipe/_schedules/pp_collective_emitter.py:39-289``:
``PPCollectiveOpEmitter.insert_send_fwd / insert_recv_fwd / insert_send_bwd / insert_recv_bwd`` over torch's patched
``c10d_functional.send / recv``; the topology comes from the 1F1B instruction list).

Here the functional p2p ops are the custom ops of ``comm/functional_p2p.py`` (``vescale_b200::p2p_send / p2p_recv``), and they
are *differentiable*: the backward of a send receives the gradient from the destination and the backward of a receive sends
the gradient back.  So only the FORWARD program is emitted — autograd derives the backward communication from it, in exactly
the reverse order, and there are no separate backward graphs to keep consistent with the forward ones.

``PPCollectiveOpEmitter`` builds, for this rank, an ``fx.GraphModule`` whose nodes are, per micro-batch and local chunk::

    x   = placeholder                         (virtual stage 0)      |  p2p_recv(like, shape, src, group)   (fed by another rank)
    y   = call_module chunk_c(x...)                                   |  fed directly when the producer chunk is local (V placement)
    tok = p2p_send(y_k, dst, group)  per output                       |  loss = loss_fn(y, label_m)          (last virtual stage)

and whose outputs are the losses (last stage) and the send tokens (every other stage); ``GraphPipeProgram`` runs it and drives
the backward from those outputs, in exactly the reverse of the forward order.  With one chunk per rank the order is GPipe's
(all micro-batches forward, then backward): every message flows down the chain in the forward pass and up in the backward
pass, so blocking, tag-less, rendezvous-style p2p (NCCL) cannot dead-lock and the ranks overlap across micro-batches.  With
several chunks per rank the program walks micro-batch by micro-batch (a rank that waits for its own micro-batch to come back
around cannot have queued sends in front of that receive), which is safe on every back end but serial; the interpreted engine
(``engine.py``) is the fast path for interleaved / V schedules.  Shapes of the received tensors are static graph constants;
``infer_stage_meta`` finds them once by passing example values down the chain of stage owners.

``PipeEngine`` uses this path when ``plan.mode == ModeType.GRAPH_EAGER``; the graph is also what a capture-based back end
(CUDA graphs over NCCL p2p, ``torch.export``) consumes — every node is a module call or a functional custom op with a fake kernel.
"""
from __future__ import annotations

import operator
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.fx as fx
import torch.nn as nn

from ...comm import functional_p2p as fp2p
from .plan import PipelineParallelPlan
from .schedule import stage_placement

__all__ = ["PPCollectiveOpEmitter", "GraphPipeProgram", "infer_stage_meta", "read_fg"]

Meta = Tuple[Tuple[int, ...], torch.dtype]


class StageMeta(list):
    """``[(shape, dtype), ...]`` of a virtual stage's outputs; ``is_tuple``: the stage returns a tuple (even of one)."""

    is_tuple = False


class _Loss(nn.Module):
    def __init__(self, fn: Callable, M: int):
        super().__init__()
        self.fn, self.M = fn, M

    def forward(self, out, label):
        return self.fn(out, label) / self.M


def _as_tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x,)


def infer_stage_meta(pipe_module, plan: PipelineParallelPlan, pp_rank: int, pp_group, example_microbatch: Sequence[torch.Tensor], device) -> Dict[int, List[Meta]]:
    """``{virtual stage: [(shape, dtype) of each output]}`` for every virtual stage, agreed on by all pipeline ranks.  The owner of
    stage v runs it once (no grad) on the example micro-batch (v = 0) or on zeros shaped like stage v-1's outputs and publishes the
    result; one ``all_gather_object`` per virtual stage, at build time only."""
    P, V = plan.num_stages, plan.virtual_chunks
    place = stage_placement(P, V, plan.schedule_type)
    metas: Dict[int, List[Meta]] = {}
    world = dist.get_world_size(pp_group) if pp_group is not None or dist.is_initialized() else 1
    for v, (r, c) in enumerate(place):
        mine = None
        if r == pp_rank:
            xs = tuple(example_microbatch) if v == 0 else tuple(torch.zeros(s, dtype=d, device=device) for s, d in metas[v - 1])
            with torch.no_grad():
                raw = pipe_module(*xs, chunk_id=c)
            mine = ([(tuple(o.shape), o.dtype) for o in _as_tuple(raw)], isinstance(raw, (tuple, list)))
        if world > 1:
            got = [None] * world
            dist.all_gather_object(got, mine, group=pp_group)
            mine = next(g for g in got if g is not None)
        sm = StageMeta((tuple(s), d) for s, d in mine[0])
        sm.is_tuple = bool(mine[1])
        metas[v] = sm
    return metas


def read_fg(fg: fx.GraphModule):
    """(number of placeholders, number of outputs) of a captured graph — what an emitter needs to know to splice communication
    nodes around it (legacy ``pp_collective_emitter.py:35-43``).  A graph that returns a single value counts one output."""
    n_in, n_out = 0, None
    for node in fg.graph.nodes:
        if node.op == "placeholder":
            n_in += 1
        elif node.op == "output":
            ret = node.args[0]
            n_out = len(ret) if isinstance(ret, (tuple, list)) else 1
    return n_in, n_out


class PPCollectiveOpEmitter:
    """Emits this rank's forward program.  ``gen_pp_collective_topo`` exposes the peers per chunk the way the reference's emitter
    does (``fwd_recv_srcs / fwd_send_dsts``; the backward peers are the same lists mirrored, and are never needed explicitly)."""

    def __init__(self, pipe_module, plan: PipelineParallelPlan, pp_rank: int, pp_group=None, loss_fn: Optional[Callable] = None):
        self.pm, self.plan, self.rank, self.group, self.loss_fn = pipe_module, plan, pp_rank, pp_group, loss_fn
        self.P, self.V = plan.num_stages, plan.virtual_chunks
        self.NV = self.P * self.V
        self.place = stage_placement(self.P, self.V, plan.schedule_type)
        self.fwd_recv_srcs: Dict[int, Optional[int]] = {}
        self.fwd_send_dsts: Dict[int, Optional[int]] = {}
        self.gen_pp_collective_topo()

    def gen_pp_collective_topo(self):
        """Per local chunk: the rank its inputs come from / its outputs go to (None = the micro-batch / the loss, or a local chunk)."""
        for v, (r, c) in enumerate(self.place):
            if r != self.rank:
                continue
            src = self.place[v - 1][0] if v > 0 else None
            dst = self.place[v + 1][0] if v + 1 < self.NV else None
            self.fwd_recv_srcs[c] = None if src == self.rank else src
            self.fwd_send_dsts[c] = None if dst == self.rank else dst
        return {"fwd_recv_srcs": dict(self.fwd_recv_srcs), "fwd_send_dsts": dict(self.fwd_send_dsts),
                "bwd_recv_srcs": dict(self.fwd_send_dsts), "bwd_send_dsts": dict(self.fwd_recv_srcs)}

    def gen_pp_collective_topo_from_schedule_engine(self, engine_or_programs):
        """The same peer sets derived from what a schedule actually does rather than from the placement: walk this rank's instruction
        program (a generator / emitter with ``get_instruction_list``, or ``{rank: program}``) and collect every operator its
        communication instructions compile to (legacy ``pp_collective_emitter.py:57-101``)."""
        prog = engine_or_programs[self.rank] if isinstance(engine_or_programs, dict) else engine_or_programs.get_instruction_list(self.rank)
        from .instruction_base import CompilePPCollectiveKind

        sets = {"fwd_recv_srcs": set(), "fwd_send_dsts": set(), "bwd_recv_srcs": set(), "bwd_send_dsts": set()}
        for ins in prog:
            for op in ins.compile():
                send = op.kind is CompilePPCollectiveKind.SEND
                sets[("bwd_" if op.is_backward else "fwd_") + ("send_dsts" if send else "recv_srcs")].add(op.dst if send else op.src)
        return {k: sorted(v) for k, v in sets.items()}

    # ------------------------------------------------------------------ emission
    def emit(self, num_microbatches: int, metas: Dict[int, List[Meta]], device, n_inputs: int = 1, with_labels: bool = True) -> "GraphPipeProgram":
        M = num_microbatches
        root = nn.Module()
        g = fx.Graph()
        gid = fp2p._gid(self.group)
        my_v = sorted(v for v, (r, _) in enumerate(self.place) if r == self.rank)
        owns_first, owns_last = 0 in my_v, (self.NV - 1) in my_v
        for v in my_v:
            root.add_module(f"chunk_{self.place[v][1]}", self.pm.chunk(self.place[v][1]))
        # placeholders first (fx requires it): micro-batch inputs on the first stage, labels on the last
        x_ph = [[g.placeholder(f"x{m}_{k}") for k in range(n_inputs)] for m in range(M)] if owns_first else None
        y_ph = [g.placeholder(f"label{m}") for m in range(M)] if (owns_last and with_labels and self.loss_fn is not None) else None
        likes: Dict[torch.dtype, fx.Node] = {}

        def like_node(dtype: torch.dtype) -> fx.Node:
            if dtype not in likes:
                name = f"like_{str(dtype).replace('torch.', '')}"
                # zero-element anchor: its dtype / device type the received tensor; requires_grad so that the receive takes part in autograd
                t = torch.empty(0, dtype=dtype, device=device)
                # (a buffer, not a parameter: optimizers and checkpoints must not see it)
                root.register_buffer(name, t.requires_grad_(dtype.is_floating_point or dtype.is_complex), persistent=False)
                likes[dtype] = g.get_attr(name)
            return likes[dtype]

        if y_ph is not None:
            root.add_module("loss", _Loss(self.loss_fn, M))
        produced: Dict[Tuple[int, int], List[fx.Node]] = {}  # (m, v) -> output nodes, for chunks fed locally
        tokens: Dict[Tuple[int, int], List[fx.Node]] = {}
        losses: List[Optional[fx.Node]] = [None] * M
        finals: List[Optional[fx.Node]] = [None] * M
        # GPipe order with one chunk per rank; micro-batch-major with several (see the module docstring)
        order = [(m, v) for v in my_v for m in range(M)] if len(my_v) == 1 else [(m, v) for m in range(M) for v in my_v]
        stamp = "inserted by graph_emitter.PPCollectiveOpEmitter"
        for m, v in order:
            c = self.place[v][1]
            if v == 0:
                ins = list(x_ph[m])
            elif self.place[v - 1][0] == self.rank:
                ins = produced.pop((m, v - 1))
            else:
                src = self.place[v - 1][0]
                ins = []
                for k, (shape, dtype) in enumerate(metas[v - 1]):
                    n = g.create_node("call_function", torch.ops.vescale_b200.p2p_recv.default, (like_node(dtype), list(shape), src, gid), {}, name=f"pp_recv_fwd_m{m}_v{v}_{k}")
                    n.meta["stack_trace"] = stamp
                    ins.append(n)
            out = g.create_node("call_module", f"chunk_{c}", tuple(ins), {}, name=f"stage_m{m}_v{v}")
            outs = [g.call_function(operator.getitem, (out, k)) for k in range(len(metas[v]))] if metas[v].is_tuple else [out]
            if v == self.NV - 1:
                finals[m] = out
                if y_ph is not None:
                    losses[m] = g.create_node("call_module", "loss", (out, y_ph[m]), {}, name=f"loss_m{m}")
            elif self.place[v + 1][0] == self.rank:
                produced[(m, v)] = outs
            else:
                dst = self.place[v + 1][0]
                for k, o in enumerate(outs):
                    t = g.create_node("call_function", torch.ops.vescale_b200.p2p_send.default, (o, dst, gid), {}, name=f"pp_send_fwd_m{m}_v{v}_{k}")
                    t.meta["stack_trace"] = stamp
                    tokens.setdefault((m, v), []).append(t)
        # outputs in program order, so that the backward can be driven in exactly the reverse order
        steps = []
        for m, v in order:
            if v == self.NV - 1 and losses[m] is not None:
                steps.append(losses[m])
            elif (m, v) in tokens:
                steps.append(tokens[(m, v)])
            else:
                steps.append(None)  # fed a local chunk: its backward is reached through the consumer's
        g.output({"steps": steps, "losses": [l for l in losses if l is not None], "outputs": [f for f in finals if f is not None]})
        gm = fx.GraphModule(root, g, class_name=f"PipeRankProgram{self.rank}")
        return GraphPipeProgram(gm, self, M, owns_first, y_ph is not None, [("F", m, v) for m, v in order])


class GraphPipeProgram:
    """A rank's emitted program: ``forward(microbatches, labels)`` runs the graph (communication included), ``backward()`` walks the
    micro-batches in reverse — the functional p2p ops carry the gradients across ranks."""

    def __init__(self, gm: fx.GraphModule, emitter: PPCollectiveOpEmitter, M: int, takes_inputs: bool, takes_labels: bool, order):
        self.gm, self.emitter, self.M, self.takes_inputs, self.takes_labels, self.order = gm, emitter, M, takes_inputs, takes_labels, order
        self._last = None

    @property
    def graph(self) -> fx.Graph:
        return self.gm.graph

    def comm_nodes(self) -> List[fx.Node]:
        return [n for n in self.gm.graph.nodes if n.op == "call_function" and n.target in (torch.ops.vescale_b200.p2p_send.default, torch.ops.vescale_b200.p2p_recv.default)]

    def forward(self, microbatches: Optional[Sequence] = None, labels: Optional[Sequence] = None):
        args: List = []
        if self.takes_inputs:
            for mb in microbatches:
                args += list(_as_tuple(mb))
        if self.takes_labels:
            args += list(labels)
        self._last = self.gm(*args)
        return self._last

    def backward(self) -> None:
        res = self._last
        assert res is not None, "forward() first"
        for step in reversed(res["steps"]):
            if step is None:
                continue
            if isinstance(step, torch.Tensor):  # a loss
                step.backward()
                continue
            toks = [t for t in step if t.requires_grad]
            if toks:
                torch.autograd.backward(toks, [torch.zeros_like(t) for t in toks])
        self._last = None

    def run(self, microbatches=None, labels=None, forward_only: bool = False):
        with torch.set_grad_enabled(not forward_only):
            res = self.forward(microbatches, labels)
        losses = [l.detach() for l in res["losses"]]
        outs = list(res["outputs"])
        if not forward_only:
            self.backward()
        return losses, outs
