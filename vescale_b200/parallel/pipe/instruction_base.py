"""Pipeline schedules lowered to explicit per-rank *instruction programs* and a small VM that runs them (legacy
``pipe/_schedules/instruction_base.py`` — ``BaseInstruction``, ``PipelineSchema``, ``InstructionBuilder``, ``CommPacket``,
``Status`` — and the instruction sets of ``pipedream_flush.py`` / ``looping_bfs.py`` / ``zero_bubble_v.py``).

``PipeEngine`` interprets the F / B / W rows of the list scheduler directly.  This module is the other classic representation:
communication is spelled out as instructions of its own —

    RECV_FORWARD(m, v)  FORWARD_STEP(m, v)  SEND_FORWARD(m, v)  RECV_BACKWARD ... BACKWARD_STEP ... SEND_BACKWARD ...
    WEIGHT_GRAD_STEP (zero-bubble)   DRAIN_SEND_REQS   DEALLOCATE_OUTPUT_TENSOR

so that a program can be printed, diffed between ranks, edited (insert a user instruction with ``register_instruction``) and
executed.  ``PipelineSchema`` turns a plan into the time x stage status grid, ``InstructionBuilder`` lowers one stage's row of
the grid into a program (with the peephole fusion of a send and the receive that follows it towards the same neighbour) and
``InstructionVM`` executes it on a ``PipeModule``.

Communication: each directed pair of ranks is one *stream*; a sender emits messages in program order and the receiver consumes
the stream in that same order, stashing messages that arrive before they are needed (``StageLink``).  NCCL matches p2p calls
per pair in issue order (it has no tags), so this is the only order both ends can agree on without negotiating — and since
every message a receiver has to skip past was sent before the one it wants, waiting for them cannot deadlock.  Sends never
block (requests are parked and drained by ``DRAIN_SEND_REQS``)."""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist

from ...profiler import ndtimeit, ndtimeit_p2p, predefined
from .plan import PipelineParallelPlan, PipelineScheduleType
from .schedule import INSTRUCTION_REGISTRY, Instr, StageDeps, build_schedule, register_instruction, stage_placement

Shape = Union[List[int], torch.Size]  # a p2p tensor shape as schedules pass it around

__all__ = ["Shape", "register_instruction", "StageDeps", "Status", "CommPacket", "BaseInstruction", "PipelineSchema", "InstructionBuilder", "InstructionVM", "StageLink", "INSTRUCTION_SET", "get_linear_pp_module_dep2",
           "CompilePPCollectiveKind", "CompilePPCollectiveOperator", "VESCALE_INTRUCTION_BUILDER", "switch_dtensor", "registed_functions", "InstructionGenerator",
           "RECV_FORWARD", "RECV_BACKWARD", "SEND_FORWARD", "SEND_BACKWARD", "SEND_FORWARD_RECV_BACKWARD", "SEND_BACKWARD_RECV_FORWARD", "FORWARD_STEP", "BACKWARD_STEP",
           "WEIGHT_GRAD_STEP", "DRAIN_SEND_REQS", "DEALLOCATE_OUTPUT_TENSOR", "BUBBLE"]


class Status(enum.Enum):
    """What a stage does in one slot of the schedule grid."""
    FORWARD = "F"
    BACKWARD = "B"
    WEIGHT = "W"
    BUBBLE = "."


@dataclass(frozen=True)
class CommPacket:
    """One tensor bundle in flight: who produces it, who consumes it, under which key.  The optional fields describe a stage boundary
    whose two sides live on different (tensor-parallel) meshes (legacy ``instruction_base.py:75-83``): which input slot of the consumer it
    feeds, and the placements on either side — ``cross_mesh_send`` / ``cross_mesh_recv`` (``_schedules/common.py``) use them."""
    kind: str = "F"  # "F" activation / "B" gradient
    microbatch: int = -1
    vstage: int = -1  # virtual stage that PRODUCED it
    src: int = -1
    dst: int = -1
    cur_mesh: object = None
    peer_mesh: object = None
    input_id: int = 0
    peer_stage: int = -1
    peer_sharding: Optional[tuple] = None
    cur_sharding: Optional[tuple] = None
    is_kwargs: bool = False

    @property
    def key(self) -> Tuple[str, int, int]:
        return (self.kind, self.microbatch, self.vstage)


def get_linear_pp_module_dep2(module_list: Sequence, device_mesh_list: Sequence) -> StageDeps:
    """Dependency table of a plain chain of stage modules (stage i feeds stage i + 1)."""
    assert len(module_list) >= 1 and len(device_mesh_list) >= 1
    return StageDeps(len(module_list))


# ---- instructions ------------------------------------------------------------------------------------------------------------------------
INSTRUCTION_SET: Dict[str, type] = {}


@dataclass
class BaseInstruction:
    """One step of a rank's program.  ``run(vm)`` does the work; ``name`` identifies it in dumps and in the user registry
    (``register_instruction(name)`` handlers run INSTEAD of ``run`` when they return something other than ``None``)."""
    microbatch: int = -1
    vstage: int = -1
    chunk: int = 0
    peer: int = -1

    name = "BASE"

    handler = None  # class-level: name of the registered function that does the work (schedule-specific instruction sets)

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        INSTRUCTION_SET.setdefault(cls.name, cls)  # the first definition of a name is the canonical one (schedule modules refine, not replace)

    def run(self, vm: "InstructionVM") -> None:
        """Schedule-specific instruction sets name the registered function that implements them (``handler``): re-registering that
        name (``register_instruction("vescale_1f1b_forward_step")``) swaps the behaviour for every program of that schedule."""
        h = type(self).handler
        if h is None:
            raise NotImplementedError
        return INSTRUCTION_REGISTRY[h](vm, self)

    def compile(self) -> List["CompilePPCollectiveOperator"]:
        """The p2p operators this instruction stands for, as the graph emitter wants them (legacy ``BaseInstruction.compile``)."""
        out = []
        for kind, m, v, peer, is_send in _wire_ops(self):
            out.append(CompilePPCollectiveOperator(CompilePPCollectiveKind.SEND if is_send else CompilePPCollectiveKind.RECV, src=None if is_send else peer, dst=peer if is_send else None, is_backward=(kind == "B")))
        return out

    def dump(self) -> str:
        where = f" peer={self.peer}" if self.peer >= 0 else ""
        return f"{self.name}(mb={self.microbatch}, v={self.vstage}, chunk={self.chunk}{where})" if self.microbatch >= 0 else self.name

    __str__ = dump


@dataclass
class RECV_FORWARD(BaseInstruction):  # noqa: N801
    name = "RECV_FORWARD"

    def run(self, vm):
        with ndtimeit_p2p(predefined.RECV_FORWARD, peer=self.peer):
            vm.inbox_f[(self.microbatch, self.vstage)] = vm.link.recv(("F", self.microbatch, self.vstage - 1), self.peer)


@dataclass
class RECV_BACKWARD(BaseInstruction):  # noqa: N801
    name = "RECV_BACKWARD"

    def run(self, vm):
        with ndtimeit_p2p(predefined.RECV_BACKWARD, peer=self.peer):
            vm.inbox_b[(self.microbatch, self.vstage)] = vm.link.recv(("B", self.microbatch, self.vstage + 1), self.peer)


@dataclass
class SEND_FORWARD(BaseInstruction):  # noqa: N801
    name = "SEND_FORWARD"

    def run(self, vm):
        with ndtimeit_p2p(predefined.SEND_FORWARD, peer=self.peer):
            vm.link.send(("F", self.microbatch, self.vstage), vm.outbox_f.pop((self.microbatch, self.vstage)), self.peer)


@dataclass
class SEND_BACKWARD(BaseInstruction):  # noqa: N801
    name = "SEND_BACKWARD"

    def run(self, vm):
        with ndtimeit_p2p(predefined.SEND_BACKWARD, peer=self.peer):
            vm.link.send(("B", self.microbatch, self.vstage), vm.outbox_b.pop((self.microbatch, self.vstage)), self.peer)


@dataclass
class SEND_FORWARD_RECV_BACKWARD(BaseInstruction):  # noqa: N801
    """Fused steady-state pair towards the next stage: activation of ``microbatch`` out, gradient of ``recv_microbatch`` in."""
    recv_microbatch: int = -1
    recv_vstage: int = -1
    name = "SEND_FORWARD_RECV_BACKWARD"

    def run(self, vm):
        with ndtimeit_p2p(predefined.SEND_FORWARD_RECV_BACKWARD, peer=self.peer):
            vm.link.send(("F", self.microbatch, self.vstage), vm.outbox_f.pop((self.microbatch, self.vstage)), self.peer)
            vm.inbox_b[(self.recv_microbatch, self.recv_vstage)] = vm.link.recv(("B", self.recv_microbatch, self.recv_vstage + 1), self.peer)

    def dump(self):
        return f"{self.name}(send mb={self.microbatch} v={self.vstage}, recv mb={self.recv_microbatch} v={self.recv_vstage}, peer={self.peer})"


@dataclass
class SEND_BACKWARD_RECV_FORWARD(BaseInstruction):  # noqa: N801
    recv_microbatch: int = -1
    recv_vstage: int = -1
    name = "SEND_BACKWARD_RECV_FORWARD"

    def run(self, vm):
        with ndtimeit_p2p(predefined.SEND_BACKWARD_RECV_FORWARD, peer=self.peer):
            vm.link.send(("B", self.microbatch, self.vstage), vm.outbox_b.pop((self.microbatch, self.vstage)), self.peer)
            vm.inbox_f[(self.recv_microbatch, self.recv_vstage)] = vm.link.recv(("F", self.recv_microbatch, self.recv_vstage - 1), self.peer)

    def dump(self):
        return f"{self.name}(send mb={self.microbatch} v={self.vstage}, recv mb={self.recv_microbatch} v={self.recv_vstage}, peer={self.peer})"


@dataclass
class FORWARD_STEP(BaseInstruction):  # noqa: N801
    name = "FORWARD_STEP"

    def run(self, vm):
        vm.forward_step(self.microbatch, self.vstage, self.chunk)


@dataclass
class BACKWARD_STEP(BaseInstruction):  # noqa: N801
    name = "BACKWARD_STEP"

    def run(self, vm):
        vm.backward_step(self.microbatch, self.vstage, self.chunk)


@dataclass
class WEIGHT_GRAD_STEP(BaseInstruction):  # noqa: N801
    name = "WEIGHT_GRAD_STEP"

    def run(self, vm):
        vm.weight_step(self.microbatch, self.vstage, self.chunk)


@dataclass
class DRAIN_SEND_REQS(BaseInstruction):  # noqa: N801
    keep: int = 0
    name = "DRAIN_SEND_REQS"

    def run(self, vm):
        vm.link.drain_sends(self.keep)


@dataclass
class DEALLOCATE_OUTPUT_TENSOR(BaseInstruction):  # noqa: N801
    """The output of (microbatch, vstage) has been sent: keep its autograd graph, free its storage (the data is only needed by
    the consumer stage; backward needs ``grad_fn`` only).  Skipped for tensors that are views or are still referenced as inputs."""
    name = "DEALLOCATE_OUTPUT_TENSOR"

    def run(self, vm):
        vm.deallocate_output(self.microbatch, self.vstage)


@dataclass
class BUBBLE(BaseInstruction):  # noqa: N801
    name = "BUBBLE"

    def run(self, vm):
        pass


# ---- p2p operators for graph mode ----------------------------------------------------------------------------------------------------------
class CompilePPCollectiveKind(enum.Enum):
    SEND = 1
    RECV = 2
    BORADCAST = 3  # (sic, the reference's spelling) one source, several destinations: a stage boundary that fans out across meshes
    UNKNOWN = 4


class CompilePPCollectiveOperator:
    """What a communication instruction turns into when a rank's program is compiled to a graph (``graph_emitter.py``): a send to
    ``dst``, a receive from ``src``, or a broadcast from ``src`` to the ranks ``dst`` (which contain ``src``).  Hashable, so an emitter
    can de-duplicate the process groups / buffers it creates per distinct operator."""

    def __init__(self, kind: CompilePPCollectiveKind, src: Optional[int] = None, dst: Union[int, Sequence[int], None] = None, is_backward: bool = False):
        if kind is CompilePPCollectiveKind.SEND:
            if not isinstance(dst, int):
                raise ValueError("a SEND names one destination rank")
        elif kind is CompilePPCollectiveKind.RECV:
            if not isinstance(src, int):
                raise ValueError("a RECV names one source rank")
        elif kind is CompilePPCollectiveKind.BORADCAST:
            if not isinstance(src, int) or isinstance(dst, int) or dst is None or src not in list(dst):
                raise ValueError("a BORADCAST names a source rank and a list of ranks that contains it")
            dst = tuple(dst)
        else:
            raise ValueError(f"cannot compile a collective of kind {kind}")
        self.kind, self.src, self.dst, self.is_backward = kind, src, dst, bool(is_backward)

    def _key(self):
        return (self.kind, self.src, self.dst, self.is_backward)

    def __hash__(self) -> int:
        return hash(self._key())

    def __eq__(self, other) -> bool:
        return isinstance(other, CompilePPCollectiveOperator) and self._key() == other._key()

    def __repr__(self):
        arrow = f"-> {self.dst}" if self.kind is CompilePPCollectiveKind.SEND else (f"<- {self.src}" if self.kind is CompilePPCollectiveKind.RECV else f"{self.src} => {self.dst}")
        return f"{self.kind.name}{'(bwd)' if self.is_backward else ''} {arrow}"


def switch_dtensor(fn: Callable) -> Callable:
    """Decorator for instruction bodies that move tensors over the wire: DTensor arguments go in as their local shards, and a result
    that corresponds to a DTensor argument comes back wrapped with that argument's mesh and placements (legacy
    ``instruction_base.py:42-55``)."""
    import functools

    from ...dtensor.api import DTensor

    @functools.wraps(fn)
    def wrap(*args, **kwargs):
        specs = [(a.device_mesh, a.placements) if isinstance(a, DTensor) else None for a in args]
        out = fn(*[a.to_local() if isinstance(a, DTensor) else a for a in args], **{k: (v.to_local() if isinstance(v, DTensor) else v) for k, v in kwargs.items()})
        live = [s for s in specs if s is not None]
        if not live or out is None:
            return out
        def rewrap(t, spec):
            return DTensor.from_local(t, spec[0], spec[1], run_check=False) if isinstance(t, torch.Tensor) and not isinstance(t, DTensor) else t
        if isinstance(out, (tuple, list)):
            return type(out)(rewrap(t, live[min(i, len(live) - 1)]) for i, t in enumerate(out))
        return rewrap(out, live[0])

    return wrap


registed_functions = INSTRUCTION_REGISTRY  # the reference's name (and spelling) for the registry dict


# ---- schedule grid ---------------------------------------------------------------------------------------------------------------------------
class PipelineSchema:
    """Plan -> schedule.  ``rows[rank]`` are the timed F / B / W ops of the list scheduler; ``grid()`` quantises them into the
    slot x stage status table people draw (``_stage_view``), ``batch_view`` lists for every micro-batch where and when it runs."""

    def __init__(self, plan: PipelineParallelPlan, num_microbatches: int, knobs=None):
        self.plan, self.batches = plan, int(num_microbatches)
        self.P, self.V = plan.num_stages, plan.virtual_chunks
        self.place = stage_placement(self.P, self.V, plan.schedule_type)
        self.rows: List[List[Instr]] = self._gen_schedule(knobs)

    @property
    def name(self) -> str:
        return self.plan.schedule_type.name.lower()

    def _gen_schedule(self, knobs=None) -> List[List[Instr]]:
        return build_schedule(self.plan, self.batches, knobs)

    @property
    def schedules(self) -> List[List[Instr]]:
        return self.rows

    def grid(self) -> List[List[Tuple[Status, int, int]]]:
        """slot -> stage -> (status, microbatch, chunk); a slot is the shortest op duration."""
        ops = [i for r in self.rows for i in r]
        if not ops:
            return []
        dt = min(i.end - i.start for i in ops if i.end > i.start)
        n_slots = int(round(max(i.end for i in ops) / dt))
        g = [[(Status.BUBBLE, -1, -1) for _ in range(self.P)] for _ in range(n_slots)]
        for r, row in enumerate(self.rows):
            for i in row:
                for s in range(int(round(i.start / dt)), max(int(round(i.start / dt)) + 1, int(round(i.end / dt)))):
                    if s < n_slots:
                        g[s][r] = (Status(i.kind), i.microbatch, i.chunk)
        return g

    def _stage_view(self) -> str:
        lines = []
        g = self.grid()
        for r in range(self.P):
            cells = [f"{st.value}{mb}" + (f"'{c}" if self.V > 1 else "") if st != Status.BUBBLE else "." for st, mb, c in (slot[r] for slot in g)]
            lines.append(f"stage {r}: " + " ".join(f"{c:>5}" for c in cells))
        return "\n".join(lines)

    def batch_view(self) -> Dict[int, List[Tuple[float, int, str, int]]]:
        out: Dict[int, List] = {m: [] for m in range(self.batches)}
        for r, row in enumerate(self.rows):
            for i in row:
                out[i.microbatch].append((i.start, r, i.kind, i.vstage))
        return {m: sorted(v) for m, v in out.items()}

    def __str__(self):
        return self._stage_view()


class InstructionBuilder:
    """Lower a schema to per-rank instruction programs.

    ``fuse=True`` merges ``SEND_FORWARD ; RECV_BACKWARD`` (same neighbour) into ``SEND_FORWARD_RECV_BACKWARD`` and the mirror pair
    into ``SEND_BACKWARD_RECV_FORWARD`` — 1F1B's steady state.  ``drain_every``: insert ``DRAIN_SEND_REQS(keep=...)`` after every
    n-th backward so send buffers are retired while the pipeline runs.  ``deallocate``: emit ``DEALLOCATE_OUTPUT_TENSOR`` after sends
    of activations."""

    def __init__(self, fuse: bool = True, drain_every: int = 1, keep_sends: int = 4, deallocate: bool = False):
        self.fuse, self.drain_every, self.keep_sends, self.deallocate = fuse, drain_every, keep_sends, deallocate
        self.programs: Dict[int, List[BaseInstruction]] = {}

    def build(self, schema: PipelineSchema, rank: int) -> List[BaseInstruction]:
        place, NV = schema.place, schema.P * schema.V
        prog: List[BaseInstruction] = []
        n_b = 0
        for ins in schema.rows[rank]:
            m, v, c = ins.microbatch, ins.vstage, ins.chunk
            if ins.kind == "F":
                if v > 0 and place[v - 1][0] != rank:
                    prog.append(RECV_FORWARD(m, v, c, place[v - 1][0]))
                prog.append(FORWARD_STEP(m, v, c))
                if v + 1 < NV and place[v + 1][0] != rank:
                    prog.append(SEND_FORWARD(m, v, c, place[v + 1][0]))
                    if self.deallocate:
                        prog.append(DEALLOCATE_OUTPUT_TENSOR(m, v, c))
            elif ins.kind == "B":
                if v + 1 < NV and place[v + 1][0] != rank:
                    prog.append(RECV_BACKWARD(m, v, c, place[v + 1][0]))
                prog.append(BACKWARD_STEP(m, v, c))
                if v > 0 and place[v - 1][0] != rank:
                    prog.append(SEND_BACKWARD(m, v, c, place[v - 1][0]))
                n_b += 1
                if self.drain_every and n_b % self.drain_every == 0:
                    prog.append(DRAIN_SEND_REQS(keep=self.keep_sends))
            elif ins.kind == "W":
                prog.append(WEIGHT_GRAD_STEP(m, v, c))
            else:  # user instruction kinds pass through by name
                cls = INSTRUCTION_SET.get(ins.kind)
                if cls is None:
                    raise KeyError(f"unknown instruction kind {ins.kind!r}: subclass BaseInstruction with name={ins.kind!r}")
                prog.append(cls(m, v, c))
        prog.append(DRAIN_SEND_REQS(keep=0))
        if self.fuse:
            prog = self._fuse(prog)
        self.programs[rank] = prog
        return prog

    def build_all(self, schema: PipelineSchema) -> Dict[int, List[BaseInstruction]]:
        return {r: self.build(schema, r) for r in range(schema.P)}

    @staticmethod
    def _fuse(prog: List[BaseInstruction]) -> List[BaseInstruction]:
        out: List[BaseInstruction] = []
        i = 0
        while i < len(prog):
            a = prog[i]
            j = i + 1
            while j < len(prog) and isinstance(prog[j], (DEALLOCATE_OUTPUT_TENSOR, DRAIN_SEND_REQS)):
                j += 1  # bookkeeping instructions between the pair do not touch the wire
            b = prog[j] if j < len(prog) else None
            if type(a) is SEND_FORWARD and type(b) is RECV_BACKWARD and a.peer == b.peer:
                out.append(SEND_FORWARD_RECV_BACKWARD(a.microbatch, a.vstage, a.chunk, a.peer, b.microbatch, b.vstage))
                out.extend(prog[i + 1:j])
                i = j + 1
            elif type(a) is SEND_BACKWARD and type(b) is RECV_FORWARD and a.peer == b.peer:
                out.append(SEND_BACKWARD_RECV_FORWARD(a.microbatch, a.vstage, a.chunk, a.peer, b.microbatch, b.vstage))
                out.extend(prog[i + 1:j])
                i = j + 1
            else:
                out.append(a)
                i += 1
        return out


    # -- user-written programs (legacy ``instruction_base.py:436-520``) ------------------------------------------------------------------------
    # A power user can skip schedules altogether: register plain functions under names, give every stage a comma-separated list of
    # those names, and ``run(stage_id)`` calls them in order.  The functions take no arguments; they talk to each other through the
    # builder: ``builder.last`` is what the previous function returned, ``builder.pos`` its own index, and anything else the user hangs
    # on the builder (``builder.model``, ``builder.dataloader``, ``builder.topo``, ``builder.stage_id`` ...).
    def build_from_dict(self, instructions: Dict) -> None:
        if not isinstance(instructions, dict):
            raise TypeError("instructions: {stage_id: 'name,name,...' | [names]}")
        self.global_instructions_funcs, self.global_instructions_str = getattr(self, "global_instructions_funcs", {}), getattr(self, "global_instructions_str", {})
        for stage_id, names in instructions.items():
            names = [n.strip() for n in names.split(",")] if isinstance(names, str) else list(names)
            missing = [n for n in names if n not in INSTRUCTION_REGISTRY]
            if missing:
                raise KeyError(f"stage {stage_id}: instructions {missing} are not registered (register_instruction(name))")
            self.global_instructions_funcs[stage_id] = [INSTRUCTION_REGISTRY[n] for n in names]
            self.global_instructions_str[stage_id] = names

    @property
    def pos(self) -> int:
        return getattr(self, "_pos", 0)

    @property
    def last(self):
        return getattr(self, "_stack", None)

    def run(self, stage_id: int) -> List:
        """Call stage ``stage_id``'s functions in order; returns every function's result."""
        out = []
        for pos, fn in enumerate(getattr(self, "global_instructions_funcs", {}).get(stage_id, [])):
            self._pos = pos
            self._stack = fn()
            out.append(self._stack)
        return out

    def export(self, stage_id: int, *args, **kwargs):
        """``torch.export`` of one stage's user program: the functions run in order on (args, kwargs); the one called ``forward``
        is the stage module (``builder.model``)."""
        funcs, model = list(self.global_instructions_funcs[stage_id]), self.model

        class _Program(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.model = model

            def forward(self, *a, **kw):
                for f in funcs:
                    if getattr(f, "__name__", "") == "forward":
                        a, kw = (self.model(*a, **kw),), {}
                    else:
                        a, kw = f(*a, **kw)
                return a, kw

        return torch.export.export(_Program(), args, kwargs or None)

    def draw_user_instructions(self, path: Optional[str] = None) -> str:
        """The user programs as a text table (one row per stage); with ``path`` also as a picture when matplotlib is installed."""
        rows = getattr(self, "global_instructions_str", {})
        w = max((len(n) for names in rows.values() for n in names), default=1)
        text = "\n".join(f"stage {s}: " + " | ".join(n.ljust(w) for n in names) for s, names in sorted(rows.items()))
        if path is not None:
            try:
                from matplotlib import pyplot as plt
            except ImportError:
                return text
            fig, ax = plt.subplots()
            for s, names in sorted(rows.items()):
                for k, n in enumerate(names):
                    ax.add_patch(plt.Rectangle((k, -s), 1, 1, fill=False, edgecolor="black", lw=2))
                    ax.text(k + 0.5, -s + 0.5, n, ha="center", va="center")
                ax.text(-0.5, -s + 0.5, str(s), ha="center", va="center")
            ax.set_xlim(0, max(len(n) for n in rows.values()))
            ax.set_ylim(-len(rows) + 1, 1)
            ax.axis("off")
            fig.savefig(path)
            plt.close(fig)
        return text

    # -- inspection ------------------------------------------------------------------------------------------------------------------------
    def dump_instructions(self, rank: Optional[int] = None) -> str:
        ranks = [rank] if rank is not None else sorted(self.programs)
        return "\n".join(f"[rank {r}] {k:3d}: {ins.dump()}" for r in ranks for k, ins in enumerate(self.programs[r]))

    def draw_instructions(self, width: int = 6) -> str:
        """One line per rank, one cell per instruction, compute steps only (F3 / B3 / W3): the classic pipeline diagram."""
        if not self.programs and getattr(self, "global_instructions_str", None):
            return self.draw_user_instructions()
        sym = {"FORWARD_STEP": "F", "BACKWARD_STEP": "B", "WEIGHT_GRAD_STEP": "W", "FWD": "F", "BWD": "B"}
        return "\n".join(f"rank {r}: " + " ".join(f"{sym[i.name]}{i.microbatch}".ljust(width) for i in self.programs[r] if i.name in sym) for r in sorted(self.programs))

    @staticmethod
    def check_streams(programs: Dict[int, List[BaseInstruction]]) -> None:
        """Every send has exactly one matching receive and, per directed pair, the receiver's order is a permutation the stash can
        serve (always true) — what is checked is existence: an unmatched message would hang the run."""
        sent, recvd = {}, {}
        for r, prog in programs.items():
            for ins in prog:
                for kind, m, v, peer, is_send in _wire_ops(ins):
                    (sent if is_send else recvd).setdefault((r, peer) if is_send else (peer, r), []).append((kind, m, v))
        for pair in set(sent) | set(recvd):
            a, b = sorted(sent.get(pair, [])), sorted(recvd.get(pair, []))
            if a != b:
                raise AssertionError(f"stream {pair[0]} -> {pair[1]}: sends {a[:4]}... do not match receives {b[:4]}...")


def _wire_ops(ins: BaseInstruction):
    own = getattr(ins, "wire_ops", None)
    if own is not None:
        yield from own()
    elif isinstance(ins, SEND_FORWARD):
        yield ("F", ins.microbatch, ins.vstage, ins.peer, True)
    elif isinstance(ins, SEND_BACKWARD):
        yield ("B", ins.microbatch, ins.vstage, ins.peer, True)
    elif isinstance(ins, RECV_FORWARD):
        yield ("F", ins.microbatch, ins.vstage - 1, ins.peer, False)
    elif isinstance(ins, RECV_BACKWARD):
        yield ("B", ins.microbatch, ins.vstage + 1, ins.peer, False)
    elif isinstance(ins, SEND_FORWARD_RECV_BACKWARD):
        yield ("F", ins.microbatch, ins.vstage, ins.peer, True)
        yield ("B", ins.recv_microbatch, ins.recv_vstage + 1, ins.peer, False)
    elif isinstance(ins, SEND_BACKWARD_RECV_FORWARD):
        yield ("B", ins.microbatch, ins.vstage, ins.peer, True)
        yield ("F", ins.recv_microbatch, ins.recv_vstage - 1, ins.peer, False)


# ---- wire ------------------------------------------------------------------------------------------------------------------------------------
class StageLink:
    """Stream-ordered tensor bundles between pipeline ranks.  A message is a header (key + shapes / dtypes, int64) followed by its
    tensors; the receiver reads a peer's stream strictly in order and stashes what it does not need yet."""

    _DT = [torch.float32, torch.bfloat16, torch.float16, torch.float64, torch.int64, torch.int32, torch.uint8, torch.bool, torch.int8, torch.int16]
    _HDR = 64

    def __init__(self, group, ranks: Sequence[int], device, wire_dtype: Optional[torch.dtype] = None):
        self.group, self.ranks, self.device, self.wire_dtype = group, list(ranks), device, wire_dtype
        self.stash: Dict[int, Dict[Tuple, Tuple[torch.Tensor, ...]]] = {}
        self.sends: List = []
        self.bytes_sent = 0

    def _g(self, pp_rank: int) -> int:
        return self.ranks[pp_rank]

    def send(self, key: Tuple[str, int, int], tensors: Sequence[torch.Tensor], dst: int) -> None:
        kinds = {"F": 0, "B": 1}
        hdr = torch.zeros(self._HDR, dtype=torch.int64)
        hdr[0], hdr[1], hdr[2], hdr[3] = kinds[key[0]], key[1], key[2], len(tensors)
        p = 4
        payload = []
        for t in tensors:
            t = t.detach()
            if self.wire_dtype is not None and t.is_floating_point():
                t = t.to(self.wire_dtype)
            t = t.contiguous()
            hdr[p], hdr[p + 1] = self._DT.index(t.dtype), t.dim()
            hdr[p + 2:p + 2 + t.dim()] = torch.tensor(list(t.shape), dtype=torch.int64)
            p += 2 + t.dim()
            assert p <= self._HDR, "too many / too high-rank tensors for one pipeline message header"
            payload.append(t)
        hdr = hdr.to(self.device)
        self.sends.append((dist.isend(hdr, self._g(dst), group=self.group), hdr))
        for t in payload:
            self.sends.append((dist.isend(t, self._g(dst), group=self.group), t))
            self.bytes_sent += t.numel() * t.element_size()

    def _read_one(self, src: int) -> Tuple[Tuple, Tuple[torch.Tensor, ...]]:
        hdr = torch.zeros(self._HDR, dtype=torch.int64, device=self.device)
        dist.recv(hdr, self._g(src), group=self.group)
        h = hdr.tolist()
        key = ("F" if h[0] == 0 else "B", int(h[1]), int(h[2]))
        p, ts = 4, []
        for _ in range(int(h[3])):
            dt, nd = self._DT[int(h[p])], int(h[p + 1])
            shape = [int(x) for x in h[p + 2:p + 2 + nd]]
            p += 2 + nd
            buf = torch.empty(shape, dtype=dt, device=self.device)
            dist.recv(buf, self._g(src), group=self.group)
            ts.append(buf)
        return key, tuple(ts)

    def recv(self, key: Tuple[str, int, int], src: int) -> Tuple[torch.Tensor, ...]:
        st = self.stash.setdefault(src, {})
        while key not in st:
            k, ts = self._read_one(src)
            st[k] = ts
        return st.pop(key)

    def drain_sends(self, keep: int = 0) -> None:
        while len(self.sends) > keep:
            req, _ = self.sends.pop(0)
            req.wait()

    def assert_empty(self) -> None:
        left = {s: list(d) for s, d in self.stash.items() if d}
        assert not left, f"received but never consumed: {left}"


# ---- VM ---------------------------------------------------------------------------------------------------------------------------------------
class InstructionVM:
    """Runs one rank's program on a ``PipeModule``.  State per (microbatch, vstage): received inputs / gradients, produced outputs,
    saved (inputs, outputs) for backward, deferred weight-gradient closures for zero-bubble schedules."""

    def __init__(self, module, plan: PipelineParallelPlan, pp_rank: int, pp_group=None, loss_fn: Optional[Callable] = None, device=None, pp_ranks: Optional[Sequence[int]] = None):
        self.module, self.plan, self.rank, self.group, self.loss_fn = module, plan, pp_rank, pp_group, loss_fn
        self.P, self.V = plan.num_stages, plan.virtual_chunks
        self.NV = self.P * self.V
        self.place = stage_placement(self.P, self.V, plan.schedule_type)
        self.split_w = plan.schedule_type in (PipelineScheduleType.ZERO_BUBBLE, PipelineScheduleType.ZERO_BUBBLE_V)
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        ranks = list(pp_ranks) if pp_ranks is not None else (dist.get_process_group_ranks(pp_group) if pp_group is not None else list(range(self.P)))
        self.link = StageLink(pp_group, ranks, self.device, plan.p2p_tensor_dtype)
        self.builder = InstructionBuilder()
        self._programs: Dict[Tuple[int, bool], List[BaseInstruction]] = {}
        self.executed: List[str] = []

    def program(self, num_microbatches: int, forward_only: bool = False) -> List[BaseInstruction]:
        key = (num_microbatches, bool(forward_only or self.plan.forward_only))
        if key not in self._programs:
            import dataclasses

            plan = dataclasses.replace(self.plan, forward_only=True) if key[1] else self.plan
            self._programs[key] = self.builder.build(PipelineSchema(plan, num_microbatches), self.rank)
        return self._programs[key]

    def run(self, inputs: Sequence, labels: Optional[Sequence] = None, forward_only: bool = False, program: Optional[List[BaseInstruction]] = None):
        M = len(inputs)
        self.inputs, self.labels, self.forward_only, self.M = inputs, labels, forward_only, M
        self.inbox_f: Dict = {}
        self.inbox_b: Dict = {}
        self.outbox_f: Dict = {}
        self.outbox_b: Dict = {}
        self.acts: Dict = {}
        self.local_f: Dict = {}
        self.local_b: Dict = {}
        self.pending_w: Dict = {}
        self.sent_out: Dict = {}
        self.losses: List[Optional[torch.Tensor]] = [None] * M
        self.outputs: List = [None] * M
        self.executed = []
        for ins in (program if program is not None else self.program(M, forward_only)):
            h = INSTRUCTION_REGISTRY.get(ins.name)
            self.executed.append(ins.name)
            if h is not None and h(self, ins) is not None:
                continue
            ins.run(self)
        self.link.drain_sends(0)
        self.link.assert_empty()
        if all(l is None for l in self.losses):
            return None, self.outputs
        return torch.stack([l for l in self.losses if l is not None]).sum(), self.outputs

    __call__ = run

    # -- compute steps ------------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _tup(x):
        return x if isinstance(x, tuple) else (tuple(x) if isinstance(x, list) else (x,))

    def forward_step(self, m: int, v: int, c: int) -> None:
        if v == 0:
            xs = self._tup(self.inputs[m])
        elif self.place[v - 1][0] == self.rank:
            xs = self.local_f.pop((m, v - 1))
        else:
            xs = self.inbox_f.pop((m, v))
        xs = tuple(x.detach().requires_grad_(x.is_floating_point() and not self.forward_only) if isinstance(x, torch.Tensor) else x for x in xs)
        with torch.set_grad_enabled(not self.forward_only), ndtimeit(predefined.FORWARD_COMPUTE, microbatch=m, vstage=v):
            out = self.module(*xs, chunk_id=c)
            outs = self._tup(out)
            if v == self.NV - 1:
                self.outputs[m] = out
                if self.loss_fn is not None and self.labels is not None:
                    loss = self.loss_fn(out, self.labels[m]) / self.M
                    self.losses[m] = loss.detach()
                    outs = (loss,)
        self.acts[(m, v)] = (xs, outs)
        if v + 1 < self.NV:
            if self.place[v + 1][0] == self.rank:
                self.local_f[(m, v)] = tuple(o.detach() for o in outs)
            else:
                self.outbox_f[(m, v)] = outs
                self.sent_out[(m, v)] = outs

    def backward_step(self, m: int, v: int, c: int) -> None:
        xs, outs = self.acts[(m, v)]
        if v == self.NV - 1:
            gouts = tuple(torch.ones_like(o) for o in outs)
        elif self.place[v + 1][0] == self.rank:
            gouts = self.local_b.pop((m, v + 1))
        else:
            gouts = self.inbox_b.pop((m, v))
        pairs = [(o, g) for o, g in zip(outs, gouts) if isinstance(o, torch.Tensor) and o.requires_grad]
        o_t, g_t = [p[0] for p in pairs], [p[1].to(p[0].dtype) for p in pairs]
        grad_inputs = [x for x in xs if isinstance(x, torch.Tensor) and x.requires_grad]
        freed = any(o.shape != g.shape for o, g in zip(o_t, g_t))  # DEALLOCATE_OUTPUT_TENSOR shrank the storage; grad_fn is intact
        with ndtimeit(predefined.BACKWARD_COMPUTE, microbatch=m, vstage=v):
            if self.split_w:
                gi = torch.autograd.grad(o_t, grad_inputs, g_t, retain_graph=True, allow_unused=True) if grad_inputs else ()
                self.pending_w[(m, v)] = (o_t, g_t)
            elif freed:
                # the Python front end insists on output.shape == grad.shape; the engine itself only needs grad_fn
                from torch.autograd import Variable

                Variable._execution_engine.run_backward(tuple(o_t), tuple(g_t), False, False, tuple(), allow_unreachable=True, accumulate_grad=True)
            else:
                torch.autograd.backward(o_t, g_t)
            if not self.split_w:
                gi = tuple(x.grad for x in grad_inputs)
                del self.acts[(m, v)]
        self.sent_out.pop((m, v), None)
        if v > 0:
            gi = tuple(g if g is not None else torch.zeros_like(x) for g, x in zip(gi, grad_inputs))
            if self.place[v - 1][0] == self.rank:
                self.local_b[(m, v)] = gi
            else:
                self.outbox_b[(m, v)] = gi

    def weight_step(self, m: int, v: int, c: int) -> None:
        o_t, g_t = self.pending_w.pop((m, v))
        params = [p for p in self.module.chunk(c).parameters() if p.requires_grad]
        with ndtimeit(predefined.BACKWARD_COMPUTE, microbatch=m, vstage=v, part="weight-grad"):
            gs = torch.autograd.grad(o_t, params, g_t, allow_unused=True)
        for p, g in zip(params, gs):
            if g is not None:
                p.grad = g if p.grad is None else p.grad + g
        del self.acts[(m, v)]

    def deallocate_output(self, m: int, v: int) -> None:
        """Free the storage of an activation that only its consumer stage needs; its ``grad_fn`` stays for backward.  Only safe once the
        send has completed, so the sends are drained first."""
        outs = self.sent_out.pop((m, v), None)
        if outs is None:
            return
        self.link.drain_sends(0)
        for o in outs:
            if isinstance(o, torch.Tensor) and o._base is None and o.grad_fn is not None and o.numel() > 1:
                o.data = torch.empty((1,), device=o.device, dtype=o.dtype)


VESCALE_INTRUCTION_BUILDER = InstructionBuilder()  # the process-wide builder user programs hang their state on (the reference's spelling)


def __getattr__(name):  # ``InstructionGenerator`` lives with the generators; importing it here eagerly would be a cycle
    if name == "InstructionGenerator":
        from ._schedules import InstructionGenerator

        return InstructionGenerator
    raise AttributeError(name)
