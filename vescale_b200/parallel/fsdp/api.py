"""``fully_shard``: the veScale-FSDP wrapper on RaggedShard.

    for blk in model.layers: fully_shard(blk, mesh)
    fully_shard(model.embed, mesh); fully_shard(model.head, mesh); fsdp = fully_shard(model, mesh)
    opt = vescale_b200.optim.FSDPAdamW(model, lr=...)          # fused flat AdamW over unit shards
    loss = model(tokens, labels); loss.backward(); opt.step(); opt.zero_grad()

Semantics (what the reference *describes*, ``docs/texts/raggedshard.md:67-77``; designed here):
  * outside forward/backward ``module.parameters()`` are fp32 **RaggedShard DTensors** (views into the unit's
    master shard) → any torch optimizer, ``clip_grad_norm_`` and DCP work on them;
  * inside, the original parameters are plain tensors viewing the gathered unit buffer; model compute never
    goes through DTensor dispatch;
  * all-gather of unit i+1 is prefetched on a side stream during unit i's forward (unit i-1 in backward);
    reduce-scatter of unit i runs on another stream right after its backward; gathered buffers come from a
    small pool when ``reshard_after_forward`` else stay resident (ZeRO-2 style; 180 GB of HBM3e makes that
    the default for models whose bf16 copy fits comfortably);
  * gradient accumulation: ``set_requires_gradient_sync(False)`` keeps accumulating into the flat buffers.
"""
from __future__ import annotations

import contextlib
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from ...mesh import DeviceMesh, init_device_mesh
from .unit import FSDPUnit, MixedPrecisionPolicy, make_event, _NullStream

__all__ = ["checkpoint_module", "fully_shard", "FSDPState", "MixedPrecisionPolicy", "get_fsdp_state", "fsdp_units"]


from ...profiler import ndtimeit, predefined  # ndtimeline: UNSHARD_AG / GRAD_RS regions on the communication streams

class _BufferPool:
    """Recycles gathered-parameter / gradient buffers.  A buffer handed back is tagged with the event
    after which its contents are dead; the next user makes its stream wait for that event."""

    def __init__(self, device):
        self.device = device
        self.free: Dict[Tuple[int, torch.dtype, bool], List[Tuple[torch.Tensor, Any]]] = {}
        self.allocated_bytes = 0

    def get(self, numel: int, dtype, stream, alloc: Callable[[], torch.Tensor], symmetric: bool = False, wait=None) -> torch.Tensor:
        key = (numel, dtype, symmetric)
        lst = self.free.get(key)
        if lst:
            buf, evt = lst.pop(0)  # FIFO: the buffer that has been idle longest (its collectives are most likely done)
            if evt is not None and stream is not None:
                if wait is not None:
                    wait(evt)  # the caller's (timed) wait on its compute stream
                elif hasattr(stream, "wait_event"):
                    stream.wait_event(evt)
            return buf
        buf = alloc()
        self.allocated_bytes += buf.numel() * buf.element_size()
        return buf

    def put(self, buf: torch.Tensor, event, symmetric: bool = False) -> None:
        self.free.setdefault((buf.numel(), buf.dtype, symmetric), []).append((buf, event))

    def poison_free(self, stream=None) -> None:
        """Debug (``VESCALE_B200_SYMM_DEBUG=1``): NaN-fill every idle buffer once its last collective has finished, so a
        consumer that reads a pooled buffer before its producer has written it cannot go unnoticed."""
        from ...comm.symm_debug import poison

        for lst in self.free.values():
            for buf, evt in lst:
                if evt is not None and hasattr(evt, "synchronize"):
                    evt.synchronize()
                poison(buf)


class FSDPState:
    """Root-level state shared by all units of one ``fully_shard``-ed model."""

    def __init__(self, mesh: DeviceMesh, mesh_dim: int, device: torch.device):
        self.mesh = mesh
        self.mesh_dim = mesh_dim
        self.device = device
        self.units: List[FSDPUnit] = []
        self.cuda = device.type == "cuda"
        if self.cuda:
            self.ag_stream = torch.cuda.Stream(priority=-1)
            self.rs_stream = torch.cuda.Stream(priority=-1)
            from ...profiler.stream import register_comm_stream

            register_comm_stream("fsdp-ag", self.ag_stream)  # ndtimeline looks communication streams up by name / group
            register_comm_stream("fsdp-rs", self.rs_stream)
        else:
            self.ag_stream = self.rs_stream = _NullStream()
        self.pool = _BufferPool(device)
        self.requires_gradient_sync = True
        self.reshard_after_forward = True
        self.prefetch = 1
        self.in_backward = False
        self.grad_scale: Optional[float] = None  # default 1/world
        self.final_callback_queued = False
        self.post_backward_pending: List[FSDPUnit] = []
        self.exposed_wait_ms = 0.0
        self.comm = None
        self.iteration = 0
        self._handshake_iter = -1  # iteration in which a gather already exchanged "my shards are final" flags
        # exposed-communication accounting: when on, every wait of the compute stream on a communication event is bracketed
        # by two timing events; their elapsed time is the stall (≈ 0 when the collective had already finished)
        self.measure_exposed = False
        self._exposed_events: List[Tuple[Any, Any]] = []
        self._event_pool: List[Tuple[Any, Any]] = []  # pre-created timing events (prepare_exposed_measure)

    # ------------------------------------------------------------------ exposed communication
    def wait_comm_event(self, event) -> None:
        """Make the current (compute) stream wait for a communication event, timing the stall when ``measure_exposed``."""
        if not self.cuda or event is None:
            return
        s = self.cur_stream()
        if not self.measure_exposed:
            s.wait_event(event)
            return
        pool = self._event_pool
        if pool:
            a, b = pool.pop()
        else:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        s.wait_event(event)
        b.record(s)
        self._exposed_events.append((a, b))

    def prepare_exposed_measure(self, n_pairs: int) -> None:
        """Create ``n_pairs`` timing-event pairs NOW and record each once, so that a measured region never creates a CUDA event.
        ``cudaEventCreate`` is lazy in torch (first ``record``) and the driver grows its event pools in chunks; with peer access
        enabled between all GPUs of a node a pool growth maps new memory into every peer and synchronises the device.  That was
        the one-step stall (150-330 ms, always the SECOND measured step, only while ``measure_exposed`` was on) in the round-1 and
        early round-2 bench records: the host blocked in the driver until the GPU drained, then had to re-fill the launch queue."""
        if not self.cuda:
            return
        s = self.cur_stream()
        pool = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(int(n_pairs))]
        for x, y in pool:
            x.record(s)
            y.record(s)
        self._event_pool = pool

    def exposed_comm_ms(self, reset: bool = True) -> float:
        """Device time the compute stream spent stalled on all-gather / reduce-scatter completion since the last reset
        (the "exposed communication" of BASELINE.json's metric).  Call after a device synchronize."""
        tot = 0.0
        for a, b in self._exposed_events:
            tot += a.elapsed_time(b)
        if reset:
            self._exposed_events = []
        return tot

    # ------------------------------------------------------------------ stream helpers
    def cur_stream(self):
        return torch.cuda.current_stream() if self.cuda else _NullStream()

    @contextlib.contextmanager
    def on(self, stream):
        if self.cuda:
            with torch.cuda.stream(stream):
                yield
        else:
            yield

    def _need_handshake(self) -> bool:
        """One flag exchange per optimizer step is enough: a peer's signal follows its whole optimizer step in stream order, so
        every one of its shards is final; later gathers of the same step (forward and backward) skip the exchange."""
        need = self._handshake_iter != self.iteration
        self._handshake_iter = self.iteration
        return need

    # ------------------------------------------------------------------ unshard / reshard
    def launch_all_gather(self, u: FSDPUnit) -> None:
        if u.unsharded or u.ag_event is not None:
            return
        if u.world == 1:
            u.use_full(u.param_shard)
            u.ag_event = None
            u._ag_direct = True
            return
        persistent = not self.reshard_after_forward
        if persistent and getattr(u, "_persistent_full", None) is not None:
            full = u._persistent_full
        else:
            full = self.pool.get(u.S * u.world, u.param_dtype, self.ag_stream, lambda: u._alloc_full(u.param_dtype))
            if persistent:
                u._persistent_full = full
        # the shard may have just been written by the optimizer on the compute stream
        self.ag_stream.wait_stream(self.cur_stream())
        hs = self._need_handshake()
        with self.on(self.ag_stream), ndtimeit(predefined.UNSHARD_AG, stream=self.ag_stream if self.cuda else None, unit=u.name):
            u.all_gather(full, handshake=hs)
            evt = make_event(self.device)
            evt.record(self.ag_stream if self.cuda else None)
        u.ag_event = evt
        u._ag_full = full
        u._ag_direct = False

    def launch_fused_gather(self, u: FSDPUnit) -> bool:
        """Exposed (not prefetched) forward all-gather of a unit whose first GEMM weight is known: the GEMM kernel gathers
        that weight itself (``SymmUnitComm.fused_first_linear``), small parameters are pulled first by a tiny launch, and the
        rest of the unit streams in on the all-gather stream *behind* the first GEMM."""
        comm = u.comm
        name = getattr(u, "fused_first", None)
        if not name or comm is None or not getattr(comm, "symmetric", False) or u.world == 1 or self.in_backward:
            return False
        slot = getattr(u, "_fused_slot", False)
        if slot is False:
            slot = u._fused_slot = comm.fusable_slot(u, name) if name in u.param_names else None
        if slot is None:
            return False
        persistent = not self.reshard_after_forward
        if persistent and getattr(u, "_persistent_full", None) is not None:
            full = u._persistent_full
        else:
            full = self.pool.get(u.S * u.world, u.param_dtype, self.ag_stream, lambda: u._alloc_full(u.param_dtype))
            if persistent:
                u._persistent_full = full
        hs = u.refresh_param_shard() or self._need_handshake()
        for s in u.layout.slots:  # norm weights / biases: needed before the first GEMM, a few KB each
            if s is not slot and s.numel <= 65536:
                comm.all_gather(u.param_shard, full, u, only=(s.offset, s.end), handshake=hs)
                hs = False
        self.ag_stream.wait_stream(self.cur_stream())
        with self.on(self.ag_stream):
            comm.all_gather(u.param_shard, full, u, skip=(slot.offset, slot.end), handshake=hs)
            evt = make_event(self.device)
            evt.record(self.ag_stream)
        u.use_full(full)
        u._params_valid = True
        u._ag_direct = False
        param = u.params[u.param_names.index(name)]
        state = self
        others = [p for s, p in zip(u.layout.slots, u.params) if s is not slot and s.numel > 65536]

        def rest_arrived():
            """The first use of any other large weight of the unit waits for the rest of the all-gather — by then it has
            been streaming in behind the first GEMM and whatever followed it (RoPE, attention, ...)."""
            if u._pending_rest:
                u._pending_rest = False
                state.wait_comm_event(evt)
                for q in others:
                    q.__dict__.pop("_vb_pending_gather", None)

        def pending(a: torch.Tensor, out=None):
            del param._vb_pending_gather
            u._pending_param = None
            a2 = a.reshape(-1, a.shape[-1])
            if a2.dtype == torch.bfloat16 and a2.is_contiguous() and a2.shape[0] % 256 == 0:
                y = comm.fused_first_linear(a2, u, slot, full, handshake=False)
            else:  # shape the kernel does not take: gather the weight the ordinary way
                comm.all_gather(u.param_shard, full, u, only=(slot.offset, slot.end), handshake=False)
                rest_arrived()
                from ...ops import functional as Fn

                y = Fn.gemm_nt(a2, param)
            if out is not None:
                return out.copy_(y.view(out.shape))
            return y.view(*a.shape[:-1], y.shape[-1])

        def make_other(q):
            def other(a: torch.Tensor, out=None):
                rest_arrived()
                from ...ops import functional as Fn

                return Fn.gemm_nt(a, q, out)

            return other

        param._vb_pending_gather = pending
        for q in others:
            q._vb_pending_gather = make_other(q)
        u._pending_param = param
        u._pending_rest = True
        u._pending_rest_fn = rest_arrived
        u._pending_evt = evt
        return True

    def wait_all_gather(self, u: FSDPUnit) -> None:
        if u.unsharded:
            return
        if u.ag_event is None and self.launch_fused_gather(u):
            return
        if u.ag_event is None:
            self.launch_all_gather(u)
        if u.unsharded:
            return
        self.wait_comm_event(u.ag_event)
        u.ag_event = None
        u.use_full(u._ag_full)
        u._params_valid = True

    def reshard(self, u: FSDPUnit) -> None:
        pend = getattr(u, "_pending_param", None)
        if pend is not None:
            # the module never ran the GEMM that was meant to gather this weight: finish the gather the ordinary way
            del pend._vb_pending_gather
            u._pending_param = None
            slot = u._fused_slot
            u.comm.all_gather(u.param_shard, u.full_param, u, only=(slot.offset, slot.end))
            self.cur_stream().wait_event(u._pending_evt)
            raise RuntimeError(f"FSDP unit {u.name}: fuse_first_gemm={u.fused_first!r} but that weight was not the first GEMM of the forward")
        if getattr(u, "_pending_rest", False):
            u._pending_rest_fn()  # no other large weight was touched: the gather must still complete before the buffer is recycled
        if not u.unsharded or u.world == 1:
            return
        if not self.reshard_after_forward:
            return  # stays resident; contents refreshed by the next all-gather after the optimizer step
        full = u.release_full()
        evt = make_event(self.device)
        evt.record()
        self.pool.put(full, evt)

    def invalidate_params(self) -> None:
        """After an optimizer step every gathered copy is stale."""
        for u in self.units:
            if u.world == 1:
                continue
            if u.unsharded:
                full = u.release_full()
                if self.reshard_after_forward:
                    evt = make_event(self.device)
                    evt.record()
                    self.pool.put(full, evt)
            u.ag_event = None

    # ------------------------------------------------------------------ gradient reduction
    def prepare_grad_buffer(self, u: FSDPUnit) -> None:
        if u.full_grad is not None:
            for p in u.params:  # accumulation across micro-batches
                p._main_grad_initialised = True
            return
        sym = self.comm is not None and getattr(self.comm, "symmetric", False)
        if u.world == 1:
            fg = getattr(u, "_persistent_grad", None)
            if fg is None:
                fg = u._persistent_grad = torch.empty(u.S, dtype=u.param_dtype, device=u.device)
            u.attach_grad_buffer(fg, accumulate=False)
            return
        def _alloc():
            if sym and getattr(self, "symm_pool_frozen", False):
                raise RuntimeError("symmetric gradient pool exhausted after lazy_init (would need a rendezvous mid-step)")
            return u._alloc_full(u.param_dtype, symmetric=sym)

        fg = self.pool.get(u.S * u.world, u.param_dtype, self.cur_stream(), _alloc, symmetric=sym, wait=self.wait_comm_event)
        if sym:
            self.comm.wait_buffer_free(fg)  # peers may still be pull-reducing the previous tenant of this buffer
        u.attach_grad_buffer(fg, accumulate=False)

    def post_backward(self, u: FSDPUnit) -> None:
        if getattr(u, "_post_backward_done", False):
            return
        u._post_backward_done = True
        u.collect_autograd_grads()
        self.reshard(u)
        if not self.requires_gradient_sync:
            return
        scale = self.grad_scale if self.grad_scale is not None else 1.0 / u.world
        if u.world == 1:
            u.reduce_scatter(scale)
            u.rs_event = None
            u.detach_grad_buffer()  # grad_shard keeps aliasing the persistent bf16 buffer
            return
        self.rs_stream.wait_stream(self.cur_stream())
        with self.on(self.rs_stream), ndtimeit(predefined.GRAD_RS, stream=self.rs_stream if self.cuda else None, unit=u.name):
            fused = getattr(self, "fused_optimizer", None)
            if fused is not None and self.comm is not None and getattr(self.comm, "symmetric", False):
                fused.fused_update(u, scale)  # reduce-scatter ⊕ AdamW ⊕ bf16 cast in one kernel
            else:
                u.reduce_scatter(scale)
            evt = make_event(self.device)
            evt.record(self.rs_stream if self.cuda else None)
        u.rs_event = evt
        fg = u.detach_grad_buffer()
        self.pool.put(fg, evt, symmetric=self.comm is not None and getattr(self.comm, "symmetric", False))

    def finish_backward(self) -> None:
        """Runs once at the end of ``backward`` (queued autograd callback): units whose inputs needed no
        gradient never see their post-backward node, so flush them here."""
        self.final_callback_queued = False
        for u in self.units:
            if getattr(u, "_saw_forward", False) and not getattr(u, "_post_backward_done", False) and u.full_grad is not None:
                self.post_backward(u)
        for u in self.units:
            u._saw_forward = False
            u._post_backward_done = False
            u._swap.to_sharded()
        self.in_backward = False

    def wait_grads(self) -> None:
        if self.cuda:
            for u in self.units:
                if u.rs_event is not None:
                    self.wait_comm_event(u.rs_event)
                    u.rs_event = None

    def set_requires_gradient_sync(self, flag: bool) -> None:
        self.requires_gradient_sync = flag

    def lazy_init(self) -> None:
        """Before the first forward: allocate every *symmetric* pool buffer now.  Symmetric allocation is a
        rendezvous (host-blocking collective, may synchronise the device); doing it later — while another rank
        already sits in a kernel that spins on this rank's signal — can deadlock.  ``grad_pool_depth`` gradient
        buffers per distinct unit size are enough for reduce-scatter to overlap the next unit's backward."""
        if getattr(self, "_lazy_done", False):
            return
        self._lazy_done = True
        sym = self.comm is not None and getattr(self.comm, "symmetric", False)
        if not sym:
            return
        depth = getattr(self, "grad_pool_depth", 3)
        sizes = {}
        for u in self.units:
            if u.world > 1:
                sizes.setdefault((u.S * u.world, u.param_dtype), []).append(u)
        for (numel, dtype), us in sizes.items():
            for _ in range(min(depth, len(us))):
                self.pool.put(us[0]._alloc_full(dtype, symmetric=True), None, symmetric=True)
        self.symm_pool_frozen = True


class _PostBackward(torch.autograd.Function):
    """Identity on the unit's inputs; its backward runs after every gradient of the unit was produced."""

    @staticmethod
    def forward(ctx, state, unit, *xs):
        ctx.state, ctx.unit = state, unit
        return tuple(x.view_as(x) for x in xs) if len(xs) > 1 else xs[0].view_as(xs[0])

    @staticmethod
    def backward(ctx, *grads):
        ctx.state.post_backward(ctx.unit)
        return (None, None) + grads


def _unit_index(state: FSDPState, u: FSDPUnit) -> int:
    return state.units.index(u)


def _pre_forward(state: FSDPState, u: FSDPUnit, module, args, kwargs):
    idx = u._index
    state.lazy_init()
    u._saw_forward = True
    u._post_backward_done = False
    state.wait_all_gather(u)
    for k in range(1, state.prefetch + 1):
        if idx + k < len(state.units):
            state.launch_all_gather(state.units[idx + k])
    if torch.is_grad_enabled():
        # hook the inputs so that post-backward fires once every gradient of the unit exists
        tens = [a for a in args if isinstance(a, torch.Tensor) and a.requires_grad]
        if tens:
            outs = _PostBackward.apply(state, u, *tens)
            outs = (outs,) if isinstance(outs, torch.Tensor) else outs
            it = iter(outs)
            args = tuple(next(it) if isinstance(a, torch.Tensor) and a.requires_grad else a for a in args)
    return args, kwargs


def _post_forward(state: FSDPUnit, u: FSDPUnit, module, args, output):
    if torch.is_grad_enabled():
        outs = [o for o in (output if isinstance(output, (tuple, list)) else (output,)) if isinstance(o, torch.Tensor) and o.requires_grad]
        if outs:
            fired = [False]

            def pre_backward(_grad, _f=fired):
                if not _f[0]:
                    _f[0] = True
                    _pre_backward(state, u)
                return _grad

            for o in outs:
                o.register_hook(pre_backward)
    state.reshard(u)
    return output


def _pre_backward(state: FSDPState, u: FSDPUnit) -> None:
    state.in_backward = True
    if not state.final_callback_queued:
        state.final_callback_queued = True
        torch.autograd.Variable._execution_engine.queue_callback(state.finish_backward)
    state.wait_all_gather(u)
    state.prepare_grad_buffer(u)  # wgrad GEMMs of this unit write straight into it
    idx = u._index
    for k in range(1, state.prefetch + 1):
        if idx - k >= 0 and state.reshard_after_forward:
            state.launch_all_gather(state.units[idx - k])


class _ParamSwap:
    """While a unit is idle its module exposes the sharded fp32 DTensor parameters; forward/backward need the
    original (unsharded) Parameter objects.  The swap is a dict assignment per parameter."""

    def __init__(self, u: FSDPUnit, owners: List[Tuple[nn.Module, str]]):
        self.u = u
        self.owners = owners

    def to_sharded(self):
        for (m, n), sp in zip(self.owners, self.u.sharded_params):
            m._parameters[n] = sp

    def to_unsharded(self):
        for (m, n), p in zip(self.owners, self.u.params):
            m._parameters[n] = p


def checkpoint_module(module: nn.Module) -> nn.Module:
    """Full activation checkpointing of ``module`` (non-reentrant ``torch.utils.checkpoint``): call *before* ``fully_shard`` so
    that the FSDP hooks sit outside the checkpointed region — the recomputation then runs while the unit is already
    unsharded by the pre-backward hook and never triggers a gather of its own.  The Llama blocks additionally do selective
    recomputation of their cheap elementwise intermediates inside ``ops.functional``; use this for memory-bound configurations
    (e.g. 70B on 8 GPUs)."""
    from torch.utils.checkpoint import checkpoint

    if getattr(module, "_vb_checkpointed", False):
        return module
    if getattr(module, "_fsdp_unit", None) is not None:
        raise RuntimeError("checkpoint_module must be applied before fully_shard")
    orig = module.forward

    def forward(*args, **kwargs):
        if not torch.is_grad_enabled():
            return orig(*args, **kwargs)
        return checkpoint(orig, *args, use_reentrant=False, **kwargs)

    module.forward = forward
    module._vb_checkpointed = True
    return module


def get_fsdp_state(module: nn.Module) -> Optional[FSDPState]:
    return getattr(module, "_fsdp_state", None)


def fsdp_units(module: nn.Module) -> List[FSDPUnit]:
    st = None
    for m in module.modules():
        st = get_fsdp_state(m)
        if st is not None:
            break
    return st.units if st is not None else []


_STATES: Dict[int, FSDPState] = {}
from ...comm.symm import _COMM_CACHE  # noqa: E402  (shared arena cache; re-exported for FusedTP / MoE)


def fully_shard(
    module: nn.Module,
    mesh: Optional[DeviceMesh] = None,
    *,
    mesh_dim: int | str = 0,
    mp_policy: Optional[MixedPrecisionPolicy] = None,
    reshard_after_forward: Optional[bool] = None,
    prefetch: int = 1,
    comm_backend: str = "auto",
    block_rows: int = 1,
    granularity_fn=None,
    init_fn: Optional[Callable[[nn.Module], None]] = None,
    state: Optional[FSDPState] = None,
    fuse_first_gemm: bool | str | None = None,
) -> nn.Module:
    """Shard the parameters of ``module`` that are not already managed by an inner ``fully_shard``.

    Call bottom-up (inner blocks first, root last), like torch's FSDP2.  ``comm_backend``: ``"nccl"`` (c10d
    collectives; also what CPU/gloo tests use), ``"symm"`` (sm_100a symmetric-memory kernels) or ``"auto"``
    (symm on CUDA when the world has more than one rank and the extension is loaded).

    ``fuse_first_gemm``: name of the weight consumed by the module's first GEMM (``True`` = the module's
    ``fsdp_first_gemm_param`` attribute).  When this unit's forward all-gather is *exposed* (not prefetched), that weight
    is gathered by the GEMM kernel itself and the rest of the unit streams in behind it; needs ``comm_backend="symm"``
    and ``block_rows`` a multiple of 32 (forced to 32 when left at 1).
    """
    if fuse_first_gemm is True:
        fuse_first_gemm = getattr(module, "fsdp_first_gemm_param", None)
    if fuse_first_gemm and block_rows % 32:
        block_rows = 32 * block_rows
    if mesh is None:
        mesh = init_device_mesh("cuda" if torch.cuda.is_available() else "cpu", (dist.get_world_size() if dist.is_initialized() else 1,))
    md = mesh._dim_index(mesh_dim)
    mp_policy = mp_policy or MixedPrecisionPolicy()
    named = [(n, p) for n, p in module.named_parameters() if getattr(p, "_fsdp_unit", None) is None and not getattr(p, "_is_fsdp_sharded", False)]
    # find / create the shared state: inner units wrapped earlier each carry a state; merge them so the
    # whole model has one unit list in module-traversal (= forward) order
    if state is None:
        inner: List[FSDPState] = []
        for m in module.modules():
            s = get_fsdp_state(m)
            if s is not None and all(s is not t for t in inner):
                inner.append(s)
        if inner:
            state = inner[0]
            if len(inner) > 1:
                for m in module.modules():
                    if get_fsdp_state(m) is not None:
                        m._fsdp_state = state
                merged: List[FSDPUnit] = []
                for m in module.modules():
                    u_ = getattr(m, "_fsdp_unit", None)
                    if u_ is not None and all(u_ is not x for x in merged):
                        merged.append(u_)
                state.units = merged
                for i_, u_ in enumerate(merged):
                    u_._index = i_
                    u_._state = state
    dev = None
    for _, p in named:
        dev = p.device
        break
    if dev is None:
        dev = next(module.parameters()).device if any(True for _ in module.parameters()) else torch.device(mesh.device_type)
    if dev.type == "meta":
        tgt = torch.device("cuda", torch.cuda.current_device()) if mesh.device_type == "cuda" else torch.device("cpu")
        _materialize(module, [n for n, _ in named], tgt, init_fn)
        named = [(n, p) for n, p in module.named_parameters() if getattr(p, "_fsdp_unit", None) is None and not getattr(p, "_is_fsdp_sharded", False)]
        dev = tgt
    if state is None:
        state = FSDPState(mesh, md, dev)
        state.comm = _make_comm(comm_backend, mesh, md, dev)
    if reshard_after_forward is not None:
        state.reshard_after_forward = reshard_after_forward
    state.prefetch = prefetch
    module._fsdp_state = state
    if not named:
        return module
    # de-duplicate tied parameters
    seen, uniq = set(), []
    for n, p in named:
        if id(p) not in seen:
            seen.add(id(p))
            uniq.append((n, p))
    owners: List[Tuple[nn.Module, str]] = []
    for n, p in uniq:
        parts = n.split(".")
        m = module
        for a in parts[:-1]:
            m = getattr(m, a)
        owners.append((m, parts[-1]))
    u = FSDPUnit(module, uniq, mesh, md, mp_policy, name=type(module).__name__, comm=state.comm, block_rows=block_rows, granularity_fn=granularity_fn)
    u._swap = _ParamSwap(u, owners)
    u._state = state
    u.fused_first = fuse_first_gemm or None
    module._fsdp_unit = u
    # keep units in module-traversal order of the outermost wrapped module seen so far
    state.units.append(u)
    for i_, u_ in enumerate(state.units):
        u_._index = i_
    for sp in u.sharded_params:
        sp._is_fsdp_sharded = True
    u._swap.to_sharded()

    def pre(mod, args, kwargs):
        u._swap.to_unsharded()
        return _pre_forward(u._state, u, mod, args, kwargs)

    def post(mod, args, output):
        out = _post_forward(u._state, u, mod, args, output)
        if not torch.is_grad_enabled():
            u._swap.to_sharded()
        return out

    module.register_forward_pre_hook(pre, with_kwargs=True)
    module.register_forward_hook(post)
    return module


def _materialize(module: nn.Module, names: List[str], device: torch.device, init_fn) -> None:
    """Deferred init: allocate the meta parameters of this unit on ``device`` and initialise them
    (legacy ``initialize/deferred_init.py:98-182`` materialises only what the rank needs; here a unit at a
    time is materialised, sliced into the master shard, and freed)."""
    for n in names:
        parts = n.split(".")
        m = module
        for a in parts[:-1]:
            m = getattr(m, a)
        old = m._parameters[parts[-1]]
        m._parameters[parts[-1]] = nn.Parameter(torch.empty(old.shape, dtype=old.dtype, device=device), requires_grad=old.requires_grad)
    # meta buffers (routing tables, rope caches ...): allocate, then let the owning module fill them (``reset_buffers``)
    for sub in module.modules():
        touched = False
        for bn, b in list(sub._buffers.items()):
            if b is not None and b.device.type == "meta":
                sub._buffers[bn] = torch.zeros(b.shape, dtype=b.dtype, device=device)
                touched = True
        if touched and hasattr(sub, "reset_buffers"):
            sub.reset_buffers()
    if init_fn is not None:
        init_fn(module)
    elif hasattr(module, "reset_parameters"):
        module.reset_parameters()
    else:
        for m in module.modules():
            if m is not module and hasattr(m, "reset_parameters") and any(True for _ in m.parameters(recurse=False)):
                m.reset_parameters()


def _make_comm(backend: str, mesh: DeviceMesh, md: int, dev: torch.device):
    world = mesh.size(md)
    if backend == "nccl" or world == 1 or dev.type != "cuda":
        return None
    from ...ops import _ext

    if backend == "auto" and not _ext.available():
        return None
    from ...comm.symm import SymmUnitComm

    # one arena per (process group, device): every unit of every model on this mesh dim shares it
    key = (id(mesh.get_group(md)), dev.index)
    comm = _COMM_CACHE.get(key)
    if comm is None:
        comm = _COMM_CACHE[key] = SymmUnitComm(mesh, md, dev)
    return comm
