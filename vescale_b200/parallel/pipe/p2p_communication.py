"""Stage-to-stage tensor exchange between neighbouring pipeline stage *meshes* (legacy ``pipe/p2p_communication.py``): the
function-level API that hand-written schedules call —

    x = recv_forward(shape, dtype, cur_mesh, prev_mesh)            # from the previous stage
    send_forward(y, cur_mesh, next_mesh)                           # to the next stage
    g = send_forward_recv_backward(y, shape, cur_mesh, next_mesh)  # ... and the seven other combinations

A stage is a sub-mesh (its TP x DP ranks); rank ``i`` of a stage talks to rank ``i`` of the neighbouring stage (same mesh
coordinate).  One exchange moves up to four tensors (to / from the previous / next stage):

* ``batch_p2p_comm=True``: one ``batch_isend_irecv`` group (NCCL fuses it into one kernel, no ordering hazard);
* otherwise individual ``isend`` / ``irecv`` ordered by stage parity — even stages post their sends first, odd stages their
  receives — so that two neighbours never both sit in a blocking send;
* ``overlap_p2p_comm=True``: requests are not waited here but parked in the module's send / receive queues; the schedule calls
  ``drain_send_reqs()`` / ``drain_recv_reqs(kind)`` when it needs the buffers (receives are waited right before the
  tensor's first use, sends before their buffers are reused);
* ``tensor_shape=None``: shapes travel first (``_communicate_shapes``: one 8 x int64 header per tensor, same four directions); both
  neighbours of a link must make the same choice (as with the reference's ``variable_seq_lengths``).

Every call is timed with ``ndtimeit_p2p`` under the predefined metric of its combinator.  ``PipeEngine`` does not use this module
(its ``P2PContext`` prefetches receives in pair order from the static schedule); it is the building block for custom schedules
and for the instruction programs of ``instruction_base.py``."""
from __future__ import annotations

from enum import Enum
from typing import List, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist

from ...profiler import ndtimeit_p2p, predefined

__all__ = [
    "OpType", "Shape", "reset_global_counter", "drain_send_reqs", "drain_recv_reqs", "check_nan", "peer_rank", "recv_forward", "recv_backward", "send_forward",
    "send_backward", "send_forward_recv_backward", "send_backward_recv_forward", "send_forward_recv_forward", "send_backward_recv_backward",
    "send_forward_backward_recv_forward_backward", "pending_counts",
]

Shape = Union[List[int], Tuple[int, ...], torch.Size]


class OpType(Enum):
    SEND, RECV_FWD, RECV_BWD = 0, 1, 2


# requests parked by overlap_p2p_comm=True, oldest first
_SEND_REQS: List = []
_RECV_FWD_REQS: List = []
_RECV_BWD_REQS: List = []
_COUNTER = {"calls": 0, "bytes_sent": 0, "bytes_received": 0}


def reset_global_counter() -> None:
    """Forget parked requests and statistics (start of a new mini-batch / after an aborted one)."""
    _SEND_REQS.clear()
    _RECV_FWD_REQS.clear()
    _RECV_BWD_REQS.clear()
    for k in _COUNTER:
        _COUNTER[k] = 0


def pending_counts() -> Tuple[int, int, int]:
    return len(_SEND_REQS), len(_RECV_FWD_REQS), len(_RECV_BWD_REQS)


_WAITED: dict = {}  # id -> request already waited through another queue; the reference keeps the id from being reused


def _wait_all(reqs: List) -> None:
    while reqs:
        r = reqs.pop(0)
        inner = getattr(r, "req", r)
        if r is not None and id(inner) not in _WAITED:
            _WAITED[id(inner)] = inner
            r.wait()
    if not (_SEND_REQS or _RECV_FWD_REQS or _RECV_BWD_REQS):
        _WAITED.clear()


def _wait_once(reqs: Sequence) -> None:
    """A batched group hands the same request objects out for every direction; a gloo work object must not be waited twice."""
    seen = set()
    for r in reqs:
        if r is not None and id(r) not in seen:
            seen.add(id(r))
            r.wait()


def drain_send_reqs() -> None:
    """Wait for every parked send (their source tensors may be reused afterwards)."""
    _wait_all(_SEND_REQS)


def drain_recv_reqs(drain_type: str = "all") -> None:
    """Wait for parked receives: ``"forward"`` (activations), ``"backward"`` (gradients) or ``"all"``."""
    if drain_type in ("all", "forward"):
        _wait_all(_RECV_FWD_REQS)
    if drain_type in ("all", "backward"):
        _wait_all(_RECV_BWD_REQS)
    if drain_type not in ("all", "forward", "backward"):
        raise ValueError(f"drain_type must be 'all', 'forward' or 'backward', got {drain_type!r}")


def check_nan(tensor_list: Sequence[Optional[torch.Tensor]], check: bool = False) -> None:
    """Debug aid: raise when a tensor about to be sent / just received holds a NaN (costs a device sync; off by default)."""
    if not check:
        return
    for i, t in enumerate(tensor_list):
        if t is not None and t.is_floating_point() and bool(torch.isnan(t).any()):
            raise FloatingPointError(f"NaN in pipeline p2p tensor #{i} of shape {tuple(t.shape)}")


def _ranks_of(mesh) -> List[int]:
    m = getattr(mesh, "mesh", mesh)
    return [int(r) for r in torch.as_tensor(m).reshape(-1).tolist()]


def peer_rank(local_rank: int, current_device_mesh, target_device_mesh) -> int:
    """Global rank in ``target_device_mesh`` at the coordinate ``local_rank`` has in ``current_device_mesh``."""
    cur, tgt = _ranks_of(current_device_mesh), _ranks_of(target_device_mesh)
    if len(cur) != len(tgt):
        raise ValueError(f"neighbouring stages must have the same number of ranks ({len(cur)} vs {len(tgt)})")
    return tgt[cur.index(int(local_rank))]


def _my_rank(mesh) -> int:
    return int(mesh.get_rank()) if hasattr(mesh, "get_rank") else dist.get_rank()


def _stage_parity(cur_mesh, prev_mesh, next_mesh) -> int:
    """0 / 1 alternating along the pipeline, derived without global knowledge: stages are laid out in increasing rank order, so
    the parity of (first rank of the stage // stage size) alternates between neighbours."""
    ranks = _ranks_of(cur_mesh)
    return (min(ranks) // max(1, len(ranks))) % 2


def _ops_for(send_prev, recv_prev_buf, send_next, recv_next_buf, prev_rank, next_rank, group) -> List[dist.P2POp]:
    ops = []
    if send_prev is not None:
        ops.append(dist.P2POp(dist.isend, send_prev, prev_rank, group))
    if recv_prev_buf is not None:
        ops.append(dist.P2POp(dist.irecv, recv_prev_buf, prev_rank, group))
    if send_next is not None:
        ops.append(dist.P2POp(dist.isend, send_next, next_rank, group))
    if recv_next_buf is not None:
        ops.append(dist.P2POp(dist.irecv, recv_next_buf, next_rank, group))
    return ops


def _batched_p2p_ops(send_prev, recv_prev_buf, send_next, recv_next_buf, prev_rank, next_rank, group=None):
    """All four directions as one ``batch_isend_irecv`` group; returns (send reqs, recv-from-prev reqs, recv-from-next reqs).  A
    batched group completes as a whole, so the same request list stands in for every direction."""
    ops = _ops_for(send_prev, recv_prev_buf, send_next, recv_next_buf, prev_rank, next_rank, group)
    reqs = dist.batch_isend_irecv(ops) if ops else []
    has_send = send_prev is not None or send_next is not None
    return (reqs if has_send else []), (reqs if recv_prev_buf is not None else []), (reqs if recv_next_buf is not None else [])


def _p2p_ops(send_prev, recv_prev_buf, send_next, recv_next_buf, prev_rank, next_rank, parity: int, group=None):
    """Individual requests in a deadlock-free order: even stages send-next, recv-prev, send-prev, recv-next; odd stages the mirror
    image (recv-prev, send-next, recv-next, send-prev), so every send meets a receive that was posted before or with it."""
    sends, rp, rn = [], [], []
    def s(t, r):
        sends.append(dist.isend(t, r, group))
    def r_(buf, r, out):
        out.append(dist.irecv(buf, r, group))
    if parity == 0:
        if send_next is not None:
            s(send_next, next_rank)
        if recv_prev_buf is not None:
            r_(recv_prev_buf, prev_rank, rp)
        if send_prev is not None:
            s(send_prev, prev_rank)
        if recv_next_buf is not None:
            r_(recv_next_buf, next_rank, rn)
    else:
        if recv_prev_buf is not None:
            r_(recv_prev_buf, prev_rank, rp)
        if send_next is not None:
            s(send_next, next_rank)
        if recv_next_buf is not None:
            r_(recv_next_buf, next_rank, rn)
        if send_prev is not None:
            s(send_prev, prev_rank)
    return sends, rp, rn


def _communicate_shapes(send_next, send_prev, recv_prev: bool, recv_next: bool, prev_rank, next_rank, batch: bool, parity: int, device, group=None):
    """Exchange the shapes of the tensors about to travel: (ndim, d0, d1, ...) padded to 8 int64 per tensor."""
    def enc(t):
        v = torch.zeros(8, dtype=torch.int64, device=device)
        if t is not None:
            v[0] = t.dim()
            v[1:1 + t.dim()] = torch.tensor(list(t.shape), dtype=torch.int64)
        return v
    sp = enc(send_prev) if send_prev is not None else None
    sn = enc(send_next) if send_next is not None else None
    rp = torch.zeros(8, dtype=torch.int64, device=device) if recv_prev else None
    rn = torch.zeros(8, dtype=torch.int64, device=device) if recv_next else None
    fn = _batched_p2p_ops if batch else (lambda *a, **k: _p2p_ops(*a, parity=parity, **k))
    a, b, c = fn(sp, rp, sn, rn, prev_rank, next_rank, group=group)
    _wait_once(list(a) + list(b) + list(c))
    dec = lambda v: None if v is None else tuple(int(x) for x in v[1:1 + int(v[0])].tolist())  # noqa: E731
    return dec(rp), dec(rn)


def _communicate(*, tensor_send_next: Optional[torch.Tensor], tensor_send_prev: Optional[torch.Tensor], current_device_mesh, prev_device_mesh=None, next_device_mesh=None,
                 recv_prev: bool = False, recv_next: bool = False, tensor_shape: Optional[Shape] = None, recv_prev_shape: Optional[Shape] = None,
                 recv_next_shape: Optional[Shape] = None, batch_p2p_comm: bool = True, overlap_p2p_comm: bool = False, dtype: Optional[torch.dtype] = None,
                 device=None, group=None, nan_check: bool = False):
    """The one exchange every combinator is a special case of.  Returns ``(tensor_from_prev, tensor_from_next, reqs)``; ``reqs`` is
    ``None`` unless ``overlap_p2p_comm`` (then the received tensors must not be read before ``drain_recv_reqs``)."""
    me = _my_rank(current_device_mesh)
    prev_rank = peer_rank(me, current_device_mesh, prev_device_mesh) if prev_device_mesh is not None else None
    next_rank = peer_rank(me, current_device_mesh, next_device_mesh) if next_device_mesh is not None else None
    if (tensor_send_prev is not None or recv_prev) and prev_rank is None:
        raise ValueError("exchange with the previous stage requested but prev_device_mesh is None")
    if (tensor_send_next is not None or recv_next) and next_rank is None:
        raise ValueError("exchange with the next stage requested but next_device_mesh is None")
    ref = tensor_send_next if tensor_send_next is not None else tensor_send_prev
    if device is None:
        device = ref.device if ref is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    dtype = dtype or (ref.dtype if ref is not None else torch.get_default_dtype())
    parity = _stage_parity(current_device_mesh, prev_device_mesh, next_device_mesh)
    rp_shape = recv_prev_shape if recv_prev_shape is not None else tensor_shape
    rn_shape = recv_next_shape if recv_next_shape is not None else tensor_shape
    # tensor_shape=None means "shapes travel first" (variable sequence lengths) and BOTH ends of a link must agree on that, so the
    # handshake runs whenever no shape was given — also on a pure send, whose peer is waiting for the header
    if tensor_shape is None and ((recv_prev and rp_shape is None) or (recv_next and rn_shape is None) or (tensor_send_next is not None and recv_next_shape is None)
                                 or (tensor_send_prev is not None and recv_prev_shape is None)):
        a, b = _communicate_shapes(tensor_send_next, tensor_send_prev, recv_prev, recv_next, prev_rank, next_rank, batch_p2p_comm, parity, device, group)
        rp_shape = a if rp_shape is None else rp_shape
        rn_shape = b if rn_shape is None else rn_shape
    check_nan([tensor_send_next, tensor_send_prev], nan_check)
    send_next = tensor_send_next.contiguous() if tensor_send_next is not None else None
    send_prev = tensor_send_prev.contiguous() if tensor_send_prev is not None else None
    from_prev = torch.empty(tuple(rp_shape), dtype=dtype, device=device) if recv_prev else None
    from_next = torch.empty(tuple(rn_shape), dtype=dtype, device=device) if recv_next else None
    if batch_p2p_comm:
        sends, rp, rn = _batched_p2p_ops(send_prev, from_prev, send_next, from_next, prev_rank, next_rank, group)
    else:
        sends, rp, rn = _p2p_ops(send_prev, from_prev, send_next, from_next, prev_rank, next_rank, parity, group)
    _COUNTER["calls"] += 1
    _COUNTER["bytes_sent"] += sum(t.numel() * t.element_size() for t in (send_next, send_prev) if t is not None)
    _COUNTER["bytes_received"] += sum(t.numel() * t.element_size() for t in (from_prev, from_next) if t is not None)
    if overlap_p2p_comm:
        # keep the source tensors alive until their sends are drained
        _SEND_REQS.extend(_Keep(r, (send_next, send_prev)) for r in sends)
        _RECV_FWD_REQS.extend(rp)
        _RECV_BWD_REQS.extend(rn)
        return from_prev, from_next, list(sends) + list(rp) + list(rn)
    _wait_once(list(sends) + list(rp) + list(rn))
    if device.type == "cuda" and batch_p2p_comm:
        torch.cuda.current_stream(device).synchronize()  # a batched group signals completion per group, not per tensor
    check_nan([from_prev, from_next], nan_check)
    return from_prev, from_next, None


class _Keep:
    """A request plus references that must outlive it."""
    __slots__ = ("req", "refs")

    def __init__(self, req, refs):
        self.req, self.refs = req, refs

    def wait(self):
        self.req.wait()
        self.refs = None


# ---- the nine combinators ---------------------------------------------------------------------------------------------------------------
def recv_forward(tensor_shape: Optional[Shape], recv_dtype: Optional[torch.dtype], current_device_mesh, peer_device_mesh=None, batch_p2p_comm: bool = True, **kw):
    """Activation from the previous stage; ``None`` on the first stage (``peer_device_mesh is None``)."""
    if peer_device_mesh is None:
        return None
    with ndtimeit_p2p(predefined.RECV_FORWARD, peer=peer_rank(_my_rank(current_device_mesh), current_device_mesh, peer_device_mesh)):
        t, _, _ = _communicate(tensor_send_next=None, tensor_send_prev=None, current_device_mesh=current_device_mesh, prev_device_mesh=peer_device_mesh, recv_prev=True,
                               tensor_shape=tensor_shape, batch_p2p_comm=batch_p2p_comm, dtype=recv_dtype, **kw)
    return t


def recv_backward(tensor_shape: Optional[Shape], recv_dtype: Optional[torch.dtype], current_device_mesh, peer_device_mesh=None, batch_p2p_comm: bool = True, **kw):
    """Output gradient from the next stage; ``None`` on the last stage."""
    if peer_device_mesh is None:
        return None
    with ndtimeit_p2p(predefined.RECV_BACKWARD, peer=peer_rank(_my_rank(current_device_mesh), current_device_mesh, peer_device_mesh)):
        _, t, _ = _communicate(tensor_send_next=None, tensor_send_prev=None, current_device_mesh=current_device_mesh, next_device_mesh=peer_device_mesh, recv_next=True,
                               tensor_shape=tensor_shape, batch_p2p_comm=batch_p2p_comm, dtype=recv_dtype, **kw)
    return t


def send_forward(output_tensor: torch.Tensor, current_device_mesh, peer_device_mesh=None, tensor_shape: Optional[Shape] = None, batch_p2p_comm: bool = True, **kw) -> None:
    if peer_device_mesh is None:
        return
    with ndtimeit_p2p(predefined.SEND_FORWARD, peer=peer_rank(_my_rank(current_device_mesh), current_device_mesh, peer_device_mesh)):
        _communicate(tensor_send_next=output_tensor, tensor_send_prev=None, current_device_mesh=current_device_mesh, next_device_mesh=peer_device_mesh,
                     tensor_shape=tensor_shape, batch_p2p_comm=batch_p2p_comm, **kw)


def send_backward(input_tensor_grad: torch.Tensor, current_device_mesh, peer_device_mesh=None, tensor_shape: Optional[Shape] = None, batch_p2p_comm: bool = True, **kw) -> None:
    if peer_device_mesh is None:
        return
    with ndtimeit_p2p(predefined.SEND_BACKWARD, peer=peer_rank(_my_rank(current_device_mesh), current_device_mesh, peer_device_mesh)):
        _communicate(tensor_send_next=None, tensor_send_prev=input_tensor_grad, current_device_mesh=current_device_mesh, prev_device_mesh=peer_device_mesh,
                     tensor_shape=tensor_shape, batch_p2p_comm=batch_p2p_comm, **kw)


def send_forward_recv_backward(output_tensor: torch.Tensor, tensor_shape: Optional[Shape], recv_dtype: Optional[torch.dtype], current_device_mesh, peer_device_mesh=None,
                               batch_p2p_comm: bool = True, **kw):
    """Steady-state 1F1B on the sending side: activation out, its gradient in, same neighbour (the next stage)."""
    if peer_device_mesh is None:
        return None
    with ndtimeit_p2p(predefined.SEND_FORWARD_RECV_BACKWARD, peer=peer_rank(_my_rank(current_device_mesh), current_device_mesh, peer_device_mesh)):
        _, g, _ = _communicate(tensor_send_next=output_tensor, tensor_send_prev=None, current_device_mesh=current_device_mesh, next_device_mesh=peer_device_mesh,
                               recv_next=True, tensor_shape=tensor_shape, batch_p2p_comm=batch_p2p_comm, dtype=recv_dtype, **kw)
    return g


def send_backward_recv_forward(input_tensor_grad: torch.Tensor, tensor_shape: Optional[Shape], recv_dtype: Optional[torch.dtype], current_device_mesh, peer_device_mesh=None,
                               batch_p2p_comm: bool = True, **kw):
    """Steady-state 1F1B on the receiving side: gradient out, next activation in, same neighbour (the previous stage)."""
    if peer_device_mesh is None:
        return None
    with ndtimeit_p2p(predefined.SEND_BACKWARD_RECV_FORWARD, peer=peer_rank(_my_rank(current_device_mesh), current_device_mesh, peer_device_mesh)):
        x, _, _ = _communicate(tensor_send_next=None, tensor_send_prev=input_tensor_grad, current_device_mesh=current_device_mesh, prev_device_mesh=peer_device_mesh,
                               recv_prev=True, tensor_shape=tensor_shape, batch_p2p_comm=batch_p2p_comm, dtype=recv_dtype, **kw)
    return x


def send_forward_recv_forward(output_tensor: Optional[torch.Tensor], recv_prev: bool, tensor_shape: Optional[Shape], current_device_mesh, prev_device_mesh=None,
                              next_device_mesh=None, send_dtype=None, batch_p2p_comm: bool = True, overlap_p2p_comm: bool = False, **kw):
    """Interleaved warm-up: pass an activation on and take the next one in.  With ``overlap_p2p_comm`` returns ``(tensor, reqs)``."""
    with ndtimeit_p2p("send-forward-recv-forward"):
        x, _, reqs = _communicate(tensor_send_next=output_tensor if next_device_mesh is not None else None, tensor_send_prev=None, current_device_mesh=current_device_mesh,
                                  prev_device_mesh=prev_device_mesh, next_device_mesh=next_device_mesh, recv_prev=recv_prev and prev_device_mesh is not None,
                                  tensor_shape=tensor_shape, batch_p2p_comm=batch_p2p_comm, overlap_p2p_comm=overlap_p2p_comm, dtype=send_dtype, **kw)
    return (x, reqs) if overlap_p2p_comm else x


def send_backward_recv_backward(input_tensor_grad: Optional[torch.Tensor], recv_next: bool, tensor_shape: Optional[Shape], current_device_mesh, prev_device_mesh=None,
                                next_device_mesh=None, send_dtype=None, batch_p2p_comm: bool = True, overlap_p2p_comm: bool = False, **kw):
    """Interleaved cool-down: pass a gradient back and take the next one in."""
    with ndtimeit_p2p("send-backward-recv-backward"):
        _, g, reqs = _communicate(tensor_send_next=None, tensor_send_prev=input_tensor_grad if prev_device_mesh is not None else None, current_device_mesh=current_device_mesh,
                                  prev_device_mesh=prev_device_mesh, next_device_mesh=next_device_mesh, recv_next=recv_next and next_device_mesh is not None,
                                  tensor_shape=tensor_shape, batch_p2p_comm=batch_p2p_comm, overlap_p2p_comm=overlap_p2p_comm, dtype=send_dtype, **kw)
    return (g, reqs) if overlap_p2p_comm else g


def send_forward_backward_recv_forward_backward(output_tensor: Optional[torch.Tensor], input_tensor_grad: Optional[torch.Tensor], recv_prev: bool, recv_next: bool,
                                                tensor_shape: Optional[Shape], current_device_mesh, prev_device_mesh=None, next_device_mesh=None, send_dtype=None,
                                                batch_p2p_comm: bool = True, overlap_p2p_comm: bool = False, **kw):
    """Interleaved steady state: all four directions in one exchange.  Returns ``(activation_in, gradient_in[, reqs])``."""
    with ndtimeit_p2p("send-forward-backward-recv-forward-backward"):
        x, g, reqs = _communicate(tensor_send_next=output_tensor if next_device_mesh is not None else None,
                                  tensor_send_prev=input_tensor_grad if prev_device_mesh is not None else None, current_device_mesh=current_device_mesh,
                                  prev_device_mesh=prev_device_mesh, next_device_mesh=next_device_mesh, recv_prev=recv_prev and prev_device_mesh is not None,
                                  recv_next=recv_next and next_device_mesh is not None, tensor_shape=tensor_shape, batch_p2p_comm=batch_p2p_comm,
                                  overlap_p2p_comm=overlap_p2p_comm, dtype=send_dtype, **kw)
    return (x, g, reqs) if overlap_p2p_comm else (x, g)
