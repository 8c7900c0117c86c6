"""Out-of-band coordination service for asynchronous checkpointing (legacy ``checkpoint/utilities/server/server_lib.py`` +
``report_service.proto``).

The background half of an asynchronous save must agree across ranks on plans and write results while the TRAINING threads own the
NCCL communicators: a collective issued from a checkpoint thread on a group the training loop also uses can interleave with it and
deadlock.  This service gives the checkpoint threads their own channel: one small gRPC server per job (``start_server_in_new_process``)
and three primitives on top of it —

    gather(stub, gather_rank, rank, obj, tag)     every rank contributes ``obj``; ``gather_rank`` gets the list, others ``None``
    broadcast(stub, src_rank, rank, obj, tag)     ``src_rank`` contributes, everyone gets it
    barrier(stub, rank, tag)                      everyone waits for everyone

— matched by ``tag`` (default: the call site, so that distinct calls never mix even when ranks drift apart), blocking server-side
until the group of ``world_size`` is complete.  ``get_server_status`` shows which tag is waiting for which ranks: the first thing to
look at when a save hangs.  Wire format: gRPC generic handlers, pickled payloads (trusted peers of one job only)."""
from __future__ import annotations

import inspect
import multiprocessing as mp
import pickle
import socket
import threading
import time
from concurrent import futures
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

__all__ = ["Item", "ReportServicer", "serve", "start_server_in_new_process", "get_stub", "gather", "broadcast", "barrier", "get_server_status", "ReportStub"]

_SERVICE = "vescale_b200.CheckpointReport"
_ident = lambda b: b  # noqa: E731


@dataclass
class Item:
    """One in-flight rendezvous: contributions so far, who is still missing, a condition to sleep on."""
    cv: threading.Condition = field(default_factory=threading.Condition)
    contents: Dict[int, Any] = field(default_factory=dict)
    ranks: set = field(default_factory=set)
    done: int = 0  # ranks that have picked up the result (the item is deleted when all have)
    created: float = field(default_factory=time.time)


class ReportServicer:
    def __init__(self, world_size: int):
        self.world_size = int(world_size)
        self._lock = threading.Lock()
        self._items: Dict[str, Item] = {}
        self.completed = 0

    def _item(self, tag: str) -> Item:
        with self._lock:
            it = self._items.get(tag)
            if it is None:
                it = self._items[tag] = Item()
            return it

    def _retire(self, tag: str, it: Item, n: int) -> None:
        it.done += 1
        if it.done == n:
            with self._lock:
                self._items.pop(tag, None)
                self.completed += 1

    def Gather(self, req: dict) -> dict:  # noqa: N802 (rpc method names)
        tag, rank = req["tag"], int(req["rank"])
        n = int(req.get("world") or self.world_size)  # a sub-group (one pipeline stage's ranks, numbered 0..n-1) may rendezvous on its own
        it = self._item(tag)
        with it.cv:
            if rank in it.ranks:
                return {"ok": False, "error": f"rank {rank} joined rendezvous {tag!r} twice"}
            it.ranks.add(rank)
            it.contents[rank] = req.get("content")
            if len(it.ranks) == n:
                it.cv.notify_all()
            elif not it.cv.wait_for(lambda: len(it.ranks) == n, timeout=req.get("timeout")):
                missing = sorted(set(range(n)) - it.ranks)
                return {"ok": False, "error": f"rendezvous {tag!r} timed out waiting for ranks {missing}"}
            out = [it.contents[r] for r in range(n)] if req.get("with_result") else None
            self._retire(tag, it, n)
        return {"ok": True, "contents": out}

    def Broadcast(self, req: dict) -> dict:  # noqa: N802
        tag, rank, src = req["tag"], int(req["rank"]), int(req["src_rank"])
        n = int(req.get("world") or self.world_size)
        it = self._item(tag)
        with it.cv:
            it.ranks.add(rank)
            if rank == src:
                it.contents[src] = req.get("content")
                it.cv.notify_all()
            elif not it.cv.wait_for(lambda: src in it.contents, timeout=req.get("timeout")):
                return {"ok": False, "error": f"broadcast {tag!r} timed out waiting for its source rank {src}"}
            out = it.contents[src]
            self._retire(tag, it, n)
        return {"ok": True, "content": out}

    def GetStatus(self, req: dict) -> dict:  # noqa: N802
        with self._lock:
            waiting = {tag: {"have": sorted(it.ranks), "missing": sorted(set(range(self.world_size)) - it.ranks), "age_s": round(time.time() - it.created, 3)} for tag, it in self._items.items()}
        return {"ok": True, "world_size": self.world_size, "waiting": waiting, "completed": self.completed}

    def _handle(self, request: bytes, context) -> bytes:
        req = pickle.loads(request)
        try:
            resp = getattr(self, req["method"])(req)
        except Exception as e:  # noqa: BLE001
            resp = {"ok": False, "error": f"{type(e).__name__}: {e}"}
        return pickle.dumps(resp)


def _local_ip() -> str:
    try:
        return socket.gethostbyname(socket.gethostname())
    except OSError:
        return "127.0.0.1"


def serve(servicer: ReportServicer, host: str = "127.0.0.1", port: int = 0, max_workers: Optional[int] = None):
    """Start the service in this process; returns ``(server, "host:port")``.  One worker thread per rank can be parked in a
    rendezvous at a time, hence the pool size."""
    import grpc

    opts = [("grpc.max_send_message_length", 256 << 20), ("grpc.max_receive_message_length", 256 << 20)]
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers or max(8, 2 * servicer.world_size + 2)), options=opts)
    h = {"Call": grpc.unary_unary_rpc_method_handler(servicer._handle, _ident, _ident)}
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(_SERVICE, h),))
    bound = server.add_insecure_port(f"{host}:{port}")
    server.start()
    return server, f"{host}:{bound}"


def _serve_in_loop(world_size: int, conn, host: str) -> None:
    server, addr = serve(ReportServicer(world_size), host=host)
    conn.send(addr)
    conn.close()
    server.wait_for_termination()


def start_server_in_new_process(world_size: int, host: str = "127.0.0.1") -> str:
    """A daemon process that serves until the parent exits; returns its address (hand it to the other ranks, e.g. through one
    ``broadcast_object_list`` at start-up)."""
    ctx = mp.get_context("spawn")
    parent, child = ctx.Pipe()
    p = ctx.Process(target=_serve_in_loop, args=(world_size, child, host), daemon=True)
    p.start()
    from .mem_server import _recv_or_die

    addr = _recv_or_die(parent, p, "checkpoint report service")
    parent.close()
    start_server_in_new_process.processes.append(p)  # keep a handle: the process dies with us, tests terminate it explicitly
    return addr


start_server_in_new_process.processes = []  # type: ignore[attr-defined]


class ReportStub:
    def __init__(self, addr: str):
        import grpc

        opts = [("grpc.max_send_message_length", 256 << 20), ("grpc.max_receive_message_length", 256 << 20)]
        self.addr = addr
        self.channel = grpc.insecure_channel(addr, options=opts)
        self._call = self.channel.unary_unary(f"/{_SERVICE}/Call", request_serializer=_ident, response_deserializer=_ident)

    def call(self, method: str, timeout: Optional[float] = None, **kw) -> dict:
        resp = pickle.loads(self._call(pickle.dumps({"method": method, "timeout": timeout, **kw}), timeout=None if timeout is None else timeout + 5.0))
        if not resp.get("ok"):
            raise RuntimeError(resp.get("error", "checkpoint report service error"))
        return resp

    def close(self) -> None:
        self.channel.close()


def get_stub(addr: str) -> ReportStub:
    return ReportStub(addr)


def _get_tag() -> str:
    """The caller's caller: file:line of the ``gather`` / ``broadcast`` / ``barrier`` call — identical on every rank that executes the
    same code path, different for different call sites."""
    fr = inspect.stack()[2]
    return f"{fr.filename}:{fr.lineno}"


def gather(stub: ReportStub, gather_rank: int, rank: int, obj: Any, tag: Optional[str] = None, timeout: Optional[float] = None, world: Optional[int] = None) -> Optional[List[Any]]:
    tag = tag if tag is not None else _get_tag()
    return stub.call("Gather", timeout, tag=f"gather/{tag}", rank=rank, content=obj, with_result=(rank == gather_rank), world=world)["contents"]


def broadcast(stub: ReportStub, src_rank: int, rank: int, obj: Any = None, tag: Optional[str] = None, timeout: Optional[float] = None, world: Optional[int] = None) -> Any:
    tag = tag if tag is not None else _get_tag()
    return stub.call("Broadcast", timeout, tag=f"broadcast/{tag}", rank=rank, src_rank=src_rank, content=obj if rank == src_rank else None, world=world)["content"]


def barrier(stub: ReportStub, rank: int, tag: Optional[str] = None, timeout: Optional[float] = None, world: Optional[int] = None) -> None:
    tag = tag if tag is not None else _get_tag()
    stub.call("Gather", timeout, tag=f"barrier/{tag}", rank=rank, content=None, with_result=False, world=world)


def get_server_status(stub: ReportStub) -> dict:
    return stub.call("GetStatus")
