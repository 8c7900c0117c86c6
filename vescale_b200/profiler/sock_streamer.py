"""Unix-socket streaming of timeline records to one collector process per host.

``SockNDHandler`` (in every training rank) frames each flush with a small binary header and sends it to the
``NDtimelineStreamer`` process, which hands decoded records to ordinary ``NDHandler``s (e.g. one merged perfetto file per
host) — so trace serialisation and file I/O never run inside a training process.

Capability parity: legacy ``ndtimeline/sock_streamer.py:94-132`` (streamer process), ``handlers/sock_handler.py`` and
``binary_protocol.py`` (length-prefixed frames with a magic and a version); the framing here is our own:

    frame := magic "NDTL" | u8 version | u8 kind | u16 rank | u32 step | u32 payload_len | payload (UTF-8 JSON array)
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import queue
import socket
import socketserver
import struct
import threading
from typing import Callable, List, Optional, Sequence

from .exceptions import ProtocolValidationError
from .handlers import NDHandler

__all__ = ["encode_frame", "decode_frames", "SockNDHandler", "NDtimelineStreamer", "dumps_fn", "loads_fn", "encode_package", "serialize_to_package", "SOCK_PARENT_DIR", "SOCK_PATH", "SOCK_TIMEOUT_CLIENT", "MsgHandler", "internal_queue_consume"]

SOCK_PARENT_DIR = os.environ.get("VESCALE_NDTIMELINE_SOCK_DIR", "/tmp/ndtimeline")
SOCK_PATH = os.path.join(SOCK_PARENT_DIR, "ndtimeline.sock")  # default collector socket of this host
SOCK_TIMEOUT_CLIENT = 10.0  # seconds a training rank waits for the collector to accept

_MAGIC = b"NDTL"
_VERSION = 1
_HDR = struct.Struct("<4sBBHII")
KIND_RECORDS, KIND_CLOSE = 0, 1


def encode_frame(records: List[dict], rank: int, step: int, kind: int = KIND_RECORDS) -> bytes:
    payload = json.dumps(records, separators=(",", ":")).encode()
    return _HDR.pack(_MAGIC, _VERSION, kind, rank & 0xFFFF, step & 0xFFFFFFFF, len(payload)) + payload


# generic payload helpers under the reference's names (legacy ``binary_protocol.py:75-90``, ``sock_streamer.py``)
def dumps_fn(v) -> bytes:
    return json.dumps(v, separators=(",", ":")).encode()


def loads_fn(b: bytes):
    return json.loads(b)


def encode_package(payload: bytes, rank: int = 0, step: int = 0, kind: int = KIND_RECORDS) -> bytes:
    """Frame an already serialised payload."""
    return _HDR.pack(_MAGIC, _VERSION, kind, rank & 0xFFFF, step & 0xFFFFFFFF, len(payload)) + payload


def serialize_to_package(v, rank: int = 0, step: int = 0) -> bytes:
    return encode_package(dumps_fn(v), rank, step)


def decode_frames(buf: bytearray):
    """Yield (kind, rank, step, records) for every complete frame at the head of ``buf`` and consume it."""
    while len(buf) >= _HDR.size:
        magic, ver, kind, rank, step, n = _HDR.unpack_from(buf, 0)
        if magic != _MAGIC or ver != _VERSION:
            raise ProtocolValidationError("ndtimeline stream: bad frame header")
        if len(buf) < _HDR.size + n:
            return
        payload = bytes(buf[_HDR.size : _HDR.size + n])
        del buf[: _HDR.size + n]
        yield kind, rank, step, json.loads(payload)


class SockNDHandler(NDHandler):
    def __init__(self, sock_path: str, connect_timeout: float = 10.0):
        self.sock_path = sock_path
        self.sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self.sock.settimeout(connect_timeout)
        self.sock.connect(sock_path)
        self.sock.settimeout(None)
        self._lock = threading.Lock()

    def __call__(self, records, rank, step):
        clean = [{k: v for k, v in r.items() if k not in ("e0", "e1")} for r in records]
        with self._lock:
            self.sock.sendall(encode_frame(clean, rank, step))

    def close(self, rank: int = 0):
        with self._lock:
            try:
                self.sock.sendall(encode_frame([], rank, 0, KIND_CLOSE))
            finally:
                self.sock.close()


class MsgHandler(socketserver.BaseRequestHandler):
    """One connection of the collector (one training rank): read frames, validate, queue them.  Nothing else happens on this thread —
    the NDHandlers (file writes, trace merging) run on the consumer thread, so a slow handler fills the queue, not the socket, and the
    training rank's ``sendall`` keeps returning immediately."""

    def handle(self) -> None:
        from .binary_protocol import loads_fn as _loads, recv_and_validate

        carry = bytearray()
        try:
            while True:
                try:
                    kind, rank, step, payload = recv_and_validate(self.request.recv, carry)
                except (EOFError, BrokenPipeError):
                    return
                if kind == KIND_CLOSE:
                    return
                self.server.queue.put((rank, step, _loads(payload)))
        except ProtocolValidationError as e:
            self.server.queue.put(e)
        finally:
            self.server.queue.put(_CLIENT_DONE)


_CLIENT_DONE = object()


def internal_queue_consume(q: "queue.Queue", handlers: Sequence[NDHandler], expected_clients: int) -> int:
    """Consumer loop: hand queued record batches to the handlers until ``expected_clients`` connections have finished.  Returns the
    number of batches processed.  A protocol error from a connection is re-raised here, after that connection was counted."""
    done = n = 0
    err = None
    while done < expected_clients:
        item = q.get()
        if item is _CLIENT_DONE:
            done += 1
        elif isinstance(item, Exception):
            err = item
        else:
            rank, step, recs = item
            for h in handlers:
                h(recs, rank, step)
            n += 1
    if err is not None:
        raise err
    return n


class _Collector(socketserver.ThreadingMixIn, socketserver.UnixStreamServer):
    daemon_threads = True
    request_queue_size = 128


def _serve(sock_path: str, make_handlers: Callable[[], Sequence[NDHandler]], expected_clients: int, ready) -> None:
    handlers = list(make_handlers())
    if os.path.exists(sock_path):
        os.unlink(sock_path)
    srv = _Collector(sock_path, MsgHandler)
    srv.queue = queue.Queue()
    t = threading.Thread(target=srv.serve_forever, kwargs={"poll_interval": 0.05}, daemon=True)
    t.start()
    ready.set()
    try:
        internal_queue_consume(srv.queue, handlers, expected_clients)
    finally:
        srv.shutdown()
        srv.server_close()
        if os.path.exists(sock_path):
            os.unlink(sock_path)


class NDtimelineStreamer:
    """Collector process: ``NDtimelineStreamer.start(path, make_handlers, n_clients)`` on local rank 0, then every rank adds a
    ``SockNDHandler(path)`` to ``init_ndtimers``.  ``make_handlers`` must be picklable (it runs in the child)."""

    def __init__(self, proc, sock_path: str):
        self.proc, self.sock_path = proc, sock_path

    @classmethod
    def start(cls, sock_path: str, make_handlers: Callable[[], Sequence[NDHandler]], expected_clients: int = 1, timeout: float = 20.0) -> "NDtimelineStreamer":
        ctx = mp.get_context("spawn")
        ready = ctx.Event()
        proc = ctx.Process(target=_serve, args=(sock_path, make_handlers, expected_clients, ready), daemon=True)
        proc.start()
        if not ready.wait(timeout):
            proc.terminate()
            raise RuntimeError("ndtimeline streamer did not come up")
        return cls(proc, sock_path)

    def join(self, timeout: Optional[float] = 30.0) -> None:
        self.proc.join(timeout)
        if self.proc.is_alive():
            self.proc.terminate()
