"""Reading ndtimeline frames off a byte stream (legacy ``ndtimeline/binary_protocol.py``).

The frame layout is ``sock_streamer``'s (``magic "NDTL" | u8 version | u8 kind | u16 rank | u32 step | u32 payload_len | payload``);
this module is the receiving side written against a bare ``recv(n) -> bytes`` callable plus a carry-over buffer, so the same code
reads from a socket, a pipe or a file: ``recv_and_validate(sock.recv, carry)`` returns one validated frame and leaves whatever was
read beyond it in ``carry`` for the next call."""
from __future__ import annotations

import gc
import io
import json
from typing import Any, Callable, Tuple

from .exceptions import ProtocolValidationError

__all__ = ["dumps", "loads", "dumps_fn", "loads_fn", "recv_to_buf", "read_or_recv", "recv_and_validate", "MAX_PAYLOAD_LEN"]

MAX_PAYLOAD_LEN = 128 << 20


def dumps(v: Any) -> bytes:
    return json.dumps(v, separators=(",", ":")).encode()


def loads(binary: bytes) -> Any:
    """Decoding a large record list allocates many small objects; the cyclic GC has nothing to find in them, so it is paused."""
    was = gc.isenabled()
    gc.disable()
    try:
        return json.loads(binary)
    finally:
        if was:
            gc.enable()


dumps_fn, loads_fn = dumps, loads


def recv_to_buf(size: int, recv: Callable[[int], bytes], preload_data: bytearray) -> bytes:
    """Exactly ``size`` bytes: first what ``preload_data`` holds (it must hold less than ``size``), then ``recv`` until complete.
    Bytes received beyond ``size`` go back into ``preload_data``.  ``BrokenPipeError`` if the peer closes mid-frame."""
    if len(preload_data) > size:
        raise ValueError("recv_to_buf: the carry-over buffer already holds more than was asked for (use read_or_recv)")
    buf = io.BytesIO()
    buf.write(preload_data)
    remaining = size - len(preload_data)
    del preload_data[:]
    while remaining > 0:
        chunk = recv(max(8192, min(remaining, 1 << 20)))
        if not chunk:
            raise BrokenPipeError("peer closed the stream in the middle of a frame")
        if len(chunk) <= remaining:
            buf.write(chunk)
            remaining -= len(chunk)
        else:
            buf.write(chunk[:remaining])
            preload_data.extend(chunk[remaining:])
            remaining = 0
    return buf.getvalue()


def read_or_recv(size: int, recv: Callable[[int], bytes], preload_data: bytearray) -> bytes:
    if len(preload_data) >= size:
        out = bytes(preload_data[:size])
        del preload_data[:size]
        return out
    return recv_to_buf(size, recv, preload_data)


def recv_and_validate(recv_func: Callable[[int], bytes], preload_data: bytearray) -> Tuple[int, int, int, bytes]:
    """One frame: ``(kind, rank, step, payload bytes)``.  Raises ``ProtocolValidationError`` on a bad magic / version / length and
    ``EOFError`` when the stream ends cleanly BETWEEN frames."""
    from .sock_streamer import _HDR, _MAGIC, _VERSION

    if not preload_data:
        first = recv_func(8192)
        if not first:
            raise EOFError("stream closed")
        preload_data.extend(first)
    magic, ver, kind, rank, step, n = _HDR.unpack(read_or_recv(_HDR.size, recv_func, preload_data))
    if magic != _MAGIC:
        raise ProtocolValidationError(f"ndtimeline stream: bad magic {magic!r}")
    if ver != _VERSION:
        raise ProtocolValidationError(f"ndtimeline stream: protocol version {ver}, this reader speaks {_VERSION}")
    if n > MAX_PAYLOAD_LEN:
        raise ProtocolValidationError(f"ndtimeline stream: payload of {n} bytes exceeds {MAX_PAYLOAD_LEN}")
    return kind, rank, step, read_or_recv(n, recv_func, preload_data)


def __getattr__(name):  # the sending half lives with the frame definition
    if name in ("encode_package", "serialize_to_package", "encode_frame", "decode_frames"):
        from . import sock_streamer

        return getattr(sock_streamer, name)
    raise AttributeError(name)
