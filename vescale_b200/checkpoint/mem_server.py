"""In-memory checkpoint file server + report service over gRPC, and the DCP storage adapters that write to / read from it.

A training rank finishes ``checkpoint.save("mem://127.0.0.1:PORT/ckpt/step100", ...)`` as soon as its shards are in the
node-local server's memory; the server persists them to disk in the background (``persist``) and can serve them back for a
fast restart (``load`` from the same ``mem://`` address), so slow storage never blocks training.

Capability parity: legacy ``checkpoint/utilities/server/*.proto`` + ``mem_server_lib.py`` (in-memory file server: Write /
Read / Rename / Remove / Listdir / Exists) and ``server_lib.py`` (report service: gather per-rank status); the wire format
here is our own — gRPC *generic* handlers with raw-bytes messages ``u32 header_len | JSON header | payload`` (no protoc
step), streamed in 4 MB chunks.
"""
from __future__ import annotations

import io
import json
import os
import struct
import threading
import time
from concurrent import futures
from contextlib import contextmanager
from typing import Dict, Iterable, Iterator, List, Optional, Tuple

__all__ = ["MemFileServer", "MemFileClient", "MemFileSystem", "parse_mem_uri", "make_mem_writer", "make_mem_reader"]

_SERVICE = "vescale_b200.MemFile"
_CHUNK = 4 << 20
_ident = lambda b: b  # noqa: E731  (bytes in, bytes out)


def _pack(header: dict, payload: bytes = b"") -> bytes:
    h = json.dumps(header, separators=(",", ":")).encode()
    return struct.pack("<I", len(h)) + h + payload


def _unpack(msg: bytes) -> Tuple[dict, bytes]:
    (n,) = struct.unpack_from("<I", msg, 0)
    return json.loads(msg[4 : 4 + n]), msg[4 + n :]


class _Store:
    def __init__(self):
        self.files: Dict[str, bytes] = {}
        self.dirs = set()
        self.reports: Dict[str, dict] = {}
        self.lock = threading.Lock()
        self.bytes_written = 0


class MemFileServer:
    """``MemFileServer(port=0).start()``; ``.address`` is ``host:port``.  One per node (or per job)."""

    def __init__(self, host: str = "127.0.0.1", port: int = 0, max_workers: int = 8):
        import grpc

        self._grpc = grpc
        self.store = _Store()
        opts = [("grpc.max_send_message_length", 64 << 20), ("grpc.max_receive_message_length", 64 << 20)]
        self.server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers), options=opts)
        h = {
            "Write": grpc.stream_unary_rpc_method_handler(self._write, _ident, _ident),
            "Read": grpc.unary_stream_rpc_method_handler(self._read, _ident, _ident),
            "Call": grpc.unary_unary_rpc_method_handler(self._call, _ident, _ident),
        }
        self.server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(_SERVICE, h),))
        self.port = self.server.add_insecure_port(f"{host}:{port}")
        self.host = host

    @property
    def address(self) -> str:
        return f"{self.host}:{self.port}"

    def start(self) -> "MemFileServer":
        self.server.start()
        return self

    def stop(self, grace: Optional[float] = 0.5) -> None:
        self.server.stop(grace)

    # ------------------------------------------------------------------ handlers
    def _write(self, request_iterator: Iterable[bytes], context) -> bytes:
        name, parts = None, []
        for msg in request_iterator:
            hdr, payload = _unpack(msg)
            name = hdr.get("name", name)
            parts.append(payload)
        data = b"".join(parts)
        with self.store.lock:
            self.store.files[name] = data
            self.store.bytes_written += len(data)
        return _pack({"ok": True, "size": len(data)})

    def _read(self, request: bytes, context) -> Iterator[bytes]:
        hdr, _ = _unpack(request)
        with self.store.lock:
            data = self.store.files.get(hdr["name"])
        if data is None:
            yield _pack({"ok": False, "error": "not found"})
            return
        if not data:
            yield _pack({"ok": True, "size": 0})
        for off in range(0, len(data), _CHUNK):
            yield _pack({"ok": True, "size": len(data)}, data[off : off + _CHUNK])

    def _call(self, request: bytes, context) -> bytes:
        hdr, _ = _unpack(request)
        op, st = hdr["op"], self.store
        with st.lock:
            if op == "exists":
                n = hdr["name"]
                return _pack({"ok": True, "exists": n in st.files or n in st.dirs or any(k.startswith(n.rstrip("/") + "/") for k in st.files)})
            if op == "rename":
                if hdr["src"] not in st.files:
                    return _pack({"ok": False, "error": "not found"})
                st.files[hdr["dst"]] = st.files.pop(hdr["src"])
                return _pack({"ok": True})
            if op == "remove":
                st.files.pop(hdr["name"], None)
                return _pack({"ok": True})
            if op == "mkdir":
                st.dirs.add(hdr["name"].rstrip("/"))
                return _pack({"ok": True})
            if op == "listdir":
                pre = hdr["name"].rstrip("/") + "/"
                return _pack({"ok": True, "names": sorted(k for k in st.files if k.startswith(pre))})
            if op == "report":  # report service: ranks post status, anyone can read the table
                if "status" in hdr:
                    st.reports[str(hdr["who"])] = {"status": hdr["status"], "time": time.time(), **hdr.get("extra", {})}
                return _pack({"ok": True, "reports": st.reports, "files": len(st.files), "bytes": sum(len(v) for v in st.files.values())})
            if op == "persist":
                pre, dst = hdr["name"].rstrip("/") + "/", hdr["dst"]
                todo = [(k, v) for k, v in st.files.items() if k.startswith(pre)]
        if op == "persist":
            for k, v in todo:  # outside the lock: disk I/O
                out = os.path.join(dst, k[len(pre) :])
                os.makedirs(os.path.dirname(out), exist_ok=True)
                with open(out + ".tmp", "wb") as f:
                    f.write(v)
                os.replace(out + ".tmp", out)
            return _pack({"ok": True, "files": len(todo)})
        return _pack({"ok": False, "error": f"unknown op {op}"})


class MemFileClient:
    def __init__(self, address: str, timeout: float = 60.0):
        import grpc

        opts = [("grpc.max_send_message_length", 64 << 20), ("grpc.max_receive_message_length", 64 << 20)]
        self.channel = grpc.insecure_channel(address, options=opts)
        self.timeout = timeout
        self._write = self.channel.stream_unary(f"/{_SERVICE}/Write", request_serializer=_ident, response_deserializer=_ident)
        self._read = self.channel.unary_stream(f"/{_SERVICE}/Read", request_serializer=_ident, response_deserializer=_ident)
        self._call = self.channel.unary_unary(f"/{_SERVICE}/Call", request_serializer=_ident, response_deserializer=_ident)

    def write(self, name: str, data: bytes) -> int:
        def gen():
            if not data:
                yield _pack({"name": name})
            for off in range(0, len(data), _CHUNK):
                yield _pack({"name": name}, data[off : off + _CHUNK])

        hdr, _ = _unpack(self._write(gen(), timeout=self.timeout))
        return hdr["size"]

    def read(self, name: str) -> bytes:
        parts = []
        for msg in self._read(_pack({"name": name}), timeout=self.timeout):
            hdr, payload = _unpack(msg)
            if not hdr["ok"]:
                raise FileNotFoundError(name)
            parts.append(payload)
        return b"".join(parts)

    def call(self, op: str, **kw) -> dict:
        hdr, _ = _unpack(self._call(_pack({"op": op, **kw}), timeout=self.timeout))
        if not hdr.get("ok"):
            raise OSError(hdr.get("error", "mem file server error"))
        return hdr

    def exists(self, name: str) -> bool:
        return self.call("exists", name=name)["exists"]

    def listdir(self, name: str) -> List[str]:
        return self.call("listdir", name=name)["names"]

    def remove(self, name: str) -> None:
        self.call("remove", name=name)

    def rename(self, src: str, dst: str) -> None:
        self.call("rename", src=src, dst=dst)

    def persist(self, name: str, dst_dir: str) -> int:
        """Write every file under ``name/`` to ``dst_dir`` on the server's disk (atomic per file)."""
        return self.call("persist", name=name, dst=dst_dir)["files"]

    def report(self, who=None, status: Optional[str] = None, **extra) -> dict:
        kw = {"who": who, "status": status, "extra": extra} if status is not None else {}
        return self.call("report", **kw)

    def close(self) -> None:
        self.channel.close()


# --------------------------------------------------------------------------- DCP adapters
def parse_mem_uri(uri: str) -> Optional[Tuple[str, str]]:
    """``mem://host:port/some/dir`` -> (``host:port``, ``some/dir``); None for ordinary paths."""
    if not isinstance(uri, str) or not uri.startswith("mem://"):
        return None
    rest = uri[len("mem://") :]
    addr, _, path = rest.partition("/")
    return addr, path.strip("/")


class _UploadOnClose(io.BytesIO):
    def __init__(self, client: MemFileClient, name: str):
        super().__init__()
        self._client, self._name = client, name

    def close(self):
        if not self.closed:
            self._client.write(self._name, self.getvalue())
        super().close()


def _mem_fs_class():
    from torch.distributed.checkpoint.filesystem import FileSystemBase

    class MemFileSystem(FileSystemBase):
        """``torch.distributed.checkpoint`` file-system shim over a ``MemFileClient`` (paths are server-side names)."""

        def __init__(self, client: MemFileClient):
            self.client = client

        @contextmanager
        def create_stream(self, path, mode: str):
            path = str(path)
            if "w" in mode:
                s = _UploadOnClose(self.client, path)
                try:
                    yield s
                finally:
                    s.close()
            else:
                s = io.BytesIO(self.client.read(path))
                try:
                    yield s
                finally:
                    s.close()

        def concat_path(self, path, suffix: str):
            return f"{str(path).rstrip('/')}/{suffix}"

        def rename(self, path, new_path) -> None:
            self.client.rename(str(path), str(new_path))

        def init_path(self, path):
            return str(path)

        def mkdir(self, path) -> None:
            self.client.call("mkdir", name=str(path))

        @classmethod
        def validate_checkpoint_id(cls, checkpoint_id) -> bool:
            return isinstance(checkpoint_id, str)

        def exists(self, path) -> bool:
            return self.client.exists(str(path))

        def rm_file(self, path) -> None:
            self.client.remove(str(path))

        def ls(self, path):
            return self.client.listdir(str(path))

    return MemFileSystem


def MemFileSystem(client: MemFileClient):  # noqa: N802  (factory with a class-like name; the class needs torch at import)
    return _mem_fs_class()(client)


def make_mem_writer(uri: str):
    """A DCP ``StorageWriter`` that lands every file of the checkpoint in the memory server named by ``uri``."""
    from torch.distributed.checkpoint.filesystem import FileSystemWriter

    addr, path = parse_mem_uri(uri)
    client = MemFileClient(addr)
    w = FileSystemWriter(path, sync_files=False)
    w.fs = MemFileSystem(client)
    w.path = w.fs.init_path(path)
    return w


def make_mem_reader(uri: str):
    from torch.distributed.checkpoint.filesystem import FileSystemReader

    addr, path = parse_mem_uri(uri)
    client = MemFileClient(addr)
    r = FileSystemReader(path)
    r.fs = MemFileSystem(client)
    r.path = r.fs.init_path(path)
    return r
