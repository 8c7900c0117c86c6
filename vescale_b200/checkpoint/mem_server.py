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

__all__ = ["MemFileServer", "MemFileClient", "MemFileSystem", "parse_mem_uri", "make_mem_writer", "make_mem_reader", "MemFileServicer", "get_mem_server_sock_file", "get_prefix",
           "start_server", "start_server_in_new_process", "wait_until_fs_ready", "mem_open", "rename", "remove", "listdir", "exists"]

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
                pre = (hdr["name"].rstrip("/") + "/") if hdr["name"].strip("/") else ""
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


# ---- named servers + a file API on ``/local_mem/<name>/...`` paths (legacy ``mem_server_lib.py:48-307``) ------------------------------------------
# The reference addresses a node-local server by NAME (a unix socket under /var/tmp) and files as ``/local_mem/<name>/<path>``.
# Same model here: ``start_server(name)`` writes the server's TCP address to a small rendezvous file; the module-level ``open`` /
# ``rename`` / ``remove`` / ``listdir`` / ``exists`` resolve the name through it.
MemFileServicer = MemFileServer  # the reference's name for the service implementation
_PREFIX = "/local_mem"
_NAMED: Dict[str, MemFileServer] = {}
_CLIENTS: Dict[str, MemFileClient] = {}


def get_mem_server_sock_file(name: str) -> str:
    """The rendezvous file of server ``name`` (holds ``host:port``)."""
    import tempfile

    return os.path.join(tempfile.gettempdir(), f"vescale_b200_mem_server_{os.getuid()}_{name}.addr")


def get_prefix(name: str) -> str:
    return f"{_PREFIX}/{name}/"


def start_server(name: str, force: bool = False) -> MemFileServer:
    """Serve ``name`` from this process.  A live server of that name (this process or another) is an error unless ``force``."""
    sock = get_mem_server_sock_file(name)
    if os.path.exists(sock) and not force:
        try:
            MemFileClient(open_text(sock), timeout=2.0).exists("__probe__")
            raise RuntimeError(f"a mem file server named {name!r} is already running at {open_text(sock)}")
        except RuntimeError:
            raise
        except Exception:  # noqa: BLE001  stale rendezvous file of a dead server
            pass
    srv = MemFileServer().start()
    tmp = sock + f".{os.getpid()}"
    with io.open(tmp, "w") as f:
        f.write(srv.address)
    os.replace(tmp, sock)
    _NAMED[name] = srv
    _CLIENTS.pop(name, None)
    return srv


def open_text(path: str) -> str:
    with io.open(path) as f:
        return f.read().strip()


def _serve_named(name: str, conn) -> None:
    srv = start_server(name, force=True)
    conn.send(srv.address)
    conn.close()
    srv.server.wait_for_termination()


def start_server_in_new_process(name: str):
    """Serve ``name`` from a daemon child process (checkpoints then survive a crash of the training process, which is the point of
    a detached in-memory server); returns the ``multiprocessing.Process``."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    parent, child = ctx.Pipe()
    p = ctx.Process(target=_serve_named, args=(name, child), daemon=True)
    p.start()
    _recv_or_die(parent, p, f"mem file server {name!r}")
    parent.close()
    _CLIENTS.pop(name, None)
    return p


def _recv_or_die(conn, proc, what: str, timeout: float = 120.0):
    """First message of a freshly spawned server process — or a clear error if the child died before sending it (e.g. the parent's
    ``__main__`` cannot be re-imported by ``spawn``), instead of blocking forever on the pipe."""
    end = time.time() + timeout
    while time.time() < end:
        if conn.poll(0.2):
            return conn.recv()
        if not proc.is_alive():
            raise RuntimeError(f"{what}: the server process exited with code {proc.exitcode} before it came up")
    proc.terminate()
    raise TimeoutError(f"{what}: the server process did not come up within {timeout} s")


def wait_until_fs_ready(name: str, timeout: float = 120.0) -> bool:
    end = time.time() + timeout
    while time.time() < end:
        try:
            _client(name).exists("__probe__")
            return True
        except Exception:  # noqa: BLE001
            _CLIENTS.pop(name, None)
            time.sleep(0.1)
    return False


def _client(name: str) -> MemFileClient:
    c = _CLIENTS.get(name)
    if c is None:
        sock = get_mem_server_sock_file(name)
        if not os.path.exists(sock):
            raise FileNotFoundError(f"no mem file server named {name!r} (start_server / start_server_in_new_process)")
        c = _CLIENTS[name] = MemFileClient(open_text(sock))
    return c


def _split(path: str) -> Tuple[MemFileClient, str]:
    if not path.startswith(_PREFIX + "/"):
        raise ValueError(f"{path!r} is not under {_PREFIX}/<server name>/")
    name, _, rest = path[len(_PREFIX) + 1:].partition("/")
    return _client(name), rest


class _Upload(io.BytesIO):
    def __init__(self, client: MemFileClient, name: str, initial: bytes = b""):
        super().__init__()
        self._client, self._name = client, name
        if initial:
            self.write(initial)

    def close(self):
        if not self.closed:
            self._client.write(self._name, self.getvalue())
        super().close()


def mem_open(name: str, mode: str = "rb"):
    """A binary file object on the named server: ``rb`` downloads, ``wb`` uploads on close, ``ab`` appends (download + upload)."""
    client, rest = _split(name)
    if "r" in mode:
        return io.BytesIO(client.read(rest))
    if "a" in mode:
        return _Upload(client, rest, client.read(rest) if client.exists(rest) else b"")
    return _Upload(client, rest)


def rename(src: str, dst: str, overwrite: bool = False) -> None:
    client, a = _split(src)
    _, b = _split(dst)
    if not overwrite and client.exists(b):
        raise FileExistsError(dst)
    client.rename(a, b)


def remove(name: str) -> None:
    client, rest = _split(name)
    client.remove(rest)


def listdir(name: str) -> List[str]:
    client, rest = _split(name.rstrip("/") + "/")
    rest = rest.rstrip("/")
    return sorted({n[len(rest) + 1:].split("/", 1)[0] if rest and n.startswith(rest + "/") else n.split("/", 1)[0] for n in client.listdir(rest)})


def exists(name: str) -> bool:
    client, rest = _split(name)
    return client.exists(rest)


def __getattr__(attr):  # ``mem_server.open(...)`` is the reference's spelling; a module-level ``open`` would shadow the builtin in here
    if attr == "open":
        return mem_open
    raise AttributeError(attr)
