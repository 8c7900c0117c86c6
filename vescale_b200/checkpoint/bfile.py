"""One file API over every place a checkpoint can live.

The reference funnels checkpoint IO through ``bfile`` (``legacy/vescale/checkpoint/utilities/bfile.py:33-129``): a path's *schema*
picks the back end (local disk, the node-local in-memory file server; HDFS in ByteDance's internal build), and ``atomic_write``
(write to a temporary name, then rename) is what makes a crash mid-save leave the previous checkpoint readable.

Here the schemas are a registry, so a deployment adds its object store without touching the checkpointer::

    bfile.register_scheme("hdfs", MyHdfsBackend())          # open / exists / listdir / remove / rename / makedirs

Built in: ``local`` (plain paths and ``file://``) and ``mem`` (``mem://host:port/dir`` -> ``mem_server.MemFileClient``).
"""
from __future__ import annotations

import contextlib
import enum
import io
import os
import shutil
import uuid
from typing import Dict, Iterator, List, Optional

__all__ = ["FileType", "get_schema", "register_scheme", "BFile", "exists", "listdir", "remove", "rename", "makedirs", "atomic_write", "safe_atomic_write",
           "is_local_path", "local_list_folder", "read_bytes"]


class FileType(enum.Enum):
    LOCAL = "local"
    LOCAL_MEM = "mem"
    REMOTE = "remote"  # any registered non-built-in scheme


class _LocalBackend:
    kind = FileType.LOCAL

    @staticmethod
    def _p(path: str) -> str:
        return path[len("file://"):] if path.startswith("file://") else path

    def open(self, path: str, mode: str = "r"):
        return open(self._p(path), mode)

    def exists(self, path: str) -> bool:
        return os.path.exists(self._p(path))

    def listdir(self, path: str) -> List[str]:
        return sorted(os.listdir(self._p(path)))

    def remove(self, path: str) -> None:
        p = self._p(path)
        if os.path.isdir(p):
            shutil.rmtree(p, ignore_errors=True)
        elif os.path.exists(p):
            os.remove(p)

    def rename(self, src: str, dst: str, overwrite: bool = False) -> None:
        s, d = self._p(src), self._p(dst)
        if os.path.exists(d) and not overwrite:
            raise FileExistsError(d)
        os.replace(s, d)

    def makedirs(self, path: str) -> None:
        os.makedirs(self._p(path), exist_ok=True)


class _MemBackend:
    """``mem://host:port/name`` on the in-memory file server (a directory is a name prefix there)."""

    kind = FileType.LOCAL_MEM

    def __init__(self):
        self._clients: Dict[str, object] = {}

    def _c(self, path: str):
        from .mem_server import MemFileClient, parse_mem_uri

        addr, name = parse_mem_uri(path)
        if addr not in self._clients:
            self._clients[addr] = MemFileClient(addr)
        return self._clients[addr], name

    def open(self, path: str, mode: str = "r"):
        c, name = self._c(path)
        if "r" in mode:
            data = c.read(name)
            return io.BytesIO(data) if "b" in mode else io.StringIO(data.decode())
        from .mem_server import _UploadOnClose

        if "b" in mode:
            return _UploadOnClose(c, name)
        return _TextUpload(c, name)

    def exists(self, path: str) -> bool:
        c, name = self._c(path)
        return c.exists(name) or bool(c.listdir(name))

    def listdir(self, path: str) -> List[str]:
        """Direct children (the server lists full names under a prefix)."""
        c, name = self._c(path)
        pre = name.rstrip("/") + "/"
        return sorted({n[len(pre):].split("/", 1)[0] for n in c.listdir(name) if n.startswith(pre)})

    def remove(self, path: str) -> None:
        c, name = self._c(path)
        pre = name.rstrip("/") + "/"
        for full in c.listdir(name):
            if full.startswith(pre):
                c.remove(full)
        if c.exists(name):
            c.remove(name)

    def rename(self, src: str, dst: str, overwrite: bool = False) -> None:
        c, s = self._c(src)
        _, d = self._c(dst)
        if c.exists(d) and not overwrite:
            raise FileExistsError(dst)
        c.rename(s, d)

    def makedirs(self, path: str) -> None:  # directories are implicit
        return None


class _TextUpload(io.StringIO):
    def __init__(self, client, name):
        super().__init__()
        self._client, self._name = client, name

    def close(self):
        if not self.closed:
            self._client.write(self._name, self.getvalue().encode())
        super().close()


_BACKENDS: Dict[str, object] = {"local": _LocalBackend(), "file": _LocalBackend(), "mem": _MemBackend()}


def register_scheme(scheme: str, backend) -> None:
    """``backend``: an object with ``open(path, mode)``, ``exists``, ``listdir``, ``remove``, ``rename(src, dst, overwrite)``, ``makedirs``."""
    missing = [m for m in ("open", "exists", "listdir", "remove", "rename", "makedirs") if not callable(getattr(backend, m, None))]
    if missing:
        raise TypeError(f"backend for {scheme}:// lacks {missing}")
    _BACKENDS[scheme.lower()] = backend


def _scheme(path: str) -> str:
    head, sep, _ = str(path).partition("://")
    return head.lower() if sep and head.isidentifier() else "local"


def _backend(path: str):
    s = _scheme(path)
    if s not in _BACKENDS:
        raise ValueError(f"no checkpoint storage back end registered for '{s}://' (bfile.register_scheme)")
    return _BACKENDS[s]


def get_schema(path: str) -> FileType:
    return getattr(_backend(path), "kind", FileType.REMOTE)


def is_local_path(path: str) -> bool:
    return get_schema(path) in (FileType.LOCAL, FileType.LOCAL_MEM)


@contextlib.contextmanager
def BFile(name: str, mode: str = "r") -> Iterator:
    f = _backend(name).open(name, mode)
    try:
        yield f
    finally:
        f.close()


def exists(path: str) -> bool:
    return _backend(path).exists(path)


def listdir(path: str) -> List[str]:
    return list(_backend(path).listdir(path))


def remove(path: str) -> None:
    _backend(path).remove(path)


def rename(src: str, dst: str, overwrite: bool = False) -> None:
    if _scheme(src) != _scheme(dst):
        raise ValueError(f"rename across storage back ends: {src} -> {dst}")
    _backend(src).rename(src, dst, overwrite)


def makedirs(path: str) -> None:
    _backend(path).makedirs(path)


def local_list_folder(folder_path: str, recursive: bool = False) -> List[str]:
    if not recursive:
        return [os.path.join(folder_path, n) for n in sorted(os.listdir(folder_path))]
    out = []
    for root, _, files in os.walk(folder_path):
        out += [os.path.join(root, f) for f in sorted(files)]
    return out


def atomic_write(path: str, content: bytes) -> None:
    """All or nothing: readers see the old file or the new one, never a torn write."""
    tmp = f"{path}_tmp_{uuid.uuid4().hex}"
    with BFile(tmp, "wb") as f:
        f.write(content)
    rename(tmp, path, overwrite=True)


def safe_atomic_write(path: str, content: bytes) -> None:
    parent = path.rsplit("/", 1)[0] if "/" in path else ""
    if parent:
        makedirs(parent)
    atomic_write(path, content)


def read_bytes(path: str) -> bytes:
    with BFile(path, "rb") as f:
        return f.read()
