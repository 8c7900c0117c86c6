"""Build the in-tree sm_100a extension: nvcc cross-compiles on a CPU-only box.

    python csrc/build.py            # -> vescale_b200/_C.so
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
import sysconfig
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "csrc")
BUILD = os.path.join(ROOT, "build", "csrc")
OUT = os.path.join(ROOT, "vescale_b200", "_C.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
    "--use_fast_math", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-DTORCH_EXTENSION_NAME=_C",
    "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-DTORCH_EXTENSION_NAME=_C"]


def _includes():
    import torch
    from torch.utils import cpp_extension as ce

    inc = ce.include_paths(device_type="cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(cuda=True)
    inc.append(sysconfig.get_paths()["include"])
    return inc, ce.library_paths(device_type="cuda") if "device_type" in ce.library_paths.__code__.co_varnames else ce.library_paths(cuda=True), torch


def _needs(src, obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src] + deps)


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    inc, libdirs, torch = _includes()
    inc_flags = [f"-I{i}" for i in inc] + [f"-I{CSRC}"]
    abi = f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(BUILD, os.path.basename(s) + ".o")
        objs.append(o)
        if not force and not _needs(s, o, headers):
            continue
        if s.endswith(".cu"):
            cmd = ["nvcc", "-c", s, "-o", o] + NVCC_FLAGS + inc_flags + ["-Xcompiler", abi]
        else:
            cmd = ["g++", "-c", s, "-o", o] + CXX_FLAGS + inc_flags + [abi]
        jobs.append((s, cmd))

    def run(job):
        s, cmd = job
        t0 = time.time()
        p = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(BUILD, os.path.basename(s) + ".log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
        if p.returncode != 0:
            raise RuntimeError(f"compile failed: {s}\n{p.stderr[-6000:]}")
        if verbose:
            print(f"[build] {os.path.basename(s)} ok in {time.time() - t0:.1f}s")
        return p.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(OUT):
        link = ["g++", "-shared", "-o", OUT] + objs + [f"-L{d}" for d in libdirs] + ["-L/usr/local/cuda/lib64", "-lc10", "-ltorch_cpu", "-ltorch", "-lc10_cuda", "-ltorch_cuda", "-lcudart"] + [f"-Wl,-rpath,{d}" for d in libdirs]
        p = subprocess.run(link, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stderr[-4000:]}")
        if verbose:
            print(f"[build] linked {OUT}")
    return OUT


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
