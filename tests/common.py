"""Multi-process test harness: spawn ``world_size`` ranks on this host (gloo on CPU, nccl when enough GPUs),
file-store rendezvous.  Mirrors the reference's DTensorTestBase (``test/common_dtensor.py:126-345``)."""
from __future__ import annotations

import os
import sys
import tempfile
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, fn, init_file, backend, args):
    sys.path.insert(0, ROOT)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.set_num_threads(1)
    try:
        if backend == "nccl":
            torch.cuda.set_device(rank)
            # a rank that is still inside the test after this long prints where it is (a spin in a kernel, a collective a failed
            # peer never joined); the per-test pytest timeout then shows it with the captured output
            import faulthandler

            faulthandler.dump_traceback_later(int(os.environ.get("VESCALE_TEST_HANG_DUMP_S", "150")), exit=False)
        dist.init_process_group(backend, init_method=f"file://{init_file}", rank=rank, world_size=world_size)
        fn(rank, world_size, *args)
        dist.barrier()
    except Exception:
        traceback.print_exc()
        if backend == "nccl":
            # Do NOT tear the NCCL group down after a failure: the peers are blocked in (or about to enter) a collective this rank
            # will never join, and destroy_process_group() would wait for them — the failure would turn into a hang of all ranks.
            # Exiting makes mp.spawn see the failure and terminate the peers.
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(1)
        raise
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def run_distributed(fn, world_size: int = 4, *args, backend: str | None = None):
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() and torch.cuda.device_count() >= world_size else "gloo"
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "store")
        mp.spawn(_worker, args=(world_size, fn, init_file, backend, args), nprocs=world_size, join=True)


def device_type() -> str:
    return "cuda" if dist.get_backend() == "nccl" else "cpu"
