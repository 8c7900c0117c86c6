"""Host-side chunk arithmetic under the names NCCL (and the legacy emulator, ``emulator/calculate_chunk_size.py``) use.  The
implementations live with the model they belong to: ``nccl/tuning.py`` (enqueue-time decisions), ``nccl/comm.py`` (patterns and
loop counts), ``algorithms.py`` (device-side geometry); ``tuning.py`` keeps the one-formula estimate for quick use."""
from __future__ import annotations

from typing import Optional, Tuple

from .algorithms import calc_bytes_per_grain, calc_bytes_per_step, chunk_layout
from .nccl import Algo, CollInfo, Func, NcclComm, Proto, get_algo_info, init_comm
from .nccl.comm import loop_info, pattern_of
from .nccl.tuning import compute_coll
from .tuning import calculate_chunk_size

__all__ = ["calcBytePerStep", "calcBytePerGrain", "get_pattern_info", "get_loop_info", "get_info_nchannels_nthreads_proto", "topo_get_algo_info", "compute_last_chunk_size",
           "calculate_chunk_size", "chunk_layout"]

calcBytePerStep = calc_bytes_per_step  # noqa: N816  (NCCL's spelling)
calcBytePerGrain = calc_bytes_per_grain  # noqa: N816


def get_pattern_info(func: int, algo: int) -> int:
    return pattern_of(func, algo)


def get_loop_info(pattern: int, n_ranks: int) -> Tuple[int, int]:
    """(steps per loop, chunks per loop)."""
    return loop_info(pattern, n_ranks)


def topo_get_algo_info(comm: NcclComm, func: int, count: int, dtype_size: int, num_pipe_ops: int = 1) -> CollInfo:
    """Cheapest (algorithm, protocol) with channels / threads / chunk geometry filled in."""
    return get_algo_info(comm, CollInfo(int(func), count, dtype_size), num_pipe_ops)


def get_info_nchannels_nthreads_proto(func: int, count: int, dtype_size: int, n_ranks: int, n_nodes: int = 1, compcap: int = 100, comm: Optional[NcclComm] = None) -> Tuple[int, int, int, int]:
    """(algorithm, protocol, channels, threads) NCCL would launch a collective with."""
    comm = comm or init_comm(n_ranks, n_nodes, compcap, nvls=False)
    i = topo_get_algo_info(comm, func, count, dtype_size)
    return i.algo, i.proto, i.n_channels, i.n_threads


def compute_last_chunk_size(comm: NcclComm, func: int, count: int, dtype_size: int, algo: int, proto: int, n_channels: int, n_threads: int) -> int:
    """Elements of the chunk the kernel uses for tree / LL variants (``work->lastChunkSize``); 0 for Simple rings, which size the
    last loop on the device."""
    info = CollInfo(int(func), count, dtype_size, n_channels=n_channels, n_threads=n_threads, algo=int(algo), proto=int(proto))
    return compute_coll(comm, info).last_chunk_size
