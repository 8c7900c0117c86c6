"""NCCL's analytic cost model (``src/graph/tuning.cc``) and the enqueue-time decisions built on it (``src/enqueue.cc``):

    comm = init_comm(n_ranks=8)                   # runs tune_model
    info = get_algo_info(comm, CollInfo(Func.ALL_REDUCE, count=1 << 20, dtype_size=2))
    info.algo, info.proto, info.n_channels, info.n_threads, info.chunk_size, info.last_chunk_size, info.n_loops

``tune_model`` fills ``comm.latencies[func][algo][proto]`` (us) and ``comm.bandwidths[...]`` (GB/s algorithm bandwidth);
``algo_time`` evaluates ``lat + bytes / bw`` with NCCL's empirical corrections; ``get_algo_info`` picks the minimum, sheds
channels / threads for small messages and computes the chunk geometry that the device kernels (and therefore the emulator's
reduction order) use.  Legacy ``emulator/nccl/graph/tuning.py`` models Volta-Hopper; the Blackwell row here extends the same
tables with the per-channel ceilings of NVLink 5 (an assumption until NCCL publishes its own; override through
``NcclComm.bandwidths`` or the profiler-result loader when matching a particular run)."""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

from .comm import CollInfo, NcclComm, loop_info, pattern_of
from .constants import (CHUNKSTEPS, LL128_DATAELEMS, LL128_LINEELEMS, LL128_MAX_NTHREADS, LL_MAX_NTHREADS, MAX_WORK_ELEMENTS, NCCL_STEPS, NUM_ALGOS, NUM_FUNCS, NUM_PROTOS,
                        SIMPLE_MAX_NTHREADS, SLICESTEPS, THREAD_THRESHOLD, WARP_SIZE, Algo, Func, Hw, Proto, TopoPattern, compcap_index)

__all__ = ["tune_model", "algo_time", "get_algo_info", "compute_coll", "BASE_LAT", "HW_LAT", "TREE_CORRECTION"]

T, R, CD, CC, NV, NT = (int(a) for a in (Algo.TREE, Algo.RING, Algo.COLLNET_DIRECT, Algo.COLLNET_CHAIN, Algo.NVLS, Algo.NVLS_TREE))
LL, LL128, SIMPLE = (int(p) for p in (Proto.LL, Proto.LL128, Proto.SIMPLE))

# software latency of one kernel by (algorithm, protocol), us
BASE_LAT = {T: (6.8, 14.0, 0.0), R: (6.6, 14.0, 8.4), CD: (0, 0, 0), CC: (0, 0, 0), NV: (0, 0, 0), NT: (0, 0, 0)}
# per-hop hardware latency by link type, us
HW_LAT = {
    int(Hw.NVLINK): {T: (0.6, 1.25, 4.0), R: (0.6, 1.9, 3.4), CD: (0, 0, 3.7), CC: (0, 0, 2.8), NV: (0, 0, 23.0), NT: (0, 0, 23.0)},
    int(Hw.PCI): {T: (1.0, 1.9, 6.0), R: (1.0, 2.5, 5.7), CD: (0, 0, 3.7), CC: (0, 0, 2.8), NV: (0, 0, 0), NT: (0, 0, 0)},
    int(Hw.NET): {T: (5.0, 8.5, 14.0), R: (2.7, 4.0, 14.0), CD: (0, 0, 31.0), CC: (0, 0, 30.0), NV: (0, 0, 18.0), NT: (0, 0, 14.0)},
}
# bandwidth ceilings, rows = Volta / Ampere / Hopper / Blackwell, columns = 1 / 2 / 4+ nodes (GB/s)
LL_MAX_BW = ((39.0, 39.0, 20.4), (87.7, 22.5, 19.0), (141.0, 45.0, 35.0), (2 * 141.0, 2 * 45.0, 2 * 35.0))
PER_CH_MAX_RING_LL128 = ((20.0,) * 3, (20.0,) * 3, (36.7,) * 3, (2 * 36.7,) * 3)
PER_CH_MAX_TREE_LL128 = ((20.0,) * 3, (20.0,) * 3, (36.7, 36.7, 29.0), (2 * 36.7, 2 * 36.7, 2 * 29.0))
PER_CH_MAX_TREE = ((26.5, 18.5, 10.0), (24.0, 23.6, 17.8), (38.7, 41.4, 36.0), (2 * 38.7, 2 * 41.4, 2 * 36.0))
PER_CH_MAX_NVLS_TREE = ((26.5, 18.5, 10.0), (24.0, 23.6, 17.8), (38.7, 41.4, 36.0), (2 * 38.7, 2 * 41.4, 2 * 36.0))
# measured-vs-model fudge of the tree algorithm by protocol and log2(bytes / 64)
TREE_CORRECTION = (
    (1.0, 1.0, 1.0, 1.0, 0.9, 0.8, 0.7, 0.7, 0.7, 0.7, 0.6, 0.5, 0.4, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.0, 1.0, 1.0),
    (1.0, 1.0, 1.0, 1.0, 1.0, 0.9, 0.8, 0.8, 0.8, 0.7, 0.6, 0.6, 0.6, 0.6, 0.6, 0.6, 0.8, 0.9, 0.9, 0.9, 0.9, 1.0, 1.0),
    (0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.8, 0.7, 0.6, 0.6, 0.5, 0.5, 0.5, 0.5, 0.6, 0.7, 0.8, 0.7, 0.7, 0.8, 0.9, 0.9),
)


def _log2i(n: int) -> int:
    return max(0, int(n).bit_length() - 1)


def _net_overhead(comm: NcclComm) -> float:
    return 2.0 if comm.cpu_arch_amd else 1.0


def tune_model(comm: NcclComm, min_compcap: Optional[int] = None) -> NcclComm:
    n_ranks, n_nodes = comm.n_ranks, comm.n_nodes
    cc = compcap_index(min_compcap if min_compcap is not None else comm.compcap)
    col = 0 if n_nodes == 1 else (1 if n_nodes == 2 else 2)
    ppn = comm.ppn
    simple_threads = SIMPLE_MAX_NTHREADS
    comm.max_threads = {a: {LL: LL_MAX_NTHREADS, LL128: LL128_MAX_NTHREADS, SIMPLE: simple_threads} for a in range(NUM_ALGOS)}
    comm.thread_thresholds = {a: {LL: THREAD_THRESHOLD[Proto.LL], LL128: THREAD_THRESHOLD[Proto.LL128], SIMPLE: THREAD_THRESHOLD[Proto.SIMPLE]} for a in range(NUM_ALGOS)}
    comm.thread_thresholds[R][LL] *= n_ranks
    comm.thread_thresholds[CD][SIMPLE] = comm.thread_thresholds[CC][SIMPLE] = 256
    comm.latencies = {f: {a: [0.0] * NUM_PROTOS for a in range(NUM_ALGOS)} for f in range(NUM_FUNCS)}
    comm.bandwidths = {f: {a: [0.0] * NUM_PROTOS for a in range(NUM_ALGOS)} for f in range(NUM_FUNCS)}
    if n_ranks <= 1:
        return comm
    for f in range(NUM_FUNCS):
        func = Func(f)
        nsteps = 2 * (n_ranks - 1) if func == Func.ALL_REDUCE else (n_ranks - 1 if func in (Func.REDUCE_SCATTER, Func.ALL_GATHER) else n_ranks)
        n_inter = (2 * n_nodes if n_nodes > 1 else 0) if func == Func.ALL_REDUCE else (n_nodes - 1 if func in (Func.REDUCE_SCATTER, Func.ALL_GATHER) else n_nodes)
        for a in range(NUM_ALGOS):
            if func in (Func.BROADCAST, Func.REDUCE) and a != R:
                continue
            if func in (Func.REDUCE_SCATTER, Func.ALL_GATHER) and a not in (R, NV):
                continue
            g = comm.graphs.get(a)
            if g is None or g.n_channels == 0:
                continue
            for p in range(NUM_PROTOS):
                if a in (NV, NT) and p != SIMPLE:
                    continue
                collnet = a in (CD, CC)
                bw = g.bw_intra if (n_nodes <= 2 or collnet) else g.bw_inter
                if a == NV:
                    bw = min(g.bw_intra, g.bw_inter) if n_nodes > 1 else g.bw_intra
                if a == NT:
                    bw = min(g.bw_intra, g.bw_inter if n_nodes <= 2 else g.bw_inter / 2)
                bus = g.n_channels * bw
                if a == R and p == LL:
                    bus = min(LL_MAX_BW[cc][col], bus * (0.25 if (n_nodes > 1 or func in (Func.ALL_REDUCE, Func.REDUCE)) else 1.0 / 3.0))
                if a == R and p == LL128:
                    bus = min(bus * (0.7 if ppn < 2 else 0.92), g.n_channels * PER_CH_MAX_RING_LL128[cc][col])
                if a == T:
                    bus = min(bus * 0.92, g.n_channels * PER_CH_MAX_TREE[cc][col])
                if a == T and p == LL:
                    bus = min(bus / 3.8, LL_MAX_BW[cc][col])
                if a == T and p == LL128:
                    bus = min(bus * (7.0 / 9.0 if n_nodes == 1 else 120.0 / 128.0), g.n_channels * PER_CH_MAX_TREE_LL128[cc][col])
                if a == T and g.pattern == int(TopoPattern.TREE):
                    bus *= 0.85
                if a == CD and p != SIMPLE:
                    bus = 0.0
                if a == CC and p != SIMPLE:
                    bus = 0.0
                if a == CD and p == SIMPLE:
                    bus *= 0.5 if ppn > 1 else 1.0  # remote all-gather of the direct scheme shares the links
                if a == NT:
                    bus = min(bus, g.n_channels * PER_CH_MAX_NVLS_TREE[cc][col])
                # bus bandwidth -> algorithm bandwidth
                if a == R:
                    ratio = n_ranks / nsteps
                elif a in (NV, NT):
                    ratio = 5.0 / 6.0
                else:
                    ratio = 0.5
                if a == NV and func in (Func.REDUCE_SCATTER, Func.ALL_GATHER):
                    ratio = (n_ranks - 1.0) / n_ranks * n_ranks / (n_ranks - 1.0)  # the switch moves every shard once
                comm.bandwidths[f][a][p] = bus * ratio
                # latency
                lat = BASE_LAT[a][p]
                hw_intra = comm.intra_hw(a)
                intra = HW_LAT[hw_intra][a][p]
                inter = HW_LAT[int(Hw.NET)][a][p] + g.latency_inter
                if p == SIMPLE:
                    inter += g.latency_inter  # the flush
                if a == R:
                    hop = HW_LAT[hw_intra][a][p]
                    if func in (Func.REDUCE, Func.BROADCAST):
                        if g.same_channels:
                            lat += hop
                        else:
                            if p == LL:
                                hop = HW_LAT[hw_intra][T][p]  # the chain is the less favourable case
                            lat += nsteps * hop
                    else:
                        net = 0.0
                        if n_nodes > 1:
                            net = _net_overhead(comm) * (3 if p == SIMPLE else 1)
                        intra = max(intra, net)
                        lat += (nsteps - n_inter) * intra + n_inter * inter
                elif a == T:
                    lat += 2 * ((n_ranks / n_nodes - 1) * intra + _log2i(n_nodes) * inter)
                elif a == CD:
                    lat += 2 * (min(1.0, ppn - 1) * intra + (ppn - 1) * 0.4) + inter
                elif a == CC:
                    lat += 2 * (ppn - 1) * intra + inter
                elif a == NV:
                    lat = intra
                    if n_nodes > 1:
                        lat += HW_LAT[int(Hw.NET)][a][p] + inter
                elif a == NT:
                    lat += intra + 2 * _log2i(n_nodes) * inter
                comm.latencies[f][a][p] = lat
    # protocols NCCL disables outright: LL128 needs NVLink-connected Volta+ inside the node; NVLS only with NVSwitch multicast
    for f in range(NUM_FUNCS):
        for a in range(NUM_ALGOS):
            g = comm.graphs.get(a)
            if g is not None and g.type_intra > 2 and comm.bandwidths[f][a][LL128]:  # beyond NVB: PCIe path
                comm.bandwidths[f][a][LL128] = 0.0
            if a in (NV, NT) and not comm.nvls:
                comm.bandwidths[f][a] = [0.0] * NUM_PROTOS
            if a in (NT,) and n_nodes == 1:
                comm.bandwidths[f][a] = [0.0] * NUM_PROTOS  # NVLS tree only exists across nodes
    return comm


def algo_time(comm: NcclComm, func: int, algo: int, proto: int, n_bytes: int, n_channels: int = 0, num_pipe_ops: int = 1) -> float:
    """``ncclTopoGetAlgoTime``: predicted microseconds, or -1 when the combination is unavailable."""
    bw = comm.bandwidths[int(func)][int(algo)][int(proto)]
    lat = comm.latencies[int(func)][int(algo)][int(proto)]
    if bw == 0:
        return -1.0
    log_size = _log2i(n_bytes >> 6)
    if algo == T and log_size < 23:
        bw *= TREE_CORRECTION[int(proto)][log_size]
    if n_channels:
        bw = bw / comm.n_channels * n_channels
    if algo == R and proto == SIMPLE and comm.n_nodes > 1 and func == Func.ALL_REDUCE and n_bytes / (comm.n_channels * comm.n_ranks) >= 64:
        lat *= 1.9 if comm.compcap < 80 else 1.4  # plateau of inter-node rings
    lat_count = num_pipe_ops if algo == R else -(-num_pipe_ops // MAX_WORK_ELEMENTS)
    return lat * lat_count + n_bytes / (1000.0 * bw)


def compute_coll(comm: NcclComm, info: CollInfo) -> CollInfo:
    """``computeColl``: chunk geometry of an already selected (algo, proto, nChannels, nThreads)."""
    func, a, p = Func(info.func), info.algo, info.proto
    info.pattern = pattern_of(func, a)
    info.nsteps_per_loop, info.nchunks_per_loop = loop_info(info.pattern, comm.n_ranks)
    step = comm.buff_sizes[p] // NCCL_STEPS
    ring_simple = p == SIMPLE and a == R
    info.chunk_steps = CHUNKSTEPS[func] if ring_simple else 1
    info.slice_steps = SLICESTEPS[func] if ring_simple else 1
    chunk = step * info.chunk_steps
    nb, nch, ts = info.n_bytes, max(1, info.n_channels), info.dtype_size
    if a == T and p == SIMPLE:
        if info.pattern == 6:  # up-down: trade chunk size against pipeline depth
            depth = comm.tree_depth
            while nb // (nch * chunk) < depth * 8 and chunk > 131072:
                chunk //= 2
            while nb // (nch * chunk) < depth * 4 and chunk > 65536:
                chunk //= 2
            while nb // (nch * chunk) < depth and chunk > 32768:
                chunk //= 2
        info.last_chunk_size = chunk // ts
    elif a in (NV, NT):
        mx = 131072 if a == NV else 262144
        chunk = min(chunk, mx)
        while nb // (nch * comm.n_ranks // max(1, comm.n_nodes) * chunk) < 2 and chunk > 32768 and a == NV:
            chunk //= 2
        info.last_chunk_size = chunk // ts
    elif p == LL:
        slice_ = step * 8 // 16  # payload of a FIFO slot: half of every 16-byte line
        loop = nch * info.nchunks_per_loop * slice_
        last = -(-(nb - (nb // loop) * loop) // (nch * info.nchunks_per_loop))
        align = max(1, info.n_threads) * 8
        last = -(-last // align) * align
        info.last_chunk_size = last // ts
    elif a == T and p == LL128:
        nsteps = 1 + _log2i(comm.n_nodes) + 0.1 * comm.ppn
        while nb / (nch * chunk) < nsteps * 64 / comm.ppn and chunk > 131072:
            chunk //= 2
        while nb / (nch * chunk) < nsteps * 16 / comm.ppn and chunk > 32768:
            chunk //= 2
        info.last_chunk_size = chunk * LL128_DATAELEMS // (LL128_LINEELEMS * ts)
    eff = chunk
    if p == LL:
        eff //= 2
    elif p == LL128:
        eff = chunk // LL128_LINEELEMS * LL128_DATAELEMS
    info.chunk_size = chunk
    info.n_loops = max(1, -(-nb // (nch * info.nchunks_per_loop * eff))) if nb else 0
    info.proxy_steps = info.nsteps_per_loop * info.n_loops * info.chunk_steps
    return info


def get_algo_info(comm: NcclComm, info: CollInfo, num_pipe_ops: int = 1, force: Optional[Tuple[int, int]] = None) -> CollInfo:
    """``getAlgoInfo`` + ``computeColl``: choose (algo, proto), then channels / threads, then the chunk geometry.  ``force`` pins
    (algo, proto) the way ``NCCL_ALGO`` / ``NCCL_PROTO`` do."""
    best = (-1, -1, math.inf)
    if force is not None:
        best = (int(force[0]), int(force[1]), algo_time(comm, info.func, force[0], force[1], info.n_bytes, 0, num_pipe_ops))
    else:
        for a in range(NUM_ALGOS):
            if a in (CD, CC):
                continue
            for p in range(NUM_PROTOS):
                t = algo_time(comm, info.func, a, p, info.n_bytes, 0, num_pipe_ops)
                if 0 <= t < best[2]:
                    best = (a, p, t)
    if best[0] < 0:
        raise RuntimeError(f"no algorithm/protocol available for {Func(info.func).name} on {comm.n_ranks} ranks")
    info.algo, info.proto, info.time_us = best
    a, p = info.algo, info.proto
    nc = info.n_channels or comm.n_channels
    nt = comm.max_threads[a][p]
    thr = comm.thread_thresholds[a][p]
    if a in (NV, NT):
        nc = min(nc, comm.graphs[a].n_channels)
        nt = SIMPLE_MAX_NTHREADS + WARP_SIZE
    else:
        while info.n_bytes < nc * nt * thr:
            if nc >= 2:
                nc -= 1
            elif nt % 128 == 0:
                nt //= 2
            else:
                break
    if p == SIMPLE and a not in (NV, NT):
        if a == R:
            nt += WARP_SIZE  # sync warp
        if a == T:
            nt += 4 * WARP_SIZE
    nt = 3 * WARP_SIZE if nt // WARP_SIZE < 3 else nt
    info.n_channels, info.n_threads = nc, nt
    return compute_coll(comm, info)


# ---- tuning.cc's own names ------------------------------------------------------------------------------------------------------------------------
nccl_topo_tune_model, nccl_topo_get_algo_time = tune_model, algo_time
getNetOverhead = _net_overhead  # noqa: N816


def DIVUP(x: int, y: int) -> int:  # noqa: N802
    return (x + y - 1) // y


class ncclTopoCpuType:  # noqa: N801
    """(arch, vendor, model) of the host CPU as ``ncclTopoCpuType`` reports it: the tuning model only distinguishes AMD x86 (twice
    the network overhead) from everything else."""
    ARCH_X86, ARCH_POWER, ARCH_ARM = 1, 2, 3
    VENDOR_INTEL, VENDOR_AMD, VENDOR_ZHAOXIN = 1, 2, 3


def get_cpu_info():
    """``(arch, vendor)`` of this host (from ``/proc/cpuinfo`` / ``platform``)."""
    import platform

    m = platform.machine().lower()
    arch = ncclTopoCpuType.ARCH_X86 if m in ("x86_64", "amd64", "i686") else ncclTopoCpuType.ARCH_ARM if m.startswith(("aarch", "arm")) else ncclTopoCpuType.ARCH_POWER
    vendor = 0
    try:
        with open("/proc/cpuinfo") as f:
            text = f.read(8192)
        vendor = ncclTopoCpuType.VENDOR_AMD if "AuthenticAMD" in text else ncclTopoCpuType.VENDOR_INTEL if "GenuineIntel" in text else ncclTopoCpuType.VENDOR_ZHAOXIN if "Shanghai" in text else 0
    except OSError:
        pass
    return arch, vendor


def get_nthreads(name: str, env_value: int, min_threads: int, max_threads: int, default: int) -> int:
    """``getNthreads``: an ``NCCL_*NTHREADS`` override is honoured only if it is a multiple of the warp size inside
    ``[min_threads, max_threads]``; otherwise (or when unset: -2) the default applies."""
    nt = env_value
    if nt > 0:
        if nt % WARP_SIZE != 0:
            nt = max_threads
        elif nt > max_threads:
            nt = max_threads
        elif nt < min_threads:
            nt = min_threads
    else:
        nt = default
    return nt


__all__ += ["nccl_topo_tune_model", "nccl_topo_get_algo_time", "getNetOverhead", "DIVUP", "ncclTopoCpuType", "get_cpu_info", "get_nthreads"]
