"""Which algorithm / protocol / chunk size NCCL would use for a collective of a given size — the part of the emulator that
decides *which* reduction order to reproduce (legacy ``emulator/nccl/graph/tuning.py``, ``calculate_chunk_size.py``).

This is a latency-bandwidth model in the spirit of ``ncclTopoTuneModel`` with constants for one NVSwitch node
(intra-node only): time = latency(algo, proto) + bytes / bandwidth(algo, proto); the cheapest (algo, proto) wins.  The
constants are defaults, not measurements — pass your own table (e.g. read off ``NCCL_DEBUG=INFO`` tuning lines) when the
emulated run must match a particular machine.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

__all__ = ["Tuning", "select_algorithm", "calculate_chunk_size", "PROTO_EFFICIENCY"]

ALGOS = ("ring", "tree")
PROTOS = ("ll", "ll128", "simple")
# base latencies in us (nccl/src/graph/tuning.cc baseLat, NVLink hw latencies)
_BASE_LAT = {"tree": {"ll": 6.8, "ll128": 14.0, "simple": 8.4}, "ring": {"ll": 6.6, "ll128": 14.0, "simple": 8.4}}
_HW_LAT = {"tree": {"ll": 0.6, "ll128": 1.25, "simple": 4.0}, "ring": {"ll": 0.6, "ll128": 1.9, "simple": 3.4}}
PROTO_EFFICIENCY = {"ll": 0.5, "ll128": 0.92, "simple": 1.0}  # payload fraction of the wire bytes
_MAX_BYTES = {"ll": 16 << 10, "ll128": 1 << 20, "simple": 1 << 62}  # above these the protocol is not considered


@dataclass
class Tuning:
    algo: str
    proto: str
    time_us: float
    nchannels: int
    chunk_bytes: int


def calculate_chunk_size(nbytes: int, nranks: int, nchannels: int, proto: str, algo: str, buff_bytes: int = 4 << 20, steps: int = 8, chunk_steps: int = 2) -> int:
    """Bytes per (rank, channel) step: NCCL slices its ``buff_bytes`` FIFO into ``steps`` slots and sends ``chunk_steps`` of
    them per chunk, then shrinks the chunk for small messages so every channel still has work (``enqueue.cc``)."""
    stepsize = buff_bytes // steps
    if proto == "ll":
        stepsize //= 2  # half of every 8-byte line is a flag
    elif proto == "ll128":
        stepsize = stepsize * 120 // 128
    chunk = stepsize * (chunk_steps if proto == "simple" else 1)
    per = max(1, nbytes // max(1, nchannels))
    if algo == "ring":
        while chunk // 2 >= 512 and per < nranks * chunk:  # not enough data for a full ring loop: halve
            chunk //= 2
    else:
        while chunk // 2 >= 512 and per < chunk * 8:
            chunk //= 2
    return max(16, chunk // 16 * 16)


def select_algorithm(coll: str, nbytes: int, nranks: int, bus_bw_gbs: float = 360.0, max_channels: int = 16, table: Optional[Dict[Tuple[str, str], Tuple[float, float]]] = None) -> Tuning:
    """Cheapest (algorithm, protocol) for ``coll`` in {"all_reduce", "all_gather", "reduce_scatter", "broadcast"}.
    ``table[(algo, proto)] = (latency_us, bandwidth_GBps)`` overrides the built-in model."""
    best: Optional[Tuning] = None
    for algo in ALGOS:
        if algo == "tree" and coll != "all_reduce":
            continue
        for proto in PROTOS:
            if nbytes > _MAX_BYTES[proto] * nranks:
                continue
            if table and (algo, proto) in table:
                lat, bw = table[(algo, proto)]
            else:
                nsteps = 2 * (nranks - 1) if (algo == "ring" and coll == "all_reduce") else (nranks - 1 if algo == "ring" else 2 * max(1, (nranks - 1).bit_length()))
                lat = _BASE_LAT[algo][proto] + nsteps * _HW_LAT[algo][proto]
                bw = bus_bw_gbs * PROTO_EFFICIENCY[proto]
                if algo == "tree":
                    bw *= 0.92  # two trees, each moving half: slightly below the ring's bus bandwidth at large sizes
                if algo == "ring" and coll == "all_reduce":
                    bw *= nranks / (2.0 * (nranks - 1))  # algorithm bandwidth from bus bandwidth
            t = lat + nbytes / (bw * 1e3)
            nch = max(1, min(max_channels, nbytes // (64 << 10) or 1))
            cand = Tuning(algo, proto, t, nch, calculate_chunk_size(nbytes, nranks, nch, proto, algo))
            if best is None or cand.time_us < best.time_us:
                best = cand
    assert best is not None
    return best
