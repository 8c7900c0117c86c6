"""Enumerations and compile-time constants of NCCL that the tuning / chunking model needs (``src/include/nccl_common.h``,
``devcomm.h``, ``collectives.h``); legacy ``emulator/nccl/constants.py``."""
from __future__ import annotations

from enum import IntEnum

__all__ = [
    "Func", "Algo", "Proto", "Hw", "Pattern", "TopoPattern", "NUM_FUNCS", "NUM_ALGOS", "NUM_PROTOS", "WARP_SIZE", "NCCL_STEPS", "MAX_NCHANNELS",
    "LL_LINES_PER_THREAD", "LL_MAX_NTHREADS", "LL128_MAX_NTHREADS", "LL128_LINEELEMS", "LL128_DATAELEMS", "LL128_SHMEM_ELEMS_PER_THREAD", "LL128_ELEMS_PER_THREAD", "SIMPLE_MAX_NTHREADS",
    "CHUNKSTEPS", "SLICESTEPS", "THREAD_THRESHOLD", "DEFAULT_BUFFSIZE", "COMPCAP_IDX", "compcap_index", "TYPE_SIZE", "MAX_WORK_ELEMENTS", "MAX_TREE_ARITY",
]


class Func(IntEnum):
    BROADCAST = 0
    REDUCE = 1
    ALL_GATHER = 2
    REDUCE_SCATTER = 3
    ALL_REDUCE = 4
    SEND_RECV = 5


class Algo(IntEnum):
    TREE = 0
    RING = 1
    COLLNET_DIRECT = 2
    COLLNET_CHAIN = 3
    NVLS = 4
    NVLS_TREE = 5


class Proto(IntEnum):
    LL = 0
    LL128 = 1
    SIMPLE = 2


class Hw(IntEnum):
    NVLINK = 0
    PCI = 1
    NET = 2


class Pattern(IntEnum):
    """How a collective walks its topology (``ncclPattern_t``): decides steps and chunks per loop."""
    RING = 0
    RING_TWICE = 1
    PIPELINE_FROM = 2
    PIPELINE_TO = 3
    TREE_UP = 4
    TREE_DOWN = 5
    TREE_UP_DOWN = 6
    COLLNET_CHAIN = 7
    COLLNET_DIRECT = 8
    NVLS = 9
    NVLS_TREE = 10
    SEND = 11
    RECV = 12


class TopoPattern(IntEnum):
    """Graph-search patterns as they appear in an ``NCCL_GRAPH_DUMP_FILE`` (``graph.h``)."""
    BALANCED_TREE = 1
    SPLIT_TREE = 2
    TREE = 3
    RING = 4
    NVLS = 5


NUM_FUNCS, NUM_ALGOS, NUM_PROTOS = 5, 6, 3  # SEND_RECV has no tuning entry
WARP_SIZE = 32
NCCL_STEPS = 8  # FIFO slots per connection
MAX_NCHANNELS = 32
MAX_WORK_ELEMENTS = 9
MAX_TREE_ARITY = 3

# LL: 8-byte lines of 4 data + 4 flag bytes; LL128: 128-byte lines of 120 data + 8 flag bytes
LL_LINES_PER_THREAD = 8
LL_MAX_NTHREADS = 512
LL128_LINEELEMS = 16  # uint64 per line
LL128_DATAELEMS = 15
LL128_MAX_NTHREADS = 640
LL128_SHMEM_ELEMS_PER_THREAD = 8
LL128_ELEMS_PER_THREAD = 120
SIMPLE_MAX_NTHREADS = 512

# (chunk steps, slice steps) of the Simple-protocol ring kernels: a chunk is CHUNKSTEPS FIFO slots, sent SLICESTEPS at a time
CHUNKSTEPS = {Func.ALL_REDUCE: NCCL_STEPS // 2, Func.ALL_GATHER: NCCL_STEPS // 2, Func.REDUCE_SCATTER: NCCL_STEPS // 2, Func.BROADCAST: 1, Func.REDUCE: 1}
SLICESTEPS = {Func.ALL_REDUCE: NCCL_STEPS // 4, Func.ALL_GATHER: NCCL_STEPS // 4, Func.REDUCE_SCATTER: NCCL_STEPS // 4, Func.BROADCAST: 1, Func.REDUCE: 1}

# bytes per thread below which enqueue sheds channels, then threads
THREAD_THRESHOLD = {Proto.LL: 8, Proto.LL128: 8, Proto.SIMPLE: 64}

DEFAULT_BUFFSIZE = {Proto.LL: LL_LINES_PER_THREAD * LL_MAX_NTHREADS * NCCL_STEPS * 16, Proto.LL128: LL128_ELEMS_PER_THREAD * LL128_MAX_NTHREADS * NCCL_STEPS * 8,
                    Proto.SIMPLE: 1 << 22}

# rows of the per-architecture bandwidth tables
COMPCAP_IDX = {"volta": 0, "ampere": 1, "hopper": 2, "blackwell": 3}


def compcap_index(compcap: int) -> int:
    """70 → Volta row, 80 → Ampere, 90 → Hopper, 100+ → Blackwell."""
    if compcap >= 100:
        return COMPCAP_IDX["blackwell"]
    if compcap >= 90:
        return COMPCAP_IDX["hopper"]
    if compcap >= 80:
        return COMPCAP_IDX["ampere"]
    return COMPCAP_IDX["volta"]


TYPE_SIZE = {"int8": 1, "uint8": 1, "float8_e4m3fn": 1, "float8_e5m2": 1, "float16": 2, "bfloat16": 2, "int32": 4, "float32": 4, "int64": 8, "float64": 8}


# ---- NCCL's own spellings and integer helpers (``devcomm.h`` / ``align.h``) ------------------------------------------------------------------
NcclFunc, NcclAlgo, NcclProto, NcclPattern = Func, Algo, Proto, Pattern
ALIGN_SIZE = 4096  # buffers are carved out of one allocation on page boundaries


def div_up(x: int, y: int) -> int:
    return (x + y - 1) // y


def round_up(x: int, y: int) -> int:
    return (x + y - 1) // y * y


def align_up(x: int, a: int) -> int:
    """``a`` a power of two."""
    if a & (a - 1):
        raise ValueError("align_up: the alignment must be a power of two")
    return (x + a - 1) & ~(a - 1)


def log2i(n: int) -> int:
    """floor(log2(n)); 0 for n <= 1."""
    return max(0, int(n).bit_length() - 1)


__all__ += ["NcclFunc", "NcclAlgo", "NcclProto", "NcclPattern", "ALIGN_SIZE", "div_up", "round_up", "align_up", "log2i"]
