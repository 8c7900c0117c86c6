"""A host-side model of the NCCL library decisions the emulator depends on: which (algorithm, protocol) a collective of a given
size runs, over how many channels and threads, and with which chunk size — because those choose the summation order the emulator
has to reproduce (legacy ``emulator/nccl/``).  Numbers follow the public NCCL sources (``src/graph/tuning.cc``, ``src/enqueue.cc``,
``src/include/collectives.h`` of the 2.19-2.28 line) plus a Blackwell (sm_100) row for the NVLink 5 / NVSwitch node this framework
targets.  Nothing here talks to a GPU."""
from .constants import *  # noqa: F401,F403
from .comm import CollInfo, NcclComm, TopoGraph, init_comm  # noqa: F401
from .tuning import algo_time, get_algo_info, tune_model  # noqa: F401
from .profiler_result import NcclProfilerResult, parse_nccl_debug_log  # noqa: F401
