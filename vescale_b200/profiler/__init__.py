"""ndtimeline: multi-rank CUDA-event timeline with a simulated global clock
(parity: ``legacy/vescale/ndtimeline`` — timer.py, api.py, handlers/*, pool.py, predefined.py)."""
from .timer import (  # noqa: F401
    DeviceTimerMeta,
    NDMetricLevel,
    NDTimerManager,
    NDTimerManagerSingleton,
    flush,
    inc_step,
    init_ndtimers,
    is_initialized,
    ndtimeit,
    ndtimeit_coll,
    ndtimeit_p2p,
    ndtimer,
    set_global_step,
    wait,
)
from .handlers import DeviceTimerStreamRecord, parse_record, ChromeTraceNDHandler, DoNothingNDHandler, LocalRawNDHandler, LocalTimelineNDHandler, LoggingNDHandler, NDHandler, ParserNDHandler  # noqa: F401
from .sock_streamer import (  # noqa: F401
    SOCK_PARENT_DIR, SOCK_PATH, SOCK_TIMEOUT_CLIENT, NDtimelineStreamer, SockNDHandler, decode_frames, dumps_fn, encode_frame, encode_package, loads_fn,
    serialize_to_package,
)
from .pool import CudaEventPool, DefaultEventPool  # noqa: F401
from .stream import get_nccl_coll_stream, get_nccl_p2p_stream, register_comm_stream  # noqa: F401
from .world_info import TopoInfo, TrainingInfo, WorldInfo  # noqa: F401
from . import exceptions  # noqa: F401
from .logger import NDTimelineLogger, get_logger  # noqa: F401

logger = get_logger()

LOCAL_LOGGING_PATH = SOCK_PARENT_DIR  # default directory of LocalRawNDHandler / LocalTimelineNDHandler output (legacy ``variables.py``)
DEFAULT_CUDA_EVENT_POOL_SIZE = 20
NDTIMELINE_FLUSH_SEPCIAL = "special"  # step tag of an out-of-band flush (spelling as in the reference)
NDTIMELINE_INNER_GLOBAL_STEP_KEY = "_inner_global_step"  # record key of the step counter maintained by inc_step / set_global_step
NDTIMELINE_STREAM_KEY = "stream_key"  # tag naming the stream a region was timed on
from . import predefined  # noqa: F401
from . import chrome_trace_event  # noqa: F401,E402
