"""ndtimeline: multi-rank CUDA-event timeline with a simulated global clock
(parity: ``legacy/vescale/ndtimeline`` — timer.py, api.py, handlers/*, pool.py, predefined.py)."""
from .timer import (  # noqa: F401
    NDMetricLevel,
    NDTimerManager,
    flush,
    inc_step,
    init_ndtimers,
    is_initialized,
    ndtimeit,
    ndtimeit_p2p,
    ndtimer,
    set_global_step,
    wait,
)
from .handlers import ChromeTraceNDHandler, LocalRawNDHandler, LocalTimelineNDHandler, LoggingNDHandler, NDHandler, ParserNDHandler  # noqa: F401
from .sock_streamer import NDtimelineStreamer, SockNDHandler, decode_frames, encode_frame  # noqa: F401
from . import predefined  # noqa: F401
