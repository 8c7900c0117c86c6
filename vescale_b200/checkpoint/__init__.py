"""vescale_b200.checkpoint — ``save`` / ``load`` of model + optimizer state with on-load resharding.

Layout on disk is plain torch DCP (``.metadata`` + ``__{rank}_{i}.distcp``) under ``path/model`` and
``path/optimizer`` (legacy ``vescale.checkpoint`` directory convention), produced through the DTensor
``__create_write_items__ / __create_chunk_list__ / __get_tensor_shard__`` protocol, so RaggedShard tensors are
saved *without communication* as a few axis-aligned boxes per rank and can be reloaded under any other
placement, mesh shape or world size (reference ``vescale/dtensor/vescale_utils/checkpoint.py``,
``docs/texts/raggedshard.md:93-95``).

    vescale_b200.checkpoint.save(path, {"model": model, "optimizer": optimizer}, async_checkpoint=True)
    vescale_b200.checkpoint.load(path, {"model": model, "optimizer": optimizer})

``path`` may be ``mem://host:port/dir``: the shards then land in a ``MemFileServer`` (gRPC, node-local memory) that persists
them to disk in the background and serves fast restarts.  ``load(..., broadcast_checkpoint=True)`` makes one rank read the
replicated entries and broadcast them.

Parity: ``legacy/vescale/checkpoint/__init__.py``, ``api/vescale_checkpointer.py:71-249``, planner cache / balanced dedup
``planner/common.py:65-132``, ``utilities/server/*`` (in-memory file server, report service).
"""
from .api import VeScaleCheckpointer, load, save, wait_for_async  # noqa: F401
from .pinned_pool import PinnedPool  # noqa: F401
from .meta_type import MODEL_STR, OPTIMIZER_STR, STATE_DICT_TYPE, CheckpointState, Stateful, SupportedStrategy  # noqa: F401
from .mem_server import MemFileClient, MemFileServer  # noqa: F401
from . import bfile  # noqa: F401
from .logger import get_vescale_checkpoint_logger  # noqa: F401
from .recorder import TorchCheckpointRecorder  # noqa: F401
from .planner import PlanLRUCache, VeScaleLoadPlanner, VeScaleSavePlanner, custom_dedup_tensors  # noqa: F401
from .state_dict_io import CheckpointException, ServiceComm, load_state_dict, save_state_dict  # noqa: F401
from .sync_queue import SynchronizedQueue  # noqa: F401
from .version import __version__ as CHECKPOINT_FORMAT_VERSION  # noqa: F401
