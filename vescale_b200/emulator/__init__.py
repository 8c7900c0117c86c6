"""Single-process emulation of NCCL collectives with NCCL's reduction order, plus a global-view DTensor
redistribute built on them — for validating distributed numerics bit-for-bit without GPUs.
Parity: ``legacy/vescale/emulator`` (all_reduce ring/tree, all_gather, reduce_scatter, all_to_all, ProcessGroup,
mesh_collectives, comm_api, comm_primitive, topo, nccl/graph/tuning, calculate_chunk_size, emulator_instrumentation)."""
from .collectives import EmulatorProcessGroup, ring_all_reduce, ring_reduce_scatter, tree_all_reduce, double_tree_all_reduce, all_gather, all_to_all, nccl_chunking  # noqa: F401
from .comm_api import distribute_tensor, redistribute_dtensor, full_tensor, mesh_all_reduce, mesh_all_gather, mesh_reduce_scatter, mesh_all_to_all, mesh_broadcast, mesh_scatter  # noqa: F401
from . import distributed  # noqa: F401
from .distributed import ProcessGroup, ReduceOp  # noqa: F401
from .device_mesh import DeviceMesh, init_device_mesh  # noqa: F401
from .comm_primitive import P2R, P2S, R2P, R2R, R2S, S2R, S2S  # noqa: F401
from .emulator_instrumentation import EmulatorInstrumentation, map_over_ranks  # noqa: F401
from .topo import BinaryTree, DoubleTree, Ring, btree, double_tree, parse_graph_dump  # noqa: F401
from .tuning import Tuning, calculate_chunk_size, select_algorithm  # noqa: F401
from . import algorithms, chunk_math, nccl, primitives  # noqa: F401,E402
from .algorithms import chunk_layout, run_all_to_all, run_broadcast, run_ring_all_gather, run_ring_all_reduce, run_ring_reduce_scatter, run_tree_all_reduce  # noqa: F401,E402
from .primitives import Point2PointPrimitive, RingPrimitive, Traffic, TreePrimitive  # noqa: F401,E402
