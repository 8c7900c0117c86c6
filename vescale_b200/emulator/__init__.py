"""Single-process emulation of NCCL collectives with NCCL's reduction order, plus a global-view DTensor
redistribute built on them — for validating distributed numerics bit-for-bit without GPUs.
Parity: ``legacy/vescale/emulator`` (all_reduce ring/tree, all_gather, reduce_scatter, all_to_all, ProcessGroup,
mesh_collectives, comm_api)."""
from .collectives import EmulatorProcessGroup, ring_all_reduce, ring_reduce_scatter, tree_all_reduce, all_gather, all_to_all, nccl_chunking  # noqa: F401
from .comm_api import distribute_tensor, redistribute_dtensor, full_tensor, mesh_all_reduce, mesh_all_gather, mesh_reduce_scatter  # noqa: F401
