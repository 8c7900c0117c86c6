"""Small helpers of the emulator front end (legacy ``emulator/utils.py``)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from .distributed import ReduceOp, _op_name

__all__ = ["flatten_tensors", "restore_tensors", "torch_reduce_op_to_emulator", "emulator_reduce_op_to_torch"]


def flatten_tensors(tensor_list: Sequence[torch.Tensor]) -> Tuple[List[torch.Tensor], List[torch.Size]]:
    """Every tensor as a 1-D view plus the shapes needed to undo it."""
    return [t.reshape(-1) for t in tensor_list], [t.shape for t in tensor_list]


def restore_tensors(flattened_list: Sequence[torch.Tensor], original_shapes) -> List[torch.Tensor]:
    return [t.reshape(tuple(s)) for t, s in zip(flattened_list, original_shapes)]


def torch_reduce_op_to_emulator(torch_reduce_op) -> str:
    return _op_name(torch_reduce_op)


def emulator_reduce_op_to_torch(reduce_op):
    import torch.distributed as dist

    return {ReduceOp.SUM: dist.ReduceOp.SUM, ReduceOp.PRODUCT: dist.ReduceOp.PRODUCT, ReduceOp.MAX: dist.ReduceOp.MAX, ReduceOp.MIN: dist.ReduceOp.MIN,
            ReduceOp.AVG: dist.ReduceOp.AVG}[_op_name(reduce_op)]
