"""The emulator's reduction kernel (legacy ``emulator/reduce_kernel.py``; NCCL ``reduce_kernel.h``): ONE definition of what
``reduce(a, b)`` means for every op, used by the step-level primitives, the closed-form collectives and the process-group front end.

Bit-exactness against NCCL hinges on three details kept here: the operand ORDER (``reduceCopy`` folds its sources left to right,
local buffer first — floating-point addition does not commute with rounding in a chain), the ACCUMULATION TYPE (NCCL reduces bf16 /
fp16 pairs in fp32 and rounds once per pairwise step, it does not carry a wide accumulator across steps), and ``avg`` being a
``sum`` followed by one division by the group size at the very end (``ncclDevPreMulSum`` / post-div, depending on the version: the
division happens once, never per step)."""
from __future__ import annotations

from typing import Sequence

import torch

__all__ = ["ReduceOp", "op_name", "reduce_pair", "reduce_sources", "finalize"]


class ReduceOp:
    """Reduction selector; also callable: ``ReduceOp.SUM`` is a name, ``ReduceOp("max")(a, b)`` applies it."""

    SUM, PRODUCT, MAX, MIN, AVG = "sum", "product", "max", "min", "avg"

    def __init__(self, op="sum"):
        self.op = op_name(op)

    def __call__(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        return reduce_pair(a, b, self.op)

    def __repr__(self) -> str:
        return f"ReduceOp({self.op})"


def op_name(op) -> str:
    """``"sum"`` / ``ReduceOp.SUM`` / ``torch.distributed.ReduceOp.SUM`` / a ``ReduceOp`` instance -> canonical lower-case name."""
    if isinstance(op, ReduceOp):
        return op.op
    if isinstance(op, str):
        return op.lower()
    name = getattr(op, "name", None) or str(op)
    return {"SUM": "sum", "PRODUCT": "product", "MAX": "max", "MIN": "min", "AVG": "avg"}.get(name.split(".")[-1].upper(), "sum")


_WIDE = {torch.bfloat16: torch.float32, torch.float16: torch.float32}


def reduce_pair(a: torch.Tensor, b: torch.Tensor, op: str) -> torch.Tensor:
    """One pairwise step ``a (op) b`` in the input dtype, computed in fp32 for half types and rounded back (one rounding per step)."""
    op = op_name(op)
    wide = _WIDE.get(a.dtype)
    x, y = (a.to(wide), b.to(wide)) if wide is not None else (a, b)
    if op in ("sum", "avg"):
        r = x + y
    elif op == "max":
        r = torch.maximum(x, y)
    elif op == "min":
        r = torch.minimum(x, y)
    elif op == "product":
        r = x * y
    else:
        raise ValueError(f"unknown reduce op {op!r}")
    return r.to(a.dtype) if wide is not None else r


def reduce_sources(srcs: Sequence[torch.Tensor], op: str) -> torch.Tensor:
    """Left-to-right fold over the sources of one primitive call (the local buffer comes first)."""
    acc = srcs[0].clone()
    for s in srcs[1:]:
        acc = reduce_pair(acc, s, op)
    return acc


def finalize(t: torch.Tensor, op: str, group_size: int) -> torch.Tensor:
    """What happens once after the last step: ``avg`` divides by the group size; every other op is already final."""
    return t / group_size if op_name(op) == "avg" else t
