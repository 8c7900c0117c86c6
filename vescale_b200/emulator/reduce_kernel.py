"""Reduction selector of the emulator (legacy ``emulator/reduce_kernel.py``)."""
from .distributed import ReduceOp  # noqa: F401
