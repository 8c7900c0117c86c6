"""Explicit DTensor collectives (legacy ``vescale.dtensor.api``: vescale_all_gather / all_reduce / reduce_scatter); the
implementation lives next to ``redistribute_dtensor`` in ``api.py``."""
from .api import vescale_all_gather, vescale_all_reduce, vescale_reduce_scatter  # noqa: F401

__all__ = ["vescale_all_gather", "vescale_all_reduce", "vescale_reduce_scatter"]
