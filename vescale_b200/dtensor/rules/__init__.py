"""Sharding rules.  Importing this package registers every rule with the propagator."""
from . import common, pointwise, math, matrix, view, tensor, conv, dim_maps  # noqa: F401
