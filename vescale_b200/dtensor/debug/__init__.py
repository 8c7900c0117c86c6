"""Debug helpers (reference ``vescale/dtensor/debug``): CommDebugMode, visualize_sharding."""
from ._comm_mode import CommDebugMode  # noqa: F401
from ._visualize import visualize_sharding  # noqa: F401
from ..sharding_prop import propagator as _propagator


def _get_sharding_prop_cache_info():
    return _propagator.cache_info()


__all__ = ["CommDebugMode", "visualize_sharding"]
