from .api import parallelize_module, is_dmodule, PlacementsInterface, DModule  # noqa: F401
