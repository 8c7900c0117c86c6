from .api import parallelize_module, is_dmodule, PlacementsInterface, DModule  # noqa: F401
from . import _factory  # noqa: F401,E402
