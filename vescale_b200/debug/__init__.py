from .debug_log import DebugLogger, set_vescale_debug_mode  # noqa: F401
from .pdb import ForkedPdb  # noqa: F401
