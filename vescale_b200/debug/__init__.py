from .debug_log import DebugLogger, set_vescale_debug_mode, update_vescale_debug_mode_from_env  # noqa: F401
from .pdb import ForkedPdb  # noqa: F401
