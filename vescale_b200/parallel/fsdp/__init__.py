from .api import FSDPState, MixedPrecisionPolicy, checkpoint_module, fsdp_units, fully_shard, get_fsdp_state  # noqa: F401
from .layout import ParamSlot, UnitLayout, row_granularity  # noqa: F401
from .unit import FSDPUnit  # noqa: F401
