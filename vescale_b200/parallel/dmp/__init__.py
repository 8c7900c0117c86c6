from .dmp import auto_parallelize_module, set_plan_overriding_policy, get_plan_overriding_policy  # noqa: F401
from .registry import register_policy, get_policy, register_provider  # noqa: F401
from .policies import megatron  # noqa: F401
