from .dmp import PlanGenerator, auto_parallelize_module, generate_plan, get_plan_overriding_policy, set_plan_overriding_policy  # noqa: F401
from .registry import register_policy, get_policy, register_provider  # noqa: F401
from .policies import megatron  # noqa: F401
