from .layer import GroupedExperts, MoEConfig, MoELayer, TopKRouter, all_to_all_uneven, ragged_token_placement  # noqa: F401
from .api import BasicExpertsAllocator, BasicTokenDispatcher, ExpertsAllocator, MoEOptimizer, TokenDispatcher, is_experts_parallized, is_moe, parallelize_experts  # noqa: F401
from .hijack import EPExperts, hijack_moe_block, is_hijackable  # noqa: F401
from .param_buffer import MoELayerParamBuffer  # noqa: F401
from .scheduler import ExpertsAllocation, MoEScheduler, MoETask, ScheduledMoELayer  # noqa: F401
