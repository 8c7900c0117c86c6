from .fsdp_adamw import FSDPAdamW  # noqa: F401
from .base_optimizer import BasicOptimizer, BasicOptimizerHook, GradOptimizerHookBase  # noqa: F401
from .clip_grads import clip_grad_norm_fp32, get_grad_norm_fp32  # noqa: F401
from .distributed_optimizer import DistributedOptimizer, OptimizerStateSpec  # noqa: F401
