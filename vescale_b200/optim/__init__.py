from .fsdp_adamw import FSDPAdamW  # noqa: F401
