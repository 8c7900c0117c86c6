"""Small shared utilities: method patching (reference ``vescale/utils/monkey_patch.py``), the environment-flag registry
(SURVEY §5.6 — the reference scatters these over modules), MFU accounting used by the examples and ``bench.py``."""
from .env import FLAGS, describe_flags, flag  # noqa: F401
from .mfu import llama_mfu, mixtral_flops_per_token, model_tflops  # noqa: F401
from .monkey_patch import patch_method, unpatch_all  # noqa: F401
from .compat import deprecated_function, switch_dtensor_for_torch_export  # noqa: F401
