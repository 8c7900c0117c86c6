"""Hand-written sm_100a kernels (csrc/*.cu) and their differentiable Python front-ends."""
from . import _ext  # noqa: F401
from .functional import *  # noqa: F401,F403
from .functional import norm_linear, swiglu_linear, rope_tables, gemm_nn, gemm_tn  # noqa: F401
