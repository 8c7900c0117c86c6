"""Communication backends: c10d collectives (plumbing + baseline) and sm_100a symmetric-memory kernels."""
from . import collectives  # noqa: F401
