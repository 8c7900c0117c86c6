from .deferred_init import deferred_init, is_deferred, materialize_dparameter, materialize_dtensor, materialize_module  # noqa: F401
