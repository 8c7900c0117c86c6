"""Type vocabulary of the checkpoint API (capability parity: legacy ``checkpoint/api/meta_type.py``).

``save(path, state)`` / ``load(path, state)`` take a *checkpoint state*: a mapping from a role name to an object that can
produce and consume a state dict.  The two conventional roles are ``"model"`` and ``"optimizer"``; any other key is written
as an opaque extra state."""
from __future__ import annotations

import enum
from typing import Any, Dict, Protocol, TypeVar, runtime_checkable

__all__ = ["CheckpointState", "Stateful", "SupportedStrategy", "STATE_DICT_TYPE", "MODEL_STR", "OPTIMIZER_STR", "STATE_DICT_STR"]

MODEL_STR, OPTIMIZER_STR, STATE_DICT_STR = "model", "optimizer", "state_dict"
STATE_DICT_TYPE = Dict[str, Any]


@runtime_checkable
class Stateful(Protocol):
    """Anything with the ``state_dict`` / ``load_state_dict`` pair (modules, optimizers, schedulers, data-loader cursors)."""

    def state_dict(self) -> STATE_DICT_TYPE: ...

    def load_state_dict(self, state_dict: STATE_DICT_TYPE) -> None: ...


_S = TypeVar("_S", bound=Stateful)
CheckpointState = Dict[str, _S]


class SupportedStrategy(enum.Enum):
    """Which framework produced the optimizer state being resharded."""

    Megatron = 0
    FSDP = 1
    VeScale = 2
