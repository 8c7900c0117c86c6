"""Error types of the timeline subsystem (legacy ``ndtimeline/exceptions.py``)."""


class ProtocolValidationError(ValueError):
    """A frame on the collector socket does not follow the binary protocol (bad magic / version / length)."""


class NDHandlerError(RuntimeError):
    """A record handler failed; the flusher thread reports it instead of killing the training process."""
