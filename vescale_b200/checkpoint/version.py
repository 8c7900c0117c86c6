"""Checkpoint format version (legacy ``checkpoint/version.py``).  The on-disk layout is torch DCP's; this number versions the
conventions ON TOP of it (directory names ``model`` / ``optimizer`` / ``optimizer/pp_{rank}``, the ``__extras__`` entry, FlatPiece
boxes for optimizer state)."""
__version__ = "0.2.0"
