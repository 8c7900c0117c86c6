"""``torchdistx`` names the reference's users import next to ``vescale.initialize`` (the reference vendors a patched torchdistX,
``legacy/patches/patched_torchdistX_9c1b9f.patch``).  Here deferred construction is the meta device plus an init recorder
(``vescale_b200/initialize/deferred_init.py``); this package only forwards the familiar entry points to it."""
from . import deferred_init, fake  # noqa: F401
