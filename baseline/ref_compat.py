"""torch-2.11 compatibility shim for the UNMODIFIED reference installed in ``baseline/_ref``.

The reference (volcengine/veScale @ 20cf5c7) pins ``torch==2.7.1`` and imports torch-private DTensor symbols that
torch 2.11 renamed or moved into C++.  This module lives OUTSIDE ``baseline/_ref`` and only (a) re-creates the missing
private names inside torch's own modules and (b) adapts the two torch-side behaviour changes the reference's
RaggedShard path trips over.  No file of the reference is edited; no code of ``vescale_b200`` is imported.

    import baseline.ref_compat as rc; rc.install(); import vescale
"""
from __future__ import annotations

import os
import sys

REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_installed = False


def install() -> None:
    global _installed
    if _installed:
        return
    import torch.distributed.tensor._op_schema as s

    # torch 2.7 names (torch/distributed/tensor/_op_schema.py @ v2.7.1) that 2.11 dropped / renamed
    if not hasattr(s, "_is_inplace_op"):
        s._is_inplace_op = lambda op: op._schema.name[-1] == "_"
    if not hasattr(s, "_is_out_variant_op"):
        s._is_out_variant_op = lambda op: "out" in op._schema.overload_name
    if not hasattr(s, "PlacementStrategy"):
        s.PlacementStrategy = s.OpSpec
    import functools

    import torch.distributed.tensor._redistribute as tr

    if "is_backward" not in tr.redistribute_local_tensor.__code__.co_varnames:
        _orig_redist = tr.redistribute_local_tensor

        @functools.wraps(_orig_redist)
        def redistribute_local_tensor(local_tensor, current_spec, target_spec, *, async_op=False, is_backward=False, **kw):
            # torch 2.7 signature: (local, cur, tgt, *, async_op, is_backward); 2.11 dropped is_backward
            return _orig_redist(local_tensor, current_spec, target_spec, async_op=async_op, **kw)

        tr.redistribute_local_tensor = redistribute_local_tensor
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "vescale" or k.startswith("vescale.")]:
        del sys.modules[k]
    import vescale.dtensor.placement_types as pt
    from torch.distributed.tensor.placement_types import Placement

    # torch 2.11 made Placement a C++ class whose __init__ must run; the reference's frozen dataclass never calls it.
    for cls in (pt.RaggedShard, getattr(pt, "_StridedRaggedShard", None)):
        if cls is None or getattr(cls.__init__, "_compat", False):
            continue
        orig = cls.__init__

        def __init__(self, *a, __orig=orig, **k):
            Placement.__init__(self)
            __orig(self, *a, **k)

        __init__._compat = True
        cls.__init__ = __init__
    _installed = True
