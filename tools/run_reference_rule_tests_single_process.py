"""Run reference DTensorTestBase test methods in ONE process against a fake mesh (rules only look at specs)."""
import sys, types, traceback, importlib, unittest
sys.path[:0] = ["/root/repo", "/tmp/vescale_legacy_reftests", "/tmp/vescale_legacy_reftests/stubs"]
import torch
import vescale
from vescale_b200.mesh import DeviceMesh as _DM
modname, clsname = sys.argv[1], sys.argv[2]
mod = importlib.import_module(modname)
mod.DeviceMesh = lambda dev, t, **kw: _DM("cpu", t, _rank=0)
cls = getattr(mod, clsname)
ok = bad = 0
for name in sorted(n for n in dir(cls) if n.startswith("test_")):
    obj = cls.__new__(cls)
    unittest.TestCase.__init__(obj, name)
    obj.rank = 0
    type(obj).device_type = property(lambda self: "cpu")
    fn = getattr(cls, name)
    while hasattr(fn, "__wrapped__"):
        fn = fn.__wrapped__
    try:
        fn(obj); ok += 1; print("PASS", name)
    except Exception:
        bad += 1; print("FAIL", name); traceback.print_exc(limit=3)
print(ok, "passed", bad, "failed")
