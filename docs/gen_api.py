"""Regenerates docs/API.md from the public names (``__all__``) and docstring first lines of the main modules."""
import importlib
import inspect
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

MODULES = [
    "vescale_b200", "vescale_b200.dtensor", "vescale_b200.dtensor.debug", "vescale_b200.dtensor.random", "vescale_b200.comm.collectives", "vescale_b200.comm.symm",
    "vescale_b200.comm.symm_collectives", "vescale_b200.comm.fused_tp", "vescale_b200.comm.symm_debug", "vescale_b200.parallel.fsdp", "vescale_b200.parallel.ddp",
    "vescale_b200.parallel.dmodule", "vescale_b200.parallel.dmp", "vescale_b200.parallel.pipe", "vescale_b200.parallel.moe", "vescale_b200.parallel.context", "vescale_b200.optim",
    "vescale_b200.checkpoint", "vescale_b200.profiler", "vescale_b200.emulator", "vescale_b200.initialize", "vescale_b200.devicemesh_api", "vescale_b200.model.patch",
    "vescale_b200.models", "vescale_b200.ops", "vescale_b200.ops.fp8", "vescale_b200.utils", "vescale_b200.debug",
]


def main():
    out = io.StringIO()
    out.write("# API index (generated from the modules' public names and the first line of their docstrings)\n\nRegenerate: `python docs/gen_api.py`.\n")
    for mn in MODULES:
        m = importlib.import_module(mn)
        names = getattr(m, "__all__", None) or [n for n in dir(m) if not n.startswith("_")]
        doc = (inspect.getdoc(m) or "").strip().splitlines()
        out.write(f"\n## `{mn}`\n\n{doc[0] if doc else ''}\n\n")
        seen = set()
        for n in names:
            if n in seen:
                continue
            seen.add(n)
            try:
                o = getattr(m, n)
            except Exception:  # noqa: BLE001
                continue
            if inspect.ismodule(o):
                continue
            own = o.__dict__.get("__doc__") if inspect.isclass(o) else getattr(o, "__doc__", None)
            d = (inspect.cleandoc(own) if own else "").splitlines()
            first = d[0] if d and not isinstance(o, (int, float, str, dict, list, tuple)) else ""
            kind = "class" if inspect.isclass(o) else "def" if callable(o) else "const"
            try:
                sig = str(inspect.signature(o)) if callable(o) and not inspect.isclass(o) else ""
            except Exception:  # noqa: BLE001
                sig = ""
            if len(sig) > 110:
                sig = sig[:107] + "...)"
            out.write(f"- **{n}**{'`' + sig + '`' if sig else ''} — *{kind}*{': ' + first if first else ''}\n")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "API.md")
    open(path, "w").write(out.getvalue())
    print(path, len(out.getvalue().splitlines()), "lines")


if __name__ == "__main__":
    main()
