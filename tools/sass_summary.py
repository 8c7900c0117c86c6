#!/usr/bin/env python
"""Per-kernel counts of the tensor-core / TMA / TMEM / mbarrier / cluster / system-scope / multimem SASS mnemonics of
vescale_b200/_C.so (the listing committed as profiles/sass_*.txt).   python tools/sass_summary.py > profiles/sass_rN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "vescale_b200", "_C.so")
PAT = re.compile(
    r"\b(UTC[A-Z]*MMA[\w.]*|UTCCP[\w.]*|UTCBAR[\w.]*|UTCATOMSWS[\w.]*|UTMALDG[\w.]*|UTMASTG[\w.]*|UTMAREDG[\w.]*|UBLKCP[\w.]*|LDTM[\w.]*|STTM[\w.]*|SYNCS[\w.]*|"
    r"UCGABAR_\w+|LDGMC[\w.]*|STGMC[\w.]*|REDGMC[\w.]*|MULTIMEM[\w.]*|ERRBAR|MEMBAR\.\w+\.SYS|LDG\.E[\w.]*\.STRONG\.SYS|STG\.E[\w.]*\.STRONG\.SYS|ATOMG[\w.]*\.SYS[\w.]*|"
    r"HMMA[\w.]*|QMMA[\w.]*)"
)


def _strip_params(name: str) -> str:
    depth = 0
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return name[:i]
    return name


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = PAT.search(line)
        if m:
            kernels[cur][m.group(1)] += 1
    names = list(kernels)
    try:
        dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dem))
    except Exception:  # noqa: BLE001
        pass
    print("# cuobjdump -sass vescale_b200/_C.so (sm_100a): per-kernel counts of the tensor-core / TMA / TMEM / mbarrier / cluster /")
    print("# system-scope / multimem mnemonics (tools/sass_summary.py; kernels without any of them are omitted)")
    for k, c in kernels.items():
        if not c:
            continue
        name = demangle.get(k, k)
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = _strip_params(re.sub(r"^void ", "", name).replace("<unnamed>::", ""))
        print(f"\n== {name}")
        for mn in sorted(c):
            print(f"{c[mn]:7d} {mn}")


if __name__ == "__main__":
    sys.exit(main())
