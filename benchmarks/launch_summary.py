"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import collections
import csv
import re
import sys


def main(path, top=30, skip_first=0):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    rows = list(csv.DictReader(lines))
    for row in rows[skip_first:]:
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v * 1e6 if unit in ("s", "second") else v
        name = row["Kernel Name"].replace("<unnamed>::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"<.*", "", name)[:80]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"total {T/1e3:.2f} ms over {sum(cnt.values())} launches")
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:top]:
        print(f"{v/1e3:9.3f} ms {100*v/T:5.1f}%  n={cnt[k]:5d}  avg={v/cnt[k]:9.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
