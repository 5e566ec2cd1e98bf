"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), v, r["Grid Size"]))
    return rows


def main():
    rows = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for name, v, _ in rows:
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"{len(rows)} launches, {tot:.1f} us total (cold-cache, serialised: compare SHARES)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d}  avg={t / n:8.1f} us  {k[:100]}")
    if "--list" in sys.argv:
        for name, v, grid in rows:
            print(f"{v:9.1f} {name[:70]} grid={grid}")


if __name__ == "__main__":
    main()
