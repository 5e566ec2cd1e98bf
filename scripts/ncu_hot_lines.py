"""Top CUDA source lines by warp-stall samples from an .ncu-rep captured with --import-source on (-lineinfo build)."""
import collections
import csv
import os
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    path = "?"
    hdr = None
    agg = collections.defaultdict(lambda: [0.0, 0.0, ""])
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            path = os.path.basename(r[1])
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            si = hdr.index("Warp Stall Sampling (All Samples)")
            ii = hdr.index("Instructions Executed")
            key = (path, int(r[0]))
            agg[key][0] += float(r[si] or 0)
            agg[key][1] += float(r[ii] or 0)
            agg[key][2] = r[1]
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
    for (p, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * v[0] / tot:5.1f}% smp {100 * v[1] / toti:5.1f}% inst | {p}:{ln:<4} {v[2].strip()[:100]}")


if __name__ == "__main__":
    main()
