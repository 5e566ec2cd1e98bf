"""Per-kernel counts of the Blackwell-native SASS mnemonics in libtcvn.so (cuobjdump -sass): UTC*MMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTMALDG / UTMASTG (TMA), UTCBAR (tcgen05.commit), HMMA (legacy mma.sync: should be absent).
usage: python scripts/sass_summary.py [libtcvn.so] > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dune_transformercvn_b200", "libtcvn.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "HGMMA"]
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "")
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for p in pats:
        if re.search(r"\b" + p + r"\b|\b" + p + r"\.", line):
            counts[cur][p] += 1
print(f"# cuobjdump -sass {os.path.basename(so)} (sm_100a): Blackwell-native mnemonics per kernel; kernels without any are omitted")
tot = collections.Counter()
for k, c in counts.items():
    if sum(c.values()) == 0:
        continue
    tot.update(c)
    print(f"{k[:110]:110s} " + "  ".join(f"{p}={c[p]}" for p in pats if c[p]))
print("TOTAL " + "  ".join(f"{p}={tot[p]}" for p in pats))
print(f"kernels in the library: {len(counts)}")
