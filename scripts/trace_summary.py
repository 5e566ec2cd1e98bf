"""Per-role timeline of one CTA of umma_gemm_kernel from the clock stamps the role-trace patch writes
(scripts/patches/conv1_role_trace.patch; build with it, run `TCVN_C1_TRACE=out.txt [TCVN_C1_TRACE_K=128] python
scripts/profile_cnn.py 194 2 --sparse`).  Rows of the file = roles, columns = successive events of CTA 0 in clock64() ticks.
usage: python scripts/trace_summary.py out.txt [sm_clock_ghz=1.92]

What to read off it ("what actually bounded conv1"):
  producer issue(i) -> transform 'data landed'(i)      = latency of a stage's TMA load under load
  transform landed(i) -> done(i)                        = the in-place BN+PReLU of one 16 KB chunk
  MMA 'stage picked up' period                          = chunk throughput; x stages = how long a stage lives
  epilogue: accumulator full -> released -> store       = how long a group holds a TMEM accumulator; gaps between a group's
                                                          store and its next 'full' mean the epilogue is NOT the limit
"""
import sys

import numpy as np

NAMES = ["producer: stage issued (chunk)", "MMA: accumulator free (tile)", "MMA: stage picked up (chunk)",
         "transform: data landed (chunk)", "transform: done (chunk)",
         "epilogue 0: accumulator full", "epilogue 0: staging free", "epilogue 0: accumulator released", "epilogue 0: store issued",
         "epilogue 1: accumulator full", "epilogue 1: staging free", "epilogue 1: accumulator released", "epilogue 1: store issued"]


def main():
    ghz = float(sys.argv[2]) if len(sys.argv) > 2 else 1.92
    rows = [list(map(int, l.split())) for l in open(sys.argv[1]) if l.strip() and not l.startswith("#")]
    a = np.array(rows[:len(NAMES)], dtype=np.float64)
    t0 = a[a > 0].min()
    a = np.where(a > 0, (a - t0) / (ghz * 1e3), np.nan)   # microseconds
    for name, r in zip(NAMES, a):
        v = r[~np.isnan(r)]
        if len(v) < 12:
            continue
        period = (v[-1] - v[10]) / (len(v) - 11)
        print(f"{name:36s} period {period:6.3f} us   first events: " + " ".join(f"{x:6.2f}" for x in v[:12]))
    issue, landed, done, picked = a[0], a[3], a[4], a[2]
    n = int(min(np.sum(~np.isnan(x)) for x in (issue, landed, done, picked)))
    if n > 12:
        print(f"load latency (issue -> landed), chunks 10..{n - 1}: mean {np.nanmean(landed[10:n] - issue[10:n]):.2f} us, "
              f"max {np.nanmax(landed[10:n] - issue[10:n]):.2f} us")
        print(f"transform (landed -> done): mean {np.nanmean(done[10:n] - landed[10:n]):.2f} us")
        print(f"hand-off (transform done -> MMA picked up): mean {np.nanmean(picked[10:n] - done[10:n]):.2f} us")


if __name__ == "__main__":
    main()
