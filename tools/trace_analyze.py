"""Analyse the clock64 timeline written by GB_TC_FUSED_TRACE (CTA 0 of conv1_pw2_pool_kernel): python tools/trace_analyze.py <file>"""
import sys
import numpy as np
rows = [list(map(int, l.split())) for l in open(sys.argv[1])]
T = {}
for r in rows:
    T[(r[0], r[1])] = np.array(r[2:], dtype=np.int64)
def col(role, k, lo=8, hi=88):
    return np.array([T[(role, p)][k] for p in range(lo, hi)], dtype=np.float64)
t0 = min(T[(1, 8)][0], T[(0, 8)][0])
print("planes 8..87 of CTA 0; cycles (SM clock)")
m0, m1, m2, m3, m4 = (col(1, k) for k in range(5))
print("MMA warp per plane: total %.0f | pw wait+issue %.0f | acce wait %.0f | full(TMA) wait %.0f | issue+commit %.0f" %
      (np.diff(m0).mean(), (m1 - m0).mean(), (m2 - m1).mean(), (m3 - m2).mean(), (m4 - m3).mean()))
p0, p1 = col(0, 0), col(0, 1)
print("producer per plane: total %.0f | empty wait %.0f" % (np.diff(p0).mean(), (p1 - p0).mean()))
e0, e1, e2, e3 = (col(2, k) for k in range(4))
print("epilogue per plane: total %.0f | accf wait %.0f | step1 %.0f | to d2 ready %.0f | rest %.0f" %
      (np.diff(e0).mean(), (e1 - e0).mean(), (e2 - e1).mean(), (e3 - e2).mean(), (np.roll(e0, -1) - e3)[:-1].mean()))
# relative lags: TMA issue (p1) -> data seen by MMA warp (m3) for the same plane
print("TMA issue -> MMA warp sees the slab: %.0f (same plane)" % (m3 - p1).mean())
print("MMA issue done (m4, plane c) -> epilogue sees accf of plane c-1 (e1, index c-1): %.0f" % (e1[1:] - m4[1:] + 0 * 1).mean() if True else "")
for p in range(8, 20):
    print(p, "MMA", (T[(1, p)][:5] - t0).tolist(), "EPI", (T[(2, p)][:4] - t0).tolist(), "PROD", (T[(0, p)][:2] - t0).tolist())
