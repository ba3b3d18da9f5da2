"""Analyse the clock64 timeline written by GB_TC_FUSED_TRACE=<file> (CTA 0 of the fused conv kernel, second launch of the process):
    python tools/trace_analyze.py <file>
Rows: role (0 producer, 1 convolution issuer, 2 first epilogue warp), plane, 8 stamps.
v2 kernel (default): issuer stamps 0 window start, 1 after the first 12 MMAs, 2 after the NEXT window's barrier waits, 4 after
the last MMA + commits; epilogue stamps 0 before the accumulator wait, 1 accumulator ready, 2 operand staged, 3 pointwise result ready.
Two-CTAs-per-SM kernel (GB_TC_FUSED_V2=0): issuer stamps 0 start, 1 pointwise MMAs issued, 2 slot waits done, 3 TMA box there, 4 issued."""
import sys
import numpy as np
rows = [list(map(int, l.split())) for l in open(sys.argv[1])]
T = {(r[0], r[1]): np.array(r[2:], dtype=np.int64) for r in rows}
n = max(p for (role, p) in T if role == 1 and T[(role, p)][0] > 0) + 1
v2 = all(T[(1, p)][3] == 0 for p in range(min(n, 24)))
per = 24
m0 = np.array([T[(1, p)][0] for p in range(n)], dtype=np.float64)
print("%s kernel, %d planes traced; cycles (SM clock)" % ("v2" if v2 else "two-CTA", n))
d = np.diff(m0)
for i in range(0, n - 1, per):
    print("issuer period per plane:", d[i:i + per].astype(int).tolist())
lo, hi = per, (n // per) * per
def seg(role, k0, k1):
    a = np.array([T[(role, p)][k1] - T[(role, p)][k0] for p in range(lo, hi)], dtype=np.float64)
    return a.reshape(-1, per).mean(0).astype(int).tolist()
if hi > lo:
    if v2:
        print("issuer: first 12 MMAs      ", seg(1, 0, 1))
        print("issuer: next window's waits", seg(1, 1, 2))
        print("issuer: last 6 + commits   ", seg(1, 2, 4))
    else:
        print("issuer: pointwise wait+issue", seg(1, 0, 1))
        print("issuer: slot waits          ", seg(1, 1, 2))
        print("issuer: TMA box wait        ", seg(1, 2, 3))
        print("issuer: issue + commit      ", seg(1, 3, 4))
    e = [p for p in range(lo, hi) if T[(2, p)][1] > 0]
    if e:
        a = np.array([[T[(2, p)][k] for k in range(4)] for p in e], dtype=np.float64)
        print("epilogue (traced warp's planes): accumulator wait %.0f | step 1 %.0f | to pointwise result %.0f" %
              ((a[:, 1] - a[:, 0]).mean(), (a[:, 2] - a[:, 1]).mean(), (a[:, 3] - a[:, 2]).mean()))
