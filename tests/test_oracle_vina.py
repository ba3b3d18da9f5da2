"""CPU checks of the Vina oracle restatement (oracle/vina_ref.c).  The reference's tests hold no absolute numbers for
these functions ("parity unpinned"); what can be checked on the CPU are the published closed forms and internal
consistency of the tabulation."""
import numpy as np
from oracle.vina import VinaOracle, lib

W = np.array([-0.035579, -0.005156, 0.840245, -0.035069, -0.587439, 5 * 0.05846 / 0.1 - 1], np.float32)


def test_terms_closed_form():
    L = lib()
    import ctypes as C
    w = W.ctypes.data_as(C.POINTER(C.c_float))
    # C_H (2) - C_H (2): hydrophobic pair, no h-bond. R = 3.8
    for r in (3.0, 3.8, 4.5, 5.0, 7.9):
        d = r - 3.8
        want = (W[0] * np.exp(-(d / 0.5) ** 2) + W[1] * np.exp(-((d - 3) / 2) ** 2) + W[2] * (d * d if d < 0 else 0) +
                W[3] * (1 if d <= 0.5 else 0 if d >= 1.5 else (1.5 - d)))
        assert abs(L.gvo_eval_terms(w, 2, 2, r) - want) < 2e-7
    # N donor (7) - O acceptor (13): h-bond ramp between -0.7 and 0, not hydrophobic. R = 3.5
    for r in (2.5, 2.8, 3.2, 3.5, 4.0):
        d = r - 3.5
        hb = 1 if d <= -0.7 else 0 if d >= 0 else d / -0.7
        want = W[0] * np.exp(-(d / 0.5) ** 2) + W[1] * np.exp(-((d - 3) / 2) ** 2) + W[2] * (d * d if d < 0 else 0) + W[4] * hb
        assert abs(L.gvo_eval_terms(w, 7, 13, r) - want) < 2e-7
        assert L.gvo_eval_terms(w, 7, 13, r) == L.gvo_eval_terms(w, 13, 7, r)


def test_linear_tables():
    o = VinaOracle()
    assert o.n == 2051
    fast, se, sd = o.table(2, 13)
    rs = np.sqrt(np.arange(o.n + 2, dtype=np.float32) / np.float32(32)).astype(np.float32)
    assert np.allclose(fast[:-1], (se[:-1] + se[1:]) / 2, atol=1e-7)
    i = np.arange(1, o.n - 1)
    want = (se[i + 1] - se[i - 1]) / ((rs[i + 1] - rs[i - 1]) * rs[i])
    assert np.allclose(sd[i], want, rtol=1e-5, atol=1e-7) and sd[0] == 0 and sd[-1] == 0
    # eval_deriv interpolates, eval_fast is piecewise constant
    r2 = 20.3
    e, dor = o.eval_deriv(2, 13, r2)
    k = int(32 * r2); rem = np.float32(32 * r2) - k
    assert abs(e - (se[k] + rem * (se[k + 1] - se[k]))) < 1e-7 and o.eval_fast(13, 2, r2) == fast[k]
    # the interpolated table tracks the exact terms, and dor ~ (dE/dr)/r
    assert abs(e - o.exact(2, 13, r2)) < 5e-4
    h = 1e-3
    num = (o.exact(2, 13, (np.sqrt(r2) + h) ** 2) - o.exact(2, 13, (np.sqrt(r2) - h) ** 2)) / (2 * h) / np.sqrt(r2)
    assert abs(dor - num) < 2e-3


def test_cache_populate_and_trilinear_eval_are_consistent():
    from gnina_b200 import synth
    o = VinaOracle()
    rx, rt = synth.make_receptor(300, box=24)
    begin, end, n = [-6, -6, -6], [6, 6, 6], [8, 8, 8]
    g = o.cache_populate(begin, end, n, rx, rt, 2)
    assert g.shape == (9, 9, 9)
    # a grid point equals the direct sum of eval_fast over receptor atoms within the cutoff
    p = np.array([-6 + 1.5 * 3, -6 + 1.5 * 2, -6 + 1.5 * 5], np.float32)
    r2 = ((rx - p) ** 2).sum(1)
    want = sum(o.eval_fast(int(t), 2, float(q)) for t, q in zip(rt, r2) if q <= 64)
    assert abs(g[5, 2, 3] - want) < 1e-5
    # trilinear: exact at nodes, linear along an edge, penalty outside, analytic derivative matches differences
    grids = {2: g}
    e, _ = VinaOracle.cache_eval(grids, begin, end, n, p[None], [2], 1e3, 1000.0)
    cap = lambda x: x * 1000.0 / (1000.0 + x) if x > 0 else x
    assert abs(e - cap(g[5, 2, 3])) < 1e-5
    q = p + np.array([0.4, 0.7, 0.2], np.float32)
    e0, d0 = VinaOracle.cache_eval(grids, begin, end, n, q[None], [2], 1e3, 1000.0)
    for ax in range(3):
        hh = 1e-2
        qp, qm = q.copy(), q.copy(); qp[ax] += hh; qm[ax] -= hh
        fd = (VinaOracle.cache_eval(grids, begin, end, n, qp[None], [2], 1e3, 1000.0)[0] -
              VinaOracle.cache_eval(grids, begin, end, n, qm[None], [2], 1e3, 1000.0)[0]) / (2 * hh)
        assert abs(fd - d0[0, ax]) < 2e-3 * max(1.0, abs(fd))
    out = np.array([[7.5, 0, 0]], np.float32)   # 1.5 A outside in +x
    eo, do = VinaOracle.cache_eval(grids, begin, end, n, out, [2], 1e3, 1000.0)
    ein, _ = VinaOracle.cache_eval(grids, begin, end, n, np.array([[6.0, 0, 0]], np.float32), [2], 1e3, 1000.0)
    assert abs(eo - (ein + 1.5e3)) < 1e-2 and do[0, 0] == 1e3
    # hydrogens are skipped
    eh, dh = VinaOracle.cache_eval(grids, begin, end, n, q[None], [1], 1e3, 1000.0)
    assert eh == 0 and not dh.any()


def test_final_scoring_pieces():
    from gnina_b200 import synth
    o = VinaOracle()
    rx, rt = synth.make_receptor(400, box=26)
    lx, lt = synth.make_ligand(14, 3)
    e = o.naive_exact(rx, rt, lx, lt)
    # direct restatement in numpy (per-atom partial sums, curl with v = 1000)
    tot = 0.0
    for a, t1 in zip(lx, lt):
        if t1 <= 1:
            continue
        r2 = ((rx - a) ** 2).sum(1)
        s = sum(o.exact(int(t1), int(t2), float(q)) for t2, q in zip(rt, r2) if q < 64 and t2 > 1)
        tot += s * 1000.0 / (1000.0 + s) if s > 0 else s
    assert abs(e - tot) < 1e-4 * max(1.0, abs(tot))
    w = 0.1 * (W[5] + 1)
    assert abs(o.num_tors_div(-7.5, 4.0) - (-7.5 / (1 + w * 4.0 / 5.0))) < 1e-6
    assert o.num_tors_div(0.0, 3.0) == 0.0


def test_splines_interpolate_the_terms_with_zero_end_slopes():
    o = VinaOracle()
    tab = o.spline_table(2, 13)
    assert tab.shape == (80, 4)
    for i in (0, 17, 33, 38, 60):           # knots are reproduced, derivative tracks the exact terms
        r = i * 0.1
        e, dor = o.spline_eval_deriv(2, 13, r * r if i else 1e-12)
        assert abs(e - o.exact(2, 13, r * r)) < 2e-6 * max(1.0, abs(e))
    assert abs(tab[0, 2]) < 1e-5             # zero first derivative at r = 0 (c coefficient of the first interval)
    r = 4.41
    e, dor = o.spline_eval_deriv(13, 2, r * r)
    h = 1e-3
    num = (o.exact(2, 13, (r + h) ** 2) - o.exact(2, 13, (r - h) ** 2)) / (2 * h) / r
    assert abs(e - o.exact(2, 13, r * r)) < 1e-5 and abs(dor - num) < 1e-4
    assert o.spline_eval_deriv(2, 13, 64.0) == (0.0, 0.0)
