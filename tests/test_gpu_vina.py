"""GPU parity of the Vina scoring rows (V2, V4, V5, V12) through the C ABI vs the CPU oracle restatement.
north_star tolerance: 1e-6 on the Vina score (taken relative to max(1, |score|))."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-6


@pytest.fixture(scope="module")
def system():
    from gnina_b200 import synth
    rx, rt = synth.make_receptor(1500, box=44)
    lx, lt, offs = synth.make_screen(25, seed=4, trans_box=14)
    return rx, rt, lx, lt, offs


def test_tables_bit_identical_to_oracle():
    from gnina_b200.vina import VinaScorer
    from oracle.vina import VinaOracle
    v, o = VinaScorer(), VinaOracle()
    assert v.n == o.n == 2051
    for t1, t2 in ((2, 2), (2, 13), (7, 13), (13, 7), (12, 12), (4, 17), (23, 9)):
        for a, b in zip(v.table(t1, t2), o.table(t1, t2)):
            assert np.array_equal(a, b)


def test_cache_populate_matches_oracle(system):
    from gnina_b200.vina import VinaScorer
    from oracle.vina import VinaOracle
    rx, rt, *_ = system
    v, o = VinaScorer(), VinaOracle()
    v.set_receptor(rx, rt)
    begin, end, n = [-9, -9, -9], [9, 9, 9], [24, 24, 20]
    v.cache_build(begin, end, n, [2, 7, 13])
    for t in (2, 7, 13):
        g, ref = v.cache_grid(t), o.cache_populate(begin, end, n, rx, rt, t)
        assert g.shape == ref.shape
        assert np.abs(g - ref).max() <= TOL * max(1.0, np.abs(ref).max())


def test_cache_eval_and_deriv_match_oracle(system):
    from gnina_b200.vina import VinaScorer
    from oracle.vina import VinaOracle
    rx, rt, lx, lt, offs = system
    v, o = VinaScorer(), VinaOracle()
    v.set_receptor(rx, rt)
    begin, end, n = [-9, -9, -9], [9, 9, 9], [48, 48, 48]   # 0.375 A like box_granularity (main.cpp:622)
    needed = sorted(set(int(t) for t in lt if t > 1))
    v.cache_build(begin, end, n, needed)
    grids = {t: v.cache_grid(t) for t in needed}
    # push a few poses partly out of the box to exercise the penalty branch
    lx2 = lx.copy(); lx2[offs[3]:offs[4]] += np.array([8.0, 0, 0], np.float32)
    e, d = v.cache_eval(lx2, lt, offs)
    for p in range(len(offs) - 1):
        sl = slice(offs[p], offs[p + 1])
        er, dr = VinaOracle.cache_eval(grids, begin, end, n, lx2[sl], lt[sl], 1e3, 1000.0)
        assert abs(e[p] - er) <= TOL * max(1.0, abs(er))
        assert np.abs(d[sl] - dr).max() <= TOL * max(1.0, np.abs(dr).max())
    e_only, _ = v.cache_eval(lx2, lt, offs, deriv=False)
    assert np.abs(e_only - e).max() <= 1e-5 * max(1.0, np.abs(e).max())


def test_exact_final_score_matches_oracle(system):
    from gnina_b200.vina import VinaScorer
    from oracle.vina import VinaOracle
    rx, rt, lx, lt, offs = system
    v, o = VinaScorer(), VinaOracle()
    v.set_receptor(rx, rt)
    tors = (np.arange(len(offs) - 1) % 7).astype(np.float32)
    e, aff = v.score_exact(lx, lt, offs, tors)
    for p in range(len(offs) - 1):
        sl = slice(offs[p], offs[p + 1])
        er = o.naive_exact(rx, rt, lx[sl], lt[sl])
        assert abs(e[p] - er) <= TOL * max(1.0, abs(er))
        assert abs(aff[p] - o.num_tors_div(er, float(tors[p]))) <= TOL * max(1.0, abs(er))
    # empty batch and hydrogen-only pose
    assert v.score_exact(np.zeros((0, 3), np.float32), np.zeros(0, np.int32), [0])[0].shape == (0,)
    eh, _ = v.score_exact(lx[:2], np.array([1, 0], np.int32), [0, 2])
    assert eh[0] == 0.0


def test_spline_tables_and_spline_pair_terms_match_oracle():
    """V3: precalculate_splines coefficients equal the oracle's; eval_deriv with the spline pair terms matches too"""
    from gnina_b200 import synth
    from gnina_b200.vina import VinaScorer
    from oracle.vina import VinaOracle
    from oracle.vina_mc import DockOracle
    v, o = VinaScorer(), VinaOracle()
    for t1, t2 in ((2, 2), (2, 13), (7, 13), (12, 12), (4, 17)):
        a, b = v.spline_table(t1, t2), o.spline_table(t1, t2)
        assert a.shape == b.shape == (80, 4) and np.abs(a - b).max() <= 1e-6 * max(1.0, np.abs(b).max())
    rx, rt = synth.make_receptor(700, box=32)
    lig = synth.make_flexible_ligand()
    begin, end, n = [-9.0] * 3, [9.0] * 3, [48, 48, 48]
    needed = sorted(set(int(t) for t in lig["types"] if t > 1))
    v.set_receptor(rx, rt); v.cache_build(begin, end, n, needed); v.set_ligand(lig)
    d = DockOracle(o, {t: v.cache_grid(t) for t in needed}, begin, end, n, lig)
    X = np.stack([d.random_conf(50 + i, [-3, -3, -3], [3, 3, 3])[0] for i in range(16)])
    e_lin, _ = v.eval_deriv(X)
    v.set_precalc(True); d.use_splines(True)
    e, g = v.eval_deriv(X)
    assert np.abs(e - e_lin).max() > 0           # the spline tables are really in use
    for i, x in enumerate(X):
        er, gr = d.eval_deriv(x)
        assert abs(e[i] - er) <= 2e-5 * max(1.0, abs(er)) and np.abs(g[i] - gr).max() <= 2e-4 * max(1.0, np.abs(gr).max())
