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


@pytest.mark.parametrize("seed", [1, 20220616, 987654321])
def test_gninacheck_random_molecules(seed):
    """The reference's `gninacheck` tests (test/gnina/test_cache.cu, test_gpucode.cpp) hold no numbers: they draw random
    molecules (test_utils.cpp make_mol: every smina type incl. hydrogens and metals, overlapping atoms allowed), build the
    cache over the ligand's bounding box (granularity 0.375, slope 10, v = 10) and require two live implementations to
    agree to 0.01.  Same generator, same set-up, device vs oracle, at the north star's 1e-6 instead of 0.01:
      V4  cache::populate         on the box grid (sub-sampled in the oracle: the scalar C loop is slow)
      V5  cache::eval_deriv       energy and forces of the ligand, slope 10, v = 10
      V6  eval_interacting_pairs  all atom pairs of a random molecule as one rigid body (test_eval_intra)"""
    from gnina_b200 import synth
    from gnina_b200.vina import VinaScorer
    from oracle.vina import VinaOracle
    from oracle.vina_mc import DockOracle
    rs = np.random.RandomState(seed)                      # seed logged by pytest's parametrisation
    lig_xyz, lig_t = synth.make_gninacheck_mol(rs, min_atoms=20, max_atoms=60, max_x=5, max_y=5, max_z=5)
    lo, hi = lig_xyz.min(0), lig_xyz.max(0)
    center, span = (hi + lo) / 2, hi - lo
    n = np.ceil(span / 0.375).astype(np.int32)             # test_cache.cu:77-82
    begin = (center - 0.375 * n / 2).astype(np.float32)
    end = (begin + 0.375 * n).astype(np.float32)
    rec_xyz, rec_t = synth.make_gninacheck_mol(rs, 0, 10, 500, *(np.abs(np.stack([lo, hi])).max(0) + 8.0))
    v, o = VinaScorer(), VinaOracle()
    v.set_receptor(rec_xyz, rec_t)
    needed = sorted(set(int(t) for t in lig_t if t > 1))   # get_movable_atom_types: heavy types only
    v.cache_build(begin, end, n, needed)
    grids = {t: v.cache_grid(t) for t in needed}
    t0 = needed[len(needed) // 2]
    ref = o.cache_populate(begin, end, n, rec_xyz, rec_t, t0)
    assert np.abs(grids[t0] - ref).max() <= TOL * max(1.0, np.abs(ref).max())
    offs = np.array([0, len(lig_t)], np.int32)
    e, d = v.cache_eval(lig_xyz, lig_t, offs, slope=10.0, v=10.0)
    er, dr = VinaOracle.cache_eval(grids, begin, end, n, lig_xyz, lig_t, 10.0, 10.0)
    assert abs(e[0] - er) <= TOL * max(1.0, abs(er))
    assert np.abs(d - dr).max() <= TOL * max(1.0, np.abs(dr).max())
    # V6: one rigid segment, every pair (i < j), curl cap 10 on the pairs like single_point_calc's v
    na = len(lig_t)
    pa, pb = np.triu_indices(na, 1)
    lig = dict(local_xyz=lig_xyz - lig_xyz[0], types=lig_t, seg_parent=np.array([-1], np.int32), seg_begin=np.array([0], np.int32),
               seg_end=np.array([na], np.int32), seg_rel_origin=np.zeros((1, 3), np.float32), seg_rel_axis=np.zeros((1, 3), np.float32),
               pair_a=pa.astype(np.int32), pair_b=pb.astype(np.int32), gyration_radius=1.0)
    v.set_ligand(lig)
    dk = DockOracle(o, grids, begin, end, n, lig, slope=10.0)
    conf = np.zeros((1, 7), np.float32); conf[0, :3] = lig_xyz[0]; conf[0, 3] = 1.0
    caps = (10.0, 10.0, 10.0)
    e2, g2 = v.eval_deriv(conf, caps, slope=10.0)
    e2r, g2r = dk.eval_deriv(conf[0], caps)
    assert abs(e2[0] - e2r) <= TOL * max(1.0, abs(e2r))
    assert np.abs(g2[0] - g2r).max() <= 4e-6 * max(1.0, np.abs(g2r).max())
