"""The C restatement (oracle/vina_ref.c, vina_mc_ref.c) against oracle/_ref LIVE: the reference's own Vina sources compiled where
they lie under /root/reference (oracle/Makefile.ref; built by __graft_entry__.build()).  Fresh random inputs on every row, other
ligand topologies than the committed fixture (tests/golden/vina_ref_kat.npz, checked by test_oracle_vina_golden.py everywhere).
Skipped where neither the library nor the reference exists."""
import numpy as np
import pytest
from gnina_b200 import synth
from oracle import vina_refbuild as R
from oracle.vina import VinaOracle, lib as vlib
from oracle.vina_mc import DockOracle

pytestmark = pytest.mark.skipif(not (R.available() or R.build()), reason="oracle/_ref is not built and /root/reference is absent")

BEGIN, END, N = [-9.7] * 3, [10.55] * 3, [54, 54, 54]     # 0.375 A spacing (main/main.cpp:622), not aligned to 3 A


@pytest.fixture(scope="module")
def libm():
    vlib().gvo_use_libm(1)
    yield
    vlib().gvo_use_libm(0)


def _setup(lig, n_rec=500, seed=5):
    rx, rt = synth.make_receptor(n_rec, box=30, seed=seed)
    sf, vo = R.RefScoring(), VinaOracle()
    rm = R.RefModel(lig, rx, rt)
    lo, ro, ra = rm.export()
    lig2 = dict(lig); lig2["local_xyz"], lig2["seg_rel_origin"], lig2["seg_rel_axis"] = lo, ro, ra
    cg = R.RefGrid.cache(sf, R.LINEAR, rm, BEGIN, END, N, 1e3)
    needed = sorted(set(int(t) for t in lig["types"] if t > 1))
    grids = {t: vo.cache_populate(BEGIN, END, N, rx, rt, t) for t in needed}
    for t in needed:
        assert np.array_equal(grids[t], cg.grid(t)), "cache::populate, type %d" % t
    return sf, vo, rm, cg, DockOracle(vo, grids, BEGIN, END, N, lig2, slope=1e3), lig2, rx, rt


def _confs(rs, lig, T, k, spread=4.0):
    X = np.tile(lig["conf0"], (k, 1)).astype(np.float32)
    X[:, :3] += rs.uniform(-spread, spread, (k, 3))
    q = rs.randn(k, 4); X[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    X[:, 7:] = rs.uniform(-np.pi, np.pi, (k, T))
    return X.astype(np.float32)


@pytest.mark.parametrize("lig_kw", [dict(n_heavy=24, n_tors=5, n_branch=3, seed=11), dict(n_heavy=31, n_tors=8, n_branch=4, seed=3),
                                     dict(n_heavy=14, n_tors=2, n_branch=2, seed=29)])
def test_evaluation_minimisation_and_chains_are_bit_identical(lig_kw, libm):
    lig = synth.make_flexible_ligand(**lig_kw)
    sf, vo, rm, cg, d, lig2, rx, rt = _setup(lig, seed=lig_kw["seed"])
    rs = np.random.RandomState(lig_kw["seed"])
    X = _confs(rs, lig, d.T, 24)
    for x in X:                                          # V5-V8
        assert np.array_equal(d.coords(x), rm.set(x))
        for caps in ((1000, 1000, 1000), (10, 1.5, 10)):
            e, g = d.eval_deriv(x, caps); er, gr = R.model_eval_deriv(rm, sf, R.LINEAR, cg, x, caps)
            assert e == er and np.array_equal(g, gr)
    rm.set(X[0])
    assert d.gyration_radius(X[0]) == rm.gyration_radius()   # model::gyration_radius of the conformation the model holds
    for x in X[:12]:                                     # V9
        for it, caps in ((2, (10, 10, 10)), (15, (1000, 1000, 1000))):
            e, xo, g, _ = d.bfgs(x, it, caps); er, xr, gr = R.bfgs(rm, sf, R.LINEAR, cg, x, it, caps)
            assert e == er and np.array_equal(xo, xr) and np.array_equal(g, gr)
    for x in X[:8]:                                      # V9, --minimize: accurate line search, early termination
        for acc, et, it in ((True, False, 40), (True, True, 400), (False, True, 40)):
            e, xo, g, _ = d.bfgs(x, it, accurate=acc, early_term=et)
            er, xr, gr = R.bfgs(rm, sf, R.LINEAR, cg, x, it, accurate=acc, early_term=et)
            assert e == er and np.array_equal(xo, xr) and np.array_equal(g, gr)
    maxit = (25 + len(lig["types"])) // 3                # V10: main/main.cpp:454
    for seed in (17, 4242):
        x0, st = R.random_conf(rm, seed, [-4] * 3, [4] * 3)
        er, xr = R.mc(rm, sf, R.LINEAR, cg, seed, [-4] * 3, [4] * 3, 40, maxit, lig["conf0"])
        e, x = d.mc_ex(st, [-4] * 3, [4] * 3, 40, maxit, init_conf=x0, state_conf=lig["conf0"])
        assert len(e) == len(er) and np.array_equal(e, er) and np.array_equal(x, xr)


@pytest.mark.parametrize("seed", [0, 2, 4, 5])
def test_random_torsion_trees(seed, libm):
    """nodes with several children and nested branches (synth.make_tree_ligand): kinematics, energy + change, quasi_newton and one whole
    Monte-Carlo chain of the restatement against the reference; the model -> topology adapter and both host-side minimisers' kinematics
    on the same trees"""
    from gnina_b200 import minimize
    lig = synth.make_tree_ligand(seed)
    sf, vo, rm, cg, d, lig2, rx, rt = _setup(lig, seed=seed + 50)
    t = rm.adapter_topology()
    for k in ("seg_parent", "seg_begin", "seg_end", "pair_a", "pair_b"):
        assert np.array_equal(t[k], np.asarray(lig[k], np.int32)), k
    rs = np.random.RandomState(seed)
    X = _confs(rs, lig, d.T, 12)
    minimize.set_transcendentals(*_host_libm())
    try:
        tree = minimize.TorsionTree(lig2)
        coords, so, sa = tree.set_conf(X)
        forces = rs.uniform(-10, 10, coords.shape).astype(np.float32)
        change = tree.derivative(coords, forces, so, sa)
        for i, x in enumerate(X):
            c = rm.set(x)
            assert np.array_equal(d.coords(x), c) and np.array_equal(coords[i], c)
            assert np.array_equal(change[i], rm.tree_derivative(forces[i]))
            e, g = d.eval_deriv(x, (10, 1.5, 10)); er, gr = R.model_eval_deriv(rm, sf, R.LINEAR, cg, x, (10, 1.5, 10))
            assert e == er and np.array_equal(g, gr)
    finally:
        minimize.set_transcendentals()
    for x in X[:6]:
        for acc in (False, True):
            e, xo, g, _ = d.bfgs(x, 20, accurate=acc); er, xr, gr = R.bfgs(rm, sf, R.LINEAR, cg, x, 20, accurate=acc)
            assert e == er and np.array_equal(xo, xr) and np.array_equal(g, gr)
    maxit = (25 + len(lig["types"])) // 3
    x0, st = R.random_conf(rm, 31337, [-4] * 3, [4] * 3)
    er, xr = R.mc(rm, sf, R.LINEAR, cg, 31337, [-4] * 3, [4] * 3, 40, maxit, lig["conf0"])
    e, x = d.mc_ex(st, [-4] * 3, [4] * 3, 40, maxit, init_conf=x0, state_conf=lig["conf0"])
    assert len(e) == len(er) and np.array_equal(e, er) and np.array_equal(x, xr)


def test_ligand_with_hydrogens(libm):
    """polar and non-polar hydrogens among the movable atoms: skipped by the grid term, zero force, absent from the gyration radius and
    from the containers' RMSD (lib/cache.cpp:65-83, lib/model.cpp:1002-1014, get_heavy_atom_movable_coords) -- evaluation and whole chains"""
    lig = dict(synth.make_flexible_ligand())
    ty = np.array(lig["types"]).copy(); ty[5] = 1; ty[17] = 1; ty[26] = 0; lig["types"] = ty
    sf, vo, rm, cg, d, lig2, rx, rt = _setup(lig, seed=9)
    for x in _confs(np.random.RandomState(9), lig, d.T, 16):
        e, g = d.eval_deriv(x, (10, 1.5, 10)); er, gr = R.model_eval_deriv(rm, sf, R.LINEAR, cg, x, (10, 1.5, 10))
        assert e == er and np.array_equal(g, gr)
    maxit = (25 + len(lig["types"])) // 3
    for seed in (5, 901):
        x0, st = R.random_conf(rm, seed, [-4] * 3, [4] * 3)
        er, xr = R.mc(rm, sf, R.LINEAR, cg, seed, [-4] * 3, [4] * 3, 50, maxit, lig["conf0"])
        e, x = d.mc_ex(st, [-4] * 3, [4] * 3, 50, maxit, init_conf=x0, state_conf=lig["conf0"])
        assert len(e) == len(er) and np.array_equal(e, er) and np.array_equal(x, xr)


def test_add_minus_forces_consumes_its_list_compactly():
    """what the DLScorer adapter inherits: model::add_minus_forces (lib/model.cu:247-259) gives list entry j to the j-th
    NON-hydrogen movable atom, although CNNTorchScorer::getGradient indexes the list by movable-atom index"""
    import ctypes as C
    lig = dict(synth.make_flexible_ligand())
    ty = np.array(lig["types"]).copy(); ty[2] = 1; lig["types"] = ty
    rm = R.RefModel(lig)
    fp = C.POINTER(C.c_float)
    R.lib().gref_add_minus_forces.argtypes = [C.c_void_p, fp, C.c_int, fp]
    f = np.zeros((rm.na, 3), np.float32); f[:, 0] = np.arange(rm.na)
    out = np.zeros((rm.na, 3), np.float32)
    R.lib().gref_add_minus_forces(rm.p, f.ctypes.data_as(fp), rm.na, out.ctypes.data_as(fp))
    assert out[:6, 0].tolist() == [0.0, 1.0, 0.0, 2.0, 3.0, 4.0]


def test_stateless_chain_variant_is_not_the_reference(libm):
    """what round 1's kernels did -- constant gyration radius, energies re-evaluated at the returned conformation -- leaves the
    reference's trajectory within a few steps: the model-state rules are part of the algorithm, not noise"""
    lig = synth.make_flexible_ligand()
    sf, vo, rm, cg, d, lig2, rx, rt = _setup(lig)
    x0, st = R.random_conf(rm, 99, [-4] * 3, [4] * 3)
    er, xr = R.mc(rm, sf, R.LINEAR, cg, 99, [-4] * 3, [4] * 3, 80, 17, lig["conf0"])
    e, x = d.mc_ex(st, [-4] * 3, [4] * 3, 80, 17, init_conf=x0, state_conf=None)
    assert not (len(e) == len(er) and np.array_equal(e, er))


def test_non_cache_and_exact_scoring(libm):
    lig = synth.make_flexible_ligand(n_heavy=20, n_tors=4, n_branch=2, seed=8)
    sf, vo, rm, cg, d, lig2, rx, rt = _setup(lig, seed=8)
    rs = np.random.RandomState(8)
    X = _confs(rs, lig, d.T, 24, spread=8.0)
    nn = R.RefGrid.naive(sf, R.EXACT, rm)
    try:
        for slope in (10.0, 100.0, 1000.0):              # refine_structure's slopes (main/main.cpp:145-154)
            nc = R.RefGrid.non_cache(sf, R.LINEAR, rm, BEGIN, END, N, slope)
            d.use_noncache(rx, rt); d.set_box(BEGIN, END, slope)
            for x in X:
                e, g = d.eval_deriv(x); er, gr = R.model_eval_deriv(rm, sf, R.LINEAR, nc, x)
                assert e == er and np.array_equal(g, gr)
                rm.set(x)
                assert nc.within() == d.within(x)
    finally:
        d.use_noncache(None); d.set_box(None)
    nc = R.RefGrid.non_cache(sf, R.LINEAR, rm, BEGIN, END, N, 1e3)
    for x in X[:8]:
        c = rm.set(x)
        assert nn.eval(1000.0) == vo.naive_exact(rx, rt, c, lig["types"], 1000.0)                       # --score_only's energy
        assert nc.eval(1000.0) == vo.noncache_eval(rx, rt, c, lig["types"], BEGIN, END, 1e3, 1000.0)   # the docking branch's
    for e, nt in ((-7.25, 0.0), (-7.25, 3.5), (-11.0, 10.5), (2.0, 7.0)):
        assert sf.num_tors_div(e, nt) == vo.num_tors_div(e, nt)


def test_custom_weights_table_factor_and_the_minimize_flavour(libm):
    """other term weights (--custom_scoring style) and another table factor: grids and model::eval_deriv stay bit-identical.  The
    --minimize flavour -- precalculate_splines (factor 10) for the receptor terms of non_cache AND the intramolecular pairs
    (main/main.cpp:1162-1165,1387) -- agrees to the round-off of the spline coefficients (Thomas solve in double vs Eigen's float inverse)"""
    lig = synth.make_flexible_ligand()
    rx, rt = synth.make_receptor(500, box=30)
    w6 = np.float32([-0.03, -0.006, 0.9, -0.04, -0.5, 1.5])
    sf, vo = R.RefScoring(factor_linear=48.0, weights6=w6), VinaOracle(weights6=w6, factor=48.0)
    rm = R.RefModel(lig, rx, rt)
    lo, ro, ra = rm.export()
    lig2 = dict(lig); lig2["local_xyz"], lig2["seg_rel_origin"], lig2["seg_rel_axis"] = lo, ro, ra
    needed = sorted(set(int(t) for t in lig["types"] if t > 1))
    cg = R.RefGrid.cache(sf, R.LINEAR, rm, BEGIN, END, N, 1e3)
    grids = {t: vo.cache_populate(BEGIN, END, N, rx, rt, t) for t in needed}
    assert all(np.array_equal(grids[t], cg.grid(t)) for t in needed)
    d = DockOracle(vo, grids, BEGIN, END, N, lig2, slope=1e3)
    rs = np.random.RandomState(5)
    X = _confs(rs, lig, d.T, 16)
    for x in X:
        e, g = d.eval_deriv(x); er, gr = R.model_eval_deriv(rm, sf, R.LINEAR, cg, x)
        assert e == er and np.array_equal(g, gr)
    assert sf.num_tors_div(-7.5, 3.5) == vo.num_tors_div(-7.5, 3.5)
    sf2, vo2 = R.RefScoring(), VinaOracle()
    nc = R.RefGrid.non_cache(sf2, R.SPLINES, rm, BEGIN, END, N, 1e3)
    d2 = DockOracle(vo2, {}, BEGIN, END, N, lig2, slope=1e3)
    d2.use_noncache(rx, rt); d2.set_box(BEGIN, END, 1e3); d2.use_splines(True)
    for x in X:
        e, g = d2.eval_deriv(x); er, gr = R.model_eval_deriv(rm, sf2, R.SPLINES, nc, x)
        assert abs(e - er) <= 1e-6 * max(1.0, abs(er)) and np.abs(g - gr).max() <= 2e-6 * max(1.0, np.abs(gr).max())


def test_autobox_is_movable_atoms_box():
    """docking.autobox vs model::movable_atoms_box (lib/model.cpp:751-776): --autobox_ligand / the per-ligand box of --minimize"""
    from gnina_b200 import docking
    for seed in (0, 3, 5):
        lig = synth.make_tree_ligand(seed)
        rm = R.RefModel(lig)
        x = _confs(np.random.RandomState(seed), lig, rm.T, 1, spread=20.0)[0]
        c = rm.set(x)
        for add in (4.0, 0.0, 7.3):
            b, e, n = docking.autobox(c, add)
            br, er, nr = rm.movable_atoms_box(add)
            assert np.array_equal(n, nr) and np.array_equal(b, br) and np.array_equal(e, er)


def test_refine_structure_composed_from_reference_parts(libm):
    """refine_structure lives in main/main.cpp:131-171 (not a library source): its loop -- slope 10, 100, ...; quasi_newton on the
    non_cache field; m.set; stop when non_cache::within -- is replayed here with the REFERENCE's quasi_newton / non_cache / within and
    compared with the restatement's gvo_refine_structure"""
    lig = synth.make_flexible_ligand(n_heavy=22, n_tors=4, n_branch=3, seed=41)
    sf, vo, rm, cg, d, lig2, rx, rt = _setup(lig, seed=41)
    rs = np.random.RandomState(41)
    box_b, box_e, box_n = [-5.3] * 3, [5.2] * 3, [28, 28, 28]
    X = _confs(rs, lig, d.T, 10, spread=7.0)
    nc = R.RefGrid.non_cache(sf, R.LINEAR, rm, box_b, box_e, box_n, 1e3)
    maxit = (25 + len(lig["types"])) // 3
    left_box = 0
    try:
        d.use_noncache(rx, rt); d.set_box(box_b, box_e, 1e3)
        for x in X:
            xr, slope, er, ok, passes = x.copy(), 10.0, 0.0, False, 0
            for p in range(5):
                nc.set_slope(slope)
                er, xr, _ = R.bfgs(rm, sf, R.LINEAR, nc, xr, maxit)
                rm.set(xr); passes += 1
                ok = nc.within()
                if ok:
                    break
                slope *= 10
            e, xo, ne, ok_o = d.refine_structure(x, maxit)
            assert ok_o == ok and e == er and np.array_equal(xo, xr)
            left_box += passes > 1
    finally:
        d.use_noncache(None); d.set_box(None)
    assert left_box >= 1          # the escalation of the slope was exercised


def test_parallel_mc_fan_out_and_merge(libm):
    """V11: the REFERENCE's parallel_mc::operator() (lib/parallel_mc.cpp:183-214) on its own thread pool (lib/parallel.h) -- task seeds
    random_int(0, 1000000, generator), one chain per task, merge_output_containers with min_rmsd forced to 2, final sort -- against the
    restatement's chains merged by the host-side rule (Python statement and the library's gb_vina_merge_outputs)"""
    from gnina_b200 import docking
    lig = synth.make_flexible_ligand()
    sf, vo, rm, cg, d, lig2, rx, rt = _setup(lig)
    c1, c2 = [-5] * 3, [5] * 3
    maxit, S = (25 + len(lig["types"])) // 3, 20
    heavy = np.flatnonzero(np.asarray(lig["types"]) > 1)
    for seed, tasks, steps in ((4242, 4, 40), (777, 6, 60)):
        er, xr = R.parallel_mc(rm, sf, R.LINEAR, cg, seed, c1, c2, tasks, steps, maxit, lig["conf0"], (BEGIN, END, N), num_threads=3,
                               num_saved_mins=S)
        oc = docking.OutputContainer(2.0, S)
        E, X, C, n_out = np.zeros((tasks, S), np.float32), np.zeros((tasks, S, 7 + d.T), np.float32), np.zeros((tasks, S, len(heavy), 3), np.float32), []
        for c, ts in enumerate(R.task_seeds(seed, tasks)):
            x0, st = R.random_conf(rm, ts, c1, c2)
            e, x = d.mc_ex(st, c1, c2, steps, maxit, num_saved_mins=S, init_conf=x0, state_conf=lig["conf0"])
            n_out.append(len(e))
            for k in range(len(e)):
                hv = d.coords(x[k])[heavy]
                oc.add(e[k], hv, x[k])
                E[c, k], X[c, k], C[c, k] = e[k], x[k], hv
        em, xm = np.float32([o["e"] for o in oc.items]), np.stack([o["conf"] for o in oc.items])
        assert len(er) == len(em) and np.array_equal(er, em) and np.array_equal(xr, xm)
        nat = docking.merge_chains_native(E, X, C, np.int32(n_out), S)
        assert np.array_equal(np.float32([o["e"] for o in nat]), er) and np.array_equal(np.stack([o["conf"] for o in nat]), xr)


def test_non_cache_cnn_host_logic_matches_the_reference():
    """S3: the product's host-side gb::NonCacheCNNT (include/gnina_b200.hpp) against the REFERENCE's non_cache_cnn::eval / eval_deriv
    (lib/non_cache_cnn.cpp:33-54,79-169) around the same analytic stand-in for the network: out-of-box penalties of the search box
    and of the CNN box (adjust_center / set_bounding_box), hydrogens, and the empirical mixing of --cnn_mix_emp_force /
    --cnn_mix_emp_energy / --cnn_empirical_weight.  Same sums in another order: 1e-6."""
    lig = dict(synth.make_flexible_ligand())
    ty = np.array(lig["types"]).copy(); ty[5] = 1; ty[17] = 1; lig["types"] = ty        # two polar hydrogens
    rx, rt = synth.make_receptor(600, box=30)
    sf = R.RefScoring(); rm = R.RefModel(lig, rx, rt)
    rs = np.random.RandomState(0)
    begin, end, n = [-6.1] * 3, [5.9] * 3, [32, 32, 32]
    outside = 0
    for x in _confs(rs, lig, rm.T, 30, spread=9.0):
        c = rm.set(x)
        outside += bool(((c < -6.1) | (c > 5.9)).any())
        for mf, me in ((False, False), (True, False), (True, True), (False, True)):
            for deriv in (True, False):
                er, fr, em, fm = R.noncache_cnn_compare(rm, sf, R.LINEAR, begin, end, n, slope=10.0, dim=12.0, res=0.5, k=0.02,
                                                        target=(0.5, -0.3, 0.2), mix_force=mf, mix_energy=me, weight=0.7, deriv=deriv)
                assert abs(er - em) <= 1e-6 * max(1.0, abs(er)), (mf, me, deriv, er, em)
                assert np.abs(fr - fm).max() <= 1e-6 * max(1.0, np.abs(fr).max())
                if deriv:
                    assert (fm[[5, 17]] == 0).all() and (fr[[5, 17]] == 0).all()       # hydrogens carry no force
    assert outside >= 5


def _routed(grad, heavy):
    """the analytic network's by-atom gradient (zero for hydrogens) as the reference's scorer leaves it in the model: consumed compactly
    over the heavy atoms by model::add_minus_forces (minimize.route_forces_like_the_reference)"""
    from gnina_b200 import minimize
    g = np.asarray(grad, np.float32).copy()
    g[:, ~heavy] = 0
    return minimize.route_forces_like_the_reference(g, heavy)


def _host_libm():
    import ctypes
    m = ctypes.CDLL("libm.so.6")

    def wrap(name):
        f = getattr(m, name); f.argtypes = [ctypes.c_float]; f.restype = ctypes.c_float
        return lambda a: np.array([f(float(v)) for v in np.asarray(a, np.float32).ravel()], np.float32).reshape(np.shape(a))
    return wrap("sinf"), wrap("cosf"), wrap("acosf")


def test_lock_step_minimiser_reproduces_the_reference_pose_by_pose():
    """gnina_b200/minimize.py (BASELINE config 5: --minimize --cnn_scoring all over many poses): ALL poses advance together, one batched
    energy call per round, and every pose ends exactly where the REFERENCE's quasi_newton + non_cache_cnn takes it when it minimises that
    pose alone (lib/quasi_newton.cpp:49-83, bfgs.h with the accurate and the fast line search, --minimize_early_term, both out-of-box
    penalties, torsion-tree kinematics) -- energies and conformations bit for bit, around the same analytic stand-in for the network"""
    from gnina_b200 import minimize
    lig = dict(synth.make_flexible_ligand())
    ty = np.array(lig["types"]).copy(); ty[5] = 1; lig["types"] = ty
    rx, rt = synth.make_receptor(300, box=30)
    sf = R.RefScoring(); rm = R.RefModel(lig, rx, rt)
    lo, ro, ra = rm.export()
    lig2 = dict(lig); lig2["local_xyz"], lig2["seg_rel_origin"], lig2["seg_rel_axis"] = lo, ro, ra
    tree = minimize.TorsionTree(lig2)
    heavy = np.asarray(lig["types"]) >= 2
    k, target = np.float32(0.02), np.float32([0.5, -0.3, 0.2])
    begin, end, nn, slope, dim = [-8.0] * 3, [8.0] * 3, [43, 43, 43], 10.0, 20.0
    X = _confs(np.random.RandomState(3), lig, tree.T, 16, spread=6.0)
    minimize.set_transcendentals(*_host_libm())              # the reference build executes this host's sinf / cosf / acosf
    try:
        centers = minimize.heavy_centers(tree.set_conf(X)[0], heavy)
        half = np.float32(dim) / np.float32(2)
        calls = [0]

        def energy(coords, idx):
            calls[0] += 1
            d = coords - target
            loss = np.zeros(len(coords), np.float32)
            for a in np.flatnonzero(heavy):
                for j in range(3):
                    loss = (loss + k * d[:, a, j] * d[:, a, j]).astype(np.float32)
            c = centers[idx][:, None, :]
            return minimize.with_box_penalties(loss, _routed(2 * k * d, heavy), coords, heavy, (begin, end), (c - half, c + half), slope)
        for acc, et, iters in ((True, False, 10000), (False, True, 200)):
            calls[0] = 0
            e, x, ev, rounds = minimize.minimize_poses(tree, energy, X, maxiters=iters, accurate=acc, early_term=et)
            for i in range(len(X)):
                er, xr = R.minimize_cnn(rm, sf, R.LINEAR, begin, end, nn, X[i], iters, slope=slope, dim=dim, res=0.5, k=float(k),
                                        target=target, accurate=acc, early_term=et)
                assert er == e[i] and np.array_equal(xr, x[i]), (acc, et, i)
            assert calls[0] == rounds + 1 and calls[0] < 0.5 * ev.sum()      # batched: far fewer energy calls than evaluations
    finally:
        minimize.set_transcendentals()


def test_cpp_lock_step_minimiser_reproduces_the_reference_pose_by_pose():
    """the C++ host side of the same driver (include/gnina_b200_minimize.hpp: gb::minimize_poses over gb::LigandTree, energy =
    gb::NonCacheCNNT around the analytic stand-in, topology from the integration adapter b200::B200Ligand) on 48 poses at once: every
    pose ends bit for bit where the reference's quasi_newton + non_cache_cnn takes it alone, in all four line-search / termination modes"""
    lig = dict(synth.make_flexible_ligand())
    ty = np.array(lig["types"]).copy(); ty[5] = 1; lig["types"] = ty
    rx, rt = synth.make_receptor(300, box=30)
    sf = R.RefScoring(); rm = R.RefModel(lig, rx, rt)
    begin, end, nn, slope, dim, k, target = [-8.0] * 3, [8.0] * 3, [43, 43, 43], 10.0, 20.0, 0.02, np.float32([0.5, -0.3, 0.2])
    X = _confs(np.random.RandomState(3), lig, rm.T, 48, spread=6.0)
    for acc, et, iters in ((True, False, 10000), (True, True, 10000), (False, False, 40), (False, True, 200)):
        e, x, ev, rounds, calls = R.lockstep_minimize_cpp(rm, begin, end, nn, X, iters, slope=slope, dim=dim, res=0.5, k=k, target=target,
                                                          accurate=acc, early_term=et)
        for i in range(len(X)):
            er, xr = R.minimize_cnn(rm, sf, R.LINEAR, begin, end, nn, X[i], iters, slope=slope, dim=dim, res=0.5, k=k, target=target,
                                    accurate=acc, early_term=et)
            assert er == e[i] and np.array_equal(xr, x[i]), (acc, et, i)
        assert calls == rounds + 1 and calls < 0.2 * ev.sum()


def test_lock_step_cnn_refinement_reproduces_refine_structure():
    """--cnn_scoring refinement: refine_structure (main/main.cpp:131-171) on ig = non_cache_cnn for many poses at once
    (minimize.refine_structure_poses) vs the reference's parts replayed pose by pose: the slope escalates 10, 100, ... while a spring
    pulls the ligand out of both boxes; energies (max_fl for a pose that stays outside), conformations and the within flag coincide"""
    from gnina_b200 import minimize
    lig = dict(synth.make_flexible_ligand())
    ty = np.array(lig["types"]).copy(); ty[5] = 1; lig["types"] = ty
    rx, rt = synth.make_receptor(300, box=30)
    sf = R.RefScoring(); rm = R.RefModel(lig, rx, rt)
    lo, ro, ra = rm.export()
    lig2 = dict(lig); lig2["local_xyz"], lig2["seg_rel_origin"], lig2["seg_rel_axis"] = lo, ro, ra
    tree = minimize.TorsionTree(lig2)
    heavy = np.asarray(lig["types"]) >= 2
    k, target = np.float32(0.2), np.float32([14.0, 0.0, 0.0])
    begin, end, nn, dim = [-7.0] * 3, [7.0] * 3, [38, 38, 38], 12.0
    X = _confs(np.random.RandomState(7), lig, tree.T, 10, spread=3.0)
    minimize.set_transcendentals(*_host_libm())
    try:
        centers = minimize.heavy_centers(tree.set_conf(X)[0], heavy)
        half = np.float32(dim) / np.float32(2)

        def make_energy(slope):
            def energy(coords, idx):
                d = coords - target
                loss = np.zeros(len(coords), np.float32)
                for a in np.flatnonzero(heavy):
                    for j in range(3):
                        loss = (loss + k * d[:, a, j] * d[:, a, j]).astype(np.float32)
                c = centers[idx][:, None, :]
                return minimize.with_box_penalties(loss, _routed(2 * k * d, heavy), coords, heavy, (begin, end), (c - half, c + half), slope)
            return energy
        within = minimize.within_boxes(heavy, [(centers - half, centers + half), (np.float32(begin), np.float32(end))])
        e, x, inside, ev = minimize.refine_structure_poses(tree, make_energy, within, X, 30)
        for i in range(len(X)):
            er, xr, ins = R.refine_cnn(rm, sf, R.LINEAR, begin, end, nn, X[i], 30, dim=dim, res=0.5, k=float(k), target=target)
            assert er == e[i] and np.array_equal(xr, x[i]) and ins == bool(inside[i]), i
        assert ev.max() > 3 * ev.min()          # some poses needed several passes
    finally:
        minimize.set_transcendentals()


@pytest.mark.parametrize("seed", [3, 91])
def test_gninacheck_random_molecules_against_the_compiled_reference(seed, libm):
    """the reference's own `gninacheck` inputs (test/gnina/test_utils.cpp:13-44; synth.make_gninacheck_mol): atoms of EVERY smina type --
    hydrogens, metals, the generic types -- at uniform random positions, overlaps allowed, as receptor and as a one-torsion ligand with
    pairs across the torsion.  gninacheck itself compares the reference's CPU and GPU flavours to 0.01; here cache::populate for every
    type the ligand needs, model::eval_deriv, non_cache::eval and the exact terms of the restatement are bit-identical to the CPU one."""
    rs = np.random.RandomState(seed)
    rx, rt = synth.make_gninacheck_mol(rs, 0, 200, 500, 14, 14, 14)
    lx, lt = synth.make_gninacheck_mol(rs, 0, 20, 60, 5, 5, 5)
    na, h = len(lt), len(lt) // 2
    heavy = [i for i in range(na) if lt[i] > 1]
    pa = [a for a in heavy if a < h for b in heavy if b >= h][:200]
    pb = [b for a in heavy if a < h for b in heavy if b >= h][:200]
    lig = dict(xyz0=lx, types=lt, seg_parent=np.array([-1, 0], np.int32), seg_begin=np.array([0, h], np.int32),
               seg_end=np.array([h, na], np.int32), axis_root=np.array([0, h - 1], np.int32), pair_a=np.array(pa, np.int32),
               pair_b=np.array(pb, np.int32), conf0=np.array([*lx[0], 1, 0, 0, 0, 0], np.float32),
               gyration_radius=synth.gyration_radius(lx, lt, lx[0]))
    sf, vo = R.RefScoring(), VinaOracle()
    rm = R.RefModel(lig, rx, rt)
    lig["local_xyz"], lig["seg_rel_origin"], lig["seg_rel_axis"] = rm.export()
    cg = R.RefGrid.cache(sf, R.LINEAR, rm, BEGIN, END, N, 1e3)
    needed = sorted(set(int(t) for t in lt if t > 1))
    assert len(needed) >= 18                             # metals and generic types among them
    grids = {t: vo.cache_populate(BEGIN, END, N, rx, rt, t) for t in needed}
    for t in needed:
        assert np.array_equal(grids[t], cg.grid(t)), "cache::populate, type %d" % t
    d = DockOracle(vo, grids, BEGIN, END, N, lig, slope=1e3)
    X = _confs(rs, lig, 1, 16, spread=3.0)
    nn, nc = R.RefGrid.naive(sf, R.EXACT, rm), R.RefGrid.non_cache(sf, R.LINEAR, rm, BEGIN, END, N, 1e3)
    for x in X:
        c = rm.set(x)
        assert np.array_equal(d.coords(x), c)
        e, g = d.eval_deriv(x); er, gr = R.model_eval_deriv(rm, sf, R.LINEAR, cg, x)
        assert e == er and np.array_equal(g, gr)
        assert nn.eval(1000.0) == vo.naive_exact(rx, rt, c, lt, 1000.0)
        assert nc.eval(1000.0) == vo.noncache_eval(rx, rt, c, lt, BEGIN, END, 1e3, 1000.0)
    for x in X[:4]:
        e, xo, g, _ = d.bfgs(x, 10); er, xr, gr = R.bfgs(rm, sf, R.LINEAR, cg, x, 10)
        assert e == er and np.array_equal(xo, xr) and np.array_equal(g, gr)


def test_grid_aligned_to_three_angstrom_shows_the_reference_cell_list_quirk():
    """szv_grid_cache::get (lib/szv_grid.h:124-150) sizes a 3 A cell's atom list by the brick [floor(c/3)*3, ceil(c/3)*3] of the FIRST
    probe point that touches the cell: when that coordinate is an exact multiple of 3 the brick collapses and the list misses atoms
    that are within the cut-off of other points of the cell.  The restatement sums over all atoms within the cut-off (what the code
    means); on a grid that starts at a multiple of 3 A the reference's own grids are therefore HIGHER (attractive far atoms
    missing), everywhere else they are bit-identical (every other test here).  Documented, not imitated: DESIGN.md §2."""
    lig = synth.make_flexible_ligand()
    rx, rt = synth.make_receptor(500, box=30)
    sf, vo = R.RefScoring(), VinaOracle()
    rm = R.RefModel(lig, rx, rt)
    b, e, n = [-9.0] * 3, [9.0] * 3, [36, 36, 36]
    cg = R.RefGrid.cache(sf, R.LINEAR, rm, b, e, n, 1e3)
    t = int(lig["types"][0])
    ref, mine = cg.grid(t), vo.cache_populate(b, e, n, rx, rt, t)
    assert not np.array_equal(ref, mine) and np.abs(ref - mine).max() < 0.5
    far = np.abs(ref) < 0.5                               # away from clashes the missing terms are the attractive ones
    assert (ref[far] - mine[far]).mean() > 0
