"""The docking-side integration adapters (integration/docking_b200.h) EXECUTED on the CPU inside the reference's own classes
(oracle/ref_adapters_driver.cpp): the C-ABI calls they make are served by a stand-in that honours include/gnina_b200.h's contract with
the oracle's C restatement -- the same restatement the device kernels are checked against on the GPU -- so what runs here is the
adapter code a gnina maintainer would add, in place of the reference's `cache` / final-scoring code, inside the reference's
model::eval_deriv, quasi_newton and monte_carlo.  (The CNN-side adapter, CNNB200Scorer : DLScorer, is executed the same way in
tests/test_oracle_cnn_vs_reference_build.py.)  Skipped where /root/reference is absent."""
import numpy as np
import pytest
from gnina_b200 import synth
from oracle import cnn_refbuild as CR
from oracle import vina_refbuild as R
from oracle.vina import lib as vlib

pytestmark = pytest.mark.skipif(not (CR.available() or CR.build()), reason="oracle/_ref is not built and /root/reference is absent")

BEGIN, END, N = [-9.7] * 3, [10.55] * 3, [54, 54, 54]
MAX_FL = float(np.finfo(np.float32).max)


@pytest.fixture(scope="module")
def libm():
    vlib().gvo_use_libm(1)          # the reference build executes this host's expf: the stand-in's tables then equal the reference's bit for bit
    yield
    vlib().gvo_use_libm(0)


def _confs(rs, lig, k, spread):
    X = np.tile(lig["conf0"], (k, 1)).astype(np.float32)
    X[:, :3] += rs.uniform(-spread, spread, (k, 3))
    q = rs.randn(k, 4); X[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    X[:, 7:] = rs.uniform(-np.pi, np.pi, (k, X.shape[1] - 7))
    return X.astype(np.float32)


def test_cache_b200_is_a_drop_in_igrid(libm):
    """b200::cache_b200 : igrid in place of the reference's `cache` (lib/cache.cpp) as the `ig` of model::eval_deriv, of quasi_newton and
    of a whole monte_carlo chain: energies, gradients, conformations and the chain's output container are EQUAL"""
    lig = dict(synth.make_flexible_ligand(n_heavy=20, n_tors=4, n_branch=2, seed=8))
    ty = lig["types"].copy(); ty[3] = 1; ty[11] = 0; lig["types"] = ty           # a polar and a non-polar hydrogen among the movable atoms
    rx, rt = synth.make_receptor(500, box=30, seed=5)
    sf, rm = R.RefScoring(), R.RefModel(lig, rx, rt)
    theirs = R.RefGrid.cache(sf, R.LINEAR, rm, BEGIN, END, N, 1e3)
    adapters = CR.VinaAdapters(rx, rt)
    mine = adapters.cache_b200(rm, BEGIN, END, N, 1e3, sorted(set(int(t) for t in lig["types"] if t > 1)))
    for x in _confs(np.random.RandomState(2), lig, 8, 5.0):              # some poses partly outside the grid
        for caps in ((1000, 1000, 1000), (10, 1.5, 10)):
            a, b = R.model_eval_deriv(rm, sf, R.LINEAR, theirs, x, caps), R.model_eval_deriv(rm, sf, R.LINEAR, mine, x, caps)
            assert a[0] == b[0] and np.array_equal(a[1], b[1])
        a, b = R.bfgs(rm, sf, R.LINEAR, theirs, x, 12), R.bfgs(rm, sf, R.LINEAR, mine, x, 12)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for seed in (17, 4242):
        a = R.mc(rm, sf, R.LINEAR, theirs, seed, [-4] * 3, [4] * 3, 40, 12, lig["conf0"])
        b = R.mc(rm, sf, R.LINEAR, mine, seed, [-4] * 3, [4] * 3, 40, 12, lig["conf0"])
        assert len(a[0]) == len(b[0]) > 0 and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_score_docked_b200_is_the_docking_branch_affinity(libm):
    """b200::score_docked_b200 for all kept poses at once vs main/main.cpp:340-344 pose by pose on the reference's parts: non_cache::eval
    with the search's tables on the pose the model holds, then num_tors_div; a pose refine_structure could not pull inside keeps max_fl"""
    lig = synth.make_flexible_ligand(n_heavy=18, n_tors=3, n_branch=2, seed=21)
    rx, rt = synth.make_receptor(500, box=30, seed=9)
    sf, rm = R.RefScoring(), R.RefModel(lig, rx, rt)
    nc = R.RefGrid.non_cache(sf, R.LINEAR, rm, BEGIN, END, N, 1e3)
    X = _confs(np.random.RandomState(4), lig, 6, 6.0)
    coords = np.stack([rm.set(x) for x in X])
    e_in = np.float32([-5.0, -4.0, MAX_FL, -3.0, -2.0, MAX_FL])
    out = CR.VinaAdapters(rx, rt).score_docked(rm, coords, BEGIN, END, N, 1e3, (1000, 1000, 1000), 3.5, e_in)
    for i, x in enumerate(X):
        rm.set(x)
        want = MAX_FL if e_in[i] == MAX_FL else sf.num_tors_div(nc.eval(1000.0), 3.5)
        assert out[i] == np.float32(want), i


@pytest.mark.parametrize("accurate", [False, True])
def test_refine_structure_b200_is_refine_structure(libm, accurate):
    """b200::refine_structure_b200 for ALL kept poses in one call (topology through b200::B200Ligand) vs refine_structure
    (main/main.cpp:131-171) replayed pose by pose with the reference's own quasi_newton / non_cache / within: slope 10, 100, ... until the
    pose is inside the box; conformations and energies are EQUAL, a pose that never gets inside ends with max_fl"""
    lig = synth.make_flexible_ligand(n_heavy=22, n_tors=4, n_branch=3, seed=41)
    rx, rt = synth.make_receptor(500, box=30, seed=41)
    sf, rm = R.RefScoring(), R.RefModel(lig, rx, rt)
    box_b, box_e, box_n = [-5.3] * 3, [5.2] * 3, [28, 28, 28]
    X = _confs(np.random.RandomState(41), lig, 8, 7.0)
    maxit = (25 + len(lig["types"])) // 3
    nc = R.RefGrid.non_cache(sf, R.LINEAR, rm, box_b, box_e, box_n, 1e3)
    want_e, want_x, escalated = [], [], 0
    for x in X:
        xr, slope, er, ok = x.copy(), 10.0, 0.0, False
        for p in range(5):
            nc.set_slope(slope)
            er, xr, _ = R.bfgs(rm, sf, R.LINEAR, nc, xr, maxit, accurate=accurate)
            rm.set(xr)
            ok = nc.within()
            if ok:
                break
            slope *= 10
            escalated += 1
        want_e.append(er if ok else MAX_FL); want_x.append(xr)
    rm.set(X[0])
    e, x = CR.VinaAdapters(rx, rt).refine_structure(rm, X, box_b, box_e, box_n, (1000, 1000, 1000), maxit, accurate=accurate)
    assert escalated >= 1
    for i in range(len(X)):
        assert e[i] == np.float32(want_e[i]) and np.array_equal(x[i], want_x[i]), i


def test_parallel_mc_b200_is_parallel_mc(libm):
    """b200::parallel_mc_b200::operator() vs the REFERENCE's parallel_mc::operator() (lib/parallel_mc.cpp:183-214, own thread pool): task
    seeds from the caller's generator, Monte-Carlo parameters, every chain, conformation -> heavy-atom coordinates, the merge through the
    library's own host-side gb_vina_merge_outputs, the output container.  The stand-in runs the restatement's chain and draws each chain's
    START the reference's way (the device library draws its own starts, which is why the two are compared as search results on the GPU):
    with that, the merged containers are EQUAL -- energies and conformations."""
    lig = synth.make_flexible_ligand()
    rx, rt = synth.make_receptor(500, box=30, seed=5)
    sf, rm = R.RefScoring(), R.RefModel(lig, rx, rt)
    theirs = R.RefGrid.cache(sf, R.LINEAR, rm, BEGIN, END, N, 1e3)
    adapters = CR.VinaAdapters(rx, rt)
    mine = adapters.cache_b200(rm, BEGIN, END, N, 1e3, sorted(set(int(t) for t in lig["types"] if t > 1)))
    c1, c2 = [-5] * 3, [5] * 3
    maxit, S = (25 + len(lig["types"])) // 3, 20
    for seed, tasks, steps in ((4242, 4, 40), (777, 6, 60)):
        er, xr = R.parallel_mc(rm, sf, R.LINEAR, theirs, seed, c1, c2, tasks, steps, maxit, lig["conf0"], (BEGIN, END, N), num_threads=3,
                               num_saved_mins=S)
        rm.set(lig["conf0"])
        e, x = adapters.parallel_mc(rm, seed, c1, c2, tasks, steps, maxit, lig["conf0"], num_saved_mins=S)
        assert len(e) == len(er) > 0 and np.array_equal(e, er) and np.array_equal(x, xr)
    del mine
