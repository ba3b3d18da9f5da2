"""GPU parity of the docking inner loop (V6-V11) through the C ABI vs the CPU oracle (oracle/vina_mc_ref.c)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from gnina_b200 import synth
    from gnina_b200.vina import VinaScorer
    from oracle.vina import VinaOracle
    from oracle.vina_mc import DockOracle
    rx, rt = synth.make_receptor(900, box=34)
    lig = synth.make_flexible_ligand()
    begin, end, n = [-10.0] * 3, [10.0] * 3, [53, 53, 53]          # ~0.377 A spacing
    needed = sorted(set(int(t) for t in lig["types"] if t > 1))
    v = VinaScorer()
    v.set_receptor(rx, rt)
    v.cache_build(begin, end, n, needed)
    v.set_ligand(lig)
    grids = {t: v.cache_grid(t) for t in needed}                    # the device grids (themselves checked in test_gpu_vina)
    d = DockOracle(VinaOracle(), grids, begin, end, n, lig)
    return v, d, lig


def _confs(d, k, seed0=1):
    return np.stack([d.random_conf(seed0 + i, [-4, -4, -4], [4, 4, 4])[0] for i in range(k)])


def test_eval_deriv_matches_oracle(setup):
    v, d, lig = setup
    X = np.concatenate([lig["conf0"][None], _confs(d, 40)])
    for caps in ((1000, 1000, 1000), (10, 1.5, 10)):
        e, g, c = v.eval_deriv(X, caps, coords=True)
        for i, x in enumerate(X):
            er, gr = d.eval_deriv(x, caps)
            assert abs(e[i] - er) <= 2e-5 * max(1.0, abs(er)), (i, e[i], er)
            assert np.abs(g[i] - gr).max() <= 2e-4 * max(1.0, np.abs(gr).max())
            assert np.abs(c[i] - d.coords(x)).max() < 2e-5


def test_bfgs_from_identical_starts(setup):
    """One BFGS iteration (gradient, search direction, line search with up to 10 evaluations, Hessian update) must
    coincide with the oracle; over many iterations float-order differences inside an evaluation get amplified by the
    line-search decisions (measured: 100 % identical after 1 iteration, 96 % after 3, 50 % after 12), so long runs are
    compared on what matters: both are descents of the same quality."""
    v, d, lig = setup
    X = _confs(d, 24, seed0=100)
    e0, _ = v.eval_deriv(X)
    for iters, frac in ((1, 0.9), (2, 0.8)):
        e, Xo, g, ne = v.bfgs(X, iters)
        same = 0
        for i in range(len(X)):
            er, xr, gr, ner = d.bfgs(X[i], iters)
            if iters == 1:   # same number of line-search evaluations, same accepted step, same energy
                ok = abs(e[i] - er) <= 1e-4 * max(1.0, abs(er)) and ne[i] == ner and np.abs(Xo[i] - xr).max() < 1e-2
            else:
                ok = abs(e[i] - er) <= 1e-2 * max(1.0, abs(er))
            same += ok
        assert same >= frac * len(X)
    e, Xo, g, ne = v.bfgs(X, 12)
    ref = np.array([d.bfgs(x, 12)[0] for x in X])
    for i in range(len(X)):
        assert e[i] <= e0[i] + 1e-4 * max(1.0, abs(e0[i]))                       # never worse than the start
        assert abs(d.eval_deriv(Xo[i])[0] - e[i]) <= 1e-4 * max(1.0, abs(e[i]))  # returned conf has the returned energy
    assert np.median(e) <= np.median(ref) + 0.05 * abs(np.median(ref)) + 0.5
    assert (ne >= 2).all()


def test_monte_carlo_chains(setup):
    v, d, lig = setup
    seeds = np.arange(1, 33, dtype=np.uint32) * 7919
    e, X, n_out = v.mc(seeds, [-4, -4, -4], [4, 4, 4], num_steps=25, maxiters=8, num_saved_mins=6)
    assert (n_out >= 1).all() and (n_out <= 6).all()
    best_ref = []
    for c in range(len(seeds)):
        k = n_out[c]
        assert np.all(np.diff(e[c, :k]) >= 0)                                     # sorted container
        assert abs(d.eval_grid(X[c, 0]) - e[c, 0]) <= 1e-3 * max(1.0, abs(e[c, 0]))  # energies belong to the poses
        if c < 8:
            best_ref.append(d.mc(int(seeds[c]), [-4, -4, -4], [4, 4, 4], 25, 8, 6)[0][0])
    # same generator, same algorithm: chains coincide until float noise flips a decision; compare the search quality
    # (the synthetic receptor has no pocket, so energies are positive; only the relative quality is meaningful)
    assert np.median(e[:8, 0]) <= np.median(best_ref) + 0.1 * abs(np.median(best_ref)) + 0.5
