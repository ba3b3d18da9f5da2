"""CPU checks of the docking inner-loop oracle (oracle/vina_mc_ref.c): kinematics, analytic gradient vs finite
differences, BFGS descent, Monte-Carlo chain bookkeeping."""
import numpy as np
import pytest
from gnina_b200 import synth
from oracle.vina import VinaOracle
from oracle.vina_mc import DockOracle


@pytest.fixture(scope="module")
def dock():
    rx, rt = synth.make_receptor(600, box=30)
    lig = synth.make_flexible_ligand()
    vo = VinaOracle()
    begin, end, n = [-10.0] * 3, [10.0] * 3, [40, 40, 40]
    needed = sorted(set(int(t) for t in lig["types"] if t > 1))
    grids = {t: vo.cache_populate(begin, end, n, rx, rt, t) for t in needed}
    return DockOracle(vo, grids, begin, end, n, lig), lig


def test_conf0_reproduces_coordinates_and_rigid_motion(dock):
    d, lig = dock
    assert np.abs(d.coords(lig["conf0"]) - lig["xyz0"]).max() < 1e-5
    c = lig["conf0"].copy()
    c[:3] += [1.0, -2.0, 0.5]
    assert np.abs(d.coords(c) - (lig["xyz0"] + [1.0, -2.0, 0.5])).max() < 1e-5
    # a torsion moves only the atoms of its sub-tree and preserves bond lengths to the parent
    c = lig["conf0"].copy(); c[7 + 2] = 0.9
    moved = np.abs(d.coords(c) - lig["xyz0"]).max(1) > 1e-4
    b = lig["seg_begin"]
    assert not moved[: b[3]].any() and moved[b[3] + 1: lig["seg_end"][len(b) - 2]].any()
    assert abs(np.linalg.norm(d.coords(c)[b[3]] - d.coords(c)[b[3] - 1]) - np.linalg.norm(lig["xyz0"][b[3]] - lig["xyz0"][b[3] - 1])) < 1e-5


def test_change_is_the_gradient_of_the_energy(dock):
    d, lig = dock
    x = lig["conf0"].copy(); x[:3] = [0.5, -0.3, 0.8]; x[7:] = np.linspace(-1, 1, d.T)
    e, g = d.eval_deriv(x)
    # the reference's "derivative" is an independently interpolated table (precalculate_linear smooth.second) and a
    # trilinear grid slope, not the exact derivative of the interpolated energy: agreement to a few % of |g|max
    h, tol = 2e-3, 0.05 * np.abs(g).max()
    for k in range(3):                       # position
        xp, xm = x.copy(), x.copy(); xp[k] += h; xm[k] -= h
        fd = (d.eval_deriv(xp)[0] - d.eval_deriv(xm)[0]) / (2 * h)
        assert abs(fd - g[k]) < tol
    for k in range(d.T):                     # torsions
        xp, xm = x.copy(), x.copy(); xp[7 + k] += h; xm[7 + k] -= h
        fd = (d.eval_deriv(xp)[0] - d.eval_deriv(xm)[0]) / (2 * h)
        assert abs(fd - g[6 + k]) < tol


def test_bfgs_descends_and_mc_keeps_sorted_minima(dock):
    d, lig = dock
    x0, _ = d.random_conf(7, [-3, -3, -3], [3, 3, 3])
    e0, _ = d.eval_deriv(x0)
    e1, x1, g1, ne = d.bfgs(x0, 30)
    assert e1 <= e0 and ne >= 2
    assert abs(d.eval_deriv(x1)[0] - e1) < 1e-4 * max(1.0, abs(e1))
    es, xs = d.mc(123, [-3, -3, -3], [3, 3, 3], num_steps=15, maxiters=8, num_saved_mins=5)
    assert 1 <= len(es) <= 5 and np.all(np.diff(es) >= 0)
    assert abs(d.eval_grid(xs[0]) - es[0]) < 1e-4 * max(1.0, abs(es[0]))
    es2, _ = d.mc(123, [-3, -3, -3], [3, 3, 3], num_steps=15, maxiters=8, num_saved_mins=5)
    assert np.array_equal(es, es2)           # deterministic in the seed
