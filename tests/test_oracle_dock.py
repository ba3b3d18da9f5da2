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


# ---- non_cache + refine_structure (lib/non_cache.cpp, main/main.cpp:131-171): oracle groundwork for the "next" row ----
@pytest.fixture(scope="module")
def noncache():
    rx, rt = synth.make_receptor(600, box=30)
    lig = synth.make_flexible_ligand()
    vo = VinaOracle()
    begin, end, n = [-6.0] * 3, [6.0] * 3, [32, 32, 32]
    needed = sorted(set(int(t) for t in lig["types"] if t > 1))
    grids = {t: vo.cache_populate(begin, end, n, rx, rt, t) for t in needed}
    d = DockOracle(vo, grids, begin, end, n, lig, slope=10.0)
    return d, lig, rx, rt, grids, (begin, end, n)


def test_noncache_atom_term_vs_cache_grid_and_bounds(noncache):
    d, lig, rx, rt, grids, (begin, end, n) = noncache
    d.use_noncache(rx, rt)
    t = int([q for q in lig["types"] if q > 1][0])
    g = grids[t].reshape(n[2] + 1, n[1] + 1, n[0] + 1)          # x fastest
    # at grid NODES the cache holds sum eval_fast (piecewise-constant midpoint table), non_cache sums the linearly
    # interpolated table: same quantity up to the table resolution (1/32 in r^2); compare where no clash dominates
    sp = (np.array(end) - np.array(begin)) / np.array(n)
    rs = np.random.RandomState(0)
    rel = []
    for _ in range(200):
        ijk = rs.randint(2, 30, 3)
        xyz = np.array(begin) + ijk * sp
        e_nc, _ = d.noncache_atom(t, xyz, v=3.0e38)                       # no curl
        e_c = g[ijk[2], ijk[1], ijk[0]]
        if abs(e_c) < 5:
            rel.append(abs(e_nc - e_c) / max(0.05, abs(e_c)))
    assert len(rel) > 50 and np.median(rel) < 2e-2 and np.percentile(rel, 90) < 0.2
    # derivative = gradient of the energy (finite differences of the oracle's own energy)
    x0 = np.array([1.3, -0.7, 2.1], np.float32)
    e0, dv = d.noncache_atom(t, x0)
    for k in range(3):
        h = np.zeros(3, np.float32); h[k] = 5e-3
        fd = (d.noncache_atom(t, x0 + h)[0] - d.noncache_atom(t, x0 - h)[0]) / 1e-2
        assert abs(fd - dv[k]) < 0.05 * max(1.0, np.abs(dv).max())
    # outside the box: the atom is evaluated at the clamped position + slope * L1 distance, force -+ slope
    inside = np.array([5.9, 0.0, 0.0], np.float32); edge = np.array([6.0, 0.0, 0.0], np.float32)
    out = np.array([7.5, 0.0, 0.0], np.float32)
    e_edge, d_edge = d.noncache_atom(t, edge)
    e_out, d_out = d.noncache_atom(t, out)
    assert abs(e_out - (e_edge + 10.0 * 1.5)) < 1e-4 * max(1.0, abs(e_out))
    assert abs(d_out[0] - (d_edge[0] + 10.0)) < 1e-4 * max(1.0, abs(d_out[0])) and np.allclose(d_out[1:], d_edge[1:], atol=1e-5)
    assert np.isfinite(d.noncache_atom(t, inside)[0])
    d.use_noncache(None)


def test_refine_structure_descends_and_pulls_the_ligand_into_the_box():
    rx, rt = synth.make_receptor(600, box=30)
    lig = synth.make_flexible_ligand()
    vo = VinaOracle()
    begin, end, n = [-10.0] * 3, [10.0] * 3, [40, 40, 40]
    needed = sorted(set(int(t) for t in lig["types"] if t > 1))
    grids = {t: vo.cache_populate(begin, end, n, rx, rt, t) for t in needed}
    d = DockOracle(vo, grids, begin, end, n, lig, slope=10.0)
    d.use_noncache(rx, rt)
    x = lig["conf0"].copy(); x[:3] = [0.4, -0.2, 0.6]
    e0, _ = d.eval_deriv(x)
    e1, x1, ne, ok = d.refine_structure(x, 17)
    # well inside the box the first BFGS run (slope 10) already ends within: one run, a descent
    assert ok and ne >= 2 and e1 <= e0 + 1e-4 * max(1.0, abs(e0))
    assert abs(d.eval_deriv(x1)[0] - e1) < 1e-3 * max(1.0, abs(e1))
    assert np.abs(x1 - d.bfgs(x, 17)[1]).max() < 1e-6            # == quasi_newton with the field's slope (10)
    # start outside the box: the out-of-box slope escalates (10, 100, ...) and drags every heavy atom to the box; the
    # synthetic receptor has no pocket, so the ligand leans on the wall: what remains outside shrinks with the slope
    far = lig["conf0"].copy(); far[:3] = [14.0, 0.0, 0.0]
    assert not d.within(far)
    before = np.clip(np.abs(d.coords(far)) - 10.0, 0, None).max()
    e2, x2, ne2, ok2 = d.refine_structure(far, 40)
    after = np.clip(np.abs(d.coords(x2)) - 10.0, 0, None).max()
    assert before > 3.0 and after < 0.05 and ne2 > ne
    assert ok2 == d.within(x2)
    # the non-cache energy of a pose tracks the cache energy (trilinear interpolation of eval_fast vs direct sums)
    d.use_noncache(None)
    e_cache, _ = d.eval_deriv(x1)
    d.use_noncache(rx, rt)
    e_nc, _ = d.eval_deriv(x1)
    assert abs(e_nc - e_cache) < 0.15 * max(1.0, abs(e_nc))
    d.use_noncache(None)
