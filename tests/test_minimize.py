"""gnina_b200/minimize.py on the CPU without the reference: the lock-step quasi-Newton driver and its torsion-tree kinematics
(the bit-for-bit comparison with the reference's quasi_newton + non_cache_cnn is tests/test_oracle_vs_reference_build.py)."""
import numpy as np
from gnina_b200 import minimize, synth


def _setup(n=12, seed=5):
    lig = synth.make_flexible_ligand()
    tree = minimize.TorsionTree(lig)
    rs = np.random.RandomState(seed)
    X = np.tile(lig["conf0"], (n, 1)).astype(np.float32)
    X[:, :3] += rs.uniform(-5, 5, (n, 3))
    q = rs.randn(n, 4); X[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    X[:, 7:] = rs.uniform(-np.pi, np.pi, (n, tree.T))
    return lig, tree, X


def _quadratic(target, k=np.float32(0.02)):
    def energy(coords, idx):
        d = (coords - target).astype(np.float32)
        return (k * d * d).sum(axis=(1, 2)).astype(np.float32), (2 * k * d).astype(np.float32)
    return energy


def test_kinematics_reproduce_the_start_pose_and_rigid_motion():
    lig, tree, X = _setup()
    c0 = tree.set_conf(lig["conf0"][None])[0][0]
    assert np.abs(c0 - lig["xyz0"]).max() < 1e-5
    x = lig["conf0"].copy(); x[:3] += [1.0, -2.0, 0.5]
    assert np.abs(tree.set_conf(x[None])[0][0] - (lig["xyz0"] + [1.0, -2.0, 0.5])).max() < 1e-5


def test_change_is_the_gradient_of_the_energy():
    """heterotree::derivative turns atom forces into d energy / d (position, rotation vector, torsions): finite differences"""
    lig, tree, X = _setup(n=3)
    energy = _quadratic(np.float32([0.3, -0.2, 0.1]))
    for x in X:
        coords, so, sa = tree.set_conf(x[None])
        e, mf = energy(coords, np.arange(1))
        g = tree.derivative(coords, mf, so, sa)[0]
        h = np.float32(2e-3)
        for k in list(range(3)) + [6 + t for t in range(tree.T)]:
            p = np.zeros(6 + tree.T, np.float32); p[k] = 1
            ep = energy(tree.set_conf(minimize.conf_increment(x[None], p[None], np.array([h], np.float32), tree.T))[0], np.arange(1))[0][0]
            em = energy(tree.set_conf(minimize.conf_increment(x[None], p[None], np.array([-h], np.float32), tree.T))[0], np.arange(1))[0][0]
            assert abs((ep - em) / (2 * h) - g[k]) <= 2e-2 * max(1.0, np.abs(g).max())


def test_lock_step_equals_pose_by_pose_and_descends():
    lig, tree, X = _setup()
    energy = _quadratic(np.float32([0.5, -0.3, 0.2]))
    e0 = energy(tree.set_conf(X)[0], np.arange(len(X)))[0]
    for acc, et in ((True, False), (False, True)):
        e, x, ev, rounds = minimize.minimize_poses(tree, energy, X, maxiters=300, accurate=acc, early_term=et)
        assert (e <= e0).all() and (e < 0.5 * e0).all() and (ev >= 2).all()
        assert np.allclose(energy(tree.set_conf(x)[0], np.arange(len(X)))[0], e, rtol=1e-5)      # the returned conf has the returned energy
        for i in (0, 5, 11):
            e1, x1, _, _ = minimize.minimize_poses(tree, energy, X[i:i + 1], maxiters=300, accurate=acc, early_term=et)
            assert e1[0] == e[i] and np.array_equal(x1[0], x[i])                                  # no cross-talk between poses
        assert rounds < ev.sum()                                                                   # batched energy calls


def test_box_penalties():
    coords = np.float32([[[0, 0, 0], [3, -4, 0], [9, 9, 9]]])
    heavy = np.array([True, True, False])
    pen, d = minimize.box_penalty(coords, heavy, [-2, -2, -2], [2, 2, 2], 10.0)
    assert pen.tolist() == [[0.0, 30.0, 0.0]] and d[0, 1].tolist() == [10.0, -10.0, 0.0] and (d[0, 2] == 0).all()
    e, f = minimize.with_box_penalties(np.float32([1.5]), np.ones((1, 3, 3), np.float32), coords, heavy, ([-2] * 3, [2] * 3),
                                       (np.float32([-1] * 3), np.float32([1] * 3)), 10.0)
    assert e[0] == np.float32(1.5 + 30.0 + 50.0) and f[0, 1].tolist() == [21.0, -19.0, 1.0] and (f[0, 2] == 0).all()
