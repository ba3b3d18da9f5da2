"""Replay of the reference's own CNN-minimisation test, test/gnina/test_min.py, against the C ABI.

The reference test docks nothing: it minimises a ligand under the parameter-free "overlay" model
test/gnina/data/overlap.pt (score = mean(rec_density * lig_density), loss = -log score) with `--minimize
--cnn_scoring=refinement/all` and asserts that the ligand ends within 0.1 A of the receptor atoms -- the only pin the
reference holds on the gradient chain (autograd through the model, GridMaker::backward, lib/torch_model.cpp:197-221).
Here the same two cases (C.xyz / C1.xyz and CC.xyz / CC2.xyz) are minimised with scipy's BFGS over the ligand's rigid-body
coordinates, energy and gradient coming from gb_cnn_score_grad; tests/golden/overlap*.gbw and overlap_kat.npz were made from
the reference's .pt files by tests/golden/make_overlap_golden.py."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
C_TYPE = 2  # AliphaticCarbonXSHydrophobe: every carbon type maps to the model's single channel


def _scorer(golden_dir, name):
    from gnina_b200 import CNNScorer
    s = CNNScorer(cnn_models=[os.path.join(golden_dir, name + ".gbw")], precision=0)
    return s


@pytest.mark.parametrize("name", ["overlap", "overlap_smallr"])
def test_overlay_model_matches_reference_pt(golden_dir, name):
    """score and ligand-atom gradient against the reference's TorchScript module (fp64 autograd + the oracle's
    GridMaker::backward), for the start poses of test_min.py"""
    kat = np.load(os.path.join(golden_dir, "overlap_kat.npz"))
    s = _scorer(golden_dir, name)
    for case in ("C", "CC"):
        rec, lig = kat["%s_%s_rec" % (name, case)], kat["%s_%s_lig" % (name, case)]
        s.set_receptor(rec, np.full(len(rec), C_TYPE, np.int32))
        out = s.score_grad_batch(lig, np.full(len(lig), C_TYPE, np.int32), [0, len(lig)])
        want = kat["%s_%s_score" % (name, case)]
        assert abs(out[0][0] - want) <= 1e-4 * want
        assert abs(out[2][0] + np.log(want)) <= 1e-4                       # loss = -log score (apply_logistic_loss)
        g = kat["%s_%s_lig_grad" % (name, case)]
        assert np.abs(out[4] - g).max() <= 2e-4 * max(1.0, np.abs(g).max())


def _minimise(s, n_atoms, pack, unpack):
    from scipy.optimize import minimize
    types = np.full(n_atoms, C_TYPE, np.int32)

    def f(p):
        xyz = unpack(p).astype(np.float32)
        out = s.score_grad_batch(xyz, types, [0, len(xyz)])
        return float(out[2][0]), pack(p, np.asarray(out[4], np.float64))
    r = minimize(f, pack(None, None), jac=True, method="BFGS", options={"gtol": 1e-4, "maxiter": 300})
    return unpack(r.x), r


def test_single_atom_is_pulled_onto_the_receptor_atom(golden_dir):
    """test_min.py: -r data/C.xyz -l data/C1.xyz --cnn_scoring=refinement --minimize -> are_similar(C.xyz, out), 0.1 A"""
    s = _scorer(golden_dir, "overlap")
    s.set_receptor(np.zeros((1, 3), np.float32), np.array([C_TYPE], np.int32))
    start = np.array([[1.0, 1.0, 1.0]])
    xyz, r = _minimise(s, 1, lambda p, g: start.ravel().copy() if p is None else g.ravel(), lambda p: p.reshape(1, 3))
    assert np.linalg.norm(xyz[0]) < 0.1, (xyz, r.message)


def test_two_atoms_are_pulled_onto_the_receptor_pair(golden_dir):
    """test_min.py: -r data/CC.xyz -l data/CC2.xyz --cnn_scoring=all --minimize: the rigid C-C ligand (bond 1.6 A) must end
    on the receptor's two carbons (a bijection within 0.1 A).  Rigid body = midpoint + bond direction (theta, phi)."""
    s = _scorer(golden_dir, "overlap")
    rec = np.array([[0, 0, 0], [1.6, 0, 0]], np.float32)
    s.set_receptor(rec, np.full(2, C_TYPE, np.int32))
    h = 0.8

    def unpack(p):
        m, th, ph = p[:3], p[3], p[4]
        u = np.array([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
        return np.stack([m - h * u, m + h * u])

    def pack(p, g):
        if p is None:   # CC2.xyz: (1, -3.2, 1), (1, -1.6, 1): midpoint (1, -2.4, 1), direction +y
            return np.array([1.0, -2.4, 1.0, np.pi / 2, np.pi / 2])
        th, ph = p[3], p[4]
        du_dth = np.array([np.cos(th) * np.cos(ph), np.cos(th) * np.sin(ph), -np.sin(th)])
        du_dph = np.array([-np.sin(th) * np.sin(ph), np.sin(th) * np.cos(ph), 0.0])
        d = h * (g[1] - g[0])
        return np.concatenate([g[0] + g[1], [d @ du_dth, d @ du_dph]])
    xyz, r = _minimise(s, 2, pack, unpack)
    d = np.linalg.norm(xyz[:, None, :] - rec[None], axis=2)
    assert (d.min(axis=1) < 0.1).all() and len(set(d.argmin(axis=1))) == 2, (xyz, r.message)
