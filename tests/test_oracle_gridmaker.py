"""Pins oracle/gridmaker_ref.c to the reference's own voxeliser goldens (test/gninagrid/files/*.binmap,
tolerance of the reference comparator: 1e-4 abs, test/gninagrid/compare_bin.py:24) and checks the unpinned
backward against finite differences of the pinned forward."""
import os
import numpy as np
import pytest
from oracle import gridmaker as gm


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "gridmaker_golden.npz"))


def _cc_system(gold):
    nr, r2c = gm.parse_typemap(str(gold["recmap"]))
    nl, l2c = gm.parse_typemap(str(gold["ligmap"]))
    assert (nr, nl) == (16, 19)
    xyz = gold["cc_xyz"]
    types = np.array([2, 2], np.int32)
    rc, rr = gm.type_atoms(types, r2c, 0)
    lc, lr = gm.type_atoms(types, l2c, nr)
    # receptors that are not .pdbqt are round-tripped through PDBQT text (3 decimals):
    # gninasrc/lib/molgetter.cpp:80,156-164 -> parse_pdbqt.cpp:550-575
    rxyz = np.round(xyz.astype(np.float64), 3).astype(np.float32)
    return (xyz.mean(0), np.concatenate([rxyz, xyz]), np.concatenate([rc, lc]), np.concatenate([rr, lr]), nr + nl)


@pytest.mark.parametrize("tag,res,dim", [("cc48", 0.5, 23.5), ("ccsmall33", 0.25, 8.0)])
def test_forward_matches_reference_golden(gold, tag, res, dim):
    center, xyz, ch, rad, nch = _cc_system(gold)
    out = gm.grid_forward(center, xyz, ch, rad, nch, res, dim)
    assert tuple(out.shape) == tuple(gold[tag + "_shape"])
    ref = np.zeros(out.size, np.float32)
    ref[gold[tag + "_idx"]] = gold[tag + "_val"]
    ref = ref.reshape(out.shape)
    assert np.abs(out - ref).max() < 1e-6          # reference's own bar is 1e-4
    assert np.count_nonzero(out) == np.count_nonzero(ref)
    assert sorted(set(np.nonzero(ref)[0])) == [0, 16]  # rec channel 0, lig channel 16+0


def test_typemap_merges_names_on_a_line():
    from gnina_b200.model_blob import load_model
    b = load_model("crossdock_default2018")
    nr, r2c = gm.parse_typemap(b.recmap)
    nl, l2c = gm.parse_typemap(b.ligmap)
    assert (nr, nl) == (14, 14)
    assert r2c[19] == r2c[20] == r2c[18] == r2c[17] == 4      # Bromine Iodine Chlorine Fluorine share a line
    assert l2c[19] == l2c[20] == 4 and l2c[18] == 5 and l2c[17] == 6
    assert r2c[0] == -1 and r2c[1] == -1                         # hydrogens carry no channel
    assert l2c[23] == l2c[24] == l2c[26] == 13


def test_density_function_values():
    # rho(0)=1, rho(r)=e^-2, rho(1.5 r)=0, continuous at r
    c = np.zeros(3, np.float32)
    for d, want in [(0.0, 1.0), (1.9, np.exp(-2.0)), (1.9 * 1.25, np.exp(-2) * (4 * 1.25 ** 2 - 12 * 1.25 + 9)),
                    (1.9 * 1.5, 0.0)]:
        xyz = np.array([[d - 11.75, -11.75, -11.75]], np.float32)  # so grid point (0,0,0) is at distance d
        g = gm.grid_forward(c, xyz, np.array([0], np.int32), np.array([1.9], np.float32), 1)
        assert abs(g[0, 0, 0, 0] - want) < 2e-6


def test_backward_matches_finite_difference():
    rs = np.random.RandomState(0)
    xyz = (rs.rand(5, 3).astype(np.float32) - 0.5) * 6
    ch = np.array([0, 1, 0, 2, 1], np.int32)
    rad = np.array([1.9, 1.7, 1.8, 2.0, 1.5], np.float32)
    c = np.zeros(3, np.float32)
    w = rs.randn(3, 48, 48, 48).astype(np.float32)
    ana = gm.grid_backward(c, xyz, ch, rad, w)
    h = 1e-2
    for a in range(5):
        for d in range(3):
            p = xyz.copy(); p[a, d] += h
            m = xyz.copy(); m[a, d] -= h
            fd = (np.sum(gm.grid_forward(c, p, ch, rad, 3).astype(np.float64) * w) -
                  np.sum(gm.grid_forward(c, m, ch, rad, 3).astype(np.float64) * w)) / (2 * h)
            assert abs(fd - ana[a, d]) < 2e-2 * max(1.0, abs(fd))


def test_batch_driver_equals_single(gold):
    from gnina_b200 import synth
    rx, rt = synth.make_receptor(400, box=30)
    lx0, lt0 = synth.make_ligand(12, 2)
    lx, offs = synth.make_poses(lx0, 3, trans_box=6)
    n, t2c = gm.parse_typemap(str(gold["ligmap"]))
    rc, rr = gm.type_atoms(rt, t2c, 0)
    lc, lr = gm.type_atoms(np.tile(lt0, 3), t2c, n)
    out = gm.grid_forward_batch(rx, rc, rr, lx, lc, lr, offs, 2 * n, n_threads=2)
    for p in range(3):
        sl = slice(offs[p], offs[p + 1])
        one = gm.grid_forward(gm.center_of(lx[sl]), np.concatenate([rx, lx[sl]]), np.concatenate([rc, lc[sl]]),
                              np.concatenate([rr, lr[sl]]), 2 * n)
        assert np.array_equal(one, out[p])
