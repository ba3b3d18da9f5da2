"""oracle/pipeline.py (the restatement of TorchModel::forward + CNNTorchScorer::score the GPU parity tests check the device against)
against KNOWN ANSWERS OF THE REFERENCE'S OWN CODE: tests/golden/cnn_ref_kat.npz was produced by lib/torch_model.cpp,
lib/cnn_torch_scorer.cpp and lib/dl_scorer.cpp compiled where they lie under /root/reference, running the reference's TorchScript
files with libtorch on the CPU (tests/golden/make_cnn_ref_golden.py, oracle/Makefile.ref target `cnn`; libmolgrid, third party and
absent, is the stand-in over oracle/gridmaker_ref.c).  Runs on any box: the fixture travels, the reference does not.

Where the fixture was generated the restatement is bit-identical to it (tests/test_oracle_cnn_vs_reference_build.py asserts that
live); here the tolerance covers another host's convolution kernels: 2e-5 relative on the outputs, 2e-5 of max |g| on the forces."""
import os
import numpy as np
import pytest
import torch
from gnina_b200 import model_blob, scorer
from oracle import pipeline

KAT = os.path.join(os.path.dirname(__file__), "golden", "cnn_ref_kat.npz")
SINGLE = ["crossdock_default2018", "general_default2018_3", "redock_default2018_2", "dense", "dense_1_3_PT_KD_3", "default2017",
          "all_default_to_default_1_3_1"]


@pytest.fixture(scope="module")
def kat():
    return np.load(KAT)


def _oracle(names, kat, pose, types, centers=None):
    oms = [pipeline.OracleModel(model_blob.load_model(n)) for n in names]
    xyz = kat["coords"][pose]
    off = np.array([0, len(types)], np.int32)
    s, a, l, v, g = pipeline.score_grad(oms, kat["rec_xyz"], kat["rec_types"], xyz, types, off, centers=centers, dtype=torch.float32)
    return np.float32([s[0], a[0], l[0], v[0]]), g


def _close(o, ref, g, forces):
    assert np.allclose(o, ref, rtol=2e-5, atol=1e-7), (o, ref)
    assert np.abs(g - forces).max() <= 2e-5 * max(np.abs(forces).max(), 1e-3)


@pytest.mark.parametrize("name", SINGLE)
def test_single_models(kat, name):
    """make_coordset + type maps of the model file, centre = mean over the ligand atoms, rec + lig merge, module forward, softmax /
    cross-entropy head, autograd backward + GridMaker::backward, gradient split (torch_model.cpp:120-224) for every architecture"""
    for pose in (0, 1):
        o, g = _oracle([name], kat, pose, kat["lig_types"])
        _close(o, kat["single_%s_out" % name][pose], g, kat["single_%s_forces" % name][pose])


@pytest.mark.parametrize("key,names", [("default_ensemble", []), ("fast", ["fast"]), ("default1_0", ["default1.0"])])
def test_name_logic_and_ensemble_arithmetic(kat, key, names):
    """CNNTorchScorer's constructor (default ensemble, `fast`, `default1.0`; cnn_torch_scorer.cpp:24-46) and score(): mean score in a
    double, float affinity / loss sums, population variance of the affinities, forces summed over the models in the model's float
    minus_forces and scaled once by 1/cnt (:117-192) -- the product's expand_model_names must name the same models"""
    o, g = _oracle(scorer.expand_model_names(names), kat, 0, kat["lig_types"])
    _close(o, kat["alias_%s_out" % key], g, kat["alias_%s_forces" % key])
    if key != "fast":
        assert kat["alias_%s_out" % key][3] > 0                      # several models: a variance


def test_hydrogens_centre_and_box(kat):
    """setLigand hands every movable atom to the network, hydrogens included (dl_scorer.cpp:72-87); the typer leaves them untyped;
    getGradient builds a by-atom list and model::add_minus_forces consumes it compactly over the non-hydrogen atoms (model.cu:247-259):
    the j-th heavy atom receives the gradient of movable atom j.  set_center_from_model = mean of the heavy atoms' coordinates,
    set_bounding_box = centre -+ dimension / 2 (dl_scorer.cpp:196-217, cnn_torch_scorer.cpp:229-241)."""
    ty = kat["lig_types_h"]
    o, g = _oracle(["crossdock_default2018"], kat, 0, ty)
    assert np.all(g[ty < 2] == 0)
    heavy = np.flatnonzero(ty > 1)
    compact = np.zeros_like(g)
    compact[heavy] = g[:len(heavy)]
    _close(o, kat["hyd_out"], compact, kat["hyd_forces"])
    assert np.abs(g - kat["hyd_forces"]).max() > 0.1                 # ... which is NOT the by-atom gradient
    xyz = kat["coords"][0]
    c = np.zeros(3, np.float32)
    for p in xyz[heavy]:
        c = (c + p).astype(np.float32)
    c = (c / np.float32(len(heavy))).astype(np.float32)
    assert np.array_equal(c, kat["hyd_center"])
    assert np.allclose(kat["hyd_box_begin"], c - 23.5 / 2, atol=2e-6) and np.allclose(kat["hyd_box_end"], c + 23.5 / 2, atol=2e-6)
    assert list(kat["hyd_box_n"]) == [47, 47, 47]


def test_given_centre_and_model_file(kat):
    """--cnn_center fixes the grid centre (cnn_torch_scorer.cpp:131-134); --cnn_models loads a TorchScript file: the reference's own
    test network test/gnina/data/overlap.pt (converted blob: tests/golden/overlap.gbw, tests/golden/make_overlap_golden.py)"""
    o, _ = _oracle(["crossdock_default2018"], kat, 0, kat["lig_types"], centers=kat["center_given"][None])
    assert np.allclose(o[:3], kat["center_given_out"][:3], rtol=2e-5, atol=1e-7)
    blob = os.path.join(os.path.dirname(__file__), "golden", "overlap.gbw")
    if not os.path.exists(blob):
        pytest.skip("tests/golden/overlap.gbw is not committed")
    om = pipeline.OracleModel(model_blob.load_model(blob))
    off = np.array([0, len(kat["lig_types"])], np.int32)
    s, a, l, v, g = pipeline.score_grad([om], kat["rec_xyz"], kat["rec_types"], kat["coords"][0], kat["lig_types"], off, dtype=torch.float32)
    _close(np.float32([s[0], a[0], l[0], v[0]]), kat["overlap_out"], g, kat["overlap_forces"])
