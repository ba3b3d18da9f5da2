"""oracle/pipeline.py against the REFERENCE's own CNN scoring host code LIVE: lib/torch_model.cpp, lib/cnn_torch_scorer.cpp and
lib/dl_scorer.cpp compiled where they lie under /root/reference (oracle/Makefile.ref target `cnn`, oracle/ref_cnn_driver.cpp) and run on
the reference's TorchScript files with libtorch (CPU).  Fresh inputs, other ligands than the committed fixture
(tests/golden/cnn_ref_kat.npz, checked by test_oracle_cnn_ref_golden.py everywhere).  Skipped where /root/reference is absent.

Both sides run in this process on the same convolution kernels, so single-model outputs and forces are compared for EQUALITY."""
import numpy as np
import pytest
import torch
from gnina_b200 import model_blob, scorer, synth
from oracle import cnn_refbuild as CR
from oracle import pipeline
from oracle import vina_refbuild as R

pytestmark = pytest.mark.skipif(not (CR.available() or CR.build()), reason="oracle/_ref (CNN half) is not built and /root/reference is absent")


def _case(seed, n_heavy=18, n_tors=3):
    lig = synth.make_flexible_ligand(n_heavy=n_heavy, n_tors=n_tors, n_branch=2, seed=seed)
    rx, rt = synth.make_receptor(500, box=24, seed=seed + 1)
    rm = R.RefModel(lig, rx, rt)
    rs = np.random.RandomState(seed)
    x = lig["conf0"].copy()
    x[:3] += rs.uniform(-1.5, 1.5, 3)
    q = rs.randn(4); x[3:7] = q / np.linalg.norm(q)
    x[7:] = rs.uniform(-np.pi, np.pi, len(x) - 7)
    return lig, rx, rt, rm, rm.set(x.astype(np.float32))


def _oracle(names, rx, rt, xyz, types):
    oms = [pipeline.OracleModel(model_blob.load_model(n)) for n in names]
    s, a, l, v, g = pipeline.score_grad(oms, rx, rt, xyz, types, np.array([0, len(types)], np.int32), dtype=torch.float32)
    return (float(np.float32(s[0])), float(np.float32(a[0])), float(np.float32(l[0])), float(np.float32(v[0]))), g


@pytest.mark.parametrize("name,seed", [("crossdock_default2018_KD_4", 3), ("general_default2018", 4), ("dense_3", 5), ("default2017", 6),
                                       ("all_default_to_default_1_3_2", 7)])
def test_one_model_outputs_and_forces_are_identical(name, seed):
    lig, rx, rt, rm, xyz = _case(seed)
    s = CR.RefCNNScorer(names=[name])
    for _ in range(3):                                   # TorchScript's profiling runs do not change the numbers
        r = s.score(rm, True)
        o, g = _oracle([name], rx, rt, xyz, lig["types"])
        assert r[:4] == o and np.array_equal(r[4], g)
    r0 = s.score(rm, False)                              # score only: same outputs, forces cleared
    assert r0[:4] == o and not r0[4].any()


def test_default_ensemble_mean_variance_and_summed_forces():
    lig, rx, rt, rm, xyz = _case(11, n_heavy=22, n_tors=5)
    r = CR.RefCNNScorer(names=[]).score(rm, True)
    o, g = _oracle(scorer.expand_model_names([]), rx, rt, xyz, lig["types"])
    assert r[:4] == o and r[3] > 0 and np.array_equal(r[4], g)


def test_prefix_ensembles_name_the_same_models():
    """`<prefix>_ensemble` walks the reference's name table (boost::unordered_map: unspecified order, so the float sums may associate
    differently): same members means same mean and variance to float round-off"""
    lig, rx, rt, rm, xyz = _case(12)
    for ens in ("redock_default2018_ensemble", "crossdock_default2018_KD_ensemble"):
        names = scorer.expand_model_names([ens])
        assert len(names) >= 5
        r = CR.RefCNNScorer(names=[ens]).score(rm, False)
        o, _ = _oracle(names, rx, rt, xyz, lig["types"])
        assert np.allclose(r[:4], o, rtol=2e-6, atol=1e-7), (ens, r[:4], o)


def test_unknown_names_are_usage_errors_on_both_sides():
    with pytest.raises(RuntimeError, match="Invalid model name"):
        CR.RefCNNScorer(names=["nonesuch"])
    with pytest.raises(FileNotFoundError, match="Invalid model name"):
        model_blob.load_model("nonesuch")
