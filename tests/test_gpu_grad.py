"""GPU parity of the gradient path (G2 + N5 + S1's accumulation): gb_cnn_score_grad vs autograd through the
reference's own TorchScript model (tests/golden/grad_kat.npz) and vs finite differences of the library's own loss."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "cnn_kat.npz"))


def test_ligand_gradient_matches_reference_autograd(kat, golden_dir):
    from gnina_b200 import CNNScorer
    g = np.load(os.path.join(golden_dir, "grad_kat.npz"))
    n = int(g["n_poses"])
    offs = kat["pose_offsets"][:n + 1]
    s = CNNScorer([str(g["model"])], precision=0)
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    sc, aff, loss, var, grad = s.score_grad_batch(kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]], offs)
    assert np.abs(loss - g["loss"]).max() < 1e-4
    scale = np.abs(g["lig_grad"]).max()
    assert np.abs(grad - g["lig_grad"]).max() < 2e-4 * scale
    assert np.abs(grad[kat["lig_types"][:offs[-1]] <= 1]).max() == 0.0
    # forward outputs of the gradient call equal the plain scoring call in validation mode
    s.set_option("precision", 0)
    plain = s.score_batch(kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]], offs)
    assert np.abs(plain[0] - sc).max() < 1e-6 and np.abs(plain[1] - aff).max() < 1e-5


def test_gradient_is_the_derivative_of_the_loss(kat):
    """central finite differences of the library's own (fp32) loss along random ligand displacements"""
    from gnina_b200 import CNNScorer
    s = CNNScorer(["crossdock_default2018"], precision=0)
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    offs = kat["pose_offsets"][:2]
    x, t = kat["lig_xyz"][:offs[-1]].copy(), kat["lig_types"][:offs[-1]]
    center = x.mean(0)[None]                      # hold the grid fixed while atoms move
    _, _, _, _, grad = s.score_grad_batch(x, t, offs, center)
    rs = np.random.RandomState(0)
    for _ in range(3):
        d = rs.randn(*x.shape).astype(np.float32)
        d /= np.linalg.norm(d)
        h = 2e-2
        lp = s.score_batch(x + h * d, t, offs, center)[2][0]
        lm = s.score_batch(x - h * d, t, offs, center)[2][0]
        fd = (lp - lm) / (2 * h)
        an = float((grad * d).sum())
        assert abs(fd - an) < 3e-2 * max(abs(an), 0.05)


def test_ensemble_gradient_is_mean_of_model_gradients(kat):
    from gnina_b200 import CNNScorer
    names = ["crossdock_default2018", "crossdock_default2018_KD_4"]
    offs = kat["pose_offsets"][:3]
    x, t = kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]]
    gs = []
    for nm in names:
        s = CNNScorer([nm], precision=0)
        s.set_receptor(kat["rec_xyz"], kat["rec_types"])
        gs.append(s.score_grad_batch(x, t, offs)[4])
    e = CNNScorer(names, precision=0)
    e.set_receptor(kat["rec_xyz"], kat["rec_types"])
    out = e.score_grad_batch(x, t, offs)
    assert np.abs(out[4] - 0.5 * (gs[0] + gs[1])).max() < 1e-5 * max(1.0, np.abs(out[4]).max())
    assert out[3].max() > 0          # affinity variance of a 2-model ensemble


def test_gradient_rejects_unsupported_models():
    from gnina_b200 import CNNScorer, capi
    s = CNNScorer(["default2017"])
    s.set_receptor(np.zeros((1, 3), np.float32), np.array([2], np.int32))
    with pytest.raises(capi.GbError, match="default2018 and dense"):
        s.score_grad_batch(np.zeros((1, 3), np.float32), np.array([2], np.int32), [0, 1])


def test_dense_gradient_matches_reference_autograd(kat, golden_dir):
    """dense family (N2) backward: max-pool argmax routing, folded BatchNorm, concat accumulation, global max"""
    from gnina_b200 import CNNScorer
    g = np.load(os.path.join(golden_dir, "grad_kat_dense.npz"))
    n = int(g["n_poses"])
    offs = kat["pose_offsets"][:n + 1]
    x, t = kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]]
    s = CNNScorer([str(g["model"])])           # default precision: a dense member selects the fp32 kernels
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    sc, aff, loss, var, grad = s.score_grad_batch(x, t, offs)
    assert np.abs(loss - g["loss"]).max() < 1e-4
    scale = np.abs(g["lig_grad"]).max()
    assert np.abs(grad - g["lig_grad"]).max() < 5e-4 * scale
    assert np.abs(grad[t <= 1]).max() == 0.0


def test_default_ensemble_gradient_matches_oracle(kat):
    """`--cnn_scoring refinement` with gnina's default ensemble (2 dense + 1 default2018): mean gradient (S1)"""
    import torch
    from gnina_b200 import CNNScorer, model_blob
    from oracle import pipeline
    n = 1
    offs = kat["pose_offsets"][:n + 1]
    x, t = kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]]
    s = CNNScorer([])
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    out = s.score_grad_batch(x, t, offs)
    oms = [pipeline.OracleModel(model_blob.load_model(nm)) for nm in s.model_names]
    ref = pipeline.score_grad(oms, kat["rec_xyz"], kat["rec_types"], x, t, offs, dtype=torch.float32)
    assert abs(out[0][0] - ref[0][0]) < 2e-5 and abs(out[1][0] - ref[1][0]) < 1e-4 and abs(out[3][0] - ref[3][0]) < 1e-3
    scale = np.abs(ref[4]).max()
    assert np.abs(out[4] - ref[4]).max() < 5e-4 * scale


# ---- fast mode: tcgen05 backward-data convolutions, fp16 gradients with loss scaling (gb_cnn_tc_grad.cu) ----------
FAST_GRAD_TOL = 2e-2     # of max |gradient| on the reference's own vectors (measured 4.7e-3); fp16 operands end to end
FAST_GRAD_TOL_CLASH = 6e-2   # synthetic screens put ligands INSIDE receptor atoms (loss 5..17): ReLU masks of the fp16
                             # forward flip for near-zero units; measured max 4.0e-2 of the batch max, median 4e-3
FAST_LOSS_TOL = 5e-3         # golden vectors
FAST_LOSS_TOL_CLASH = 1.5e-2 # absolute, on losses of 5..17 (measured 6.7e-3)


def test_fast_gradient_matches_reference_autograd(kat, golden_dir):
    from gnina_b200 import CNNScorer
    g = np.load(os.path.join(golden_dir, "grad_kat.npz"))
    n = int(g["n_poses"])
    offs = kat["pose_offsets"][:n + 1]
    s = CNNScorer([str(g["model"])], precision=1)
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    x, t = kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]]
    sc, aff, loss, var, grad = s.score_grad_batch(x, t, offs)
    assert np.abs(loss - g["loss"]).max() < FAST_LOSS_TOL
    scale = np.abs(g["lig_grad"]).max()
    assert np.abs(grad - g["lig_grad"]).max() < FAST_GRAD_TOL * scale
    assert np.abs(grad[t <= 1]).max() == 0.0
    # the forward outputs of the gradient call are the fast scoring path's up to fp16 round-off: scoring runs the fused
    # unit1_conv/unit2_conv/pool kernel (pooling in fp32 straight from the accumulator), the gradient call keeps Y1 and pools
    # fp16-rounded unit2 outputs
    plain = s.score_batch(x, t, offs)
    assert np.abs(plain[0] - sc).max() < 1e-4 and np.abs(plain[1] - aff).max() < 1e-3


def test_fast_gradient_matches_validation_path_on_a_ragged_multi_chunk_batch(kat):
    from gnina_b200 import CNNScorer, synth
    xyz, types, offs = synth.make_screen(301, seed=5, trans_box=6.0)   # odd pose count: half-filled pose group
    names = ["crossdock_default2018", "crossdock_default2018_KD_4"]
    ref = CNNScorer(names, precision=0)
    fast = CNNScorer(names, precision=1, max_batch=128)              # 3 chunks, the last one ragged
    for s in (ref, fast):
        s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    a = ref.score_grad_batch(xyz, types, offs)
    b = fast.score_grad_batch(xyz, types, offs)
    assert np.abs(a[2] - b[2]).max() < FAST_LOSS_TOL_CLASH
    ga, gb = a[4], b[4]
    scale = np.abs(ga).max()
    assert np.isfinite(gb).all()
    assert np.abs(ga - gb).max() < FAST_GRAD_TOL_CLASH * scale
    per_pose = np.array([np.abs(ga[offs[i]:offs[i + 1]] - gb[offs[i]:offs[i + 1]]).max() /
                         max(np.abs(ga[offs[i]:offs[i + 1]]).max(), 1e-6) for i in range(len(offs) - 1)])
    assert np.median(per_pose) < 1e-2
    cos = float((ga * gb).sum() / (np.linalg.norm(ga) * np.linalg.norm(gb)))
    assert cos > 0.999
    # a second call on the same handle (buffers reused, smaller batch) is unaffected by stale workspace contents
    k = 7
    b2 = fast.score_grad_batch(xyz[:offs[k]], types[:offs[k]], offs[:k + 1])
    assert np.abs(b2[4] - gb[:offs[k]]).max() < 1e-6 * max(scale, 1.0)


def test_receptor_gradient_matches_reference_autograd(kat, golden_dir):
    """getReceptorGradient (flexible-residue atoms): d loss / d receptor atoms of a single pose, validation and fast
    kernels, against autograd through the reference .pt + the oracle's GridMaker::backward"""
    from gnina_b200 import CNNScorer, capi
    g = np.load(os.path.join(golden_dir, "grad_kat.npz"))
    offs = kat["pose_offsets"]
    for prec, tol in ((0, 2e-4), (1, FAST_GRAD_TOL)):
        s = CNNScorer([str(g["model"])], precision=prec)
        s.set_receptor(kat["rec_xyz"], kat["rec_types"])
        for p in range(int(g["n_poses"])):
            x, t = kat["lig_xyz"][offs[p]:offs[p + 1]], kat["lig_types"][offs[p]:offs[p + 1]]
            out = s.score_grad_batch(x, t, [0, len(t)], receptor=True)
            want = g["rec_grad"][p]
            scale = np.abs(want).max()
            assert out[5].shape == want.shape
            assert np.abs(out[5] - want).max() < tol * scale, (prec, p)
            assert np.abs(out[5][kat["rec_types"] <= 1]).max() == 0.0 if (kat["rec_types"] <= 1).any() else True
            assert np.abs(out[4] - g["lig_grad"][offs[p]:offs[p + 1]]).max() < tol * np.abs(g["lig_grad"]).max()
        with pytest.raises(capi.GbError, match="single pose"):
            s.score_grad_batch(kat["lig_xyz"][:offs[2]], kat["lig_types"][:offs[2]], offs[:3], receptor=True)
    # a clone carries the receptor and answers the same
    c = s.fresh_copy()
    x, t = kat["lig_xyz"][:offs[1]], kat["lig_types"][:offs[1]]
    assert np.array_equal(c.score_grad_batch(x, t, [0, len(t)], receptor=True)[5], s.score_grad_batch(x, t, [0, len(t)], receptor=True)[5])
