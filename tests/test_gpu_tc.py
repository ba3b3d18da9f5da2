"""GPU parity tests of the fast path (fused pooled voxeliser + tcgen05 fp16 convolutions), through the C ABI.

Stated tolerance of the fast mode (fp16 operands, fp32 accumulation): |dCNNscore| <= 2e-3, |dCNNaffinity| <= 1e-2
against the fp64 run of the reference's own TorchScript model (the reference accepts 1e-3 between its own CPU and
GPU paths, test/gnina/test_cnn.py:43)."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL_SCORE, TOL_AFF = 2e-3, 1e-2


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "cnn_kat.npz"))


def _fast(names):
    from gnina_b200 import CNNScorer
    s = CNNScorer(names)
    s.set_option("precision", 1)
    return s


def test_fast_path_is_the_default_for_default2018():
    from gnina_b200 import CNNScorer
    assert CNNScorer(["crossdock_default2018"]).get_option("precision") == 1


@pytest.mark.parametrize("name", ["crossdock_default2018", "crossdock_default2018_KD_4", "all_default_to_default_1_3_1"])
def test_fast_scores_match_reference_pt(kat, name):
    s = _fast([name])
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    pose, aff, loss, var = s.score_batch(kat["lig_xyz"], kat["lig_types"], kat["pose_offsets"])
    assert np.abs(pose - kat[name + "_pose_f64"]).max() < TOL_SCORE
    assert np.abs(aff - kat[name + "_aff_f64"]).max() < TOL_AFF


def test_fast_intermediates_match_oracle(kat):
    import tc_layout as tl
    from gnina_b200 import model_blob
    from oracle import pipeline
    name, n = "crossdock_default2018", 3
    offs = kat["pose_offsets"][:n + 1]
    lx, lt = kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]]
    blob = model_blob.load_model(name)
    ref = tl.oracle_intermediates(blob, pipeline.OracleModel(blob).grids(kat["rec_xyz"], kat["rec_types"], lx, lt, offs))
    s = _fast([name])
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    s.score_batch(lx, lt, offs)
    # scoring path: the pooled grid is laid out in row groups of 8 poses for the fused unit1_conv kernel
    x0, border = tl.decode_chunk_planar(s.debug_read("x0"), n, 24, 8, 32)
    assert border == 0.0 and np.abs(x0[:, 28:]).max() == 0.0
    assert np.abs(x0[:, :28] - ref["x0"]).max() < 2e-3          # fp16 rounding of densities <= ~3
    x2, border = tl.decode_chunk_planar(s.debug_read("x2"), n, 12, 2, 32)
    assert border == 0.0
    for tag, got in (("x2", x2), ("y3", tl.decode_channels_last(s.debug_read("y3"), n, 12, 64)),
                     ("y5", tl.decode_channels_last(s.debug_read("y5"), n, 6, 128))):
        scale = np.abs(ref[tag]).max()
        assert np.abs(got - ref[tag]).max() < 4e-3 * scale, tag


def test_fast_matches_fp32_validation_mode_on_many_ragged_poses():
    """size-independent property: both precisions of the library agree pose by pose on a larger ragged batch"""
    from gnina_b200 import CNNScorer, synth
    rx, rt = synth.make_receptor(2000, box=50)
    lx, lt, offs = synth.make_screen(37, seed=5, trans_box=12)
    a = CNNScorer(["crossdock_default2018"], precision=0)
    b = CNNScorer(["crossdock_default2018"], precision=1)
    for s in (a, b):
        s.set_receptor(rx, rt)
    ra, rb = a.score_batch(lx, lt, offs), b.score_batch(lx, lt, offs)
    assert np.abs(ra[0] - rb[0]).max() < TOL_SCORE and np.abs(ra[1] - rb[1]).max() < TOL_AFF
    b.set_option("max_batch", 5)   # odd chunk sizes exercise the half-filled pose groups
    rc = b.score_batch(lx, lt, offs)
    assert np.abs(rc[0] - rb[0]).max() < 1e-6 and np.abs(rc[1] - rb[1]).max() < 1e-5


@pytest.mark.parametrize("name", ["dense_1_3", "dense_1_3_PT_KD_3"])
def test_dense_fast_scores_match_reference_pt(kat, name):
    s = _fast([name])
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    pose, aff, loss, var = s.score_batch(kat["lig_xyz"], kat["lig_types"], kat["pose_offsets"])
    assert np.abs(pose - kat[name + "_pose_f64"]).max() < TOL_SCORE
    assert np.abs(aff - kat[name + "_aff_f64"]).max() < TOL_AFF


def test_default2017_fast_path_matches_reference_pt(kat):
    """N3: default2017 (35 channels, max pooling, three 3x3x3 convolutions) on the tensor-core path -- 48-channel max-pool
    voxeliser, conv3_tc_kernel<48,24>, max-pool re-layout kernels -- against the reference's own default2017.pt (fp64)"""
    from gnina_b200 import CNNScorer
    s = CNNScorer(["default2017"])
    assert s.get_option("precision") == 1                      # the fast path is the default now
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    pose, aff, loss, var = s.score_batch(kat["lig_xyz"], kat["lig_types"], kat["pose_offsets"])
    assert np.abs(pose - kat["default2017_pose_f64"]).max() < TOL_SCORE
    assert np.abs(aff - kat["default2017_aff_f64"]).max() < TOL_AFF
    v = CNNScorer(["default2017"], precision=0)
    v.set_receptor(kat["rec_xyz"], kat["rec_types"])
    ref = v.score_batch(kat["lig_xyz"], kat["lig_types"], kat["pose_offsets"])
    assert np.abs(pose - ref[0]).max() < TOL_SCORE and np.abs(aff - ref[1]).max() < TOL_AFF
    # ragged batch sizes (pose groups of two for the 12^3 and 6^3 layouts)
    n = 5
    offs = kat["pose_offsets"][:n + 1]
    sub = s.score_batch(kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]], offs)
    assert np.abs(sub[0] - pose[:n]).max() < 1e-6


def test_dense_fast_intermediates_match_oracle(kat):
    import tc_layout as tl
    from gnina_b200 import model_blob
    from oracle import pipeline
    name, n = "dense_1_3", 3
    offs = kat["pose_offsets"][:n + 1]
    lx, lt = kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]]
    blob = model_blob.load_model(name)
    ref = tl.oracle_intermediates_dense(blob, pipeline.OracleModel(blob).grids(kat["rec_xyz"], kat["rec_types"], lx, lt, offs))
    s = _fast([name])
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    s.score_batch(lx, lt, offs)
    x0, border = tl.decode_chunk_planar(s.debug_read("x0"), n, 24, 1, 32)
    assert border == 0.0 and np.abs(x0[:, :28] - ref["x0"]).max() < 2e-3       # max-pooled fp16 grid
    for tag, D, G, C in (("b0", 24, 1, 96), ("b1", 12, 2, 160), ("b2", 6, 2, 224)):
        got, border = tl.decode_chunk_planar(s.debug_read(tag), n, D, G, C)
        assert border == 0.0
        assert np.abs(got - ref[tag]).max() < 4e-3 * np.abs(ref[tag]).max(), tag


def test_default_ensemble_fast_vs_validation_mode(kat):
    """gnina's default ensemble (2 dense + 1 default2018) entirely on the tensor-core path vs the fp32 kernels"""
    from gnina_b200 import CNNScorer
    a, b = CNNScorer([], precision=0), CNNScorer([], precision=1)
    assert b.get_option("precision") == 1
    for s in (a, b):
        s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    ra = a.score_batch(kat["lig_xyz"], kat["lig_types"], kat["pose_offsets"])
    rb = b.score_batch(kat["lig_xyz"], kat["lig_types"], kat["pose_offsets"])
    assert np.abs(ra[0] - rb[0]).max() < TOL_SCORE and np.abs(ra[1] - rb[1]).max() < TOL_AFF
    assert np.abs(ra[3] - rb[3]).max() < 2e-2      # affinity variance across the three models


def test_full_size_batch_properties():
    """BASELINE config-2 size (10k poses): size-independent properties of the fast path — batch-order invariance
    (bit-exact), duplicate poses score identically, and a random subset agrees with the fp32 validation kernels."""
    from gnina_b200 import CNNScorer, synth
    rx, rt = synth.make_receptor()
    lx0, lt0 = synth.make_ligand()
    n, na = 10000, len(lt0)
    lx, offs = synth.make_poses(lx0, n, seed=1)
    lt = np.tile(lt0, n)
    s = CNNScorer(["crossdock_default2018"], precision=1)
    s.set_receptor(rx, rt)
    a = s.score_batch(lx, lt, offs)
    perm = np.random.RandomState(0).permutation(n)
    lxp = lx.reshape(n, na, 3)[perm].reshape(-1, 3)
    b = s.score_batch(lxp, lt, offs)
    assert np.array_equal(a[0][perm], b[0]) and np.array_equal(a[1][perm], b[1])
    dup = np.concatenate([lx[:na], lx[:na], lx[5 * na:6 * na]])
    d = s.score_batch(dup, lt[:3 * na], offs[:4])
    assert d[0][0] == d[0][1] == a[0][0] and d[0][2] == a[0][5]
    sub = np.sort(perm[:48])
    v = CNNScorer(["crossdock_default2018"], precision=0)
    v.set_receptor(rx, rt)
    ref = v.score_batch(lx.reshape(n, na, 3)[sub].reshape(-1, 3), lt[:48 * na], offs[:49])
    assert np.abs(ref[0] - a[0][sub]).max() < TOL_SCORE and np.abs(ref[1] - a[1][sub]).max() < TOL_AFF
    assert np.isfinite(a[0]).all() and (a[0] >= 0).all() and (a[0] <= 1).all()
    # ... and with the CPU oracle (fp64 run of the same graph on the oracle's grids) on poses spread over the batch,
    # including the last row group (10000 = 1250 groups of 8) and the last chunk
    import torch
    from gnina_b200 import model_blob
    from oracle import pipeline
    pick = np.array([0, 7, 8, 2047, 2048, 4999, 9991, 9999])
    om = pipeline.OracleModel(model_blob.load_model("crossdock_default2018"))
    want = om.score(rx, rt, lx.reshape(n, na, 3)[pick].reshape(-1, 3), lt[:len(pick) * na], offs[:len(pick) + 1], dtype=torch.float64)
    assert np.abs(want[0] - a[0][pick]).max() < TOL_SCORE and np.abs(want[1] - a[1][pick]).max() < TOL_AFF


@pytest.mark.parametrize("n", [1, 7, 9, 17, 100])
def test_fused_conv1_kernel_equals_unfused_path(kat, n):
    """The fused unit1_conv + unit2_conv + pool kernel (row-group tiles, tensor-map TMA, second MMA in the epilogue) and
    the separate kernels (GB_TC_FUSED=0 semantics, reached here through the gradient call that keeps activations)
    compute the same network: scores agree to fp16 round-off for ragged group counts (partial last group of 8)."""
    from gnina_b200 import CNNScorer, synth
    rx, rt = synth.make_receptor(1500, box=44)
    lx0, lt0 = synth.make_ligand(22, 3, seed=4)
    lx, offs = synth.make_poses(lx0, n, trans_box=10, seed=n)
    lt = np.tile(lt0, n)
    s = _fast(["crossdock_default2018"])
    s.set_receptor(rx, rt)
    fused = s.score_batch(lx, lt, offs)
    unfused = s.score_grad_batch(lx, lt, offs)          # forward with kept activations: conv1, pointwise, pool as 3 kernels
    assert np.abs(fused[0] - unfused[0]).max() < 5e-4 and np.abs(fused[1] - unfused[1]).max() < 2e-3
    v = CNNScorer(["crossdock_default2018"], precision=0)
    v.set_receptor(rx, rt)
    ref = v.score_batch(lx, lt, offs)
    assert np.abs(fused[0] - ref[0]).max() < TOL_SCORE and np.abs(fused[1] - ref[1]).max() < TOL_AFF


def test_fused_kernel_variants_agree(kat, monkeypatch):
    """The three builds of the fused unit1_conv + unit2_conv + pool kernel -- two CTAs per SM (0), one CTA per SM with ghost TMEM
    slots (1, default), and the same on CTA pairs with tcgen05 cta_group::2 (2) -- are the same arithmetic: the single-CTA and
    pair variants agree bit for bit, the older kernel to fp16 round-off (it accumulates the wrapped planes in one place)."""
    from gnina_b200 import synth
    rx, rt = synth.make_receptor(1500, box=44)
    lx0, lt0 = synth.make_ligand(22, 3, seed=4)
    n = 77                                     # ragged: 10 groups of 8 (390 items, even) ... and an odd item count below
    lx, offs = synth.make_poses(lx0, n, trans_box=10, seed=3)
    lt = np.tile(lt0, n)
    s = _fast(["crossdock_default2018"])
    s.set_receptor(rx, rt)
    out = {}
    for v in ("1", "2", "0"):
        monkeypatch.setenv("GB_TC_FUSED_V2", v)
        out[v] = s.score_batch(lx, lt, offs)
        k = 5                                  # 1 group: 39 items, the pair variant's odd tail
        out[v + "s"] = s.score_batch(lx[:offs[k]], lt[:offs[k]], offs[:k + 1])
    monkeypatch.delenv("GB_TC_FUSED_V2")
    for suf in ("", "s"):
        assert np.array_equal(out["1" + suf][0], out["2" + suf][0]) and np.array_equal(out["1" + suf][1], out["2" + suf][1])
        assert np.abs(out["1" + suf][0] - out["0" + suf][0]).max() < 5e-5 and np.abs(out["1" + suf][1] - out["0" + suf][1]).max() < 5e-4


def test_conv3_conv5_organisations_agree(kat, monkeypatch):
    """unit3_conv / unit5_conv on conv3_tc_kernel (two CTAs per SM, 8-slot ring) and on conv3_tc_v2_kernel (one CTA per SM,
    slot = plane) issue the same MMAs in the same order: identical scores, ragged pose counts included."""
    from gnina_b200 import synth
    rx, rt = synth.make_receptor(1500, box=44)
    lx0, lt0 = synth.make_ligand(22, 3, seed=4)
    s = _fast(["crossdock_default2018"])
    s.set_receptor(rx, rt)
    for n in (1, 3, 77, 300):
        lx, offs = synth.make_poses(lx0, n, trans_box=10, seed=n)
        lt = np.tile(lt0, n)
        out = {}
        for v in ("0", "3"):
            monkeypatch.setenv("GB_TC_CONV_V2", v)
            out[v] = s.score_batch(lx, lt, offs)
        monkeypatch.delenv("GB_TC_CONV_V2")
        assert np.array_equal(out["0"][0], out["3"][0]) and np.array_equal(out["0"][1], out["3"][1])


def test_dense_ensemble_matches_reference_pt(golden_dir):
    """BASELINE config 4: `--cnn dense_ensemble` = 20 models (15 dense + 5 default2018 architecture), ensemble
    statistics of CNNTorchScorer::score against the reference's own .pt files (tests/golden/ensemble_kat.npz)."""
    from gnina_b200 import CNNScorer
    kat = np.load(os.path.join(golden_dir, "cnn_kat.npz"))
    e = np.load(os.path.join(golden_dir, "ensemble_kat.npz"))
    n = int(e["n_poses"])
    offs = kat["pose_offsets"][:n + 1]
    x, t = kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]]
    for prec, tol_s, tol_a, tol_v in ((1, 2e-3, 1e-2, 2e-2), (0, 2e-5, 1e-4, 5e-4)):
        s = CNNScorer(["dense_ensemble"], precision=prec)
        assert len(s.model_names) == 20
        s.set_receptor(kat["rec_xyz"], kat["rec_types"])
        sc, aff, loss, var = s.score_batch(x, t, offs)
        assert np.abs(sc - e["score"]).max() < tol_s
        assert np.abs(aff - e["affinity"]).max() < tol_a
        assert np.abs(var - e["variance"]).max() < tol_v
        per_model = s.score_batch_models(x, t, offs)
        order = [list(e["models"]).index(m) for m in s.model_names]
        assert np.abs(per_model[0] - e["pose_f64"][order]).max() < tol_s
