"""GPU parity tests (run on the B200 box: pytest -m gpu).  The CUDA path is called through the C ABI
(gnina_b200.capi -> libgnina_b200.so) and compared with the CPU oracle and the committed golden fixtures."""
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "cnn_kat.npz"))


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "gridmaker_golden.npz"))


def _scorer(names, precision):
    from gnina_b200 import CNNScorer
    return CNNScorer(names, precision=precision)


def test_library_loaded_and_device():
    from gnina_b200 import capi
    assert capi.lib().gb_device_count() >= 1


def test_voxeliser_reproduces_reference_golden(gold):
    """default2017's rec/lig maps are exactly test/gninagrid/files/recmap + ligmap, so the CUDA voxeliser can be
    held against the reference's own cc_0.48.35.binmap (comparator tolerance 1e-4, compare_bin.py:24)."""
    s = _scorer(["default2017"], 0)
    xyz = gold["cc_xyz"]
    types = np.array([2, 2], np.int32)
    s.set_receptor(np.round(xyz.astype(np.float64), 3).astype(np.float32), types)
    g = s.voxelize(xyz, types, [0, 2])[0]
    ref = np.zeros(g.size, np.float32)
    ref[gold["cc48_idx"]] = gold["cc48_val"]
    ref = ref.reshape(g.shape)
    assert np.abs(g - ref).max() < 2e-6
    assert np.count_nonzero(g) == np.count_nonzero(ref)


@pytest.mark.parametrize("name", ["crossdock_default2018", "default2017"])
def test_voxeliser_matches_oracle(kat, name):
    from gnina_b200 import model_blob
    from oracle import pipeline
    s = _scorer([name], 0)
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    offs = kat["pose_offsets"][:4]
    lx, lt = kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]]
    g = s.voxelize(lx, lt, offs)
    om = pipeline.OracleModel(model_blob.load_model(name))
    ref = om.grids(kat["rec_xyz"], kat["rec_types"], lx, lt, offs)
    assert g.shape == ref.shape
    assert np.abs(g - ref).max() < 5e-6
    # typing (G0) goes through the library too
    ch, rad = s.type_atoms(lt, True)
    from oracle import gridmaker as gm
    rc, rr = gm.type_atoms(lt, om.lig_t2c, om.n_rec)
    assert np.array_equal(ch, rc) and np.array_equal(rad, rr)
    # ... and its radii are the reference's own table (xs_radius of every smina type as oracle/_ref reports it)
    import os
    ref_tab = np.load(os.path.join(os.path.dirname(__file__), "golden", "vina_ref_kat.npz"))["type_xs_radius"]
    all_t = np.arange(28, dtype=np.int32)
    for is_lig in (False, True):
        ch28, rad28 = s.type_atoms(all_t, is_lig)
        assert np.array_equal(rad28[ch28 >= 0], ref_tab[ch28 >= 0])


@pytest.mark.parametrize("name,tol_p,tol_a", [("crossdock_default2018", 2e-5, 1e-4), ("dense_1_3", 2e-5, 1e-4),
                                              ("default2017", 2e-5, 1e-4),
                                              ("all_default_to_default_1_3_1", 2e-5, 1e-4)])
def test_fp32_scores_match_reference_pt(kat, name, tol_p, tol_a):
    """fp32 validation mode vs the outputs of the reference's own TorchScript model (fp64 run of the .pt)."""
    s = _scorer([name], 0)
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    pose, aff, loss, var = s.score_batch(kat["lig_xyz"], kat["lig_types"], kat["pose_offsets"])
    assert np.abs(pose - kat[name + "_pose_f64"]).max() < tol_p
    assert np.abs(aff - kat[name + "_aff_f64"]).max() < tol_a
    assert np.abs(loss + np.log(kat[name + "_pose_f64"])).max() < 1e-3 * max(1.0, np.abs(loss).max())
    assert np.all(var == 0)


def test_default_ensemble_matches_oracle(kat):
    """default ensemble (dense_1_3, dense_1_3_PT_KD_3, crossdock_default2018_KD_4; cnn_torch_scorer.cpp:33-35):
    mean score/affinity/loss and affinity variance as CNNTorchScorer::score computes them."""
    from oracle import cnn_ref
    s = _scorer([], 0)
    assert s.model_names == ["dense_1_3", "dense_1_3_PT_KD_3", "crossdock_default2018_KD_4"]
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    sc, aff, loss, var = s.score_batch(kat["lig_xyz"], kat["lig_types"], kat["pose_offsets"])
    for i in range(len(sc)):
        ps = [kat[m + "_pose_f64"][i] for m in s.model_names]
        as_ = [kat[m + "_aff_f64"][i] for m in s.model_names]
        es, ea, el, ev = cnn_ref.ensemble(ps, as_, [-np.log(p) for p in ps])
        assert abs(sc[i] - es) < 2e-5 and abs(aff[i] - ea) < 1e-4 and abs(var[i] - ev) < 2e-4
        assert abs(loss[i] - el) < 1e-3 * max(1.0, abs(el))
    pm, am, lm = s.score_batch_models(kat["lig_xyz"], kat["lig_types"], kat["pose_offsets"])
    assert pm.shape == (3, len(sc))
    assert np.abs(pm.mean(0) - sc).max() < 1e-6


def test_edge_cases_ragged_empty_hydrogen_only_and_centers(kat):
    from gnina_b200 import model_blob
    from oracle import pipeline
    name = "crossdock_default2018"
    s = _scorer([name], 0)
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    o = kat["pose_offsets"]
    na = o[1]
    lx, lt = kat["lig_xyz"], kat["lig_types"]
    # ragged: pose0 full, pose1 = first 5 atoms of pose 1, pose2 = empty, pose3 = hydrogens only
    xs = np.concatenate([lx[:na], lx[na:na + 5], np.zeros((0, 3), np.float32), lx[2 * na:2 * na + 3]])
    ts = np.concatenate([lt[:na], lt[na:na + 5], np.zeros(0, np.int32), np.array([1, 1, 0], np.int32)])
    offs = np.array([0, na, na + 5, na + 5, na + 8], np.int32)
    centers = np.array([[0, 0, 0], [1, 2, 3], [0.5, 0.5, 0.5], [-2, 1, 0]], np.float32)
    for c in (centers, None):
        if c is None:  # the empty pose has no ligand centre; give every pose atoms for the NULL-centre case
            offs2 = np.array([0, na, na + 5, na + 8], np.int32)
            got = s.score_batch(xs, ts, offs2, None)
            om = pipeline.OracleModel(model_blob.load_model(name))
            want = om.score(kat["rec_xyz"], kat["rec_types"], xs, ts, offs2, None, torch.float64)
        else:
            got = s.score_batch(xs, ts, offs, c)
            om = pipeline.OracleModel(model_blob.load_model(name))
            want = om.score(kat["rec_xyz"], kat["rec_types"], xs, ts, offs, c, torch.float64)
        assert np.abs(got[0] - want[0]).max() < 2e-5
        assert np.abs(got[1] - want[1]).max() < 1e-4
    assert s.score_batch(np.zeros((0, 3), np.float32), np.zeros(0, np.int32), [0])[0].shape == (0,)


def test_single_pose_score_and_fresh_copy(kat):
    name = "crossdock_default2018"
    s = _scorer([name], 0)
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    o = kat["pose_offsets"]
    one = s.score(kat["lig_xyz"][:o[1]], kat["lig_types"][:o[1]])
    assert abs(one[0] - kat[name + "_pose_f64"][0]) < 2e-5 and abs(one[1] - kat[name + "_aff_f64"][0]) < 1e-4
    c = s.fresh_copy()
    two = c.score(kat["lig_xyz"][:o[1]], kat["lig_types"][:o[1]])
    assert one == two
    # batch > max_batch chunks identically
    s.set_option("max_batch", 4)
    many = s.score_batch(kat["lig_xyz"], kat["lig_types"], o)
    s.set_option("max_batch", 2)
    many2 = s.score_batch(kat["lig_xyz"], kat["lig_types"], o)
    assert np.array_equal(many[0], many2[0]) and np.array_equal(many[1], many2[1])


def test_errors_are_reported_not_fatal():
    from gnina_b200 import CNNScorer, usage_error
    with pytest.raises(usage_error, match="Invalid model name"):
        CNNScorer(["no_such_model"])


def test_cnn_rotation_mechanism(kat):
    """G3 (--cnn_rotation): every (model, rotation) evaluation equals scoring inputs rotated on the host about the grid
    centre with the matrix the library reports; the ensemble statistics run over all model x rotation evaluations.
    (The random STREAM is the library's own -- libmolgrid's generator is not reproducible -- the mechanism is pinned.)"""
    from gnina_b200 import CNNScorer, capi
    names = ["crossdock_default2018", "crossdock_default2018_KD_4"]
    offs = kat["pose_offsets"][:3]
    x, t = kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]]
    rec, rt = kat["rec_xyz"], kat["rec_types"]
    R = 3
    s = CNNScorer(names, precision=0)
    s.set_option("cnn_rotation", R); s.set_option("rotation_seed", 7)
    s.set_receptor(rec, rt)
    per = s.score_batch_models(x, t, offs)
    assert per[0].shape == (len(names) * R, 2)
    ens = s.score_batch(x, t, offs)
    assert np.abs(ens[0] - per[0].mean(0)).max() < 1e-6 and np.abs(ens[1] - per[1].mean(0)).max() < 1e-5
    assert np.abs(ens[3] - per[1].var(0)).max() < 1e-4 and np.abs(ens[2] - per[2].mean(0)).max() < 1e-5
    plain = CNNScorer(names, precision=0)
    plain.set_receptor(rec, rt)
    base = plain.score_batch_models(x, t, offs)
    for m in range(len(names)):
        assert np.array_equal(per[0][m * R], base[0][m])                      # rotation 0 = the unrotated evaluation
    seen = []
    for r in (1, 2):
        for p in (0, 1):
            M = s.rotation(r, p)
            assert np.abs(M @ M.T - np.eye(3)).max() < 1e-6 and abs(np.linalg.det(M) - 1) < 1e-6
            assert np.abs(M - np.eye(3)).max() > 1e-2
            seen.append(M)
            lp = x[offs[p]:offs[p + 1]].astype(np.float64)
            c = x[offs[p]:offs[p + 1]].mean(0, dtype=np.float32).astype(np.float64)
            q = CNNScorer(names, precision=0)
            q.set_receptor((c + (rec.astype(np.float64) - c) @ M.T.astype(np.float64)).astype(np.float32), rt)
            out = q.score_batch_models((c + (lp - c) @ M.T.astype(np.float64)).astype(np.float32), t[offs[p]:offs[p + 1]],
                                       [0, offs[p + 1] - offs[p]], [c.astype(np.float32)])
            for m in range(len(names)):
                assert abs(out[0][m, 0] - per[0][m * R + r, p]) < 5e-5, (r, p, m)
                assert abs(out[1][m, 0] - per[1][m * R + r, p]) < 2e-4, (r, p, m)
    assert np.abs(seen[0] - seen[1]).max() > 1e-2 and np.abs(seen[0] - seen[2]).max() > 1e-2   # per pose, per rotation
    assert np.array_equal(s.rotation(0, 1), np.eye(3, dtype=np.float32))
    # fast mode runs the same rotations
    f = CNNScorer(names, precision=1)
    f.set_option("cnn_rotation", R); f.set_option("rotation_seed", 7)
    f.set_receptor(rec, rt)
    pf = f.score_batch_models(x, t, offs)
    assert np.abs(pf[0] - per[0]).max() < 2e-3 and np.abs(pf[1] - per[1]).max() < 1e-2
    c2 = f.fresh_copy()
    assert np.array_equal(c2.score_batch(x, t, offs)[0], f.score_batch(x, t, offs)[0])          # clones keep the option
    with pytest.raises(capi.GbError, match="out of range"):
        f.set_option("cnn_rotation", 25)
    # gradients: the rotated evaluation's gradient is taken in the rotated frame and rotated back (Transform::backward,
    # torch_model.cpp:204-206); the result is the mean over model x rotation evaluations (cnn_torch_scorer.cpp:164-179)
    one = [names[0]]
    g2 = CNNScorer(one, precision=0)
    g2.set_option("cnn_rotation", 2); g2.set_option("rotation_seed", 11)
    g2.set_receptor(rec, rt)
    xp, tp = x[:offs[1]], t[:offs[1]]
    c = xp.mean(0, dtype=np.float32).astype(np.float64)
    M = g2.rotation(1, 0).astype(np.float64)
    got = g2.score_grad_batch(xp, tp, [0, len(tp)], [c.astype(np.float32)], receptor=True)
    g0 = CNNScorer(one, precision=0)
    g0.set_receptor(rec, rt)
    a0 = g0.score_grad_batch(xp, tp, [0, len(tp)], [c.astype(np.float32)], receptor=True)
    g1 = CNNScorer(one, precision=0)
    g1.set_receptor((c + (rec.astype(np.float64) - c) @ M.T).astype(np.float32), rt)
    a1 = g1.score_grad_batch((c + (xp.astype(np.float64) - c) @ M.T).astype(np.float32), tp, [0, len(tp)], [c.astype(np.float32)],
                             receptor=True)
    for k in (4, 5):                                   # ligand and receptor gradients
        want = 0.5 * (a0[k] + a1[k].astype(np.float64) @ M)          # row vectors: R^T g' = g' R
        assert np.abs(got[k] - want).max() < 2e-4 * max(1e-3, np.abs(want).max()), k
    assert abs(got[2][0] - 0.5 * (a0[2][0] + a1[2][0])) < 1e-5


@pytest.mark.parametrize("prec", [0, 1])
def test_empty_grid_reproduces_the_reference_zero_grid_answer(kat, prec):
    """SURVEY.md 8c known answer of the reference's own crossdock_default2018.pt on an all-zero grid: pose 0.92233,
    affinity 1.46472.  Reached three ways through the whole pipeline: no receptor at all + untyped ligand, a grid
    centred far away from every atom, and a receptor replaced by hydrogens."""
    from gnina_b200 import CNNScorer
    name = "crossdock_default2018"
    zp = float(np.exp(kat[name + "_zero_logp_f64"][0, 1])), float(kat[name + "_zero_aff_f64"].reshape(-1)[0])
    assert abs(zp[0] - 0.92233) < 1e-5 and abs(zp[1] - 1.46472) < 1e-5
    tp, ta = (2e-5, 1e-4) if prec == 0 else (2e-3, 1e-2)
    o = kat["pose_offsets"]
    x, t = kat["lig_xyz"][:o[2]], kat["lig_types"][:o[2]]
    s = CNNScorer([name], precision=prec)
    # (1) receptor with zero atoms, ligand of hydrogens only
    s.set_receptor(np.zeros((0, 3), np.float32), np.zeros(0, np.int32))
    got = s.score_batch(x, np.ones_like(t), o[:3])
    assert np.abs(got[0] - zp[0]).max() < tp and np.abs(got[1] - zp[1]).max() < ta
    # (2) real receptor and ligand, grid centred 500 A away
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    got = s.score_batch(x, t, o[:3], np.full((2, 3), 500.0, np.float32))
    assert np.abs(got[0] - zp[0]).max() < tp and np.abs(got[1] - zp[1]).max() < ta
    # (3) back to a normal call on the same handle: nothing stale
    ref = CNNScorer([name], precision=prec)
    ref.set_receptor(kat["rec_xyz"], kat["rec_types"])
    assert np.array_equal(s.score_batch(x, t, o[:3])[0], ref.score_batch(x, t, o[:3])[0])
    # (4) receptor of hydrogens only
    s.set_receptor(kat["rec_xyz"], np.ones_like(kat["rec_types"]))
    got = s.score_batch(x, np.zeros_like(t), o[:3])
    assert np.abs(got[0] - zp[0]).max() < tp and np.abs(got[1] - zp[1]).max() < ta


def test_chunk_boundaries_and_big_ligands(kat):
    """a batch that crosses the default 2048-pose device chunk by one pose, and a 150-heavy-atom ligand (larger than
    any list floor): identical to scoring the same poses alone"""
    from gnina_b200 import CNNScorer, synth
    s = CNNScorer(["crossdock_default2018"], precision=1)
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    xyz, types, offs = synth.make_screen(2049, seed=9, trans_box=6.0)
    big = s.score_batch(xyz, types, offs)
    for p in (0, 2047, 2048):
        one = s.score_batch(xyz[offs[p]:offs[p + 1]], types[offs[p]:offs[p + 1]], [0, offs[p + 1] - offs[p]])
        assert one[0][0] == big[0][p] and one[1][0] == big[1][p]
    lx, lt = synth.make_ligand(150, 10, seed=3)
    v0 = CNNScorer(["crossdock_default2018"], precision=0)
    v0.set_receptor(kat["rec_xyz"], kat["rec_types"])
    a = s.score_batch(lx, lt, [0, len(lt)])
    b = v0.score_batch(lx, lt, [0, len(lt)])
    assert abs(a[0][0] - b[0][0]) < 3e-3 and abs(a[1][0] - b[1][0]) < 1e-2 * max(1.0, abs(b[1][0]))   # fast-mode tolerance


@pytest.mark.gpu
@pytest.mark.parametrize("ens,n", [("crossdock_default2018_ensemble", 15), ("general_default2018_ensemble", 10), ("redock_default2018_ensemble", 15)])
def test_every_packaged_default2018_ensemble_matches_the_oracle(kat, ens, n):
    """All 64 embedded models of the reference are packaged: the ensemble names expand over the reference's table
    (cnn_torch_scorer.cpp:37-60) and every member's fast-path result agrees with the CPU restatement on its own blob."""
    from gnina_b200 import model_blob
    from oracle import pipeline
    s = _scorer([ens], 1)
    assert len(s.model_names) == n
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    k = 3
    offs = kat["pose_offsets"][:k + 1]
    lx, lt = kat["lig_xyz"][:offs[k]], kat["lig_types"][:offs[k]]
    sc, aff, loss, var = s.score_batch(lx, lt, offs)
    pm, am, lm = s.score_batch_models(lx, lt, offs)
    assert pm.shape == (n, k) and np.abs(pm.mean(0) - sc).max() < 1e-6 and np.abs(am.mean(0) - aff).max() < 1e-5
    for mi in (0, n - 1):   # first and last member against the oracle
        om = pipeline.OracleModel(model_blob.load_model(s.model_names[mi]))
        po, ao, _ = om.score(kat["rec_xyz"], kat["rec_types"], lx, lt, offs)
        assert np.abs(pm[mi] - po).max() < 2e-3 and np.abs(am[mi] - ao).max() < 1e-2
