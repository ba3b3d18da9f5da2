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


class _OracleNetwork:
    """the CNNScorer interface minimize.cnn_energy uses, served by oracle/pipeline.py on the CPU"""

    def __init__(self, names, rx, rt):
        self.oms = [pipeline.OracleModel(model_blob.load_model(n)) for n in names]
        self.rx, self.rt, self.calls = rx, rt, 0

    def model_info(self, i):
        class Info:
            dimension, resolution = 23.5, 0.5
        return Info

    def score_grad_batch(self, xyz, types, offs):
        self.calls += 1
        return pipeline.score_grad(self.oms, self.rx, self.rt, xyz, types, offs, dtype=torch.float32)


def _host_libm():
    import ctypes
    m = ctypes.CDLL("libm.so.6")

    def wrap(name):
        f = getattr(m, name); f.argtypes = [ctypes.c_float]; f.restype = ctypes.c_float
        return lambda a: np.array([f(float(v)) for v in np.asarray(a, np.float32).ravel()], np.float32).reshape(np.shape(a))
    return wrap("sinf"), wrap("cosf"), wrap("acosf")


@pytest.mark.parametrize("hydrogens", [False, True])
def test_config5_minimisation_with_real_networks(hydrogens):
    """BASELINE config 5 end to end on the CPU: the REFERENCE's quasi_newton (accurate line search, what --minimize selects) over its
    non_cache_cnn over its CNNTorchScorer running the real crossdock_default2018 TorchScript file, one pose at a time -- against this
    repo's lock-step minimiser (gnina_b200/minimize.py: minimize_poses + cnn_energy) over the restated network, all poses in one batch.
    Energies and conformations are EQUAL.  With hydrogens among the ligand atoms that needs the reference's force routing
    (add_minus_forces consumes getGradient's by-atom list compactly, lib/model.cu:247-259): with the true per-atom gradient the
    minimiser ends elsewhere -- at lower energies, which is what one expects of the correct gradient."""
    from gnina_b200 import minimize as M
    begin, end, nn = [-9.7] * 3, [10.55] * 3, [54] * 3
    lig = dict(synth.make_flexible_ligand(n_heavy=14, n_tors=3, n_branch=2, seed=8))
    if hydrogens:
        ty = lig["types"].copy(); ty[3] = 1; ty[8] = 0; lig["types"] = ty
    rx, rt = synth.make_receptor(400, box=24, seed=5)
    sf, rm = R.RefScoring(), R.RefModel(lig, rx, rt)
    lig2 = dict(lig); lig2["local_xyz"], lig2["seg_rel_origin"], lig2["seg_rel_axis"] = rm.export()
    tree = M.TorsionTree(lig2)
    rs = np.random.RandomState(1)
    X = np.tile(lig["conf0"], (3, 1)).astype(np.float32)
    X[:, :3] += rs.uniform(-1, 1, (3, 3)); X[:, 7:] = rs.uniform(-1, 1, (3, X.shape[1] - 7))
    X = X.astype(np.float32)
    s = CR.RefCNNScorer(names=["crossdock_default2018"])
    ref = [R.minimize_dl(rm, sf, R.LINEAR, begin, end, nn, x, 4, s.dl(), accurate=True) for x in X]
    M.set_transcendentals(*_host_libm())                 # the reference build executes this host's sinf / cosf / acosf
    try:
        net = _OracleNetwork(["crossdock_default2018"], rx, rt)
        energy = M.cnn_energy(net, lig["types"], (np.float32(begin), np.float32(end)), slope=10.0)
        e, x, ev, rounds = M.minimize_poses(tree, energy, X, maxiters=4, accurate=True)
        for i in range(len(X)):
            assert float(e[i]) == ref[i][0] and np.array_equal(x[i], ref[i][1])
        assert net.calls == rounds + 1 and net.calls < ev.sum()          # one batched network call per round
        if hydrogens:
            net2 = _OracleNetwork(["crossdock_default2018"], rx, rt)
            true_grad = M.cnn_energy(net2, lig["types"], (np.float32(begin), np.float32(end)), slope=10.0, reference_force_routing=False)
            e2, x2, _, _ = M.minimize_poses(tree, true_grad, X, maxiters=4, accurate=True)
            assert not np.array_equal(x2, x) and e2.sum() < e.sum()
    finally:
        M.set_transcendentals()


@pytest.mark.parametrize("fresh_copy", [False, True])
def test_the_integration_adapter_is_a_drop_in_for_cnn_torch_scorer(fresh_copy):
    """integration/cnn_b200_scorer.h (CNNB200Scorer : DLScorer -- the class a gnina maintainer adds) EXECUTED, not only compiled: its
    twelve C-ABI calls are served by a stand-in that honours include/gnina_b200.h's contract with the reference's own TorchModel as the
    network, so that the adapter's code -- name resolution against the packaged blobs, setLigand / setReceptor reuse, receptor upload,
    centre option, by-atom gradient scatter + add_minus_forces, set_center_from_model / set_bounding_box, fresh_copy -- runs inside the
    reference's classes.  Scores, forces (hydrogens among the ligand atoms), boxes and a whole quasi_newton + non_cache_cnn minimisation
    equal CNNTorchScorer's."""
    lig = dict(synth.make_flexible_ligand(n_heavy=14, n_tors=3, n_branch=2, seed=8))
    ty = lig["types"].copy(); ty[3] = 1; ty[8] = 0; lig["types"] = ty
    rx, rt = synth.make_receptor(400, box=24, seed=5)
    sf, rm = R.RefScoring(), R.RefModel(lig, rx, rt)
    x = lig["conf0"].copy(); x[:3] += [0.4, -0.2, 0.3]
    for names, kw in (([], {}), (["fast"], {}), (["crossdock_default2018"], dict(cnn_center=[0.5, 0.25, -0.5])), (["default2017"], {})):
        ref = CR.RefCNNScorer(names=names, **kw)
        mine = CR.RefCNNScorer.adapter(names=names, fresh_copy=fresh_copy, **kw)
        rm.set(x)
        for grad in (True, False):
            a, b = ref.score(rm, grad), mine.score(rm, grad)
            assert a[:4] == b[:4] and np.array_equal(a[4], b[4]), (names, grad)
        ca, cb = ref.center_and_box(rm), mine.center_and_box(rm)
        assert all(np.array_equal(p, q) for p, q in zip(ca, cb))
    with pytest.raises(RuntimeError, match="Invalid model name"):
        CR.RefCNNScorer.adapter(names=["nonesuch"])
    begin, end, nn = [-9.7] * 3, [10.55] * 3, [54] * 3
    ref, mine = CR.RefCNNScorer(names=["crossdock_default2018"]), CR.RefCNNScorer.adapter(names=["crossdock_default2018"], fresh_copy=fresh_copy)
    ea, xa = R.minimize_dl(rm, sf, R.LINEAR, begin, end, nn, x, 3, ref.dl(), accurate=True)
    eb, xb = R.minimize_dl(rm, sf, R.LINEAR, begin, end, nn, x, 3, mine.dl(), accurate=True)
    assert ea == eb and np.array_equal(xa, xb)


def test_cnn_refinement_with_real_networks():
    """--cnn_scoring refinement: refine_structure (main/main.cpp:131-171) on ig = non_cache_cnn over the reference's CNNTorchScorer (real
    crossdock_default2018 file; fast line search, ssd_par's iteration count), pose by pose, vs minimize.refine_structure_poses +
    cnn_energy over the restated network for all poses at once: energies, conformations and the `within` flag are EQUAL"""
    from gnina_b200 import minimize as M
    begin, end, nn = [-6.1] * 3, [6.3] * 3, [34] * 3
    lig = dict(synth.make_flexible_ligand(n_heavy=14, n_tors=3, n_branch=2, seed=8))
    ty = lig["types"].copy(); ty[3] = 1; lig["types"] = ty
    rx, rt = synth.make_receptor(400, box=24, seed=5)
    sf, rm = R.RefScoring(), R.RefModel(lig, rx, rt)
    lig2 = dict(lig); lig2["local_xyz"], lig2["seg_rel_origin"], lig2["seg_rel_axis"] = rm.export()
    tree = M.TorsionTree(lig2)
    heavy = np.asarray(lig["types"]) >= 2
    rs = np.random.RandomState(5)
    X = np.tile(lig["conf0"], (3, 1)).astype(np.float32)
    X[:, :3] += rs.uniform(-3, 3, (3, 3)); X[:, 7:] = rs.uniform(-1, 1, (3, X.shape[1] - 7))
    X = X.astype(np.float32)
    maxit = (25 + len(lig["types"])) // 3
    s = CR.RefCNNScorer(names=["crossdock_default2018"])
    ref = [R.refine_dl(rm, sf, R.LINEAR, begin, end, nn, x, maxit, s.dl()) for x in X]
    M.set_transcendentals(*_host_libm())
    try:
        net = _OracleNetwork(["crossdock_default2018"], rx, rt)
        centers = M.heavy_centers(tree.set_conf(X)[0], heavy)
        half = np.float32(23.5) / np.float32(2)
        within = M.within_boxes(heavy, [(centers - half, centers + half), (np.float32(begin), np.float32(end))])
        e, x, inside, ev = M.refine_structure_poses(
            tree, lambda slope: M.cnn_energy(net, lig["types"], (np.float32(begin), np.float32(end)), slope=slope, cnn_center=centers),
            within, X, maxit)
        for i in range(len(X)):
            assert float(e[i]) == ref[i][0] and np.array_equal(x[i], ref[i][1]) and bool(inside[i]) == ref[i][2], i
    finally:
        M.set_transcendentals()


@pytest.mark.parametrize("name", ["crossdock_default2018", "default2017", "dense"])
def test_receptor_of_every_smina_type(name):
    """setReceptor (lib/dl_scorer.cpp:93-193) hands EVERY fixed atom to the typer -- hydrogens, metals, the generic types of the
    gninacheck generator (test/gnina/test_utils.cpp:13-44), overlapping atoms; the model file's recmap decides which of them get a
    channel.  Outputs and forces equal the restatement's."""
    rs = np.random.RandomState(77)
    rx, rt = synth.make_gninacheck_mol(rs, 0, 300, 400, 11, 11, 11)
    assert len(set(rt.tolist())) >= 26
    lig = synth.make_flexible_ligand(n_heavy=16, n_tors=3, n_branch=2, seed=13)
    rm = R.RefModel(lig, rx, rt)
    xyz = rm.set(lig["conf0"])
    r = CR.RefCNNScorer(names=[name]).score(rm, True)
    o, g = _oracle([name], rx, rt, xyz, lig["types"])
    assert r[:4] == o and np.array_equal(r[4], g)


def test_every_packaged_model_equals_its_torchscript_source():
    """all built-in models (gnina_b200/weights/*.gbw, converted by tools/extract_models.py) against the TorchScript files the reference
    embeds, each run by the reference's own TorchModel: metadata (grid dimension / resolution, both type maps, head flags) and weights
    of every blob reproduce the source's score, affinity and loss on a pose"""
    lig, rx, rt, rm, xyz = _case(31)
    names = scorer.builtin_models()
    assert len(names) >= 64
    worst = 0.0
    for name in names:
        s = CR.RefCNNScorer(names=[name])
        r = s.score(rm, False)
        blob = model_blob.load_model(name)
        assert s.grid() == (np.float32(blob.dimension), np.float32(blob.resolution))
        o, _ = _oracle_score_only(blob, rx, rt, xyz, lig["types"])
        assert np.allclose(r[:3], o, rtol=1e-6, atol=1e-7), (name, r[:3], o)
        worst = max(worst, float(np.abs(np.float32(r[:3]) - np.float32(o)).max()))
    assert worst <= 1e-5


def _oracle_score_only(blob, rx, rt, xyz, types):
    om = pipeline.OracleModel(blob)
    p, a, l = om.score(rx, rt, xyz, types, np.array([0, len(types)], np.int32))
    return (float(p[0]), float(a[0]), float(l[0])), None


def test_product_cpp_classes_are_non_cache_cnn():
    """S3 with the real network: the product's own C++ host classes -- gb::CNNScorer + gb::NonCacheCNN (include/gnina_b200.hpp), their
    C-ABI calls served by the stand-in over the reference's TorchModel -- against the reference's non_cache_cnn::eval / eval_deriv over
    its CNNTorchScorer, on poses inside and partly outside a tight search box, hydrogens among the ligand atoms: energy (loss + both
    out-of-box penalties) and the forces the minimiser sees are EQUAL; with the routing switched off the forces are the by-atom ones"""
    begin, end, nn = [-3.1] * 3, [3.3] * 3, [18] * 3
    lig = dict(synth.make_flexible_ligand(n_heavy=16, n_tors=3, n_branch=2, seed=13))
    ty = lig["types"].copy(); ty[2] = 1; ty[9] = 0; lig["types"] = ty
    rx, rt = synth.make_receptor(400, box=24, seed=5)
    sf, rm = R.RefScoring(), R.RefModel(lig, rx, rt)
    s = CR.RefCNNScorer(names=["crossdock_default2018"])
    rs = np.random.RandomState(9)
    outside = 0
    for k in range(4):
        x = lig["conf0"].copy(); x[:3] += rs.uniform(-2.5, 2.5, 3)
        xyz = rm.set(x.astype(np.float32))
        outside += bool((np.abs(xyz[ty > 1]) > 3.3).any())
        for deriv in (True, False):
            e, f, c = R.noncache_dl_eval(rm, sf, R.LINEAR, begin, end, nn, s.dl(), slope=10.0, deriv=deriv)
            rm.set(x.astype(np.float32))
            e2, f2 = CR.product_noncache_cnn(["crossdock_default2018"], rm, begin, end, nn, c, slope=10.0, deriv=deriv)
            assert e == e2 and (not deriv or np.array_equal(f, f2)), (k, deriv)
        _, f3 = CR.product_noncache_cnn(["crossdock_default2018"], rm, begin, end, nn, c, slope=10.0, reference_force_routing=False)
        assert not np.array_equal(f3, f2)
    assert outside >= 2


def test_config5_through_the_product_cpp_host_code():
    """the C++ side of config 5 -- gb::CNNScorer, gb::LigandTree (topology through b200::B200Ligand), gb::CnnBatchEnergy and
    gb::minimize_poses of include/gnina_b200_minimize.hpp, their C-ABI calls served by the stand-in over the reference's TorchModel -- on
    all poses at once vs the reference's quasi_newton + non_cache_cnn + CNNTorchScorer pose by pose (hydrogens among the ligand atoms, so
    the force routing of gb::CnnBatchEnergy is exercised): energies and conformations are EQUAL"""
    begin, end, nn = [-9.7] * 3, [10.55] * 3, [54] * 3
    lig = dict(synth.make_flexible_ligand(n_heavy=14, n_tors=3, n_branch=2, seed=8))
    ty = lig["types"].copy(); ty[3] = 1; ty[8] = 0; lig["types"] = ty
    rx, rt = synth.make_receptor(400, box=24, seed=5)
    sf, rm = R.RefScoring(), R.RefModel(lig, rx, rt)
    rs = np.random.RandomState(1)
    X = np.tile(lig["conf0"], (3, 1)).astype(np.float32)
    X[:, :3] += rs.uniform(-1, 1, (3, 3)); X[:, 7:] = rs.uniform(-1, 1, (3, X.shape[1] - 7))
    X = X.astype(np.float32)
    s = CR.RefCNNScorer(names=["crossdock_default2018"])
    for accurate, iters in ((True, 4), (False, 6)):
        ref = [R.minimize_dl(rm, sf, R.LINEAR, begin, end, nn, x, iters, s.dl(), accurate=accurate) for x in X]
        e, x, ev, rounds = CR.product_lockstep_minimize(["crossdock_default2018"], rm, begin, end, X, iters, accurate=accurate)
        for i in range(len(X)):
            assert float(e[i]) == ref[i][0] and np.array_equal(x[i], ref[i][1]), (accurate, i)
        assert rounds < ev.sum()


def test_adapter_expands_prefix_ensembles_like_the_reference():
    """`<prefix>_ensemble` through the adapter (gb::expand_model_names over the packaged blobs, sorted) vs CNNTorchScorer's walk over its
    name table (unordered): the same members -- equal mean and variance up to the association of the float sums"""
    lig, rx, rt, rm, xyz = _case(41)
    for ens in ("general_default2018_ensemble", "dense_1_3_ensemble"):
        a = CR.RefCNNScorer(names=[ens]).score(rm, False)
        b = CR.RefCNNScorer.adapter(names=[ens]).score(rm, False)
        assert a[3] > 0 and np.allclose(a[:4], b[:4], rtol=2e-6, atol=1e-7), (ens, a[:4], b[:4])
