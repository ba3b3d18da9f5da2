"""The C restatement of the Vina rows (oracle/vina_ref.c, vina_mc_ref.c) against KNOWN ANSWERS OF THE REFERENCE'S OWN CODE:
tests/golden/vina_ref_kat.npz was produced by oracle/_ref = the reference's sources compiled where they lie under
/root/reference (tests/golden/make_vina_ref_golden.py, oracle/Makefile.ref).  This is the pin of the Vina half of the oracle;
it runs on any box (the fixture travels, the reference does not).

Bit-identity is asserted with gvo_use_libm(1): the reference executes this host's sinf / cosf / expf, and the fixture was
generated in this same image.  The default mode of the restatement (correctly rounded transcendentals, what the device kernels
reproduce) is held to north_star's 1e-6."""
import os
import numpy as np
import pytest
from oracle.vina import VinaOracle, lib as vlib
from oracle.vina_mc import DockOracle

KAT = os.path.join(os.path.dirname(__file__), "golden", "vina_ref_kat.npz")


@pytest.fixture(scope="module")
def kat():
    return np.load(KAT)


@pytest.fixture(scope="module")
def vo():
    return VinaOracle()


def _lig(k):
    lig = {q: k["lig_" + q] for q in ("xyz0", "types", "seg_parent", "seg_begin", "seg_end", "axis_root", "pair_a", "pair_b", "conf0",
                                      "local_xyz", "seg_rel_origin", "seg_rel_axis")}
    lig["gyration_radius"] = 0.0
    return lig


@pytest.fixture(scope="module")
def dock(kat, vo):
    lig = _lig(kat)
    needed = sorted(set(int(t) for t in lig["types"] if t > 1))
    grids = {t: vo.cache_populate(kat["begin"], kat["end"], kat["n"], kat["rec_xyz"], kat["rec_types"], t) for t in needed}
    return DockOracle(vo, grids, kat["begin"], kat["end"], kat["n"], lig, slope=1e3), lig


@pytest.fixture()
def libm():
    vlib().gvo_use_libm(1)
    yield
    vlib().gvo_use_libm(0)


def test_smina_type_table(kat):
    """G0: names, xs_radius and the donor / acceptor / hydrophobe flags of the 28 smina types (lib/atom_constants.h:101-133) in the
    voxeliser oracle's typing and in the Vina oracle's terms"""
    import ctypes as C
    from oracle import gridmaker as gm
    for t in range(28):
        assert gm.lib().gbo_smina_name(t).decode() == str(kat["type_names"][t])
        assert np.float32(gm.lib().gbo_smina_radius(t)) == kat["type_xs_radius"][t]
        r, h, d, a = C.c_float(), C.c_int(), C.c_int(), C.c_int()
        vlib().gvo_type_props.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        vlib().gvo_type_props(t, C.byref(r), C.byref(h), C.byref(d), C.byref(a))
        f = int(kat["type_flags"][t])
        assert np.float32(r.value) == kat["type_xs_radius"][t] and (h.value, d.value, a.value) == (f & 1, (f >> 1) & 1, (f >> 2) & 1)


def test_terms_and_tables_are_bit_identical(kat, vo):
    """V1 weighted_terms::eval_fast, V2 precalculate_linear (eval_fast and eval_deriv), precalculate_exact"""
    w = np.array([-0.035579, -0.005156, 0.840245, -0.035069, -0.587439], np.float32)
    from oracle.vina import _f
    t, r, r2 = kat["tab_t"], kat["tab_r"], kat["tab_r2"]
    terms = np.float32([vlib().gvo_eval_terms(_f(w), int(a), int(b), float(x)) for (a, b), x in zip(t, r)])
    assert np.array_equal(terms, kat["tab_terms"])
    lin = np.float32([vo.eval_deriv(int(a), int(b), float(x)) for (a, b), x in zip(t, r2)])
    assert np.array_equal(lin, kat["tab_linear"])
    fast = np.float32([vo.eval_fast(int(a), int(b), float(x)) for (a, b), x in zip(t, r2)])
    assert np.array_equal(fast, kat["tab_linear_fast"])
    ex = np.float32([vo.exact(int(a), int(b), float(x)) for (a, b), x in zip(t, r2)])
    assert np.array_equal(ex, kat["tab_exact"])


def test_splines_agree_to_round_off(kat, vo):
    """V3: the restatement solves the clamped-spline system with the Thomas algorithm in double, the reference inverts the dense
    matrix in float (Eigen, lib/splines.h:44-79): same spline, different round-off"""
    sp = np.float32([vo.spline_eval_deriv(int(a), int(b), float(x)) for (a, b), x in zip(kat["tab_t"], kat["tab_r2"])])
    ref = kat["tab_splines"]
    assert np.abs(sp[:, 0] - ref[:, 0]).max() <= 2e-6
    assert np.abs(sp[:, 1] - ref[:, 1]).max() <= 1e-4 * max(1.0, np.abs(ref[:, 1]).max())


def test_cache_populate_is_bit_identical(kat, vo):
    """V4 cache::populate: every point of three 41^3 affinity grids"""
    for t, g in zip(kat["grid_types"], kat["grids"]):
        mine = vo.cache_populate(kat["begin"], kat["end"], kat["n"], kat["rec_xyz"], kat["rec_types"], int(t))
        assert np.array_equal(mine, g)


def test_set_conf_and_eval_deriv(kat, dock, libm):
    """V5-V8: model::set and model::eval_deriv (grid term + curl, intramolecular pairs, tree derivative): bit-identical"""
    d, lig = dock
    for i, x in enumerate(kat["confs"]):
        assert np.array_equal(d.coords(x), kat["coords"][i])
        for name, caps in (("full", (1000, 1000, 1000)), ("hunt", (10, 1.5, 10))):
            e, g = d.eval_deriv(x, caps)
            assert e == kat["e_" + name][i] and np.array_equal(g, kat["g_" + name][i])


def test_eval_deriv_with_correctly_rounded_transcendentals(kat, dock):
    """the default mode (what the device reproduces) differs from this host's libm in the last bit of ~1 % of the sines: atoms move by
    <= 1e-5 A; energies agree to 1e-6 (north_star) for poses inside the grid and to 3e-5 where atoms are outside (out-of-box
    penalty = 1e3 x distance, main/main.cpp:466)"""
    d, lig = dock
    inside = 0
    for i, x in enumerate(kat["confs"]):
        c = d.coords(x)
        assert np.abs(c - kat["coords"][i]).max() <= 2e-5
        e, g = d.eval_deriv(x)
        within = bool(((c > kat["begin"]) & (c < kat["end"])).all())
        inside += within
        assert abs(e - kat["e_full"][i]) <= (1e-6 if within else 3e-5) * max(1.0, abs(kat["e_full"][i]))
        assert np.abs(g - kat["g_full"][i]).max() <= 1e-5 * max(1.0, np.abs(kat["g_full"][i]).max())
    assert inside >= 10


def test_non_cache_eval_deriv_and_within(kat, dock, libm):
    """non_cache::eval_deriv (lib/non_cache.cpp:126-174) and non_cache::within, the field of refine_structure"""
    d, lig = dock
    try:
        for slope in (10.0, 1000.0):
            d.use_noncache(kat["rec_xyz"], kat["rec_types"]); d.set_box(kat["begin"], kat["end"], slope)
            for i, x in enumerate(kat["nc_confs"]):
                e, g = d.eval_deriv(x)
                assert e == kat["nc_e_%d" % slope][i] and np.array_equal(g, kat["nc_g_%d" % slope][i])
                assert d.within(x) == bool(kat["nc_within"][i])
    finally:
        d.use_noncache(None); d.set_box(None)
    assert kat["nc_within"].any() and not kat["nc_within"].all()


def test_non_cache_eval_is_bit_identical(kat, vo):
    """non_cache::eval (lib/non_cache.cpp:52-83), the intermolecular energy behind the docking branch's printed Affinity
    (main/main.cpp:340-344): table terms (eval_fast), box clamp, slope x distance outside"""
    for slope in (10.0, 1000.0):
        for i, c in enumerate(kat["nc_coords"]):
            e = vo.noncache_eval(kat["rec_xyz"], kat["rec_types"], c, kat["lig_types"], kat["begin"], kat["end"], slope, 1000.0)
            assert e == kat["nc_eval_%d" % slope][i]


def test_bfgs_is_bit_identical(kat, dock, libm):
    """V9 quasi_newton / bfgs.h with the fast line search: energies and conformations after 3 and 12 iterations"""
    d, lig = dock
    for it in (3, 12):
        for name, caps in (("full", (1000, 1000, 1000)), ("hunt", (10, 10, 10))):
            for i, x in enumerate(kat["confs"][:32]):
                e, xo, g, ne = d.bfgs(x, it, caps)
                assert e == kat["bfgs%d_%s_e" % (it, name)][i] and np.array_equal(xo, kat["bfgs%d_%s_x" % (it, name)][i])


def test_minimize_flavours_are_bit_identical(kat, dock, libm):
    """V9 for --minimize: accurate_line_search (bfgs.h:107-180, after Numerical Recipes' lnsrch, with its float / double mix) and
    --minimize_early_term (:455-462)"""
    d, lig = dock
    for tag, acc, et, it in (("acc", True, False, 30), ("acc_et", True, True, 300), ("fast_et", False, True, 60)):
        for i, x in enumerate(kat["confs"][:32]):
            e, xo, g, ne = d.bfgs(x, it, accurate=acc, early_term=et)
            assert e == kat["min_%s_e" % tag][i] and np.array_equal(xo, kat["min_%s_x" % tag][i])


def test_exact_scoring(kat, vo, dock, libm):
    """V12: naive_non_cache::eval with precalculate_exact and num_tors_div bit-identical; the printed Affinity = eval_adjusted is
    conf_independent((inter + intra) - intra) in float, the restatement's conf_independent(inter): 2e-7"""
    d, lig = dock
    for i, x in enumerate(kat["confs"][:32]):
        e = vo.naive_exact(kat["rec_xyz"], kat["rec_types"], d.coords(x), lig["types"], 1000.0)
        assert e == kat["exact_inter"][i]
        assert vo.num_tors_div(e, float(kat["num_tors"][i])) == kat["num_tors_div"][i]
        aff = kat["exact_intra_affinity"][i, 1]                 # hand-built model: no bond graph, num_tors 0
        assert abs(vo.num_tors_div(e, 0.0) - aff) <= 5e-7 * max(1.0, abs(aff))


def test_monte_carlo_chains_are_bit_identical(kat, dock, libm):
    """V10 monte_carlo::operator() + mutate_conf + metropolis + add_to_output_container: WHOLE CHAINS.  The reference's code ran on
    the restatement's generator (oracle/ref_shim/boost/random.hpp); the start conformation is the reference's own draw.  The
    restatement runs in model-state mode (vina_mc_ref.c mc_impl): the final containers -- energies and conformations -- coincide."""
    d, lig = dock
    steps, maxit, S = (int(v) for v in kat["mc_params"])
    for c in range(len(kat["mc_seeds"])):
        e, x = d.mc_ex(int(kat["mc_state_after_init"][c]), kat["corner1"], kat["corner2"], steps, maxit, num_saved_mins=S,
                       init_conf=kat["mc_init_conf"][c], state_conf=lig["conf0"])
        n = int(kat["mc_n"][c])
        assert len(e) == n and np.array_equal(e, kat["mc_e"][c, :n]) and np.array_equal(x, kat["mc_x"][c, :n])


def test_output_container_rule(kat):
    """add_to_output_container (lib/coords.cpp:43-56), which monte_carlo and merge_output_containers (parallel_mc.cpp:165-181) both
    go through: the host-side restatement keeps the same entries in the same order"""
    from gnina_b200.docking import OutputContainer
    for q in range(len(kat["cont_e"])):
        oc = OutputContainer(1.0 if q % 2 else 2.0, 9)
        for e, c in zip(kat["cont_e"][q], kat["cont_coords"][q]):
            oc.add(e, c, np.zeros(7, np.float32))
        kept = kat["cont_kept"][q]; kept = kept[~np.isnan(kept)]
        assert np.array_equal(np.float32([o["e"] for o in oc.items]), kept)
