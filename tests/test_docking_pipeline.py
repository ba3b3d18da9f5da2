"""Config 3 glue (gnina_b200/docking.py): host-side containers against hand-worked cases of lib/coords.cpp:32-57,
lib/parallel_mc.cpp:165-181 and main/main.cpp:182-192 (CPU); the whole dock -> rescore pipeline on the device (GPU)."""
import numpy as np
import pytest
from gnina_b200 import docking


def _pose(shift, n=4):
    base = np.arange(3 * n, dtype=np.float32).reshape(n, 3)
    return base + np.float32(shift)


def test_rmsd_upper_bound():
    a, b = _pose(0), _pose(0)
    b[0, 0] += 2.0
    assert abs(docking.rmsd_upper_bound(a, b) - 1.0) < 1e-7          # sqrt(4 / 4 atoms)
    assert docking.rmsd_upper_bound(np.zeros((0, 3)), np.zeros((0, 3))) == 0.0


def test_add_to_output_container_rules():
    c = docking.OutputContainer(min_rmsd=1.0, max_size=3)
    c.add(-1.0, _pose(0), [0])
    c.add(-2.0, _pose(0.1), [1])        # rmsd 0.17 < 1: similar and better -> replaces
    assert [o["e"] for o in c.items] == [-2.0] and c.items[0]["conf"][0] == 1
    c.add(-1.5, _pose(0.2), [2])        # similar and worse -> dropped
    assert [o["e"] for o in c.items] == [-2.0]
    c.add(-0.5, _pose(5), [3]); c.add(-3.0, _pose(10), [4])
    assert [o["e"] for o in c.items] == [-3.0, -2.0, -0.5]           # sorted
    c.add(0.0, _pose(20), [5])          # full, worse than the worst -> dropped
    assert [o["e"] for o in c.items] == [-3.0, -2.0, -0.5]
    c.add(-1.0, _pose(30), [6])         # full, better than the worst -> replaces the worst, re-sorted
    assert [o["e"] for o in c.items] == [-3.0, -2.0, -1.0] and c.items[2]["conf"][0] == 6


def test_merge_chains_uses_rmsd_2_and_remove_redundant_is_strict():
    # two chains found the same minimum (0.6 A apart in RMSD): with the search's min_rmsd 0.5 both survive in their own
    # containers, the merge (min_rmsd forced to 2) keeps the better one only
    e = np.array([[-5.0, -1.0], [-6.0, 0.0]], np.float32)
    coords = np.stack([np.stack([_pose(0), _pose(9)]), np.stack([_pose(0.6 / np.sqrt(3)), _pose(0)])])
    confs = np.zeros((2, 2, 7), np.float32)
    confs[1, 0, 0] = 42
    m = docking.merge_chains(e, confs, coords, np.array([2, 1]), num_saved_mins=50).items
    assert [o["e"] for o in m] == [-6.0, -1.0] and m[0]["conf"][0] == 42
    items = [{"coords": _pose(0)}, {"coords": _pose(1.0 / np.sqrt(3))}, {"coords": _pose(2.0)}]
    d01 = docking.rmsd_upper_bound(items[0]["coords"], items[1]["coords"])
    kept = docking.remove_redundant(items, min_rmsd=d01)              # distance == min_rmsd -> NOT kept ('>' in main.cpp:187)
    assert len(kept) == 2 and kept[1] is items[2]
    assert len(docking.remove_redundant(items, min_rmsd=d01 * 0.99)) == 3


def test_search_box_is_the_reference_grid_dims():
    """setup_grid_dims (main/main.cpp:625-634): n = ceil(size / 0.375), the real span is centred on the requested centre"""
    b, e, n = docking.search_box([-6, -6, -6], [6, 6, 6])
    assert n.tolist() == [32, 32, 32] and np.allclose(b, -6) and np.allclose(e, 6)
    b, e, n = docking.search_box([1.0, -2.0, 0.5], [11.3, 9.0, 16.1])
    assert n.tolist() == [28, 30, 42]
    assert np.allclose((b + e) / 2, [6.15, 3.5, 8.3], atol=1e-5) and np.allclose(e - b, 0.375 * n, atol=1e-5)


def test_reference_num_steps():
    assert docking.reference_num_steps(27, 12) == 70 * 3 * (50 + 27 + 120) // 2


@pytest.mark.gpu
def test_dock_and_rescore_pipeline():
    from gnina_b200 import CNNScorer, synth
    from gnina_b200.vina import VinaScorer
    rec_xyz, rec_t = synth.make_receptor()
    lig = synth.make_flexible_ligand()
    v = VinaScorer(); v.set_receptor(rec_xyz, rec_t)
    c = CNNScorer(["crossdock_default2018"]); c.set_receptor(rec_xyz, rec_t)
    # skip_outside=False: a 12 A box is tight for this ligand; poses that refine_structure could not pull inside stay in the list
    # (e = max_fl, within = False) instead of being dropped as the reference's ranked output drops them
    poses = docking.dock_ligand(v, c, lig, [-6, -6, -6], [6, 6, 6], exhaustiveness=8, seed=3, num_steps=60, num_saved_mins=20,
                                skip_outside=False)
    assert 1 <= len(poses) <= 9
    sc = [p["cnnscore"] for p in poses]
    assert sc == sorted(sc, reverse=True) and all(0.0 <= s <= 1.0 for s in sc)       # ranked by CNNscore
    n_heavy = int((np.asarray(lig["types"]) > 1).sum())
    for i, p in enumerate(poses):
        assert p["all_coords"].shape == (len(lig["types"]), 3) and p["coords"].shape == (n_heavy, 3)   # RMSDs: heavy atoms
        assert np.isfinite(p["cnnaffinity"]) and (np.isfinite(p["e"]) or not p["within"])
        for q in poses[:i]:
            assert docking.rmsd_upper_bound(p["coords"], q["coords"]) > 1.0              # out_min_rmsd
        # the scores attached to a pose are those of its (refined) coordinates
        one = c.score_batch(p["all_coords"], lig["types"], [0, len(lig["types"])])
        assert abs(one[0][0] - p["cnnscore"]) < 1e-6
        if p["within"]:
            b, e_, _ = docking.search_box([-6, -6, -6], [6, 6, 6])
            e_aff = v.score_noncache(p["all_coords"], lig["types"], [0, len(lig["types"])], b, e_, num_tors=np.array([v.T], np.float32))[1][0]
            assert abs(e_aff - p["e"]) < 1e-5 * max(1.0, abs(p["e"]))
            # the exact terms of --score_only give nearly the same number (tables vs closed forms)
            e_ex = v.score_exact(p["all_coords"], lig["types"], [0, len(lig["types"])], num_tors=np.array([v.T], np.float32))[1][0]
            assert abs(e_ex - p["e"]) < 0.05 * max(1.0, abs(p["e"]))
            # refine_structure left every heavy atom inside the search box (non_cache::within)
            hv = p["coords"]
            assert (hv >= b - 1e-3).all() and (hv <= e_ + 1e-3).all()
    again = docking.dock_ligand(v, c, lig, [-6, -6, -6], [6, 6, 6], exhaustiveness=8, seed=3, num_steps=60, num_saved_mins=20,
                                skip_outside=False)
    assert [p["cnnscore"] for p in again] == sc                                            # same seed, same result


@pytest.mark.gpu
def test_concurrent_ligands_equal_sequential_docking():
    """dock_many keeps several ligands in flight on worker threads (own Vina handle + CNN clone each); the result of
    every ligand is bit-identical to docking it alone: no shared mutable state between handles."""
    from gnina_b200 import CNNScorer, synth
    from gnina_b200.vina import VinaScorer
    rec_xyz, rec_t = synth.make_receptor()
    ligs = [synth.make_flexible_ligand(n_heavy=18 + 2 * i, n_tors=3 + i % 3, seed=20 + i) for i in range(6)]
    kw = dict(exhaustiveness=4, num_steps=30, num_saved_mins=10, skip_outside=False)
    many = docking.dock_many(ligs, rec_xyz, rec_t, ["crossdock_default2018"], [-6, -6, -6], [6, 6, 6], n_workers=3, **kw)
    v = VinaScorer(); v.set_receptor(rec_xyz, rec_t)
    c = CNNScorer(["crossdock_default2018"]); c.set_receptor(rec_xyz, rec_t)
    for i, lig in enumerate(ligs):
        one = docking.dock_ligand(v, c, lig, [-6, -6, -6], [6, 6, 6], seed=i + 1, **kw)
        assert len(one) == len(many[i]) >= 1
        for a, b in zip(one, many[i]):
            assert a["cnnscore"] == b["cnnscore"] and a["e"] == b["e"] and np.array_equal(a["coords"], b["coords"])


def test_native_merge_equals_the_python_statement():
    """gb_vina_merge_outputs (host-only C++ in the library; no device needed) against OutputContainer on random chains
    with many near-duplicates"""
    rs = np.random.RandomState(4)
    n_chains, S, na = 12, 8, 5
    centres = rs.randn(6, na, 3) * 4
    coords = np.zeros((n_chains, S, na, 3), np.float32)
    e = np.zeros((n_chains, S), np.float32)
    n_out = rs.randint(0, S + 1, n_chains).astype(np.int32)
    for c in range(n_chains):
        for k in range(S):
            coords[c, k] = centres[rs.randint(6)] + rs.randn(na, 3) * 0.4
            e[c, k] = rs.randn()
        e[c, :n_out[c]] = np.sort(e[c, :n_out[c]])
    confs = rs.randn(n_chains, S, 9).astype(np.float32)
    for max_size in (3, 50):
        py = docking.merge_chains(e, confs, coords, n_out, max_size).items
        nat = docking.merge_chains_native(e, confs, coords, n_out, max_size)
        assert [o["e"] for o in py] == [o["e"] for o in nat] and len(py) >= 1
        for a, b in zip(py, nat):
            assert np.array_equal(a["coords"], b["coords"]) and np.array_equal(a["conf"], b["conf"])
