"""The reference-side bindings a gnina maintainer adds (integration/cnn_b200_scorer.h: `CNNB200Scorer : DLScorer`;
integration/docking_b200.h: `parallel_mc_b200`, `cache_b200 : igrid`, `refine_structure_b200`, `score_docked_b200`, `B200Ligand`)
are real code: they COMPILE against the reference's own headers (override-checked against DLScorer / igrid), and the model -> topology
conversion RUNS on reference `model` objects (oracle/_ref).  Needs /root/reference (skipped on the GPU box)."""
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/gninasrc/lib"


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is absent")
def test_adapters_compile_against_the_reference_headers():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror=overloaded-virtual", "-DNDEBUG", "-w",
           "-I" + os.path.join(ROOT, "oracle", "ref_shim"), "-I" + REF, "-I/root/reference", "-I/usr/local/cuda/include",
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "integration"),
           os.path.join(ROOT, "integration", "compile_check.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is absent")
def test_end_to_end_program_links_and_runs_the_reference_side():
    """integration/e2e_docking.cpp: gnina's own classes (linked from oracle/_ref), the adapters and libgnina_b200.so in ONE executable --
    every symbol the adapters need exists on both sides; its `cpu` mode (reference parallel_mc on the host + the topology adapter) runs"""
    from oracle import vina_refbuild as R
    assert R.build()
    import __graft_entry__  # noqa: F401  (libgnina_b200.so must exist for the link)
    from gnina_b200 import build as b
    b.build()
    subprocess.check_call(["make", "-s", "-f", "Makefile.ref", "_ref/e2e_docking"], cwd=os.path.join(ROOT, "oracle"))
    out = subprocess.check_output([os.path.join(ROOT, "oracle", "_ref", "e2e_docking"), "cpu"], text=True).splitlines()
    assert out[0].startswith("topology atoms 9 segments 3 pairs 8 heavy 9") and out[-1] == "cpu ok"
    e = [float(l.split()[-1]) for l in out if l.startswith("reference pose")]
    assert len(e) >= 3 and e == sorted(e) and e[0] < -3.0


def test_model_to_topology_adapter_on_reference_models():
    """B200Ligand(const model&) walks the reference's heterotree: atoms, segments (DFS pre-order), parents, relative origins / axes,
    interacting pairs and gyration radius come out exactly as the reference's constructors stored them"""
    from oracle import vina_refbuild as R
    if not (R.available() or R.build()):
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    from gnina_b200 import synth
    for kw in (dict(), dict(n_heavy=31, n_tors=8, n_branch=4, seed=3), dict(n_heavy=14, n_tors=2, n_branch=2, seed=29)):
        lig = synth.make_flexible_ligand(**kw)
        rm = R.RefModel(lig)
        t = rm.adapter_topology()
        lo, ro, ra = rm.export()
        for k in ("types", "seg_parent", "seg_begin", "seg_end", "pair_a", "pair_b"):
            assert np.array_equal(t[k], np.asarray(lig[k], np.int32)), k
        assert np.array_equal(t["local_xyz"], lo) and np.array_equal(t["seg_rel_origin"], ro) and np.array_equal(t["seg_rel_axis"], ra)
        assert t["gyration_radius"] == rm.gyration_radius() and t["n_heavy"] == int((np.asarray(lig["types"]) > 1).sum())
