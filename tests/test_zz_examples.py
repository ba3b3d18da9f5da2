"""examples/rescore_gninatypes.cpp: the smallest complete C++ host program over include/gnina_b200.hpp.
CPU: it compiles warning-free against the header and the library and refuses to run without a device.
GPU: its output equals the Python mirror's on the same typed-atom files."""
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
EXE = os.path.join(ROOT, "examples", "rescore_gninatypes")


def build_example():
    from gnina_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    src = EXE + ".cpp"
    deps = [src, os.path.join(ROOT, "include", "gnina_b200.hpp"), os.path.join(ROOT, "include", "gnina_b200.h")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(d) for d in deps):
        libdir = os.path.join(ROOT, "gnina_b200")
        out = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), src, "-o", EXE,
                              "-L" + libdir, "-lgnina_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
    return EXE


def test_example_compiles_and_needs_a_device(tmp_path):
    import torch
    exe = build_example()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage:" in r.stderr
    if not torch.cuda.is_available():
        r = subprocess.run([exe, os.path.join(ROOT, "gnina_b200", "weights"), "fast", str(tmp_path / "r.gninatypes"),
                            str(tmp_path / "l.gninatypes")], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr      # fails loudly, no silent fallback


@pytest.mark.gpu
def test_example_output_equals_python_mirror(golden_dir, tmp_path):
    from gnina_b200 import CNNScorer, gninatypes
    kat = np.load(os.path.join(golden_dir, "cnn_kat.npz"))
    offs = kat["pose_offsets"]
    rec = tmp_path / "rec.gninatypes"
    gninatypes.write_gninatypes(rec, kat["rec_xyz"], kat["rec_types"])
    ligs = []
    for p in range(len(offs) - 1):
        ligs.append(str(tmp_path / ("lig%d.gninatypes" % p)))
        gninatypes.write_gninatypes(ligs[-1], kat["lig_xyz"][offs[p]:offs[p + 1]], kat["lig_types"][offs[p]:offs[p + 1]])
    names = "crossdock_default2018,crossdock_default2018_KD_4"
    out = subprocess.check_output([build_example(), os.path.join(ROOT, "gnina_b200", "weights"), names, str(rec)] + ligs, text=True)
    rows = [l.split() for l in out.strip().splitlines()]
    assert [r[0] for r in rows] == ligs
    s = CNNScorer(names.split(","))
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    want = s.score_batch(kat["lig_xyz"], kat["lig_types"], offs)
    for p, r in enumerate(rows):
        assert abs(float(r[1]) - want[0][p]) < 2e-5 and abs(float(r[2]) - want[1][p]) < 2e-4 and abs(float(r[3]) - want[3][p]) < 2e-4
