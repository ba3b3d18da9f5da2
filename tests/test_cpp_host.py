"""The C++ host side (include/gnina_b200.hpp): compiles with g++ against the C ABI, mirrors the reference's
model-name expansion (CPU), and on the GPU box reproduces the oracle through CNNScorer / NonCacheCNN."""
import os
import struct
import subprocess
import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
EXE = os.path.join(ROOT, "tests", "cpp", "host_test")


def build_exe():
    from gnina_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    src = os.path.join(ROOT, "tests", "cpp", "host_test.cpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src),
                                                              os.path.getmtime(os.path.join(ROOT, "include", "gnina_b200.hpp")),
                                                              os.path.getmtime(os.path.join(ROOT, "include", "gnina_b200_minimize.hpp")),
                                                              os.path.getmtime(os.path.join(ROOT, "include", "gnina_b200_dock.hpp"))):
        libdir = os.path.join(ROOT, "gnina_b200")
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L" + libdir, "-lgnina_b200", "-Wl,-rpath," + libdir])
    return EXE


def test_cpp_host_builds_and_expands_names():
    out = subprocess.check_output([build_exe(), "--names"], text=True)
    assert out.strip() == "dense_1_3 dense_1_3_PT_KD_3 crossdock_default2018_KD_4 | all_default_to_default_1_3_1 | 2"


def test_cpp_lock_step_minimiser_on_the_host():
    """gb::minimize_poses (include/gnina_b200_minimize.hpp) with an analytic energy: every pose descends, a pose minimised alone ends at
    the same point as inside the batch, and the energy functor is called once per round, not once per evaluation (the bit-for-bit
    comparison with the reference's quasi_newton is tests/test_oracle_vs_reference_build.py)"""
    out = subprocess.check_output([build_exe(), "--minimize-host"], text=True).split()
    kv = dict(zip(out[1::2], out[2::2]))
    assert kv["descended"] == "1" and kv["same_alone"] == "1" and kv["batched"] == "1" and int(kv["evaluations"]) > 8 * 5


@pytest.mark.gpu
def test_cpp_host_scores_match_oracle(golden_dir, tmp_path):
    import torch
    from gnina_b200 import model_blob
    from oracle import pipeline
    kat = np.load(os.path.join(golden_dir, "cnn_kat.npz"))
    n = 3
    offs = kat["pose_offsets"][:n + 1].astype(np.int32)
    lx, lt = kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]].astype(np.int32)
    case = tmp_path / "case.bin"
    with open(case, "wb") as f:
        f.write(struct.pack("<iii", len(kat["rec_types"]), len(lt), n))
        f.write(kat["rec_xyz"].astype("<f4").tobytes()); f.write(kat["rec_types"].astype("<i4").tobytes())
        f.write(lx.astype("<f4").tobytes()); f.write(lt.tobytes()); f.write(offs.tobytes())
    out = subprocess.check_output([build_exe(), os.path.join(ROOT, "gnina_b200", "weights"), str(case)], text=True)
    lines = out.strip().splitlines()
    assert lines[0].startswith("usage_error: Invalid model name: no_such_model")
    name = "crossdock_default2018"
    for p in range(n):
        _, idx, sc, aff, loss = lines[1 + p].split()
        assert abs(float(sc) - kat[name + "_pose_f64"][p]) < 2e-3 and abs(float(aff) - kat[name + "_aff_f64"][p]) < 1e-2
    by_tag = {l.split()[0]: l.split() for l in lines}
    assert by_tag["batcher"][1:] == ["delivered", str(n), "batches", "2", "maxdiff"] + by_tag["batcher"][6:] and \
        float(by_tag["batcher"][6]) < 1e-6                                # gb::PoseBatcher == the direct batch call
    single = by_tag["single"]
    assert abs(float(single[1]) - kat[name + "_pose_f64"][0]) < 2e-3      # gradient call: fast fp16 forward + backward
    grad = np.load(os.path.join(golden_dir, "grad_kat.npz"))["lig_grad"][:offs[1]]
    assert abs(float(single[5]) - np.abs(grad).sum()) < 1e-2 * np.abs(grad).sum()
    # non_cache_cnn: loss + slope * (distance outside [-1,1]^3 box) over heavy atoms; the CNN box (23.5 A around the
    # origin) adds its own penalty for atoms beyond +-11.75
    om = pipeline.OracleModel(model_blob.load_model(name))
    loss0 = om.score(kat["rec_xyz"], kat["rec_types"], lx[:offs[1]], lt[:offs[1]], offs[:2], dtype=torch.float64)[2][0]
    heavy = lx[:offs[1]][lt[:offs[1]] > 1]
    pen = 10.0 * (np.clip(np.abs(heavy) - 1.0, 0, None).sum() + np.clip(np.abs(heavy) - 11.75, 0, None).sum())
    md = by_tag["multidevice"]                                            # sharded over 2 handles per device == single scorer
    assert int(md[1]) >= 2 and float(md[3]) < 1e-6 and md[5:] == ["0", "2", "5", "7", "10"]
    # empirical mixing: total = (CNN loss + w * empirical) / (1 + w), forces blended the same way (test_min.py:45-61, 1e-3)
    e_plain, e_mixed, emp, worst = (float(x) for x in by_tag["mixing"][1:5])
    assert abs(e_mixed - (e_plain + 0.5 * emp) / 1.5) < 1e-3 * max(1.0, abs(e_mixed)) and worst < 1e-5
    from oracle.vina import VinaOracle
    from oracle.vina_mc import DockOracle
    vo = VinaOracle()
    dk = DockOracle(vo, {}, [-40.0] * 3, [40.0] * 3, [8, 8, 8], dict(local_xyz=np.zeros((1, 3), np.float32), types=np.array([2], np.int32),
                    seg_parent=np.array([-1], np.int32), seg_begin=np.array([0], np.int32), seg_end=np.array([1], np.int32),
                    seg_rel_origin=np.zeros((1, 3), np.float32), seg_rel_axis=np.zeros((1, 3), np.float32),
                    pair_a=np.zeros(0, np.int32), pair_b=np.zeros(0, np.int32), gyration_radius=1.0), slope=0.0)
    dk.use_noncache(kat["rec_xyz"], kat["rec_types"])
    emp_ref = sum(dk.noncache_atom(int(t), x)[0] for x, t in zip(lx[:offs[1]], lt[:offs[1]]) if t > 1)
    assert abs(emp - emp_ref) < 1e-5 * max(1.0, abs(emp_ref))
    e, e0 = float(by_tag["noncache"][1]), float(by_tag["noncache"][2])
    assert abs(e - (loss0 + pen)) < 1e-3 * max(1.0, abs(loss0 + pen)) and abs(e0 - (loss0 + pen)) < 3e-3 * max(1.0, abs(e0))
