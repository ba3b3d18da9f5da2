"""Generates tests/golden/cnn_ref_kat.npz: known answers of the REFERENCE's own CNN scoring host code -- lib/torch_model.cpp,
lib/cnn_torch_scorer.cpp, lib/dl_scorer.cpp compiled where they lie under /root/reference (oracle/Makefile.ref target `cnn`,
oracle/ref_cnn_driver.cpp), running the reference's own TorchScript files through libtorch on the CPU; libmolgrid (third party, absent)
is the stand-in over oracle/gridmaker_ref.c.  Run here, where /root/reference exists:

    python tests/golden/make_cnn_ref_golden.py

The fixture travels; the reference does not.  tests/test_oracle_cnn_ref_golden.py checks oracle/pipeline.py against it on any box."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gnina_b200 import synth                      # noqa: E402
from oracle import vina_refbuild as R             # noqa: E402
from oracle import cnn_refbuild as CR             # noqa: E402

SINGLE = ["crossdock_default2018", "general_default2018_3", "redock_default2018_2", "dense", "dense_1_3_PT_KD_3", "default2017",
          "all_default_to_default_1_3_1"]
ALIASES = {"default_ensemble": [], "fast": ["fast"], "default1_0": ["default1.0"]}


def make_inputs():
    lig = dict(synth.make_flexible_ligand(n_heavy=20, n_tors=4, n_branch=2, seed=8))
    rx, rt = synth.make_receptor(600, box=24, seed=5)
    ty_h = lig["types"].copy(); ty_h[2] = 1; ty_h[9] = 0; ty_h[15] = 1       # polar and non-polar hydrogens among the movable atoms
    confs = np.tile(lig["conf0"], (2, 1)).astype(np.float32)
    confs[0, :3] += [0.5, -0.4, 0.3]
    confs[1, :3] += [-1.0, 0.8, 0.2]; confs[1, 3:7] = np.float32([0.8, 0.2, -0.4, 0.4]) / np.linalg.norm([0.8, 0.2, -0.4, 0.4])
    confs[1, 7:] = np.linspace(-1.1, 2.0, confs.shape[1] - 7)
    return lig, ty_h, rx, rt, confs.astype(np.float32)


def main():
    assert CR.build(), "oracle/_ref (CNN half) could not be built (needs /root/reference)"
    lig, ty_h, rx, rt, confs = make_inputs()
    out = dict(rec_xyz=rx.astype(np.float32), rec_types=rt.astype(np.int32), confs=confs, lig_types=lig["types"].astype(np.int32),
               lig_types_h=ty_h.astype(np.int32))
    rm = R.RefModel(lig, rx, rt)
    lig_h = dict(lig); lig_h["types"] = ty_h
    rm_h = R.RefModel(lig_h, rx, rt)
    coords = np.stack([rm.set(x) for x in confs])
    out["coords"] = coords
    for name in SINGLE:                                                        # one model, both poses, with gradient
        s = CR.RefCNNScorer(names=[name])
        res, mf = [], []
        for x in confs:
            rm.set(x)
            r = s.score(rm, True)
            res.append(r[:4]); mf.append(r[4])
        out["single_%s_out" % name] = np.float32(res); out["single_%s_forces" % name] = np.stack(mf)
    for key, names in ALIASES.items():                                         # the constructor's name logic + the ensemble arithmetic
        s = CR.RefCNNScorer(names=names)
        rm.set(confs[0])
        r = s.score(rm, True)
        out["alias_%s_out" % key] = np.float32(r[:4]); out["alias_%s_forces" % key] = r[4]
    # hydrogens: setLigand keeps them, the typer leaves them untyped, add_minus_forces consumes the by-atom list compactly
    s = CR.RefCNNScorer(names=["crossdock_default2018"])
    rm_h.set(confs[0])
    r = s.score(rm_h, True)
    out["hyd_out"] = np.float32(r[:4]); out["hyd_forces"] = r[4]
    c, b, e, n = s.center_and_box(rm_h)                                        # set_center_from_model, set_bounding_box
    out["hyd_center"], out["hyd_box_begin"], out["hyd_box_end"], out["hyd_box_n"] = c, b, e, n
    # --cnn_center
    cc = np.float32([0.25, -0.5, 0.75])
    s = CR.RefCNNScorer(names=["crossdock_default2018"], cnn_center=cc)
    rm.set(confs[0])
    out["center_given"] = cc; out["center_given_out"] = np.float32(s.score(rm, False)[:4])
    # --cnn_models file: the reference's own test network (test/gnina/data/overlap.pt, used by test_min.py)
    s = CR.RefCNNScorer(files=["/root/reference/test/gnina/data/overlap.pt"])
    r = s.score(rm, True)
    out["overlap_out"] = np.float32(r[:4]); out["overlap_forces"] = r[4]
    path = os.path.join(ROOT, "tests", "golden", "cnn_ref_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
