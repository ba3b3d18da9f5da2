#!/usr/bin/env python3
"""Generates the committed golden fixtures from the REFERENCE's own artefacts.  Run only in the build
container (needs /root/reference); the outputs in tests/golden/ are what travel.

 1. gridmaker_golden.npz — sparse copy (non-zero voxels) of the reference's voxeliser goldens
    test/gninagrid/files/cc_0.48.35.binmap and ccsmall_0.33.35.binmap (35 x N^3 fp32; inputs files/CC.xyz as
    receptor AND ligand, files/recmap + files/ligmap; test/gninagrid/CMakeLists.txt grid2cmp/gridbincmp).
 2. cnn_kat.npz — known answers of the reference's TorchScript models (gninasrc/lib/models/*.pt, loaded with
    torch.jit.load exactly as torch_model.cpp:55 does) on (a) the all-zero grid and (b) oracle-voxelised
    synthetic poses, in fp32 and fp64, plus the head post-processing of torch_model.cpp:188-195.
"""
import json, os, sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
from gnina_b200 import model_blob, synth  # noqa: E402
from oracle import pipeline  # noqa: E402

REF = "/root/reference"


def sparse_golden():
    out = {}
    for tag, fn, n in (("cc48", "cc_0.48.35.binmap", 48), ("ccsmall33", "ccsmall_0.33.35.binmap", 33)):
        g = np.fromfile(os.path.join(REF, "test/gninagrid/files", fn), np.float32)
        assert g.size == 35 * n ** 3
        idx = np.flatnonzero(g).astype(np.int32)
        out[tag + "_idx"] = idx
        out[tag + "_val"] = g[idx]
        out[tag + "_shape"] = np.array([35, n, n, n], np.int32)
    out["recmap"] = np.array(open(os.path.join(REF, "test/gninagrid/files/recmap")).read())
    out["ligmap"] = np.array(open(os.path.join(REF, "test/gninagrid/files/ligmap")).read())
    # files/CC.xyz heavy atoms (hydrogens carry no channel); both carbons type as AliphaticCarbonXSHydrophobe
    out["cc_xyz"] = np.array([[1.06088, 0.05280, 0.06303], [2.57294, 0.05281, 0.06302]], np.float32)
    np.savez_compressed(os.path.join(HERE, "gridmaker_golden.npz"), **out)


MODELS = ["crossdock_default2018", "dense_1.3", "dense_1.3_PT_KD_3", "crossdock_default2018_KD_4",
          "all_default_to_default_1.3_1", "default2017"]


def cnn_kat(n_poses=6):
    rec_xyz, rec_t = synth.make_receptor()
    lig_xyz0, lig_t0 = synth.make_ligand()
    lig_xyz, offs = synth.make_poses(lig_xyz0, n_poses, seed=11)
    lig_t = np.tile(lig_t0, n_poses)
    out = {"rec_xyz": rec_xyz, "rec_types": rec_t, "lig_xyz": lig_xyz, "lig_types": lig_t, "pose_offsets": offs,
           "models": np.array(MODELS)}
    for name in MODELS:
        blob = model_blob.load_model(name)
        om = pipeline.OracleModel(blob)
        grids = om.grids(rec_xyz, rec_t, lig_xyz, lig_t, offs)
        ts = torch.jit.load(os.path.join(REF, "gninasrc/lib/models", name + ".pt"), map_location="cpu")
        key = name.replace(".", "_")
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            m = ts.to(dt)
            with torch.no_grad():
                zp, za = m(torch.zeros(1, om.n_channels, 48, 48, 48, dtype=dt))
                gp, ga = m(torch.from_numpy(grids).to(dt))
            out["%s_zero_logp_%s" % (key, tag)] = zp.numpy()
            out["%s_zero_aff_%s" % (key, tag)] = za.numpy()
            out["%s_logp_%s" % (key, tag)] = gp.numpy()
            out["%s_aff_%s" % (key, tag)] = ga.numpy()
            out["%s_pose_%s" % (key, tag)] = torch.softmax(gp, 1)[:, 1].numpy()  # torch_model.cpp:189
        print(key, "pose", out[key + "_pose_f32"], "aff", out[key + "_aff_f32"])
    np.savez_compressed(os.path.join(HERE, "cnn_kat.npz"), **out)


def grad_kat(n_poses=2, name="crossdock_default2018", out_name="grad_kat.npz"):
    """Ligand-atom gradients of the reference's own TorchScript model: autograd of CE(module output, label 1) through
    the .pt (torch_model.cpp:195-199), then the oracle's GridMaker::backward."""
    from oracle import gridmaker as gm
    import torch.nn.functional as F
    k = np.load(os.path.join(HERE, "cnn_kat.npz"))
    blob = model_blob.load_model(name)
    om = pipeline.OracleModel(blob)
    ts = torch.jit.load(os.path.join(REF, "gninasrc/lib/models", name + ".pt"), map_location="cpu").double()
    offs = k["pose_offsets"][:n_poses + 1]
    rc, rr = gm.type_atoms(k["rec_types"], om.rec_t2c, 0)
    lc, lr = gm.type_atoms(k["lig_types"], om.lig_t2c, om.n_rec)
    grads, losses, rgrads = [], [], []
    for p in range(n_poses):
        sl = slice(offs[p], offs[p + 1])
        c = gm.center_of(k["lig_xyz"][sl])
        xyz = np.concatenate([k["rec_xyz"], k["lig_xyz"][sl]]); ch = np.concatenate([rc, lc[sl]]); rad = np.concatenate([rr, lr[sl]])
        g = torch.from_numpy(gm.grid_forward(c, xyz, ch, rad, om.n_channels)[None]).double().requires_grad_(True)
        out, _ = ts(g)
        loss = F.cross_entropy(out, torch.ones(1, dtype=torch.long))
        loss.backward()
        ag = gm.grid_backward(c, xyz, ch, rad, g.grad[0].numpy().astype(np.float32))
        grads.append(ag[len(k["rec_xyz"]):]); losses.append(float(loss)); rgrads.append(ag[:len(k["rec_xyz"])])
    np.savez_compressed(os.path.join(HERE, out_name), lig_grad=np.concatenate(grads), loss=np.array(losses),
                        rec_grad=np.stack(rgrads).astype(np.float32),   # [pose][receptor atom][3]: getReceptorGradient
                        n_poses=n_poses, model=np.array(name.replace(".", "_")))
    print("grad kat: |g|max", np.abs(np.concatenate(grads)).max(), "loss", losses)


def ensemble_kat(n_poses=3):
    """BASELINE.json config 4's `--cnn dense_ensemble`: every embedded model whose name starts with "dense"
    (cnn_torch_scorer.cpp:49-60) = 15 dense-architecture + 5 default2018-architecture models.  Per-model fp64 outputs
    of the reference's own .pt files on the first poses of cnn_kat.npz, and CNNTorchScorer::score's ensemble
    statistics (mean score / affinity, population variance of the affinity, cnn_torch_scorer.cpp:117-192)."""
    k = np.load(os.path.join(HERE, "cnn_kat.npz"))
    offs = k["pose_offsets"][:n_poses + 1]
    names = sorted(f[:-3] for f in os.listdir(os.path.join(REF, "gninasrc/lib/models")) if f.startswith("dense") and f.endswith(".pt"))
    om = pipeline.OracleModel(model_blob.load_model("dense_1.3"))          # all 20 share the default 14+14 maps
    grids = torch.from_numpy(om.grids(k["rec_xyz"], k["rec_types"], k["lig_xyz"][:offs[-1]], k["lig_types"][:offs[-1]], offs)).double()
    pose, aff = [], []
    for name in names:
        ts = torch.jit.load(os.path.join(REF, "gninasrc/lib/models", name + ".pt"), map_location="cpu").double()
        with torch.no_grad():
            gp, ga = ts(grids)
        pose.append(torch.softmax(gp, 1)[:, 1].numpy()); aff.append(ga.numpy().reshape(-1))
        print(name, pose[-1], aff[-1])
    pose, aff = np.array(pose), np.array(aff)
    np.savez_compressed(os.path.join(HERE, "ensemble_kat.npz"), models=np.array([n.replace(".", "_") for n in names]),
                        n_poses=n_poses, pose_f64=pose, aff_f64=aff, score=pose.mean(0), affinity=aff.mean(0),
                        variance=aff.var(0))


if __name__ == "__main__":
    if "--ensemble-only" in sys.argv:
        ensemble_kat()
        sys.exit(0)
    if "--dense-grad-only" in sys.argv:
        grad_kat(2, "dense_1.3", "grad_kat_dense.npz")   # max-pool / BatchNorm / concat backward (dense family)
        sys.exit(0)
    if "--grad-only" not in sys.argv:
        sparse_golden()
        cnn_kat()
    grad_kat()
    grad_kat(2, "dense_1.3", "grad_kat_dense.npz")
