"""Generates tests/golden/overlap*.gbw and overlap_kat.npz from the reference's own test artefacts (run here, where
/root/reference exists): test/gnina/data/overlap.pt / overlap_smallr.pt are the models of test/gnina/test_min.py, the
only reference-held pin on the gradient chain (autograd through the model + GridMaker::backward).

  python tests/golden/make_overlap_golden.py
"""
import json
import os
import sys
import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import extract_models  # noqa: E402
from gnina_b200 import model_blob  # noqa: E402
from oracle import gridmaker as gm  # noqa: E402

REF = "/root/reference/test/gnina/data"
OUT = os.path.dirname(os.path.abspath(__file__))

kat = {}
for name in ("overlap", "overlap_smallr"):
    print(extract_models.convert(os.path.join(REF, name + ".pt"), OUT))
    blob = model_blob.load_model(os.path.join(OUT, name + ".gbw"))
    ef = {"metadata": ""}
    m = torch.jit.load(os.path.join(REF, name + ".pt"), map_location="cpu", _extra_files=ef)
    meta = json.loads(ef["metadata"])
    # the two cases of test_min.py: C.xyz / C1.xyz and CC.xyz / CC2.xyz (smina type 2 = AliphaticCarbonXSHydrophobe... any
    # carbon type maps to the single channel)
    cases = {"C": (np.array([[0, 0, 0]], np.float32), np.array([[1, 1, 1]], np.float32)),
             "CC": (np.array([[0, 0, 0], [1.6, 0, 0]], np.float32), np.array([[1.0, -3.2, 1.0], [1.0, -1.6, 1.0]], np.float32))}
    nrec, rt2c = gm.parse_typemap(blob.recmap)
    nlig, lt2c = gm.parse_typemap(blob.ligmap)
    for cname, (rec, lig) in cases.items():
        rt = np.full(len(rec), 2, np.int32); lt = np.full(len(lig), 2, np.int32)
        rc, rr = gm.type_atoms(rt, rt2c, 0)
        lc, lr = gm.type_atoms(lt, lt2c, nrec)
        c = gm.center_of(lig)
        xyz = np.concatenate([rec, lig]); ch = np.concatenate([rc, lc]); rad = np.concatenate([rr, lr])
        g = gm.grid_forward(c, xyz, ch, rad, nrec + nlig, blob.resolution, blob.dimension, meta.get("radius_scaling", 1.0))
        x = torch.from_numpy(g[None].astype(np.float64)).requires_grad_(True)
        out = m.double()(x)
        score = out[0][0, 1]
        loss = -torch.log(score)
        loss.backward()
        kat["%s_%s_score" % (name, cname)] = score.item()
        kat["%s_%s_dgrid_abs_sum" % (name, cname)] = x.grad.abs().sum().item()
        ag = gm.grid_backward(c, xyz, ch, rad, x.grad[0].numpy().astype(np.float32), blob.resolution, blob.dimension,
                              meta.get("radius_scaling", 1.0))
        kat["%s_%s_lig_grad" % (name, cname)] = ag[len(rec):]
        kat["%s_%s_rec" % (name, cname)] = rec
        kat["%s_%s_lig" % (name, cname)] = lig
        print(name, cname, "score", score.item(), "lig grad", ag[len(rec):])
np.savez(os.path.join(OUT, "overlap_kat.npz"), **kat)
