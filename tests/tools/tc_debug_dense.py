#!/usr/bin/env python3
"""Stage-by-stage comparison of the dense fast path with the oracle — run on the GPU box."""
import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gnina_b200 import CNNScorer, model_blob
from oracle import pipeline
import tc_layout as tl

kat = np.load(os.path.join(ROOT, "tests/golden/cnn_kat.npz"))
name = "dense_1_3"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
offs = kat["pose_offsets"][:n + 1]
lx, lt = kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]]
blob = model_blob.load_model(name)
om = pipeline.OracleModel(blob)
ref = tl.oracle_intermediates_dense(blob, om.grids(kat["rec_xyz"], kat["rec_types"], lx, lt, offs))
s = CNNScorer([name], precision=1)
s.set_receptor(kat["rec_xyz"], kat["rec_types"])
got = s.score_batch(lx, lt, offs)
print("pose  got", got[0], "want", kat[name + "_pose_f64"][:n])
print("aff   got", got[1], "want", kat[name + "_aff_f64"][:n])


def rep(tag, a, b):
    d = np.abs(a - b)
    i = np.unravel_index(np.argmax(d), d.shape)
    print("%-10s max|d|=%.3e at %s got %.5f want %.5f |ref|max=%.3f rel-rms=%.3e" %
          (tag, d.max(), tuple(int(q) for q in i), a[i], b[i], np.abs(b).max(), np.sqrt((d ** 2).mean()) / (np.sqrt((b ** 2).mean()) + 1e-30)))


x0, b = tl.decode_chunk_planar(s.debug_read("x0"), n, 24, 1, 32)
print("x0 border", b); rep("x0max", x0[:, :28], ref["x0"])
b0, b = tl.decode_chunk_planar(s.debug_read("b0"), n, 24, 1, 96)
print("b0 border", b)
for lo, hi, tag in ((0, 32, "init"), (32, 48, "db0.0"), (48, 64, "db0.1"), (64, 80, "db0.2"), (80, 96, "db0.3")):
    rep("b0 " + tag, b0[:, lo:hi], ref["b0"][:, lo:hi])
b1, b = tl.decode_chunk_planar(s.debug_read("b1"), n, 12, 2, 160)
print("b1 border", b)
for lo, hi, tag in ((0, 96, "bott0"), (96, 112, "db1.0"), (112, 128, "db1.1"), (128, 144, "db1.2"), (144, 160, "db1.3")):
    rep("b1 " + tag, b1[:, lo:hi], ref["b1"][:, lo:hi])
b2, b = tl.decode_chunk_planar(s.debug_read("b2"), n, 6, 2, 224)
print("b2 border", b)
for lo, hi, tag in ((0, 160, "bott1"), (160, 176, "db2.0"), (176, 192, "db2.1"), (192, 208, "db2.2"), (208, 224, "db2.3")):
    rep("b2 " + tag, b2[:, lo:hi], ref["b2"][:, lo:hi])
