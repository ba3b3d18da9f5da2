#!/usr/bin/env python3
"""Stage-by-stage comparison of the fast (tcgen05) path with the oracle — run on the GPU box."""
import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gnina_b200 import CNNScorer, model_blob
from oracle import pipeline
import tc_layout as tl

kat = np.load(os.path.join(ROOT, "tests/golden/cnn_kat.npz"))
name = "crossdock_default2018"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
offs = kat["pose_offsets"][:n + 1]
lx, lt = kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]]
blob = model_blob.load_model(name)
om = pipeline.OracleModel(blob)
grid = om.grids(kat["rec_xyz"], kat["rec_types"], lx, lt, offs)
ref = tl.oracle_intermediates(blob, grid)
s = CNNScorer([name], precision=1)
s.set_receptor(kat["rec_xyz"], kat["rec_types"])
got = s.score_batch(lx, lt, offs)
print("pose  got", got[0], "want", kat[name + "_pose_f64"][:n])
print("aff   got", got[1], "want", kat[name + "_aff_f64"][:n])


def rep(tag, a, b):
    d = np.abs(a - b)
    i = np.unravel_index(np.argmax(d), d.shape)
    print("%-4s max|d|=%.3e  at %s got %.5f want %.5f  |ref|max=%.3f  rel-rms=%.3e" %
          (tag, d.max(), i, a[i], b[i], np.abs(b).max(), np.sqrt((d ** 2).mean()) / (np.sqrt((b ** 2).mean()) + 1e-30)))


x0, b0 = tl.decode_chunk_planar(s.debug_read("x0"), n, 24, 1, 32)
print("x0 border max", b0); rep("x0", x0[:, :28], ref["x0"]); print("x0 pad channels max", np.abs(x0[:, 28:]).max())
x2, b2 = tl.decode_chunk_planar(s.debug_read("x2"), n, 12, 2, 32)
print("x2 border max", b2); rep("x2", x2, ref["x2"])
rep("y3", tl.decode_channels_last(s.debug_read("y3"), n, 12, 64), ref["y3"])
x4, b4 = tl.decode_chunk_planar(s.debug_read("x4"), n, 6, 2, 64)
print("x4 border max", b4); rep("x4", x4, ref["x4"])
rep("y5", tl.decode_channels_last(s.debug_read("y5"), n, 6, 128), ref["y5"])
