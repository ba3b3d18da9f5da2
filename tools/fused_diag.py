"""Diagnostic (GPU): fused unit1_conv kernel vs the separate kernels on the same poses -- where do x2 / scores differ?"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import tc_layout as tl
from gnina_b200 import CNNScorer, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 9
rx, rt = synth.make_receptor(1500, box=44)
lx0, lt0 = synth.make_ligand(22, 3, seed=4)
lx, offs = synth.make_poses(lx0, n, trans_box=10, seed=n)
lt = np.tile(lt0, n)
s = CNNScorer(["crossdock_default2018"], precision=1)
s.set_receptor(rx, rt)
f = s.score_batch(lx, lt, offs)
x0f, b0 = tl.decode_chunk_planar(s.debug_read("x0"), n, 24, 8, 32)
x2f, b2 = tl.decode_chunk_planar(s.debug_read("x2"), n, 12, 2, 32)
u = s.score_grad_batch(lx, lt, offs)
x0u, _ = tl.decode_chunk_planar(s.debug_read("x0"), n, 24, 1, 32)
x2u, _ = tl.decode_chunk_planar(s.debug_read("x2"), n, 12, 2, 32)
print("scores fused  ", f[0][:6]); print("scores unfused", u[0][:6])
print("x0 border", b0, "max|x0f-x0u|", np.abs(x0f - x0u).max(), "x2 border", b2)
d = np.abs(x2f - x2u)
print("x2 scale", np.abs(x2u).max(), "max diff", d.max(), "mean diff", d.mean())
for name, ax in (("pose", (1, 2, 3, 4)), ("chan", (0, 2, 3, 4)), ("x", (0, 1, 3, 4)), ("y", (0, 1, 2, 4)), ("z", (0, 1, 2, 3))):
    print(name, np.round(d.max(axis=ax), 4))
