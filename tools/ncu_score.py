"""One fast-mode scoring call (voxeliser + default2018 forward) for ncu captures: python tools/ncu_score.py [n_poses] [model]"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from gnina_b200 import CNNScorer, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
model = sys.argv[2] if len(sys.argv) > 2 else "crossdock_default2018"
rec_xyz, rec_t = synth.make_receptor()
lx0, lt0 = synth.make_ligand()
lx, offs = synth.make_poses(lx0, n, seed=5)
lt = np.tile(lt0, n)
s = CNNScorer([model], precision=1)
s.set_receptor(rec_xyz, rec_t)
for _ in range(2):
    out = s.score_batch(lx, lt, offs)
print(float(out[0].sum()))
