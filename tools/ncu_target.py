"""One fast-mode gradient call (forward + backward kernels) for ncu captures: python tools/ncu_target.py [n_poses]"""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from gnina_b200 import CNNScorer, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rec_xyz, rec_t = synth.make_receptor()
lx0, lt0 = synth.make_ligand()
lx, offs = synth.make_poses(lx0, n, seed=5)
lt = np.tile(lt0, n)
s = CNNScorer(["crossdock_default2018"], precision=1)
s.set_receptor(rec_xyz, rec_t)
out = s.score_grad_batch(lx, lt, offs)
print(float(out[0].sum()), float(np.abs(out[4]).sum()))
