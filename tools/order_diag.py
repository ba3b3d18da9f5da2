"""Diagnostic (GPU): are scores independent of batch order / repeatable?  python tools/order_diag.py [n]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from gnina_b200 import CNNScorer, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rx, rt = synth.make_receptor()
lx0, lt0 = synth.make_ligand()
na = len(lt0)
lx, offs = synth.make_poses(lx0, n, seed=1)
lt = np.tile(lt0, n)
s = CNNScorer(["crossdock_default2018"], precision=1)
s.set_receptor(rx, rt)
a = s.score_batch(lx, lt, offs)
a2 = s.score_batch(lx, lt, offs)
perm = np.random.RandomState(0).permutation(n)
lxp = lx.reshape(n, na, 3)[perm].reshape(-1, 3)
b = s.score_batch(lxp, lt, offs)
b2 = s.score_batch(lxp, lt, offs)
def rep(tag, x, y):
    bad = np.flatnonzero(x != y)
    print(tag, "mismatches", len(bad), "first", bad[:12], "max abs", float(np.abs(x - y).max()), "chunks", sorted(set((bad // 2048).tolist()))[:8],
          "pos in group", sorted(set((bad % 8).tolist())))
rep("repeat same order   ", a[0], a2[0])
rep("repeat permuted     ", b[0], b2[0])
rep("order vs permuted   ", a[0][perm], b[0])
rep("affinity order/perm ", a[1][perm], b[1])
