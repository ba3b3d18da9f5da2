"""BASELINE config 5 as a whole minimisation: N poses of one ligand minimised with the CNN (--minimize --cnn_scoring all) in lock step,
gnina_b200/minimize.py driving gb_cnn_score_grad.  NOT YET RUN ON A GPU (written after the round's GPU minutes were spent; the host
logic is pinned on the CPU against the compiled reference): python tools/minimize_demo.py [poses] [maxiters]"""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np
from gnina_b200 import CNNScorer, minimize, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
maxiters = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
rec_xyz, rec_t = synth.make_receptor()
lig = synth.make_flexible_ligand()
tree = minimize.TorsionTree(lig)
rs = np.random.RandomState(5)
X = np.tile(lig["conf0"], (n, 1)).astype(np.float32)
X[:, :3] += rs.uniform(-4, 4, (n, 3))
q = rs.randn(n, 4); X[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
X[:, 7:] = rs.uniform(-np.pi, np.pi, (n, tree.T))
s = CNNScorer(["crossdock_default2018"], precision=1)
s.set_receptor(rec_xyz, rec_t)
energy = minimize.cnn_energy(s, lig["types"], ([-12] * 3, [12] * 3), slope=10.0)
t0 = time.perf_counter()
e, x, ev, rounds = minimize.minimize_poses(tree, energy, X, maxiters=maxiters, accurate=True)
dt = time.perf_counter() - t0
print(json.dumps({"row": "CNN minimisation in lock step (config 5)", "poses": n, "seconds": dt, "poses_per_s": n / dt, "rounds": rounds,
                  "evaluations": int(ev.sum()), "evaluations_per_s": float(ev.sum()) / dt, "loss_mean_start_to_end": [None, float(e.mean())]}))
