"""Quick device-only timing of the docking kernels: python tools/dock_rows.py [chains] [steps]"""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from gnina_b200 import synth
from gnina_b200.vina import VinaScorer
rec_xyz, rec_t = synth.make_receptor()
lig = synth.make_flexible_ligand()
types = np.asarray(lig["types"], np.int32)
v = VinaScorer(); v.set_receptor(rec_xyz, rec_t); v.set_ligand(lig)
c1, c2 = np.array([-6, -6, -6], np.float32), np.array([6, 6, 6], np.float32)
begin = c1 - 4; n = np.ceil((c2 + 4 - begin) / 0.375).astype(np.int32); end = begin + n * 0.375
v.cache_build(begin.tolist(), end.tolist(), n.tolist(), sorted(set(int(x) for x in types if x > 1)))
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
seeds = np.arange(1, nch + 1, dtype=np.uint32) * 7919
v.mc(seeds[:256], c1, c2, num_steps=4, maxiters=17, num_saved_mins=20)
best = 1e9
for rep in range(2):
    t0 = time.perf_counter()
    e, X, n_out = v.mc(seeds, c1, c2, num_steps=steps, maxiters=17, num_saved_mins=20)
    best = min(best, time.perf_counter() - t0)
rng = np.random.default_rng(1)
nc = 65536
T = v.T
X0 = np.zeros((nc, 7 + T), np.float32)
X0[:, :3] = rng.uniform(-5, 5, (nc, 3)); q = rng.normal(size=(nc, 4)); X0[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
X0[:, 7:] = rng.uniform(-3, 3, (nc, T))
v.eval_deriv(X0[:1024])
t0 = time.perf_counter(); v.eval_deriv(X0); te = time.perf_counter() - t0
v.bfgs(X0[:1024], maxiters=12)
t0 = time.perf_counter(); v.bfgs(X0[:16384], maxiters=12); tb = time.perf_counter() - t0
print(json.dumps({"mc_steps_per_s": nch * steps / best, "chains": nch, "steps": steps, "eval_deriv_per_s": nc / te, "bfgs12_per_s": 16384 / tb,
                  "e_checksum": float(np.sum(e[:, 0].astype(np.float64)))}))
