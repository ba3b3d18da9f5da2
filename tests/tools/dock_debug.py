import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from gnina_b200 import synth
from gnina_b200.vina import VinaScorer
from oracle.vina import VinaOracle
from oracle.vina_mc import DockOracle
rx, rt = synth.make_receptor(900, box=34)
lig = synth.make_flexible_ligand()
begin, end, n = [-10.0] * 3, [10.0] * 3, [53, 53, 53]
needed = sorted(set(int(t) for t in lig["types"] if t > 1))
v = VinaScorer(); v.set_receptor(rx, rt); v.cache_build(begin, end, n, needed); v.set_ligand(lig)
d = DockOracle(VinaOracle(), {t: v.cache_grid(t) for t in needed}, begin, end, n, lig)
X = np.stack([d.random_conf(100 + i, [-4, -4, -4], [4, 4, 4])[0] for i in range(24)])
for it in (1, 2, 3, 5, 8, 12):
    e, Xo, g, ne = v.bfgs(X, it)
    rel, dx, nes = [], [], []
    for i in range(len(X)):
        er, xr, gr, ner = d.bfgs(X[i], it)
        rel.append(abs(e[i] - er) / max(1.0, abs(er))); dx.append(np.abs(Xo[i] - xr).max()); nes.append((int(ne[i]), ner))
    print("iters", it, "frac close", np.mean(np.array(rel) < 1e-2), "median rel", np.median(rel), "max dx", np.max(dx), "evals", nes[:6])
