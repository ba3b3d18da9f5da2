"""Device kernels vs the CPU restatement on ligands with RANDOM torsion trees (synth.make_tree_ligand: nested branches, several children
per node) -- eval_deriv, quasi-Newton, one Monte-Carlo launch.  The restatement is bit-identical to the compiled reference on these trees
(tests/test_oracle_vs_reference_build.py::test_random_torsion_trees); this script is the device half, to be run in the next GPU session
and then turned into a -m gpu test: python tests/device_tree_check.py"""
import json, sys
sys.path.insert(0, '.')
import numpy as np
from gnina_b200 import synth
from gnina_b200.vina import VinaScorer
from oracle.vina import VinaOracle
from oracle.vina_mc import DockOracle

rx, rt = synth.make_receptor(900, box=34)
begin, end, n = [-10.0] * 3, [10.0] * 3, [53, 53, 53]
out = []
for seed in (0, 2, 4, 5):
    lig = synth.make_tree_ligand(seed)
    needed = sorted(set(int(t) for t in lig["types"] if t > 1))
    v = VinaScorer(); v.set_receptor(rx, rt); v.cache_build(begin, end, n, needed)
    d = DockOracle(VinaOracle(), {t: v.cache_grid(t) for t in needed}, begin, end, n, lig)
    lig["gyration_radius"] = d.gyration_radius(lig["conf0"])
    v.set_ligand(lig)
    X = np.stack([d.random_conf(1 + i, [-4] * 3, [4] * 3)[0] for i in range(24)])
    e, g, c = v.eval_deriv(X, coords=True)
    we = max(abs(e[i] - d.eval_deriv(x)[0]) / max(1.0, abs(e[i])) for i, x in enumerate(X))
    wc = max(np.abs(c[i] - d.coords(x)).max() for i, x in enumerate(X))
    eb, xb, gb_, ne = v.bfgs(X, 3)
    same = sum(abs(eb[i] - d.bfgs(x, 3)[0]) <= 1e-5 * max(1.0, abs(eb[i])) for i, x in enumerate(X))
    seeds = np.arange(1, 9, dtype=np.uint32) * 7919
    em, xm, nout, tr = v.mc(seeds, [-4] * 3, [4] * 3, num_steps=20, maxiters=8, num_saved_mins=6, trace=True)
    chains = 0
    for k in range(len(seeds)):
        _, _, trr = d.mc_ex(int(seeds[k]), [-4] * 3, [4] * 3, 20, 8, num_saved_mins=6, min_rmsd=0.5, hunt_cap=(10, 1.5, 10),
                            state_conf=lig["conf0"], trace=True)
        chains += bool((np.abs(tr[k] - trr) <= 1e-5 * np.maximum(1.0, np.abs(trr))).all())
    out.append({"seed": seed, "segments": int(len(lig["seg_parent"])), "eval_rel": float(we), "coords": float(wc), "bfgs3_same": int(same),
                "mc_chains_identical": chains})
    v.close()
print(json.dumps(out))
