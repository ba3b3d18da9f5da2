"""Config 3 at (a fraction of) the reference's Monte-Carlo length, 64 ligands in flight: python tools/config3_row.py [fraction] [half_box]
The search box is what --autobox_ligand with the default --autobox_add 4 gives for these ligands (extent ~12 A + 4 A on every side)."""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np
from gnina_b200 import docking, synth
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
hb = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
rec_xyz, rec_t = synth.make_receptor()
n_full = workers = 64
ligs = [synth.make_flexible_ligand(n_heavy=20 + (i % 8), n_tors=3 + i % 4, seed=300 + i) for i in range(n_full)]
steps_ref = [docking.reference_num_steps(len(l["types"]), 6 + len(l["seg_parent"]) - 1) for l in ligs]
steps = int(np.mean(steps_ref) * frac)
c1, c2 = [-hb] * 3, [hb] * 3
with docking.DockingPool(rec_xyz, rec_t, ["crossdock_default2018"], n_workers=workers) as pool:
    pool.dock(ligs, c1, c2, exhaustiveness=64, num_steps=50)
    t0 = time.perf_counter()
    res = pool.dock(ligs, c1, c2, exhaustiveness=64, num_steps=steps)
    dt = time.perf_counter() - t0
mc = 64.0 * steps * n_full
print(json.dumps({"row": "config 3: dock + refine + rescore, exhaustiveness 64, 64 ligands in flight", "box_half_width": hb,
                  "mc_steps_per_chain": steps, "reference_formula_steps_mean": float(np.mean(steps_ref)), "seconds": dt,
                  "mc_steps_per_s": mc / dt, "ligands_per_s_measured": n_full / dt,
                  "ligands_per_s_at_reference_length": (mc / dt) / (64.0 * float(np.mean(steps_ref))),
                  "modes_out_mean": float(np.mean([len(r) for r in res]))}))
