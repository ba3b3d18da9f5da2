"""Config 3 throughput against the number of ligands kept in flight (host threads of DockingPool)."""
import json
import sys
import time
import numpy as np
sys.path.insert(0, ".")
from gnina_b200 import docking, synth

rec_xyz, rec_t = synth.make_receptor(1000, box=32.0, seed=3)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for workers in [int(a) for a in sys.argv[2:]] or [32, 64, 128]:
    ligs = [synth.make_flexible_ligand(n_heavy=20 + (i % 8), n_tors=3 + i % 4, seed=300 + i) for i in range(workers)]
    with docking.DockingPool(rec_xyz, rec_t, ["crossdock_default2018"], n_workers=workers) as pool:
        pool.dock(ligs, [-6, -6, -6], [6, 6, 6], exhaustiveness=64, num_steps=50)
        t0 = time.perf_counter()
        res = pool.dock(ligs, [-6, -6, -6], [6, 6, 6], exhaustiveness=64, num_steps=steps)
        dt = time.perf_counter() - t0
    print(json.dumps({"in_flight": workers, "steps_per_chain": steps, "seconds": dt, "mc_steps_per_s": workers * 64 * steps / dt,
                      "ligands_per_s_at_this_length": workers / dt, "modes": float(np.mean([len(r) for r in res]))}), flush=True)
