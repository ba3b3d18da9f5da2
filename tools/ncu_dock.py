import sys
sys.path.insert(0, '/root/repo')
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
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
seeds = np.arange(1, nch + 1, dtype=np.uint32) * 7919
e, X, n_out = v.mc(seeds, c1, c2, num_steps=int(sys.argv[2]) if len(sys.argv) > 2 else 10, maxiters=17, num_saved_mins=20)
print(e[:2, :3], n_out[:8])
