"""Per-phase wall time inside concurrent docking workers: python tools/dock_pool_profile.py [workers] [ligands]"""
import sys, time, threading
sys.path.insert(0, '/root/repo')
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from gnina_b200 import CNNScorer, synth, docking
from gnina_b200.vina import VinaScorer
W = int(sys.argv[1]) if len(sys.argv) > 1 else 16
NL = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rec_xyz, rec_t = synth.make_receptor()
ligs = [synth.make_flexible_ligand(n_heavy=20 + (i % 8), n_tors=3 + i % 4, seed=100 + i) for i in range(NL)]
master = CNNScorer(["crossdock_default2018"]); master.set_receptor(rec_xyz, rec_t)
tls = threading.local(); lock = threading.Lock()
acc = {}; acc_lock = threading.Lock()
c1, c2 = np.array([-6, -6, -6], np.float32), np.array([6, 6, 6], np.float32)
def state():
    if not hasattr(tls, "v"):
        tls.v = VinaScorer(); tls.v.set_receptor(rec_xyz, rec_t)
        with lock: tls.c = master.fresh_copy()
    return tls.v, tls.c
def run(i):
    v, c = state(); lig = ligs[i]; types = np.asarray(lig["types"], np.int32)
    t = [time.perf_counter()]
    v.set_ligand(lig); t.append(time.perf_counter())
    begin = c1 - 4; n = np.ceil((c2 + 4 - begin) / 0.375).astype(np.int32); end = begin + n * 0.375
    v.cache_build(begin.tolist(), end.tolist(), n.tolist(), sorted(set(int(x) for x in types if x > 1))); t.append(time.perf_counter())
    seeds = (np.arange(1, 65, dtype=np.uint32) * 7919 + i).astype(np.uint32)
    e, X, n_out = v.mc(seeds, c1, c2, num_steps=200, maxiters=(25 + len(types)) // 3, num_saved_mins=50); t.append(time.perf_counter())
    _, _, coords = v.eval_deriv(X.reshape(-1, 7 + v.T), coords=True); t.append(time.perf_counter())
    m = docking.merge_chains_native(e, X, coords.reshape(64, 50, len(types), 3), n_out, 50); t.append(time.perf_counter())
    xyz = np.concatenate([q["coords"] for q in m]); offs = (np.arange(len(m) + 1) * len(types)).astype(np.int32)
    c.score_batch(xyz, np.tile(types, len(m)), offs); t.append(time.perf_counter())
    v.score_exact(xyz, np.tile(types, len(m)), offs, num_tors=np.full(len(m), v.T, np.float32)); t.append(time.perf_counter())
    docking.remove_redundant(m, 1.0); t.append(time.perf_counter())
    with acc_lock:
        for nm, a, b in zip(["set_ligand", "cache_build", "mc", "eval_deriv", "merge", "cnn", "exact", "redundant"], t[:-1], t[1:]):
            acc[nm] = acc.get(nm, 0.0) + (b - a)
for workers in (1, W):
    with ThreadPoolExecutor(max_workers=workers) as ex:
        list(ex.map(run, range(2 * workers)))          # warm-up: every worker's handles and workspaces
        acc.clear()
        t0 = time.perf_counter(); list(ex.map(run, range(NL))); wall = time.perf_counter() - t0
    print("workers %2d: wall %.2f s (%.1f ligands/s); per-ligand mean ms: %s" % (
        workers, wall, NL / wall, " ".join("%s %.1f" % (k, 1e3 * v / NL) for k, v in acc.items())))
