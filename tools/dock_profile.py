"""Where does a docking run spend its time?  python tools/dock_profile.py"""
import sys, time, threading
sys.path.insert(0, '/root/repo')
import numpy as np
from gnina_b200 import CNNScorer, synth, docking
from gnina_b200.vina import VinaScorer
rec_xyz, rec_t = synth.make_receptor()
lig = synth.make_flexible_ligand()
types = np.asarray(lig["types"], np.int32)
def setup():
    v = VinaScorer(); v.set_receptor(rec_xyz, rec_t)
    return v
v = setup()
c = CNNScorer(["crossdock_default2018"]); c.set_receptor(rec_xyz, rec_t)
c1, c2 = np.array([-6, -6, -6], np.float32), np.array([6, 6, 6], np.float32)
def phases(v, c, tag):
    t = [time.perf_counter()]
    v.set_ligand(lig); t.append(time.perf_counter())
    begin = c1 - 4; n = np.ceil((c2 + 4 - begin) / 0.375).astype(np.int32); end = begin + n * 0.375
    v.cache_build(begin.tolist(), end.tolist(), n.tolist(), sorted(set(int(x) for x in types if x > 1))); t.append(time.perf_counter())
    seeds = np.arange(1, 65, dtype=np.uint32) * 7919
    e, X, n_out = v.mc(seeds, c1, c2, num_steps=200, maxiters=17, num_saved_mins=50); t.append(time.perf_counter())
    _, _, coords = v.eval_deriv(X.reshape(-1, 7 + v.T), coords=True); t.append(time.perf_counter())
    coords = coords.reshape(64, 50, len(types), 3)
    m = docking.merge_chains_native(e, X, coords, n_out, 50); t.append(time.perf_counter())
    xyz = np.concatenate([q["coords"] for q in m]); offs = (np.arange(len(m) + 1) * len(types)).astype(np.int32)
    c.score_batch(xyz, np.tile(types, len(m)), offs); t.append(time.perf_counter())
    v.score_exact(xyz, np.tile(types, len(m)), offs, num_tors=np.full(len(m), v.T, np.float32)); t.append(time.perf_counter())
    names = ["set_ligand", "cache_build", "mc", "eval_deriv", "merge", "cnn", "exact"]
    print(tag, " ".join("%s %.1fms" % (nm, 1e3 * (b - a)) for nm, a, b in zip(names, t[:-1], t[1:])), "n_out mean %.1f" % n_out.mean())
phases(v, c, "warm-up:")
phases(v, c, "single :")
# raw concurrency of the MC kernel: 8 threads, own handle each
hs = [setup() for _ in range(8)]
for h in hs:
    h.set_ligand(lig)
    begin = c1 - 4; n = np.ceil((c2 + 4 - begin) / 0.375).astype(np.int32); end = begin + n * 0.375
    h.cache_build(begin.tolist(), end.tolist(), n.tolist(), sorted(set(int(x) for x in types if x > 1)))
seeds = np.arange(1, 65, dtype=np.uint32) * 7919
def run(h): h.mc(seeds, c1, c2, num_steps=200, maxiters=17, num_saved_mins=50)
run(hs[0])
t0 = time.perf_counter(); run(hs[0]); t1 = time.perf_counter() - t0
ths = [threading.Thread(target=run, args=(h,)) for h in hs]
t0 = time.perf_counter(); [t.start() for t in ths]; [t.join() for t in ths]; t8 = time.perf_counter() - t0
print("mc alone %.1f ms; 8 concurrent handles %.1f ms (%.2fx of serial)" % (1e3 * t1, 1e3 * t8, t8 / (8 * t1)))
one = np.arange(1, 513, dtype=np.uint32) * 7919
t0 = time.perf_counter(); hs[0].mc(one, c1, c2, num_steps=200, maxiters=17, num_saved_mins=50); print("512 chains in one launch %.1f ms" % (1e3 * (time.perf_counter() - t0)))
