#!/usr/bin/env python3
"""Measurement of the hot-path rows that are not bench.py's headline: one JSON line per row with the
device throughput (CUDA events are inside the library's stream for the CNN rows; wall clock around the synchronous
C-ABI calls for the Vina rows) and the CPU oracle timed beside it on a bounded sample.

  rows: cnn_gradient (G2+N5+S1, config 5 shape: 1k poses), cnn_validation_fp32 (N1 fp32 mode), default_ensemble (3 models,
        fp32 validation path for the dense members), vina_cache_build (V4), vina_cache_eval (V5), vina_exact_affinity (V12)
"""
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    import torch
    from gnina_b200 import CNNScorer, model_blob, synth
    from gnina_b200.vina import VinaScorer
    from oracle import pipeline
    from oracle.vina import VinaOracle
    assert torch.cuda.is_available(), "no CUDA device — no CPU fallback"
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    rec_xyz, rec_t = synth.make_receptor()
    lx0, lt0 = synth.make_ligand()
    out = []

    # --- CNN gradient path (config 5 shape: 1k poses with atom gradients) ---
    n = 1000
    lx, offs = synth.make_poses(lx0, n, seed=5)
    lt = np.tile(lt0, n)
    s = CNNScorer(["crossdock_default2018"])
    s.set_receptor(rec_xyz, rec_t)
    dt = timed(lambda: s.score_grad_batch(lx, lt, offs), reps=2)
    om = pipeline.OracleModel(model_blob.load_model("crossdock_default2018"))
    k = 2
    t0 = time.perf_counter()
    pipeline.score_grad([om], rec_xyz, rec_t, lx[:offs[k]], lt[:offs[k]], offs[:k + 1], dtype=torch.float32)
    cpu = k / (time.perf_counter() - t0)
    out.append({"row": "cnn_gradient (G2+N5+S1)", "value": n / dt, "unit": "poses/s", "mode": "fp32 validation kernels",
                "cpu_oracle": cpu, "cpu_sample": "%d poses, torch autograd + C gridmaker backward" % k})

    # --- fp32 validation forward and the default 3-model ensemble ---
    n2 = 512
    for names, tag in ((["crossdock_default2018"], "cnn_validation_fp32 (N1)"), ([], "default_ensemble 3 models (N1+N2, S1)")):
        e = CNNScorer(names, precision=0)
        e.set_receptor(rec_xyz, rec_t)
        dt = timed(lambda: e.score_batch(lx[:offs[n2]], lt[:offs[n2]], offs[:n2 + 1]), reps=2)
        out.append({"row": tag, "value": n2 / dt, "unit": "poses/s", "mode": "fp32 validation kernels",
                    "models": e.model_names})

    # --- Vina rows ---
    v, o = VinaScorer(), VinaOracle()
    v.set_receptor(rec_xyz, rec_t)
    begin, end, ng = [-12.0] * 3, [12.0] * 3, [64, 64, 64]     # 0.375 A spacing, 65^3 points
    needed = sorted(set(int(t) for t in lt0 if t > 1))
    dt = timed(lambda: v.cache_build(begin, end, ng, needed), reps=2)
    pts = 65 ** 3 * len(needed)
    t0 = time.perf_counter()
    o.cache_populate(begin, end, [16, 16, 16], rec_xyz, rec_t, needed[0])
    cpu = 17 ** 3 / (time.perf_counter() - t0)
    out.append({"row": "vina_cache_build (V4)", "value": pts / dt, "unit": "grid-point-types/s", "grid": "65^3 x %d types" % len(needed),
                "cpu_oracle": cpu, "cpu_sample": "17^3 points x 1 type, scalar C"})
    nv = 20000
    lxv, offv = synth.make_poses(lx0, nv, trans_box=12, seed=6)
    ltv = np.tile(lt0, nv)
    dt = timed(lambda: v.cache_eval(lxv, ltv, offv), reps=3)
    grids = {t: v.cache_grid(t) for t in needed}
    kk = 200
    t0 = time.perf_counter()
    for p in range(kk):
        VinaOracle.cache_eval(grids, begin, end, ng, lxv[offv[p]:offv[p + 1]], lt0, 1e3, 1000.0)
    cpu = kk / (time.perf_counter() - t0)
    out.append({"row": "vina_cache_eval+deriv (V5)", "value": nv / dt, "unit": "poses/s", "cpu_oracle": cpu,
                "cpu_sample": "%d poses, scalar C via ctypes" % kk})
    dt = timed(lambda: v.score_exact(lxv, ltv, offv, np.full(nv, 4.0, np.float32)), reps=3)
    kk = 20
    t0 = time.perf_counter()
    for p in range(kk):
        o.naive_exact(rec_xyz, rec_t, lxv[offv[p]:offv[p + 1]], lt0)
    cpu = kk / (time.perf_counter() - t0)
    out.append({"row": "vina_exact_affinity (V12)", "value": nv / dt, "unit": "poses/s", "cpu_oracle": cpu,
                "cpu_sample": "%d poses, scalar C" % kk, "note": "includes H2D of the poses and D2H of the energies"})
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
