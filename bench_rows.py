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
    om = pipeline.OracleModel(model_blob.load_model("crossdock_default2018"))
    k = 2
    t0 = time.perf_counter()
    pipeline.score_grad([om], rec_xyz, rec_t, lx[:offs[k]], lt[:offs[k]], offs[:k + 1], dtype=torch.float32)
    cpu = k / (time.perf_counter() - t0)
    for prec, mode, nn in ((1, "fp16 tcgen05 forward + backward", n), (0, "fp32 validation kernels", 256)):
        s = CNNScorer(["crossdock_default2018"], precision=prec)
        s.set_receptor(rec_xyz, rec_t)
        dt = timed(lambda: s.score_grad_batch(lx[:offs[nn]], lt[:offs[nn]], offs[:nn + 1]), reps=3)
        out.append({"row": "cnn_gradient (G2+N5+S1), %d poses" % nn, "value": nn / dt, "unit": "poses/s", "mode": mode,
                    "cpu_oracle": cpu, "cpu_sample": "%d poses, torch autograd + C gridmaker backward" % k})
    # larger batch: the config-5 workload batched over many ligands
    n_big = 8192
    lxb, offb = synth.make_poses(lx0, n_big, seed=6)
    ltb = np.tile(lt0, n_big)
    s = CNNScorer(["crossdock_default2018"], precision=1)
    s.set_receptor(rec_xyz, rec_t)
    dt = timed(lambda: s.score_grad_batch(lxb, ltb, offb), reps=3)
    out.append({"row": "cnn_gradient (G2+N5+S1), %d poses" % n_big, "value": n_big / dt, "unit": "poses/s",
                "mode": "fp16 tcgen05 forward + backward"})

    # --- config 4 shape: virtual screen, ragged ligands, --cnn dense_ensemble (20 models) ---
    n4 = 4096
    sx, st, so = synth.make_screen(n4, seed=3)
    e20 = CNNScorer(["dense_ensemble"], precision=1)
    e20.set_receptor(rec_xyz, rec_t)
    dt = timed(lambda: e20.score_batch(sx, st, so), reps=2)
    out.append({"row": "virtual screen, dense_ensemble (config 4 shape), %d ligands" % n4, "value": n4 / dt, "unit": "ligands/s",
                "mode": "fp16 tcgen05", "models": len(e20.model_names), "model_evals_per_s": n4 * len(e20.model_names) / dt})
    del e20

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
    # --- docking inner loop (V6-V11): one warp per conformation / chain ---
    from oracle.vina_mc import DockOracle
    lig = synth.make_flexible_ligand()
    needed2 = sorted(set(int(t) for t in lig["types"] if t > 1))
    v.cache_build(begin, end, ng, needed2)
    v.set_ligand(lig)
    d = DockOracle(o, {t: v.cache_grid(t) for t in needed2}, begin, end, ng, lig)
    X = np.stack([d.random_conf(1 + i, [-6, -6, -6], [6, 6, 6])[0] for i in range(512)])
    Xb = np.tile(X, (64, 1))                                     # 32768 conformations
    dt = timed(lambda: v.eval_deriv(Xb), reps=3)
    t0 = time.perf_counter()
    for x in X[:200]:
        d.eval_deriv(x)
    cpu = 200 / (time.perf_counter() - t0)
    out.append({"row": "dock_eval_deriv (V5+V6+V7+V8)", "value": len(Xb) / dt, "unit": "eval_deriv/s", "cpu_oracle": cpu,
                "cpu_sample": "200 conformations, scalar C", "ligand": "27 heavy atoms, 6 torsions, %d pairs" % len(lig["pair_a"])})
    res = {}
    def run_bfgs():
        res["ne"] = v.bfgs(Xb[:8192], 12)[3]
    dt = timed(run_bfgs, reps=2)
    t0 = time.perf_counter()
    ner = sum(d.bfgs(x, 12)[3] for x in X[:40])
    cdt = time.perf_counter() - t0
    out.append({"row": "dock_bfgs 12 iterations (V9)", "value": 8192 / dt, "unit": "minimisations/s",
                "device_eval_deriv_per_s": float(res["ne"].sum()) / dt, "cpu_oracle": 40 / cdt, "cpu_eval_deriv_per_s": ner / cdt})
    n_chains, steps = 4096, 40
    seeds = (np.arange(1, n_chains + 1, dtype=np.uint32) * 2654435761) & 0xFFFFFFFF
    dt = timed(lambda: v.mc(seeds, [-6, -6, -6], [6, 6, 6], steps, 12, 8), reps=1)
    t0 = time.perf_counter()
    for c in range(3):
        d.mc(int(seeds[c]), [-6, -6, -6], [6, 6, 6], steps, 12, 8)
    cpu = 3 * steps / (time.perf_counter() - t0)
    out.append({"row": "dock_monte_carlo chains (V10+V11)", "value": n_chains * steps / dt, "unit": "MC steps/s",
                "chains": n_chains, "steps_per_chain": steps, "ligands_per_s_at_exhaustiveness_64": n_chains / 64 / dt,
                "cpu_oracle": cpu, "cpu_sample": "3 chains, scalar C, 1 thread"})
    # --- config 3 glue: cache build -> 64 chains -> merge -> CNN rescoring -> exact affinity -> ranked modes ---
    from gnina_b200 import docking
    cs = CNNScorer(["crossdock_default2018"], precision=1)
    cs.set_receptor(rec_xyz, rec_t)
    ref_steps = docking.reference_num_steps(len(lig["types"]), 6 + v.T)
    st = 200
    docking.dock_ligand(v, cs, lig, [-6, -6, -6], [6, 6, 6], exhaustiveness=64, seed=1, num_steps=st)
    t0 = time.perf_counter()
    poses = docking.dock_ligand(v, cs, lig, [-6, -6, -6], [6, 6, 6], exhaustiveness=64, seed=2, num_steps=st)
    dt = time.perf_counter() - t0
    out.append({"row": "dock + rescore pipeline, one ligand at a time (config 3 glue)", "value": 1.0 / dt, "unit": "ligands/s",
                "exhaustiveness": 64, "mc_steps_per_chain": st, "reference_formula_steps": ref_steps, "modes_out": len(poses),
                "note": "64 chains = 64 warps: one ligand cannot fill the GPU; throughput needs ligands in flight concurrently "
                        "(the MC row above runs 4096 chains per launch)"})
    n_l, workers = 96, 16
    ligs = [synth.make_flexible_ligand(n_heavy=20 + (i % 8), n_tors=3 + i % 4, seed=100 + i) for i in range(n_l)]
    kw = dict(exhaustiveness=64, num_steps=st)
    with docking.DockingPool(rec_xyz, rec_t, ["crossdock_default2018"], n_workers=workers) as pool:
        pool.dock(ligs[:2 * workers], [-6, -6, -6], [6, 6, 6], **kw)          # every worker builds its tables / workspaces
        t0 = time.perf_counter()
        res = pool.dock(ligs, [-6, -6, -6], [6, 6, 6], **kw)
        dt = time.perf_counter() - t0
    out.append({"row": "dock + rescore pipeline, %d ligands in flight (config 3 glue)" % workers, "value": n_l / dt, "unit": "ligands/s",
                "exhaustiveness": 64, "mc_steps_per_chain": st, "ligands": n_l, "host_threads": workers,
                "mc_steps_per_s": n_l * 64 * st / dt, "modes_out_mean": float(np.mean([len(r) for r in res]))})
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
