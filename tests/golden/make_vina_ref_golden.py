"""Generates tests/golden/vina_ref_kat.npz: known answers of the REFERENCE's own Vina code for the docking rows (SURVEY.md §8a
V1-V12), produced by oracle/_ref = the reference's sources compiled where they lie under /root/reference (oracle/Makefile.ref,
oracle/ref_driver.cpp).  Run here, where /root/reference exists:

    python tests/golden/make_vina_ref_golden.py

The fixture travels; the reference does not.  tests/test_oracle_vina_golden.py checks the C restatement against it on any box,
tests/test_gpu_dock.py::test_device_matches_reference_known_answers checks the device kernels against it on the GPU."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gnina_b200 import synth                      # noqa: E402
from oracle import vina_refbuild as R             # noqa: E402

BEGIN, END, N = [-10.1] * 3, [9.9] * 3, [40, 40, 40]   # not aligned to multiples of 3 A: see DESIGN.md (szv_grid cells)
C1, C2 = [-5.0] * 3, [5.0] * 3


def random_confs(rs, lig, T, k, spread):
    X = np.tile(lig["conf0"], (k, 1)).astype(np.float32)
    X[:, :3] += rs.uniform(-spread, spread, (k, 3))
    q = rs.randn(k, 4); X[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    X[:, 7:] = rs.uniform(-np.pi, np.pi, (k, T))
    return X.astype(np.float32)


def main():
    assert R.build(), "oracle/_ref could not be built (needs /root/reference)"
    rs = np.random.RandomState(20240923)
    rx, rt = synth.make_receptor(600, box=30)
    lig = synth.make_flexible_ligand()
    sf = R.RefScoring()
    rm = R.RefModel(lig, rx, rt)
    lo, ro, ra = rm.export()
    T = rm.T
    out = dict(rec_xyz=rx.astype(np.float32), rec_types=rt.astype(np.int32), begin=np.float32(BEGIN), end=np.float32(END), n=np.int32(N),
               corner1=np.float32(C1), corner2=np.float32(C2))
    for k in ("xyz0", "types", "seg_parent", "seg_begin", "seg_end", "axis_root", "pair_a", "pair_b", "conf0"):
        out["lig_" + k] = np.asarray(lig[k])
    out["lig_local_xyz"], out["lig_seg_rel_origin"], out["lig_seg_rel_axis"] = lo, ro, ra

    # G0: the smina type table (names, xs_radius, donor / acceptor / hydrophobe flags) as the reference's `data` array holds it
    ti = [R.type_info(t) for t in range(28)]
    assert all(back == t for t, (_, _, _, _, back) in enumerate(ti))
    out["type_names"] = np.array([a for a, _, _, _, _ in ti]); out["type_xs_radius"] = np.float32([b for _, b, _, _, _ in ti])
    out["type_flags"] = np.int32([d for _, _, _, d, _ in ti])

    # V1 / V2 / V3 / exact: terms and tables at random (t1, t2, r)
    K = 4000
    t12 = rs.randint(0, 28, (K, 2)).astype(np.int32)
    r = rs.uniform(0.05, 7.99, K).astype(np.float32)
    r2 = (r * r).astype(np.float32)
    out["tab_t"], out["tab_r"], out["tab_r2"] = t12, r, r2
    out["tab_terms"] = np.float32([sf.terms(a, b, x) for (a, b), x in zip(t12, r)])
    out["tab_linear"] = np.float32([sf.eval_deriv(R.LINEAR, a, b, x) for (a, b), x in zip(t12, r2)])
    out["tab_linear_fast"] = np.float32([sf.eval(R.LINEAR, a, b, x) for (a, b), x in zip(t12, r2)])
    out["tab_splines"] = np.float32([sf.eval_deriv(R.SPLINES, a, b, x) for (a, b), x in zip(t12, r2)])
    out["tab_exact"] = np.float32([sf.eval(R.EXACT, a, b, x) for (a, b), x in zip(t12, r2)])

    # V4: cache::populate, three of the ligand's types in full
    cg = R.RefGrid.cache(sf, R.LINEAR, rm, BEGIN, END, N, 1e3)   # slope: main/main.cpp:466
    needed = sorted(set(int(t) for t in lig["types"] if t > 1))
    out["grid_types"] = np.int32(needed[:3])
    out["grids"] = np.stack([cg.grid(t) for t in needed[:3]])

    # V5-V8: model::set / model::eval_deriv on the cache, two sets of curl caps
    X = random_confs(rs, lig, T, 64, 4.0)
    out["confs"] = X
    out["coords"] = np.stack([rm.set(x) for x in X])
    for name, caps in (("full", (1000, 1000, 1000)), ("hunt", (10, 1.5, 10))):
        eg = [R.model_eval_deriv(rm, sf, R.LINEAR, cg, x, caps) for x in X]
        out["e_" + name] = np.float32([a for a, _ in eg]); out["g_" + name] = np.stack([b for _, b in eg])
    # non_cache (refine_structure's field), slope 10 and 1000; every second conformation far enough out to leave the box
    Xn = random_confs(rs, lig, T, 48, 4.0); Xn[1::2, :3] += rs.uniform(-5, 5, (24, 3)).astype(np.float32)
    out["nc_confs"] = Xn
    for slope in (10.0, 1000.0):
        nc = R.RefGrid.non_cache(sf, R.LINEAR, rm, BEGIN, END, N, slope)
        eg = [R.model_eval_deriv(rm, sf, R.LINEAR, nc, x, (1000, 1000, 1000)) for x in Xn]
        out["nc_e_%d" % slope] = np.float32([a for a, _ in eg]); out["nc_g_%d" % slope] = np.stack([b for _, b in eg])
        w, ev = [], []
        for x in Xn:
            rm.set(x); w.append(nc.within()); ev.append(nc.eval(1000.0))
        out["nc_within"] = np.array(w)
        out["nc_eval_%d" % slope] = np.float32(ev)      # non_cache::eval: the docking branch's final intermolecular energy
    out["nc_coords"] = np.stack([rm.set(x) for x in Xn])
    # V9: quasi_newton (bfgs.h, fast line search)
    Xb = X[:32]
    for it in (3, 12):
        for name, caps in (("full", (1000, 1000, 1000)), ("hunt", (10, 10, 10))):
            res = [R.bfgs(rm, sf, R.LINEAR, cg, x, it, caps) for x in Xb]
            out["bfgs%d_%s_e" % (it, name)] = np.float32([a for a, _, _ in res])
            out["bfgs%d_%s_x" % (it, name)] = np.stack([b for _, b, _ in res])
    # V9, the --minimize flavours: accurate_line_search (bfgs.h:107-180) and --minimize_early_term
    for tag, acc, et, it in (("acc", True, False, 30), ("acc_et", True, True, 300), ("fast_et", False, True, 60)):
        res = [R.bfgs(rm, sf, R.LINEAR, cg, x, it, (1000, 1000, 1000), accurate=acc, early_term=et) for x in Xb]
        out["min_%s_e" % tag] = np.float32([a for a, _, _ in res]); out["min_%s_x" % tag] = np.stack([b for _, b, _ in res])
    # V12: naive_non_cache::eval with precalculate_exact, num_tors_div, eval_adjusted
    nn = R.RefGrid.naive(sf, R.EXACT, rm)
    e_inter, aff = [], []
    for x in Xb:
        rm.set(x); e_inter.append(nn.eval(1000.0))
        aff.append(R.model_affinity(rm, sf, x))
    out["exact_inter"] = np.float32(e_inter); out["exact_intra_affinity"] = np.float32(aff)
    nt = rs.uniform(0, 12, 32).astype(np.float32)
    out["num_tors"], out["num_tors_div"] = nt, np.float32([sf.num_tors_div(float(e), float(k)) for e, k in zip(e_inter, nt)])
    # V10: whole Monte-Carlo chains (monte_carlo::operator()) on the shim generator
    seeds = (np.arange(1, 9) * 7919).astype(np.uint32)
    steps, maxit, S = 60, (25 + len(lig["types"])) // 3, 20
    out["mc_seeds"], out["mc_params"] = seeds, np.int32([steps, maxit, S])
    init, state, mc_e, mc_x, mc_n = [], [], np.zeros((len(seeds), S), np.float32), np.zeros((len(seeds), S, 7 + T), np.float32), []
    for c, sd in enumerate(seeds):
        x0, st = R.random_conf(rm, int(sd), C1, C2)
        init.append(x0); state.append(st)
        e, x = R.mc(rm, sf, R.LINEAR, cg, int(sd), C1, C2, steps, maxit, lig["conf0"], num_saved_mins=S)
        mc_e[c, :len(e)] = e; mc_x[c, :len(e)] = x; mc_n.append(len(e))
    out["mc_init_conf"], out["mc_state_after_init"] = np.stack(init), np.uint32(state)
    out["mc_e"], out["mc_x"], out["mc_n"] = mc_e, mc_x, np.int32(mc_n)
    # V10 / V11 containers: add_to_output_container replayed on random entry sequences (also how merge_output_containers works)
    seq_e = rs.uniform(-9, -3, (6, 40)).astype(np.float32)
    base = rs.uniform(-3, 3, (6, 5, 12, 3)).astype(np.float32)        # 5 clusters per sequence
    seq_c = np.stack([base[q, rs.randint(0, 5, 40)] + rs.normal(0, 0.4, (40, 12, 3)) for q in range(6)]).astype(np.float32)
    kept = np.full((6, 40), np.nan, np.float32)
    for q in range(6):
        k = R.container_replay(seq_e[q], seq_c[q], 1.0 if q % 2 else 2.0, 9)
        kept[q, :len(k)] = k
    out["cont_e"], out["cont_coords"], out["cont_kept"] = seq_e, seq_c, kept
    path = os.path.join(ROOT, "tests", "golden", "vina_ref_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
