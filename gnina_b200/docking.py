"""BASELINE config 3 glued together: the docking branch of gnina's `main_procedure` (main/main.cpp:312-400) over the
device kernels — everything between "ligand topology in" and "ranked poses out":

  cache::populate (V4)  ->  parallel_mc = all chains in one launch (V10/V11)  ->  merge_output_containers
  (lib/parallel_mc.cpp:165-181, min_rmsd forced to 2)  ->  refine_structure of every kept pose (main.cpp:131-171:
  quasi-Newton on non_cache with slope escalation, all poses in one launch)  ->  CNN rescoring of the refined poses
  (get_cnn_info, main.cpp:195-207, one batch call)  ->  final Vina energy (V12)  ->  sort by CNNscore
  (main.cpp:349-360)  ->  remove_redundant (main.cpp:182-192)  ->  first num_modes.

RMSDs are taken over the heavy atoms (get_heavy_atom_movable_coords), the search uses the docking branch's Monte-Carlo
settings (min_rmsd 1, hunt_cap (10,10,10), main.cpp:458-460), and the torsion penalty uses conf_independent_inputs'
num_tors (lib/terms.cpp:74-106) when the ligand dict carries it.
Host logic (containers, sorting) is numpy; nothing in this module imports the CPU oracle."""
import numpy as np


def rmsd_upper_bound(a, b):
    """lib/coords.cpp:24-30"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / len(a))) if len(a) else 0.0


class OutputContainer:
    """output_container + add_to_output_container (lib/coords.cpp:32-57): RMSD-deduplicated, energy-sorted, capped."""

    def __init__(self, min_rmsd, max_size):
        self.min_rmsd, self.max_size = float(min_rmsd), int(max_size)
        self.items = []          # dicts: e, coords, conf (+ whatever the caller attaches later)

    def add(self, e, coords, conf):
        t = {"e": float(e), "coords": np.asarray(coords, np.float32), "conf": np.asarray(conf, np.float32)}
        best, best_r = len(self.items), np.inf
        for i, o in enumerate(self.items):                       # find_closest
            r = rmsd_upper_bound(t["coords"], o["coords"])
            if i == 0 or r < best_r:
                best, best_r = i, r
        if best < len(self.items) and best_r < self.min_rmsd:    # a very similar one: keep the better
            if t["e"] < self.items[best]["e"]:
                self.items[best] = t
        elif len(self.items) < self.max_size:
            self.items.append(t)
        elif self.items and t["e"] < self.items[-1]["e"]:        # full: replace the worst
            self.items[-1] = t
        self.items.sort(key=lambda o: o["e"])                    # stable, like ptr_vector::sort on e


def merge_chains(e, confs, coords, n_out, num_saved_mins):
    """merge_output_containers over the chains' containers, in chain order; min_rmsd = 2 (parallel_mc.cpp:175).
    Pure-Python statement of the rule (tests); `merge_chains_native` is the library's host-side C++ version."""
    out = OutputContainer(2.0, num_saved_mins)
    for c in range(len(n_out)):
        for k in range(int(n_out[c])):
            out.add(e[c, k], coords[c, k], confs[c, k])
    return out


def merge_chains_native(e, confs, coords, n_out, num_saved_mins, min_rmsd=2.0):
    """gb_vina_merge_outputs: the same merge in the library (C++, host only; thousands of RMSDs per ligand would hold
    the GIL for most of a docking run otherwise) -> list of dicts like OutputContainer.items."""
    import ctypes as C
    from . import capi
    e = np.ascontiguousarray(e, np.float32); coords = np.ascontiguousarray(coords, np.float32)
    n_out = np.ascontiguousarray(n_out, np.int32)
    n_chains, S = e.shape
    na = coords.shape[2]
    kept = np.zeros(max(num_saved_mins, 1), np.int32)
    nk = C.c_int32(0)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    capi.check(capi.lib().gb_vina_merge_outputs(e.ctypes.data_as(fp), coords.ctypes.data_as(fp), n_out.ctypes.data_as(ip),
                                                n_chains, S, na, float(min_rmsd), int(num_saved_mins), kept.ctypes.data_as(ip),
                                                C.byref(nk)))
    items = []
    for f in kept[:nk.value]:
        c, k = divmod(int(f), S)
        items.append({"e": float(e[c, k]), "coords": coords[c, k].copy(), "conf": np.asarray(confs[c, k], np.float32).copy()})
    return items


def remove_redundant(items, min_rmsd):
    """main/main.cpp:182-192 (note: keeps a pose whose closest kept pose is EXACTLY min_rmsd away or further: '>')."""
    kept = []
    for it in items:
        best = min((rmsd_upper_bound(it["coords"], k["coords"]) for k in kept), default=None)
        if best is None or best > min_rmsd:
            kept.append(it)
    return kept


def reference_num_steps(n_movable_atoms, n_dof):
    """main/main.cpp:442-443"""
    return int(70 * 3 * (50 + n_movable_atoms + 10 * n_dof) // 2)


MAX_FL = 3.4028234663852886e+38


def search_box(corner1, corner2, granularity=0.375):
    """setup_grid_dims (main/main.cpp:625-634): the search box = ONE grid_dims object that serves as the affinity-grid extent, as the
    box conf::randomize draws from (corner1/2 = gd begin/end, :438-439) and as non_cache's check_bounds box: per axis
    n = ceil(size / 0.375) intervals, real_span = 0.375 n centred on the requested centre.  float32 like the reference.
    -> (begin[3], end[3], n[3])"""
    c1, c2 = np.asarray(corner1, np.float32), np.asarray(corner2, np.float32)
    center, span = (c1 + c2) / np.float32(2), c2 - c1
    n = np.ceil(span / np.float32(granularity)).astype(np.int32)
    real = np.float32(granularity) * n.astype(np.float32)
    begin = (center - real / np.float32(2)).astype(np.float32)
    return begin, (begin + real).astype(np.float32), n


def autobox(xyz, add=4.0, granularity=0.375):
    """model::movable_atoms_box (lib/model.cpp:751-776; --autobox_ligand with --autobox_add 4, and the per-ligand box of --minimize,
    main/main.cpp:1464-1466): the bounding box of the atoms widened by `add` on every side, as grid_dims.  -> (begin[3], end[3], n[3])"""
    xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
    c1 = (xyz.min(axis=0) - np.float32(add)).astype(np.float32)
    c2 = (xyz.max(axis=0) + np.float32(add)).astype(np.float32)
    center = (np.float32(0.5) * (c2 + c1)).astype(np.float32)
    n = np.ceil((c2 - c1) / np.float32(granularity)).astype(np.int32)
    real = (np.float32(granularity) * n.astype(np.float32)).astype(np.float32)
    begin = (center - real / np.float32(2)).astype(np.float32)
    return begin, (begin + real).astype(np.float32), n


def dock_ligand(vina, cnn, lig, corner1, corner2, exhaustiveness=8, seed=1, num_steps=None, maxiters=None,
                num_saved_mins=50, num_modes=9, out_min_rmsd=1.0, sort_order="cnnscore", grid_spacing=0.375, refine=True,
                skip_outside=True):
    """vina: VinaScorer with the receptor set; cnn: CNNScorer with the same receptor set; lig: ligand topology dict; corner1/2: the
    requested search box (what --center / --size or --autobox_ligand + autobox_add describe).
    -> list of dicts (conf, coords, e = final Vina affinity, search_e, cnnscore, cnnaffinity, cnnvariance), ranked."""
    types = np.asarray(lig["types"], np.int32)
    vina.set_ligand(lig)
    T = vina.T
    # ONE box for the affinity grids, the random starts and the out-of-box penalties, as in the reference (search_box above): an atom
    # that leaves the search box pays slope x distance at once, in the grid term of the search and in non_cache afterwards
    begin, end, n = search_box(corner1, corner2, grid_spacing)
    corner1, corner2 = begin, end
    vina.cache_build(begin.tolist(), end.tolist(), n.tolist(), sorted(set(int(t) for t in types if t > 1)))
    if num_steps is None:
        num_steps = reference_num_steps(len(types), 6 + T)
    if maxiters is None:
        maxiters = int((25 + len(types)) // 3)                     # ssd_par.evals, main.cpp:454
    rs = np.random.RandomState(seed)
    seeds = rs.randint(1, 1000000, size=exhaustiveness).astype(np.uint32)   # random_int(0, 1000000, generator)
    # the docking branch's Monte-Carlo settings (main/main.cpp:458,460), not the monte_carlo ctor defaults
    e, X, n_out = vina.mc(seeds, corner1, corner2, num_steps=num_steps, maxiters=maxiters, num_saved_mins=num_saved_mins,
                          min_rmsd=1.0, hunt_cap=(10, 10, 10))
    flat = X.reshape(-1, 7 + T)
    _, _, coords = vina.eval_deriv(flat, coords=True)
    heavy = np.flatnonzero(types > 1)                               # get_heavy_atom_movable_coords
    coords = coords.reshape(len(seeds), num_saved_mins, len(types), 3)
    merged = merge_chains_native(e, X, coords[:, :, heavy], n_out, num_saved_mins)
    if not merged:
        return []
    confs = np.stack([m["conf"] for m in merged])
    if refine:  # refine_structure on every kept pose, cap = authentic_v, minparm.maxiters = ssd_par.evals
        e_ref, confs, ok, _ = vina.refine(confs, maxiters, corner1, corner2)
    else:
        e_ref, ok = np.array([m["e"] for m in merged], np.float32), np.ones(len(merged), bool)
    _, _, all_coords = vina.eval_deriv(confs, coords=True)
    # one CNN batch call and one final-scoring call over all kept poses
    xyz = all_coords.reshape(-1, 3).astype(np.float32)
    offs = (np.arange(len(merged) + 1) * len(types)).astype(np.int32)
    tt = np.tile(types, len(merged))
    sc, aff, _, var = cnn.score_batch(xyz, tt, offs)
    num_tors = float(lig.get("num_tors", T))                        # conf_independent_inputs (lib/terms.cpp:74-106)
    # the docking branch scores with eval_adjusted(..., ig = nc_new) = non_cache::eval on the search box with the search's tables
    # (main/main.cpp:340-344), not with the exact terms of --score_only
    _, affin = vina.score_noncache(xyz, tt, offs, corner1, corner2, num_tors=np.full(len(merged), num_tors, np.float32))
    for i, m in enumerate(merged):
        m["search_e"] = m["e"]
        m["conf"] = confs[i]; m["all_coords"] = all_coords[i]; m["coords"] = all_coords[i][heavy]
        m["refined_e"] = float(e_ref[i]); m["within"] = bool(ok[i])
        # a pose that never entered the box keeps e = max_fl (main.cpp:163-164, :335)
        m["e"] = float(affin[i]) if ok[i] else MAX_FL
        m["cnnscore"] = float(sc[i]); m["cnnaffinity"] = float(aff[i]); m["cnnvariance"] = float(var[i])
    key = {"cnnscore": lambda o: -o["cnnscore"], "cnnaffinity": lambda o: -o["cnnaffinity"], "energy": lambda o: o["e"]}[sort_order]
    merged.sort(key=key)
    # main/main.cpp:371-378: poses that never entered the search box (e = max_fl) are skipped, not counted towards num_modes
    # (skip_outside=False keeps them, with e = max_fl and within = False, for inspection)
    ranked = remove_redundant(merged, out_min_rmsd)
    if skip_outside:
        ranked = [o for o in ranked if o["e"] < 0.1 * MAX_FL]
    return ranked[:num_modes]


class DockingPool:
    """Config 3's "1 receptor x many ligands": one ligand's 64 chains are 64 warps and cannot fill a B200, so ligands
    are kept in flight concurrently — `n_workers` host threads, each with its own VinaScorer handle (own stream, own
    affinity-grid pool) and its own CNNScorer clone (`fresh_copy`, shared weights), exactly the per-thread model copies
    of the reference's `parallel_mc` (lib/parallel_mc.cpp:145-163) one level up.  ctypes releases the GIL during the
    library calls, so the workers' kernels overlap on the device.  Worker state (tables, receptor, workspaces) lives
    as long as the pool, so a screen pays for it once.  Results come back in input order and are identical to
    sequential `dock_ligand` calls with the same seeds."""

    def __init__(self, rec_xyz, rec_types, cnn_model_names, n_workers=8, device=0):
        import threading
        from concurrent.futures import ThreadPoolExecutor
        from .scorer import CNNScorer
        self._rec = (rec_xyz, rec_types)
        self._device = device
        self._master = CNNScorer(cnn_model_names, device=device)
        self._master.set_receptor(rec_xyz, rec_types)
        self._tls = threading.local()
        self._lock = threading.Lock()
        self._ex = ThreadPoolExecutor(max_workers=max(1, int(n_workers)))

    def _state(self):
        from .vina import VinaScorer
        t = self._tls
        if not hasattr(t, "v"):
            t.v = VinaScorer(device=self._device)
            t.v.set_receptor(*self._rec)
            with self._lock:
                t.c = self._master.fresh_copy()
        return t.v, t.c

    def dock(self, ligands, corner1, corner2, seeds=None, **kw):
        def run(i):
            v, c = self._state()
            return dock_ligand(v, c, ligands[i], corner1, corner2, seed=(i + 1 if seeds is None else int(seeds[i])), **kw)
        return list(self._ex.map(run, range(len(ligands))))

    def close(self):
        self._ex.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False


def dock_many(ligands, rec_xyz, rec_types, cnn_model_names, corner1, corner2, n_workers=8, device=0, seeds=None, **kw):
    """One-shot convenience around DockingPool."""
    with DockingPool(rec_xyz, rec_types, cnn_model_names, n_workers, device) as pool:
        return pool.dock(ligands, corner1, corner2, seeds=seeds, **kw)
