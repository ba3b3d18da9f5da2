"""Host-side mirror of the smina/Vina scoring entry points over the C ABI (include/gnina_b200.h, gb_vina_*):
`precalculate_linear` tables, `cache::populate`, `cache::eval/eval_deriv`, exact final scoring ("Affinity")."""
import ctypes as C
import numpy as np
from . import capi


def _fp(a): return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None
def _ip(a): return a.ctypes.data_as(C.POINTER(C.c_int32))


class VinaScorer:
    def __init__(self, device=0, weights6=None, factor=32.0):
        L = capi.lib()
        self._h = C.c_void_p()
        w = None if weights6 is None else np.ascontiguousarray(weights6, np.float32)
        capi.check(L.gb_vina_create(device, _fp(w), factor, C.byref(self._h)))
        self.n = L.gb_vina_table_size(self._h)
        self.dims = None

    def table(self, t1, t2):
        a, b, c = (np.empty(self.n, np.float32) for _ in range(3))
        capi.check(capi.lib().gb_vina_prec_table(self._h, t1, t2, _fp(a), _fp(b), _fp(c)))
        return a, b, c

    def set_receptor(self, xyz, smina_types):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(smina_types, np.int32)
        capi.check(capi.lib().gb_vina_set_receptor(self._h, _fp(xyz), _ip(t), len(t)))

    def cache_build(self, begin, end, n, types_needed):
        b, e = np.ascontiguousarray(begin, np.float32), np.ascontiguousarray(end, np.float32)
        n = np.ascontiguousarray(n, np.int32)
        tn = np.ascontiguousarray(types_needed, np.int32)
        capi.check(capi.lib().gb_vina_cache_build(self._h, _fp(b), _fp(e), _ip(n), _ip(tn), len(tn)))
        self.dims = (int(n[2]) + 1, int(n[1]) + 1, int(n[0]) + 1)

    def cache_grid(self, t):
        out = np.empty(self.dims, np.float32)
        capi.check(capi.lib().gb_vina_cache_read(self._h, t, _fp(out)))
        return out

    @staticmethod
    def _poses(lig_xyz, lig_types, pose_offsets):
        return (np.ascontiguousarray(lig_xyz, np.float32).reshape(-1, 3), np.ascontiguousarray(lig_types, np.int32),
                np.ascontiguousarray(pose_offsets, np.int32))

    def cache_eval(self, lig_xyz, lig_types, pose_offsets, slope=1e3, v=1000.0, deriv=True):
        x, t, o = self._poses(lig_xyz, lig_types, pose_offsets)
        e = np.empty(len(o) - 1, np.float32)
        d = np.zeros((len(t), 3), np.float32) if deriv else None
        capi.check(capi.lib().gb_vina_cache_eval(self._h, _fp(x), _ip(t), _ip(o), len(o) - 1, slope, v, _fp(e), _fp(d)))
        return e, d

    def score_exact(self, lig_xyz, lig_types, pose_offsets, num_tors=None, v=1000.0):
        """-> (intermolecular exact energy, Affinity) per pose"""
        x, t, o = self._poses(lig_xyz, lig_types, pose_offsets)
        n = len(o) - 1
        nt = None if num_tors is None else np.ascontiguousarray(num_tors, np.float32)
        e, a = np.empty(n, np.float32), np.empty(n, np.float32)
        capi.check(capi.lib().gb_vina_score_exact(self._h, _fp(x), _ip(t), _ip(o), n, _fp(nt), v, _fp(e), _fp(a)))
        return e, a

    def score_noncache(self, lig_xyz, lig_types, pose_offsets, box_begin, box_end, num_tors=None, v=1000.0, slope=1e3):
        """the docking branch's final score (main/main.cpp:340-344): non_cache::eval on the search box with the search's tables, then
        num_tors_div -> (intermolecular energy, Affinity) per pose"""
        x, t, o = self._poses(lig_xyz, lig_types, pose_offsets)
        n = len(o) - 1
        nt = None if num_tors is None else np.ascontiguousarray(num_tors, np.float32)
        b, en = np.ascontiguousarray(box_begin, np.float32), np.ascontiguousarray(box_end, np.float32)
        e, a = np.empty(n, np.float32), np.empty(n, np.float32)
        capi.check(capi.lib().gb_vina_score_noncache(self._h, _fp(x), _ip(t), _ip(o), n, _fp(nt), v, slope, _fp(b), _fp(en), _fp(e), _fp(a)))
        return e, a

    def spline_table(self, t1, t2):
        n = capi.lib().gb_vina_spline_size(self._h)
        out = np.empty((n, 4), np.float32)
        capi.check(capi.lib().gb_vina_spline_table(self._h, t1, t2, _fp(out)))
        return out

    def set_precalc(self, use_splines):
        capi.check(capi.lib().gb_vina_set_precalc(self._h, int(bool(use_splines))))

    # ---- docking inner loop -------------------------------------------------------------------------------------
    def set_ligand(self, lig):
        """lig: dict as gnina_b200.synth.make_flexible_ligand returns it (the fields of gb_ligand_topology)"""
        k = lambda a, dt: np.ascontiguousarray(a, dt)
        self._lig_keep = [k(lig["local_xyz"], np.float32), k(lig["types"], np.int32), k(lig["seg_parent"], np.int32),
                          k(lig["seg_begin"], np.int32), k(lig["seg_end"], np.int32), k(lig["seg_rel_origin"], np.float32),
                          k(lig["seg_rel_axis"], np.float32), k(lig["pair_a"], np.int32), k(lig["pair_b"], np.int32)]
        a = self._lig_keep
        t = capi.LigandTopology(len(a[1]), len(a[2]), len(a[7]), _fp(a[0]), _ip(a[1]), _ip(a[2]), _ip(a[3]), _ip(a[4]), _fp(a[5]),
                                _fp(a[6]), _ip(a[7]), _ip(a[8]), float(lig["gyration_radius"]))
        capi.check(capi.lib().gb_vina_set_ligand(self._h, C.byref(t)))
        self.T, self.na = len(a[2]) - 1, len(a[1])

    def eval_deriv(self, confs, v=(1000, 1000, 1000), slope=1e3, coords=False):
        x = np.ascontiguousarray(confs, np.float32).reshape(-1, 7 + self.T)
        n = len(x)
        v = np.ascontiguousarray(v, np.float32)
        e = np.empty(n, np.float32); g = np.empty((n, 6 + self.T), np.float32)
        c = np.empty((n, self.na, 3), np.float32) if coords else None
        capi.check(capi.lib().gb_vina_eval_deriv(self._h, _fp(x), n, _fp(v), slope, _fp(e), _fp(g), _fp(c)))
        return (e, g, c) if coords else (e, g)

    def bfgs(self, confs, maxiters, v=(1000, 1000, 1000), slope=1e3, accurate=False, early_term=False):
        """quasi_newton on the cache field; accurate / early_term = the --minimize flavours (gb_vina_minimize)"""
        x = np.array(confs, np.float32).reshape(-1, 7 + self.T)
        n = len(x)
        v = np.ascontiguousarray(v, np.float32)
        e = np.empty(n, np.float32); g = np.empty((n, 6 + self.T), np.float32); ne = np.empty(n, np.int32)
        if accurate or early_term:
            mp = capi.MinimizationParams(int(maxiters), int(accurate), int(early_term))
            capi.check(capi.lib().gb_vina_minimize(self._h, _fp(x), n, C.byref(mp), _fp(v), slope, _fp(e), _fp(g), _ip(ne)))
        else:
            capi.check(capi.lib().gb_vina_bfgs(self._h, _fp(x), n, maxiters, _fp(v), slope, _fp(e), _fp(g), _ip(ne)))
        return e, x, g, ne

    def eval_deriv_noncache(self, confs, box_begin, box_end, v=(1000, 1000, 1000), slope=1e3):
        """model::eval_deriv with ig = non_cache (direct receptor sums, lib/non_cache.cpp:126-174) -> (e, change)"""
        x = np.ascontiguousarray(confs, np.float32).reshape(-1, 7 + self.T)
        n = len(x)
        v = np.ascontiguousarray(v, np.float32)
        b, en = np.ascontiguousarray(box_begin, np.float32), np.ascontiguousarray(box_end, np.float32)
        e = np.empty(n, np.float32); g = np.empty((n, 6 + self.T), np.float32)
        capi.check(capi.lib().gb_vina_eval_deriv_noncache(self._h, _fp(x), n, _fp(v), slope, _fp(b), _fp(en), _fp(e), _fp(g)))
        return e, g

    def noncache_atoms(self, xyz, smina_types, box_begin, box_end, v=1000.0):
        """per-atom empirical term of non_cache_cnn::eval_deriv (lib/non_cache_cnn.cpp:113-140) -> (e[n], deriv[n,3])"""
        x = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3); t = np.ascontiguousarray(smina_types, np.int32)
        b, en = np.ascontiguousarray(box_begin, np.float32), np.ascontiguousarray(box_end, np.float32)
        e = np.empty(len(t), np.float32); d = np.empty((len(t), 3), np.float32)
        capi.check(capi.lib().gb_vina_noncache_atoms(self._h, _fp(x), _ip(t), len(t), _fp(b), _fp(en), v, _fp(e), _fp(d)))
        return e, d

    def refine(self, confs, maxiters, box_begin, box_end, v=(1000, 1000, 1000), accurate=False, early_term=False):
        """refine_structure (main/main.cpp:131-171) -> (e, refined confs, within, n_evals); accurate / early_term: --minimize"""
        x = np.array(confs, np.float32).reshape(-1, 7 + self.T)
        n = len(x)
        v = np.ascontiguousarray(v, np.float32)
        b, en = np.ascontiguousarray(box_begin, np.float32), np.ascontiguousarray(box_end, np.float32)
        e = np.empty(n, np.float32); ok = np.empty(n, np.int32); ne = np.empty(n, np.int32)
        if accurate or early_term:
            mp = capi.MinimizationParams(int(maxiters), int(accurate), int(early_term))
            capi.check(capi.lib().gb_vina_refine_minimize(self._h, _fp(x), n, C.byref(mp), _fp(v), _fp(b), _fp(en), _fp(e), _ip(ok), _ip(ne)))
        else:
            capi.check(capi.lib().gb_vina_refine(self._h, _fp(x), n, maxiters, _fp(v), _fp(b), _fp(en), _fp(e), _ip(ok), _ip(ne)))
        return e, x, ok.astype(bool), ne

    def mc(self, seeds, corner1, corner2, num_steps, maxiters, num_saved_mins=20, temperature=1.2, amplitude=2.0, min_rmsd=0.5,
           hunt_cap=(10, 1.5, 10), slope=1e3, trace=False):
        seeds = np.ascontiguousarray(seeds, np.uint32)
        n = len(seeds)
        P = capi.McParams(num_steps, maxiters, num_saved_mins, temperature, amplitude, min_rmsd, (C.c_float * 3)(*hunt_cap))
        e = np.zeros((n, num_saved_mins), np.float32); x = np.zeros((n, num_saved_mins, 7 + self.T), np.float32)
        no = np.zeros(n, np.int32)
        c1, c2 = np.ascontiguousarray(corner1, np.float32), np.ascontiguousarray(corner2, np.float32)
        tr = np.zeros((n, num_steps), np.float32) if trace else None
        capi.check(capi.lib().gb_vina_mc_traced(self._h, C.byref(P), _fp(c1), _fp(c2), seeds.ctypes.data_as(C.POINTER(C.c_uint32)), n,
                                                slope, _fp(e), _fp(x), _ip(no), _fp(tr)))
        return (e, x, no, tr) if trace else (e, x, no)

    def close(self):
        if self._h:
            capi.lib().gb_vina_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
