"""ctypes front-end to oracle/vina_mc_ref.c (docking inner loop oracle; test infrastructure only)."""
import ctypes as C
import numpy as np
from . import gridmaker as _gm
from .vina import VinaOracle, lib as _vlib

_fp, _ip, _vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p


class _Lig(C.Structure):
    _fields_ = [("n_atoms", C.c_int), ("n_seg", C.c_int), ("n_pairs", C.c_int), ("local_xyz", _fp), ("type", _ip),
                ("seg_parent", _ip), ("seg_begin", _ip), ("seg_end", _ip), ("seg_rel_origin", _fp), ("seg_rel_axis", _fp),
                ("pair_a", _ip), ("pair_b", _ip)]


class _Field(C.Structure):
    _fields_ = [("grids", C.POINTER(_fp)), ("begin", _fp), ("end", _fp), ("n", _ip), ("slope", C.c_float), ("prec", _vp), ("splines", _vp),
                ("rec_xyz", _fp), ("rec_type", _ip), ("n_rec", C.c_int)]


class _McParams(C.Structure):
    _fields_ = [("num_steps", C.c_int), ("maxiters", C.c_int), ("num_saved_mins", C.c_int), ("temperature", C.c_float),
                ("mutation_amplitude", C.c_float), ("min_rmsd", C.c_float), ("hunt_cap", C.c_float * 3),
                ("gyration_radius", C.c_float)]


def _f(a): return a.ctypes.data_as(_fp)
def _i(a): return a.ctypes.data_as(_ip)


class DockOracle:
    """cache grids (dict type -> array, x fastest) + precalculate tables + ligand topology"""

    def __init__(self, vina_oracle, grids, begin, end, n, lig, slope=1e3):
        L = _vlib()
        L.gvo_lig_set_conf.argtypes = [C.POINTER(_Lig), _fp, _fp, _fp, _fp]
        L.gvo_lig_eval_deriv.argtypes = [C.POINTER(_Field), C.POINTER(_Lig), _fp, _fp, _fp, _fp]; L.gvo_lig_eval_deriv.restype = C.c_float
        L.gvo_lig_eval_grid.argtypes = [C.POINTER(_Field), C.POINTER(_Lig), _fp, C.c_float, _fp]; L.gvo_lig_eval_grid.restype = C.c_float
        L.gvo_bfgs.argtypes = [C.POINTER(_Field), C.POINTER(_Lig), _fp, _fp, C.c_int, _fp, C.POINTER(C.c_int)]; L.gvo_bfgs.restype = C.c_float
        L.gvo_mc_run.argtypes = [C.POINTER(_Field), C.POINTER(_Lig), C.POINTER(_McParams), _fp, _fp, C.c_uint32, _fp, _fp]
        L.gvo_mc_run_traced.argtypes = [C.POINTER(_Field), C.POINTER(_Lig), C.POINTER(_McParams), _fp, _fp, C.c_uint32, _fp, _fp, _fp]
        L.gvo_bfgs_ex.argtypes = [C.POINTER(_Field), C.POINTER(_Lig), _fp, _fp, C.c_int, _fp, C.POINTER(C.c_int), C.c_int, C.c_int]
        L.gvo_bfgs_ex.restype = C.c_float
        L.gvo_mc_run_ex.argtypes = [C.POINTER(_Field), C.POINTER(_Lig), C.POINTER(_McParams), _fp, _fp, C.c_uint32, _fp, _fp, _fp, _fp, _fp]
        L.gvo_random_conf.argtypes = [C.POINTER(C.c_uint32), _fp, _fp, C.c_int, _fp]
        self.L = L
        self.keep = []
        k = lambda a, dt: self.keep.append(np.ascontiguousarray(a, dt)) or self.keep[-1]
        self.lig = _Lig(len(lig["types"]), len(lig["seg_parent"]), len(lig["pair_a"]), _f(k(lig["local_xyz"], np.float32)),
                        _i(k(lig["types"], np.int32)), _i(k(lig["seg_parent"], np.int32)), _i(k(lig["seg_begin"], np.int32)),
                        _i(k(lig["seg_end"], np.int32)), _f(k(lig["seg_rel_origin"], np.float32)),
                        _f(k(lig["seg_rel_axis"], np.float32)), _i(k(lig["pair_a"], np.int32)), _i(k(lig["pair_b"], np.int32)))
        self.ptrs = (_fp * 28)()
        for t, g in grids.items():
            self.ptrs[t] = _f(k(g, np.float32))
        self.field = _Field(self.ptrs, _f(k(begin, np.float32)), _f(k(end, np.float32)), _i(k(n, np.int32)), slope, vina_oracle.p, None,
                            None, None, 0)
        L.gvo_noncache_atom.argtypes = [C.POINTER(_Field), C.c_int, _fp, C.c_float, _fp]; L.gvo_noncache_atom.restype = C.c_float
        L.gvo_within.argtypes = [C.POINTER(_Field), C.POINTER(_Lig), _fp, C.c_float]
        L.gvo_refine_structure.argtypes = [C.POINTER(_Field), C.POINTER(_Lig), _fp, _fp, C.c_int, _fp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.gvo_refine_structure.restype = C.c_float
        L.gvo_refine_structure_ex.argtypes = L.gvo_refine_structure.argtypes + [C.c_int, C.c_int]; L.gvo_refine_structure_ex.restype = C.c_float
        self.vo = vina_oracle
        self.T = len(lig["seg_parent"]) - 1
        self.na = len(lig["types"])
        self.gr = lig["gyration_radius"]

    def set_box(self, begin=None, end=None, slope=None):
        """replace the box [begin, end] (check_bounds / grid origin) and the out-of-box slope; None restores the
        construction-time values.  (Never save `self.field.begin` and assign it back: a ctypes pointer FIELD accessed
        through the structure aliases the structure's memory, so the 'saved' value changes with the field.)"""
        if not hasattr(self, "_box0"):
            self._box0 = (C.cast(self.field.begin, C.c_void_p).value, C.cast(self.field.end, C.c_void_p).value, self.field.slope)
        if begin is None:
            self.field.begin = C.cast(self._box0[0], _fp); self.field.end = C.cast(self._box0[1], _fp); self.field.slope = self._box0[2]
            return
        b = np.ascontiguousarray(begin, np.float32); e = np.ascontiguousarray(end, np.float32)
        self.keep += [b, e]
        self.field.begin, self.field.end = _f(b), _f(e)
        if slope is not None:
            self.field.slope = slope

    def use_noncache(self, rec_xyz=None, rec_types=None):
        """non_cache (lib/non_cache.cpp): sum the intermolecular term over the receptor atoms directly instead of the
        cache grids; begin/end become the search box of check_bounds.  None switches back to the cache."""
        if rec_xyz is None:
            self.field.rec_xyz, self.field.rec_type, self.field.n_rec = None, None, 0
            return
        x = np.ascontiguousarray(rec_xyz, np.float32); t = np.ascontiguousarray(rec_types, np.int32)
        self.keep += [x, t]
        self.field.rec_xyz, self.field.rec_type, self.field.n_rec = _f(x), _i(t), len(t)

    def noncache_atom(self, t, xyz, v=1000.0):
        """non_cache::eval_deriv for one atom of smina type t -> (energy incl. out-of-box penalty, minus-force)"""
        a = np.ascontiguousarray(xyz, np.float32); d = np.zeros(3, np.float32)
        return self.L.gvo_noncache_atom(C.byref(self.field), int(t), _f(a), v, _f(d)), d

    def within(self, conf, margin=1e-4):
        conf = np.ascontiguousarray(conf, np.float32)
        return bool(self.L.gvo_within(C.byref(self.field), C.byref(self.lig), _f(conf), margin))

    def refine_structure(self, conf, maxiters, v=(1000, 1000, 1000), accurate=False, early_term=False):
        """main/main.cpp:131-171 -> (energy, refined conf, n_evals, within)"""
        x = np.array(conf, np.float32); v = np.ascontiguousarray(v, np.float32)
        g = np.empty(6 + self.T, np.float32); ne, ok = C.c_int(), C.c_int()
        e = self.L.gvo_refine_structure_ex(C.byref(self.field), C.byref(self.lig), _f(x), _f(g), maxiters, _f(v), C.byref(ne), C.byref(ok),
                                           int(accurate), int(early_term))
        return e, x, ne.value, bool(ok.value)

    def use_splines(self, on=True):
        """precalculate_splines (factor 10) for the intramolecular pair terms instead of precalculate_linear"""
        self.field.splines = self.vo.splines() if on else None

    def coords(self, conf):
        conf = np.ascontiguousarray(conf, np.float32)
        c = np.empty((self.na, 3), np.float32); so = np.empty((self.T + 1, 3), np.float32); sa = np.empty((self.T + 1, 3), np.float32)
        self.L.gvo_lig_set_conf(C.byref(self.lig), _f(conf), _f(c), _f(so), _f(sa))
        return c

    def eval_deriv(self, conf, v=(1000, 1000, 1000)):
        conf = np.ascontiguousarray(conf, np.float32); v = np.ascontiguousarray(v, np.float32)
        g = np.empty(6 + self.T, np.float32)
        e = self.L.gvo_lig_eval_deriv(C.byref(self.field), C.byref(self.lig), _f(conf), _f(v), _f(g), None)
        return e, g

    def eval_grid(self, conf, v1=1000.0):
        conf = np.ascontiguousarray(conf, np.float32)
        return self.L.gvo_lig_eval_grid(C.byref(self.field), C.byref(self.lig), _f(conf), v1, None)

    def bfgs(self, conf, maxiters, v=(1000, 1000, 1000), accurate=False, early_term=False):
        """quasi_newton (bfgs.h:358-502); accurate = BFGSAccurateLineSearch (--minimize), early_term = --minimize_early_term"""
        x = np.array(conf, np.float32); v = np.ascontiguousarray(v, np.float32)
        g = np.empty(6 + self.T, np.float32); ne = C.c_int()
        e = self.L.gvo_bfgs_ex(C.byref(self.field), C.byref(self.lig), _f(x), _f(g), maxiters, _f(v), C.byref(ne), int(accurate),
                               int(early_term))
        return e, x, g, ne.value

    def random_conf(self, seed, c1, c2):
        s = C.c_uint32(seed)
        x = np.empty(7 + self.T, np.float32)
        self.L.gvo_random_conf(C.byref(s), _f(np.ascontiguousarray(c1, np.float32)), _f(np.ascontiguousarray(c2, np.float32)), self.T, _f(x))
        return x, s.value

    def mc(self, seed, c1, c2, num_steps, maxiters, num_saved_mins=20, temperature=1.2, amplitude=2.0, min_rmsd=0.5,
           hunt_cap=(10, 1.5, 10), trace=False):
        P = _McParams(num_steps, maxiters, num_saved_mins, temperature, amplitude, min_rmsd, (C.c_float * 3)(*hunt_cap), self.gr)
        e = np.zeros(num_saved_mins, np.float32); x = np.zeros((num_saved_mins, 7 + self.T), np.float32)
        tr = np.zeros(num_steps, np.float32)
        n = self.L.gvo_mc_run_traced(C.byref(self.field), C.byref(self.lig), C.byref(P), _f(np.ascontiguousarray(c1, np.float32)),
                                     _f(np.ascontiguousarray(c2, np.float32)), seed, _f(e), _f(x), _f(tr))
        return (e[:n], x[:n], tr) if trace else (e[:n], x[:n])

    def gyration_radius(self, conf):
        """model::gyration_radius (lib/model.cpp:1002-1014) of a conformation: heavy atoms about the root origin"""
        self.L.gvo_gyration_radius.argtypes = [C.POINTER(_Lig), _fp]; self.L.gvo_gyration_radius.restype = C.c_float
        return self.L.gvo_gyration_radius(C.byref(self.lig), _f(np.ascontiguousarray(conf, np.float32)))

    def mc_ex(self, seed, c1, c2, num_steps, maxiters, num_saved_mins=50, temperature=1.2, amplitude=2.0, min_rmsd=1.0,
              hunt_cap=(10, 10, 10), init_conf=None, state_conf=None, trace=False):
        """monte_carlo::operator() with an optional given start (init_conf, `seed` = generator state after the start was drawn)
        and, with state_conf, in MODEL-STATE mode (vina_mc_ref.c mc_impl): the mode that follows the reference's code to the letter"""
        P = _McParams(num_steps, maxiters, num_saved_mins, temperature, amplitude, min_rmsd, (C.c_float * 3)(*hunt_cap), self.gr)
        e = np.zeros(num_saved_mins, np.float32); x = np.zeros((num_saved_mins, 7 + self.T), np.float32)
        ic = None if init_conf is None else np.ascontiguousarray(init_conf, np.float32)
        sc = None if state_conf is None else np.ascontiguousarray(state_conf, np.float32)
        tr = np.zeros(num_steps, np.float32)
        n = self.L.gvo_mc_run_ex(C.byref(self.field), C.byref(self.lig), C.byref(P), _f(np.ascontiguousarray(c1, np.float32)),
                                 _f(np.ascontiguousarray(c2, np.float32)), seed, _f(e), _f(x), _f(tr),
                                 None if ic is None else _f(ic), None if sc is None else _f(sc))
        return (e[:n], x[:n], tr) if trace else (e[:n], x[:n])
