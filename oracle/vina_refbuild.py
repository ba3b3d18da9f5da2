"""ctypes front-end to oracle/_ref/libgnina_vina_ref.so = the REFERENCE's own Vina sources compiled where they lie under
/root/reference (oracle/Makefile.ref + oracle/ref_driver.cpp + the Boost/OpenBabel stand-in headers of oracle/ref_shim/).
Test infrastructure only: it pins the restatement (oracle/vina_ref.c, vina_mc_ref.c) and generates tests/golden/vina_ref_kat.npz.
`available()` is False where neither the built library nor /root/reference exists (the GPU box has the prebuilt .so)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libgnina_vina_ref.so")
_fp, _ip, _vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p
_lib = None


def build():
    """compile the reference sources (only where /root/reference is present); returns True when the library exists"""
    if os.path.isdir("/root/reference/gninasrc/lib"):
        subprocess.check_call(["make", "-s", "-f", "Makefile.ref", "-j", "16"], cwd=_HERE)
    return os.path.exists(_SO)


def available():
    return os.path.exists(_SO)


def _f(a): return a.ctypes.data_as(_fp)
def _i(a): return a.ctypes.data_as(_ip)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO, mode=os.RTLD_NOW)
        L.gref_last_error.restype = C.c_char_p
        L.gref_type_info.argtypes = [C.c_int, C.c_char_p, _fp, _fp, _ip]
        L.gref_adapter_topology.argtypes = [_vp, _ip, _fp, _ip, _ip, _ip, _ip, _fp, _fp, _ip, _ip, _fp]
        L.gref_noncache_cnn_compare.argtypes = [_vp, _vp, C.c_int, _fp, _fp, _ip, C.c_float, C.c_float, C.c_float, C.c_float, _fp, C.c_int,
                                                C.c_int, C.c_float, C.c_float, C.c_int, _fp, _fp, _fp, _fp]
        L.gref_minimize_cnn.argtypes = [_vp, _vp, C.c_int, _fp, _fp, _ip, C.c_float, C.c_float, C.c_float, C.c_float, _fp, _fp, C.c_int,
                                        C.c_int, C.c_int, _fp]
        L.gref_refine_cnn.argtypes = [_vp, _vp, C.c_int, _fp, _fp, _ip, C.c_float, C.c_float, C.c_float, _fp, _fp, C.c_int, C.c_int, C.c_int,
                                      _fp, _ip]
        L.gref_minimize_dl.argtypes = [_vp, _vp, C.c_int, _fp, _fp, _ip, C.c_float, _vp, _fp, C.c_int, C.c_int, C.c_int, _fp]
        L.gref_refine_dl.argtypes = [_vp, _vp, C.c_int, _fp, _fp, _ip, _vp, _fp, C.c_int, C.c_int, C.c_int, _fp, _ip]
        L.gref_noncache_dl_eval.argtypes = [_vp, _vp, C.c_int, _fp, _fp, _ip, C.c_float, _vp, C.c_float, C.c_int, _fp, _fp, _fp]
        L.gref_lockstep_minimize.argtypes = [_vp, _fp, _fp, _ip, C.c_float, C.c_float, C.c_float, C.c_float, _fp, _fp, C.c_int, C.c_int, C.c_int,
                                             C.c_int, _fp, _ip, _ip, _ip]
        L.gref_sf_create.argtypes = [C.c_float, C.c_float]; L.gref_sf_create.restype = _vp
        L.gref_sf_create_weights.argtypes = [C.c_float, C.c_float, _fp]; L.gref_sf_create_weights.restype = _vp
        L.gref_sf_destroy.argtypes = [_vp]
        L.gref_cutoff_sqr.argtypes = [_vp]; L.gref_cutoff_sqr.restype = C.c_float
        L.gref_terms_eval.argtypes = [_vp, C.c_int, C.c_int, C.c_float]; L.gref_terms_eval.restype = C.c_float
        L.gref_prec_eval.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_float]; L.gref_prec_eval.restype = C.c_float
        L.gref_prec_eval_deriv.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_float, _fp]
        L.gref_model_create.argtypes = [C.c_int, _fp, _ip, C.c_int, _ip, _ip, _ip, _ip, C.c_int, _ip, _ip, C.c_int, _fp, _ip]
        L.gref_model_create.restype = _vp
        L.gref_model_destroy.argtypes = [_vp]
        L.gref_model_export.argtypes = [_vp, _fp, _fp, _fp]
        L.gref_model_set.argtypes = [_vp, _fp, _fp]
        L.gref_model_put_coords.argtypes = [_vp, _fp]
        L.gref_model_get_coords.argtypes = [_vp, _fp]
        L.gref_gyration_radius.argtypes = [_vp]; L.gref_gyration_radius.restype = C.c_float
        L.gref_tree_derivative.argtypes = [_vp, _fp, _fp]
        L.gref_movable_atoms_box.argtypes = [_vp, C.c_float, C.c_float, _fp, _fp, _ip]
        for nm in ("gref_cache_create", "gref_noncache_create"):
            getattr(L, nm).argtypes = [_vp, C.c_int, _vp, _fp, _fp, _ip, C.c_float]; getattr(L, nm).restype = _vp
        L.gref_cache_grid.argtypes = [_vp, C.c_int, _fp]
        L.gref_noncache_set_slope.argtypes = [_vp, C.c_float]
        L.gref_noncache_within.argtypes = [_vp, _vp, C.c_float]
        L.gref_naive_create.argtypes = [_vp, C.c_int]; L.gref_naive_create.restype = _vp
        L.gref_grid_destroy.argtypes = [_vp]
        L.gref_ig_eval.argtypes = [_vp, _vp, C.c_float, _fp]
        L.gref_ig_eval_deriv.argtypes = [_vp, _vp, C.c_float, _fp, _fp]
        L.gref_model_eval_deriv.argtypes = [_vp, _vp, C.c_int, _vp, _fp, _fp, _fp, _fp]
        L.gref_model_affinity.argtypes = [_vp, _vp, _fp, _fp, _fp, _fp]
        L.gref_num_tors_div.argtypes = [_vp, C.c_float, C.c_float, _fp]
        L.gref_bfgs.argtypes = [_vp, _vp, C.c_int, _vp, _fp, C.c_int, _fp, _fp, _fp, C.c_int, C.c_int]
        L.gref_random_conf.argtypes = [_vp, C.c_uint32, _fp, _fp, _fp, C.POINTER(C.c_uint32)]
        L.gref_mc.argtypes = [_vp, _vp, C.c_int, _vp, _fp, _fp, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                              _fp, _fp, C.c_int, _fp, _fp, C.POINTER(C.c_int)]
        L.gref_parallel_mc.argtypes = [_vp, _vp, C.c_int, _vp, _fp, _fp, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                       C.c_float, C.c_float, _fp, _fp, _fp, _fp, _ip, C.c_int, _fp, _fp, C.POINTER(C.c_int)]
        L.gref_container_replay.argtypes = [C.c_int, C.c_int, _fp, _fp, C.c_float, C.c_int, _fp, C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _ok(rc):
    if rc:
        raise RuntimeError("reference code raised: " + lib().gref_last_error().decode())


LINEAR, SPLINES, EXACT = 0, 1, 2


def type_info(t):
    """smina type t -> (name, xs_radius, covalent_radius, flags: 1 hydrophobe | 2 donor | 4 acceptor | 8 heteroatom, round-trip type)"""
    nm = C.create_string_buffer(64); r = np.empty(2, np.float32); fl = np.empty(1, np.int32)
    back = lib().gref_type_info(t, nm, _f(r[:1]), _f(r[1:]), _i(fl))
    return nm.value.decode(), float(r[0]), float(r[1]), int(fl[0]), back


class RefScoring:
    """custom_terms + weighted_terms with gnina's default Vina weights (main/main.cpp:1324-1329) and the three precalculate
    flavours: LINEAR (factor 32, docking), SPLINES (factor 10, --minimize), EXACT (final score)"""

    def __init__(self, factor_linear=32.0, factor_splines=10.0, weights6=None):
        w = None if weights6 is None else np.ascontiguousarray(weights6, np.float32)
        self.p = lib().gref_sf_create_weights(factor_linear, factor_splines, None if w is None else _f(w))
        if not self.p:
            raise RuntimeError(lib().gref_last_error().decode())

    def __del__(self):
        try:
            lib().gref_sf_destroy(self.p)
        except Exception:
            pass

    def cutoff_sqr(self): return lib().gref_cutoff_sqr(self.p)
    def terms(self, t1, t2, r): return lib().gref_terms_eval(self.p, t1, t2, r)
    def eval(self, kind, t1, t2, r2): return lib().gref_prec_eval(self.p, kind, t1, t2, r2)

    def eval_deriv(self, kind, t1, t2, r2):
        o = np.empty(2, np.float32)
        lib().gref_prec_eval_deriv(self.p, kind, t1, t2, r2, _f(o))
        return float(o[0]), float(o[1])

    def num_tors_div(self, e, num_tors):
        o = np.empty(1, np.float32)
        _ok(lib().gref_num_tors_div(self.p, e, num_tors, _f(o)))
        return float(o[0])


class RefModel:
    """a reference `model` built by hand from a ligand description (gnina_b200.synth.make_flexible_ligand's dict: xyz0, types,
    seg_parent / seg_begin / seg_end, axis_root, pair_a / pair_b) and receptor atoms"""

    def __init__(self, lig, rec_xyz=None, rec_types=None):
        k = lambda a, dt: np.ascontiguousarray(a, dt)
        self.na, self.ns = len(lig["types"]), len(lig["seg_parent"])
        self.T = self.ns - 1
        xyz, ty = k(lig["xyz0"], np.float32), k(lig["types"], np.int32)
        sp, sb, se = (k(lig[q], np.int32) for q in ("seg_parent", "seg_begin", "seg_end"))
        ar = k(lig["axis_root"], np.int32)
        pa, pb = k(lig["pair_a"], np.int32), k(lig["pair_b"], np.int32)
        rx = k(rec_xyz if rec_xyz is not None else np.zeros((0, 3)), np.float32)
        rt = k(rec_types if rec_types is not None else np.zeros(0), np.int32)
        self.p = lib().gref_model_create(self.na, _f(xyz), _i(ty), self.ns, _i(sp), _i(sb), _i(se), _i(ar), len(pa), _i(pa), _i(pb),
                                         len(rt), _f(rx), _i(rt))
        if not self.p:
            raise RuntimeError(lib().gref_last_error().decode())

    def __del__(self):
        try:
            lib().gref_model_destroy(self.p)
        except Exception:
            pass

    def export(self):
        """-> (local_xyz, seg_rel_origin, seg_rel_axis) as the reference's constructors computed them"""
        lo, ro, ra = np.empty((self.na, 3), np.float32), np.empty((self.ns, 3), np.float32), np.empty((self.ns, 3), np.float32)
        lib().gref_model_export(self.p, _f(lo), _f(ro), _f(ra))
        return lo, ro, ra

    def set(self, conf):
        x = np.ascontiguousarray(conf, np.float32); o = np.empty((self.na, 3), np.float32)
        _ok(lib().gref_model_set(self.p, _f(x), _f(o)))
        return o

    def put_coords(self, xyz):
        lib().gref_model_put_coords(self.p, _f(np.ascontiguousarray(xyz, np.float32)))

    def coords(self):
        o = np.empty((self.na, 3), np.float32)
        lib().gref_model_get_coords(self.p, _f(o))
        return o

    def gyration_radius(self): return lib().gref_gyration_radius(self.p)

    def movable_atoms_box(self, add=4.0, granularity=0.375):
        b, e, n = np.empty(3, np.float32), np.empty(3, np.float32), np.empty(3, np.int32)
        lib().gref_movable_atoms_box(self.p, add, granularity, _f(b), _f(e), _i(n))
        return b, e, n

    def adapter_topology(self):
        """integration/docking_b200.h::B200Ligand (the model -> gb_ligand_topology adapter) run on this reference model -> dict"""
        cnt = np.zeros(4, np.int32)
        _ok(lib().gref_adapter_topology(self.p, _i(cnt), None, None, None, None, None, None, None, None, None, None))
        na, ns, npair, nh = (int(v) for v in cnt)
        o = dict(local_xyz=np.empty((na, 3), np.float32), types=np.empty(na, np.int32), seg_parent=np.empty(ns, np.int32),
                 seg_begin=np.empty(ns, np.int32), seg_end=np.empty(ns, np.int32), seg_rel_origin=np.empty((ns, 3), np.float32),
                 seg_rel_axis=np.empty((ns, 3), np.float32), pair_a=np.empty(max(npair, 1), np.int32), pair_b=np.empty(max(npair, 1), np.int32))
        gr = np.empty(1, np.float32)
        _ok(lib().gref_adapter_topology(self.p, _i(cnt), _f(o["local_xyz"]), _i(o["types"]), _i(o["seg_parent"]), _i(o["seg_begin"]),
                                        _i(o["seg_end"]), _f(o["seg_rel_origin"]), _f(o["seg_rel_axis"]), _i(o["pair_a"]), _i(o["pair_b"]),
                                        _f(gr)))
        o["pair_a"], o["pair_b"] = o["pair_a"][:npair], o["pair_b"][:npair]
        o["gyration_radius"], o["n_heavy"] = float(gr[0]), nh
        return o

    def tree_derivative(self, forces):
        f = np.ascontiguousarray(forces, np.float32); g = np.empty(6 + self.T, np.float32)
        _ok(lib().gref_tree_derivative(self.p, _f(f), _f(g)))
        return g


class RefGrid:
    """an igrid of the reference: cache (populated), non_cache or naive_non_cache"""

    def __init__(self, p, model):
        if not p:
            raise RuntimeError(lib().gref_last_error().decode())
        self.p, self.m = p, model

    @classmethod
    def cache(cls, sf, kind, model, begin, end, n, slope):
        b, e, nn = (np.ascontiguousarray(a, dt) for a, dt in ((begin, np.float32), (end, np.float32), (n, np.int32)))
        g = cls(lib().gref_cache_create(sf.p, kind, model.p, _f(b), _f(e), _i(nn), slope), model)
        g.n = nn
        return g

    @classmethod
    def non_cache(cls, sf, kind, model, begin, end, n, slope):
        b, e, nn = (np.ascontiguousarray(a, dt) for a, dt in ((begin, np.float32), (end, np.float32), (n, np.int32)))
        return cls(lib().gref_noncache_create(sf.p, kind, model.p, _f(b), _f(e), _i(nn), slope), model)

    @classmethod
    def naive(cls, sf, kind, model):
        return cls(lib().gref_naive_create(sf.p, kind), model)

    def __del__(self):
        try:
            lib().gref_grid_destroy(self.p)
        except Exception:
            pass

    def grid(self, t):
        out = np.empty((self.n[2] + 1, self.n[1] + 1, self.n[0] + 1), np.float32)
        return out if lib().gref_cache_grid(self.p, t, _f(out)) else None

    def set_slope(self, s): lib().gref_noncache_set_slope(self.p, s)
    def within(self, margin=1e-4): return bool(lib().gref_noncache_within(self.p, self.m.p, margin))

    def eval(self, v):
        e = np.empty(1, np.float32)
        _ok(lib().gref_ig_eval(self.p, self.m.p, v, _f(e)))
        return float(e[0])

    def eval_deriv(self, v):
        e = np.empty(1, np.float32); f = np.empty((self.m.na, 3), np.float32)
        _ok(lib().gref_ig_eval_deriv(self.p, self.m.p, v, _f(e), _f(f)))
        return float(e[0]), f


def noncache_cnn_compare(model, sf, kind, begin, end, n, slope=10.0, dim=23.5, res=0.5, k=0.01, target=(0, 0, 0), mix_force=False,
                         mix_energy=False, weight=1.0, v=1000.0, deriv=True):
    """the REFERENCE's non_cache_cnn::eval / eval_deriv on the model's current coordinates vs this repo's gb::NonCacheCNNT
    (include/gnina_b200.hpp) on the same atoms, both around the same analytic stand-in for the network
    -> (e_ref, forces_ref, e_mine, forces_mine)"""
    b, e, nn = (np.ascontiguousarray(a, dt) for a, dt in ((begin, np.float32), (end, np.float32), (n, np.int32)))
    tg = np.ascontiguousarray(target, np.float32)
    er, em = np.empty(1, np.float32), np.empty(1, np.float32)
    fr, fm = np.zeros((model.na, 3), np.float32), np.zeros((model.na, 3), np.float32)
    _ok(lib().gref_noncache_cnn_compare(model.p, sf.p, kind, _f(b), _f(e), _i(nn), slope, dim, res, k, _f(tg), int(mix_force),
                                        int(mix_energy), weight, v, int(deriv), _f(er), _f(fr), _f(em), _f(fm)))
    return float(er[0]), fr, float(em[0]), fm


def minimize_cnn(model, sf, kind, begin, end, n, conf, maxiters, slope=10.0, dim=23.5, res=0.5, k=0.01, target=(0, 0, 0), accurate=True,
                 early_term=False):
    """the REFERENCE's quasi_newton with ig = non_cache_cnn around the analytic stand-in for the network (one --minimize run of one
    pose) -> (e, conf)"""
    b, e_, nn = (np.ascontiguousarray(a, dt) for a, dt in ((begin, np.float32), (end, np.float32), (n, np.int32)))
    x = np.array(conf, np.float32); tg = np.ascontiguousarray(target, np.float32); e = np.empty(1, np.float32)
    _ok(lib().gref_minimize_cnn(model.p, sf.p, kind, _f(b), _f(e_), _i(nn), slope, dim, res, k, _f(tg), _f(x), maxiters, int(accurate),
                                int(early_term), _f(e)))
    return float(e[0]), x


def minimize_dl(model, sf, kind, begin, end, n, conf, maxiters, dl, slope=10.0, accurate=True, early_term=False):
    """the REFERENCE's quasi_newton with ig = non_cache_cnn around ANY DLScorer* (dl: e.g. cnn_refbuild.RefCNNScorer.dl(), the
    reference's own CNNTorchScorer) -> (e, conf)"""
    b, e_, nn = (np.ascontiguousarray(a, dt) for a, dt in ((begin, np.float32), (end, np.float32), (n, np.int32)))
    x = np.array(conf, np.float32); e = np.empty(1, np.float32)
    _ok(lib().gref_minimize_dl(model.p, sf.p, kind, _f(b), _f(e_), _i(nn), slope, dl, _f(x), maxiters, int(accurate), int(early_term), _f(e)))
    return float(e[0]), x


def noncache_dl_eval(model, sf, kind, begin, end, n, dl, slope=10.0, v=1000.0, deriv=True):
    """non_cache_cnn::eval / eval_deriv around ANY DLScorer* on the pose the model holds (after adjust_center)
    -> (e, minus_forces [n_movable, 3] or None, CNN box centre [3])"""
    b, e_, nn = (np.ascontiguousarray(a, dt) for a, dt in ((begin, np.float32), (end, np.float32), (n, np.int32)))
    e = np.empty(1, np.float32); f = np.zeros((model.na, 3), np.float32); c = np.zeros(3, np.float32)
    _ok(lib().gref_noncache_dl_eval(model.p, sf.p, kind, _f(b), _f(e_), _i(nn), slope, dl, v, int(deriv), _f(e), _f(f), _f(c)))
    return float(e[0]), (f if deriv else None), c


def refine_dl(model, sf, kind, begin, end, n, conf, maxiters, dl, accurate=False, early_term=False):
    """refine_structure replayed with the reference's parts around ANY DLScorer* -> (e or max_fl, conf, inside)"""
    b, e_, nn = (np.ascontiguousarray(a, dt) for a, dt in ((begin, np.float32), (end, np.float32), (n, np.int32)))
    x = np.array(conf, np.float32); e = np.empty(1, np.float32); ins = np.zeros(1, np.int32)
    _ok(lib().gref_refine_dl(model.p, sf.p, kind, _f(b), _f(e_), _i(nn), dl, _f(x), maxiters, int(accurate), int(early_term), _f(e), _i(ins)))
    return float(e[0]), x, bool(ins[0])


def refine_cnn(model, sf, kind, begin, end, n, conf, maxiters, dim=23.5, res=0.5, k=0.01, target=(0, 0, 0), accurate=False, early_term=False):
    """refine_structure (main/main.cpp:131-171) replayed with the reference's quasi_newton / non_cache_cnn / within around the analytic
    stand-in for the network -> (e or max_fl, conf, inside)"""
    b, e_, nn = (np.ascontiguousarray(a, dt) for a, dt in ((begin, np.float32), (end, np.float32), (n, np.int32)))
    x = np.array(conf, np.float32); tg = np.ascontiguousarray(target, np.float32); e = np.empty(1, np.float32); ins = np.zeros(1, np.int32)
    _ok(lib().gref_refine_cnn(model.p, sf.p, kind, _f(b), _f(e_), _i(nn), dim, res, k, _f(tg), _f(x), maxiters, int(accurate), int(early_term),
                              _f(e), _i(ins)))
    return float(e[0]), x, bool(ins[0])


def lockstep_minimize_cpp(model, begin, end, n, confs, maxiters, slope=10.0, dim=23.5, res=0.5, k=0.01, target=(0, 0, 0), accurate=True,
                          early_term=False):
    """this repo's C++ lock-step minimiser (include/gnina_b200_minimize.hpp: gb::minimize_poses over gb::LigandTree, energy =
    gb::NonCacheCNNT around the analytic stand-in) on all conformations at once -> (e [n], confs [n, 7+T], evals [n], rounds, energy calls)"""
    b, e_, nn = (np.ascontiguousarray(a, dt) for a, dt in ((begin, np.float32), (end, np.float32), (n, np.int32)))
    x = np.array(confs, np.float32); tg = np.ascontiguousarray(target, np.float32)
    e = np.empty(len(x), np.float32); ev = np.zeros(len(x), np.int32); rr = np.zeros(2, np.int32)
    _ok(lib().gref_lockstep_minimize(model.p, _f(b), _f(e_), _i(nn), slope, dim, res, k, _f(tg), _f(x), len(x), maxiters, int(accurate),
                                     int(early_term), _f(e), _i(ev), _i(rr[:1]), _i(rr[1:])))
    return e, x, ev, int(rr[0]), int(rr[1])


def model_eval_deriv(model, sf, kind, grid, conf, v=(1000, 1000, 1000)):
    x = np.ascontiguousarray(conf, np.float32); vv = np.ascontiguousarray(v, np.float32)
    e = np.empty(1, np.float32); g = np.empty(6 + model.T, np.float32)
    _ok(lib().gref_model_eval_deriv(model.p, sf.p, kind, grid.p, _f(vv), _f(x), _f(e), _f(g)))
    return float(e[0]), g


def model_affinity(model, sf, conf, v=(1000, 1000, 1000)):
    """-> (intramolecular energy, eval_adjusted = the printed Affinity), exact terms (main/main.cpp:219-232)"""
    x = np.ascontiguousarray(conf, np.float32); vv = np.ascontiguousarray(v, np.float32)
    a, b = np.empty(1, np.float32), np.empty(1, np.float32)
    _ok(lib().gref_model_affinity(model.p, sf.p, _f(vv), _f(x), _f(a), _f(b)))
    return float(a[0]), float(b[0])


def bfgs(model, sf, kind, grid, conf, maxiters, v=(1000, 1000, 1000), accurate=False, early_term=False):
    x = np.array(conf, np.float32); vv = np.ascontiguousarray(v, np.float32)
    e = np.empty(1, np.float32); g = np.empty(6 + model.T, np.float32)
    _ok(lib().gref_bfgs(model.p, sf.p, kind, grid.p, _f(x), maxiters, _f(vv), _f(e), _f(g), int(accurate), int(early_term)))
    return float(e[0]), x, g


def random_conf(model, seed, c1, c2):
    x = np.empty(7 + model.T, np.float32); st = C.c_uint32()
    _ok(lib().gref_random_conf(model.p, seed, _f(np.ascontiguousarray(c1, np.float32)), _f(np.ascontiguousarray(c2, np.float32)), _f(x),
                               C.byref(st)))
    return x, st.value


def mc(model, sf, kind, grid, seed, c1, c2, num_steps, maxiters, state_conf, num_saved_mins=50, temperature=1.2, amplitude=2.0,
       min_rmsd=1.0, hunt_cap=(10, 10, 10)):
    """monte_carlo::operator(); state_conf = the conformation the model object holds when the chain starts"""
    e = np.zeros(num_saved_mins, np.float32); x = np.zeros((num_saved_mins, 7 + model.T), np.float32); n = C.c_int()
    hc = np.ascontiguousarray(hunt_cap, np.float32); sc = np.ascontiguousarray(state_conf, np.float32)
    _ok(lib().gref_mc(model.p, sf.p, kind, grid.p, _f(np.ascontiguousarray(c1, np.float32)), _f(np.ascontiguousarray(c2, np.float32)), seed,
                      num_steps, maxiters, num_saved_mins, temperature, amplitude, min_rmsd, _f(hc), _f(sc), num_saved_mins, _f(e), _f(x),
                      C.byref(n)))
    return e[:n.value], x[:n.value]


def parallel_mc(model, sf, kind, grid, seed, c1, c2, num_tasks, num_steps, maxiters, state_conf, box, num_threads=2, num_saved_mins=50,
                temperature=1.2, amplitude=2.0, min_rmsd=1.0, hunt_cap=(10, 10, 10)):
    """parallel_mc::operator() (lib/parallel_mc.cpp:183-214) on the reference's own thread pool -> merged, sorted container (e, confs);
    box = (begin, end, n) of the non_cache that is passed through"""
    e = np.zeros(num_saved_mins, np.float32); x = np.zeros((num_saved_mins, 7 + model.T), np.float32); n = C.c_int()
    k = lambda a, dt: np.ascontiguousarray(a, dt)
    hc, sc = k(hunt_cap, np.float32), k(state_conf, np.float32)
    bb, be, bn = k(box[0], np.float32), k(box[1], np.float32), k(box[2], np.int32)
    _ok(lib().gref_parallel_mc(model.p, sf.p, kind, grid.p, _f(k(c1, np.float32)), _f(k(c2, np.float32)), seed, num_tasks, num_threads,
                               num_steps, maxiters, num_saved_mins, temperature, amplitude, min_rmsd, _f(hc), _f(sc), _f(bb), _f(be), _i(bn),
                               num_saved_mins, _f(e), _f(x), C.byref(n)))
    return e[:n.value], x[:n.value]


def task_seeds(seed, num_tasks):
    """the seeds parallel_mc draws for its tasks: random_int(0, 1000000, generator) on the stand-in generator (xorshift32, a + next % range)"""
    s, out = (seed or 1) & 0xFFFFFFFF, []
    for _ in range(num_tasks):
        s ^= (s << 13) & 0xFFFFFFFF; s ^= s >> 17; s ^= (s << 5) & 0xFFFFFFFF
        out.append(s % 1000001)
    return out


def container_replay(e, coords, min_rmsd, max_size):
    e = np.ascontiguousarray(e, np.float32); c = np.ascontiguousarray(coords, np.float32)
    out = np.empty(len(e), np.float32); n = C.c_int()
    _ok(lib().gref_container_replay(len(e), c.shape[1], _f(e), _f(c), min_rmsd, max_size, _f(out), C.byref(n)))
    return out[:n.value]
