"""ctypes front-end to oracle/_ref/libgnina_cnn_ref.so: the REFERENCE's own CNN scoring host code (lib/torch_model.cpp,
lib/cnn_torch_scorer.cpp, lib/dl_scorer.cpp) compiled where it lies, running the reference's own TorchScript files with libtorch on the
CPU; libmolgrid (third party, absent) is replaced by a stand-in over oracle/gridmaker_ref.c (oracle/ref_shim/libmolgrid).  TEST
INFRASTRUCTURE: pins oracle/pipeline.py and generates tests/golden/cnn_ref_kat.npz.  The same library executes the integration adapters
(integration/*.h) and the product's C++ host classes (include/*.hpp) on the CPU inside the reference's classes, their C-ABI calls
served by stand-ins that honour include/gnina_b200.h's contract (ref_cnn_driver.cpp: the reference's TorchModel as the network;
ref_adapters_driver.cpp: the C restatement of the Vina rows).  Exists only where /root/reference does (the
TorchScript files are read from there), i.e. in the build container -- not on the GPU box."""
import ctypes as C
import os
import subprocess
import numpy as np

from . import vina_refbuild as V

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libgnina_cnn_ref.so")
MODELS_DIR = "/root/reference/gninasrc/lib/models"
_LIB = None


def available():
    return os.path.exists(_SO) and os.path.isdir(MODELS_DIR)


def build():
    """-> True if the library exists afterwards (needs /root/reference)"""
    if not os.path.isdir(MODELS_DIR) or not V.build():
        return False
    from . import gridmaker
    gridmaker.lib()
    subprocess.check_call(["make", "-s", "-f", "Makefile.ref", "-j", "4", "cnn"], cwd=_HERE)
    return os.path.exists(_SO)


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def lib():
    global _LIB
    if _LIB is None:
        import torch  # noqa: F401  (libtorch's shared objects are resolved through the interpreter's copy)
        V.lib()
        L = C.CDLL(_SO, mode=C.RTLD_GLOBAL)
        fp = C.POINTER(C.c_float)
        L.gcref_last_error.restype = C.c_char_p
        L.gcref_load_models.argtypes = [C.c_char_p]
        L.gcref_scorer_create.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_uint, C.c_uint, fp]
        L.gcref_scorer_create.restype = C.c_void_p
        L.gcref_scorer_destroy.argtypes = [C.c_void_p]
        L.gcref_scorer_dl.argtypes = [C.c_void_p]; L.gcref_scorer_dl.restype = C.c_void_p
        L.gcref_adapter_create.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_char_p, fp, C.c_int]
        L.gcref_adapter_create.restype = C.c_void_p
        ip = C.POINTER(C.c_int)
        L.gadp_vina_create.argtypes = [fp, ip, C.c_int]; L.gadp_vina_create.restype = C.c_void_p
        L.gadp_vina_destroy.argtypes = [C.c_void_p]
        L.gadp_cache_b200.argtypes = [C.c_void_p, fp, fp, ip, C.c_float, ip, C.c_int]; L.gadp_cache_b200.restype = C.c_void_p
        L.gadp_score_docked.argtypes = [C.c_void_p, C.c_void_p, fp, C.c_int, fp, fp, ip, C.c_float, fp, C.c_float, fp]
        L.gadp_refine_structure.argtypes = [C.c_void_p, C.c_void_p, fp, C.c_int, fp, fp, ip, fp, C.c_int, C.c_int, C.c_int, fp]
        L.gadp_parallel_mc.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, fp, fp, C.c_int,
                                       C.c_int, fp, fp, ip]
        L.gadp_last_error.restype = C.c_char_p
        L.gcref_product_noncache_cnn.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_void_p, fp, fp, ip, C.c_float, fp, C.c_float,
                                                 C.c_int, C.c_int, fp, fp]
        L.gcref_product_lockstep_minimize.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_void_p, fp, fp, C.c_float, fp, C.c_int,
                                                      C.c_int, C.c_int, C.c_int, C.c_int, fp, ip, ip]
        L.gcref_grid_dim.argtypes = [C.c_void_p]; L.gcref_grid_dim.restype = C.c_float
        L.gcref_grid_res.argtypes = [C.c_void_p]; L.gcref_grid_res.restype = C.c_float
        L.gcref_score.argtypes = [C.c_void_p, C.c_void_p, C.c_int, fp, fp]
        L.gcref_center_and_box.argtypes = [C.c_void_p, C.c_void_p, fp, fp, fp, C.POINTER(C.c_int)]
        if L.gcref_load_models(MODELS_DIR.encode()) <= 0:
            raise RuntimeError("no TorchScript files under " + MODELS_DIR)
        _LIB = L
    return _LIB


class RefCNNScorer:
    """CNNTorchScorer<false>(cnn_options{cnn_model_names = names, cnn_models = files})"""

    def __init__(self, names=(), files=(), rotations=0, seed=0, cnn_center=None):
        L = lib()
        na = (C.c_char_p * max(1, len(names)))(*[n.encode() for n in names])
        fa = (C.c_char_p * max(1, len(files)))(*[f.encode() for f in files])
        cc = None if cnn_center is None else _f(np.ascontiguousarray(cnn_center, np.float32))
        self.p = L.gcref_scorer_create(na, len(names), fa, len(files), rotations, seed, cc)
        if not self.p:
            raise RuntimeError(L.gcref_last_error().decode())

    @classmethod
    def adapter(cls, names=(), files=(), cnn_center=None, blob_dir=None, fresh_copy=False):
        """integration/cnn_b200_scorer.h's CNNB200Scorer -- the class a gnina maintainer adds -- constructed over a stand-in for the twelve
        C-ABI entry points it calls (oracle/ref_cnn_driver.cpp: the ABI's contract with the reference's TorchModel as the network), so
        that the adapter's own code runs on the CPU inside the reference's classes.  Same methods as the reference scorer."""
        L = lib()
        self = cls.__new__(cls)
        na = (C.c_char_p * max(1, len(names)))(*[n.encode() for n in names])
        fa = (C.c_char_p * max(1, len(files)))(*[f.encode() for f in files])
        cc = None if cnn_center is None else _f(np.ascontiguousarray(cnn_center, np.float32))
        bd = blob_dir or os.path.join(os.path.dirname(_HERE), "gnina_b200", "weights")
        self.p = L.gcref_adapter_create(na, len(names), fa, len(files), bd.encode(), cc, int(fresh_copy))
        if not self.p:
            raise RuntimeError(L.gcref_last_error().decode())
        return self

    def __del__(self):
        try:
            lib().gcref_scorer_destroy(self.p)
        except Exception:
            pass

    def dl(self):
        """the scorer as a DLScorer* for vina_refbuild.minimize_dl / refine_dl"""
        return lib().gcref_scorer_dl(self.p)

    def grid(self):
        return lib().gcref_grid_dim(self.p), lib().gcref_grid_res(self.p)

    def score(self, ref_model, gradient=False):
        """score(m, compute_gradient, affinity, loss, variance) on the coordinates `ref_model` (vina_refbuild.RefModel) holds
        -> (score, affinity, loss, variance, minus_forces[n_movable,3] as left in the model)"""
        o = np.zeros(4, np.float32); mf = np.zeros((ref_model.na, 3), np.float32)
        if lib().gcref_score(self.p, ref_model.p, int(gradient), _f(o), _f(mf)):
            raise RuntimeError(lib().gcref_last_error().decode())
        return float(o[0]), float(o[1]), float(o[2]), float(o[3]), mf

    def center_and_box(self, ref_model):
        c, b, e = (np.zeros(3, np.float32) for _ in range(3)); n = (C.c_int * 3)()
        if lib().gcref_center_and_box(self.p, ref_model.p, _f(c), _f(b), _f(e), n):
            raise RuntimeError(lib().gcref_last_error().decode())
        return c, b, e, np.array(list(n))


def product_noncache_cnn(names, ref_model, begin, end, n, center, slope=10.0, v=1000.0, deriv=True, reference_force_routing=True):
    """the product's own C++ host classes -- gb::CNNScorer + gb::NonCacheCNN of include/gnina_b200.hpp -- run on the CPU over the stand-in
    C ABI (network = the reference's TorchModel) on the pose `ref_model` holds -> (e, forces [n_movable, 3] or None)"""
    L = lib()
    na = (C.c_char_p * max(1, len(names)))(*[x.encode() for x in names])
    b, e_, c = (np.ascontiguousarray(a, np.float32) for a in (begin, end, center))
    nn = np.ascontiguousarray(n, np.int32)
    e = np.empty(1, np.float32); f = np.zeros((ref_model.na, 3), np.float32)
    wd = os.path.join(os.path.dirname(_HERE), "gnina_b200", "weights")
    if L.gcref_product_noncache_cnn(na, len(names), wd.encode(), ref_model.p, _f(b), _f(e_), nn.ctypes.data_as(C.POINTER(C.c_int)), slope,
                                    _f(c), v, int(deriv), int(reference_force_routing), _f(e), _f(f)):
        raise RuntimeError(L.gcref_last_error().decode())
    return float(e[0]), (f if deriv else None)


def product_lockstep_minimize(names, ref_model, begin, end, confs, maxiters, slope=10.0, accurate=True, early_term=False,
                              reference_force_routing=True):
    """config 5 through the product's C++ host code (gb::CNNScorer, gb::LigandTree, gb::CnnBatchEnergy, gb::minimize_poses) on the CPU over
    the stand-in C ABI -> (e [n], confs [n, 7+T], evaluations [n], rounds)"""
    L = lib()
    na = (C.c_char_p * max(1, len(names)))(*[x.encode() for x in names])
    b, e_ = (np.ascontiguousarray(a, np.float32) for a in (begin, end))
    x = np.array(confs, np.float32); e = np.zeros(len(x), np.float32); ev = np.zeros(len(x), np.int32); rounds = (C.c_int * 1)()
    wd = os.path.join(os.path.dirname(_HERE), "gnina_b200", "weights")
    if L.gcref_product_lockstep_minimize(na, len(names), wd.encode(), ref_model.p, _f(b), _f(e_), slope, _f(x), len(x), maxiters, int(accurate),
                                         int(early_term), int(reference_force_routing), _f(e), ev.ctypes.data_as(C.POINTER(C.c_int)), rounds):
        raise RuntimeError(L.gcref_last_error().decode())
    return e, x, ev, rounds[0]


class VinaAdapters:
    """integration/docking_b200.h's docking-side adapters EXECUTED on the CPU (oracle/ref_adapters_driver.cpp): their C-ABI calls are served
    by a stand-in that honours include/gnina_b200.h's contract with the oracle's C restatement, so that b200::cache_b200 (an `igrid`) and
    b200::score_docked_b200 run inside the reference's own classes"""

    def __init__(self, rec_xyz, rec_types):
        rx, rt = np.ascontiguousarray(rec_xyz, np.float32), np.ascontiguousarray(rec_types, np.int32)
        self.h = lib().gadp_vina_create(_f(rx), rt.ctypes.data_as(C.POINTER(C.c_int)), len(rt))

    def __del__(self):
        try:
            lib().gadp_vina_destroy(self.h)
        except Exception:
            pass

    def cache_b200(self, ref_model, begin, end, n, slope, needed):
        """b200::cache_b200 as a vina_refbuild.RefGrid: model_eval_deriv / bfgs / mc of the reference take it in place of `cache`"""
        b, e = np.ascontiguousarray(begin, np.float32), np.ascontiguousarray(end, np.float32)
        nn, nd = np.ascontiguousarray(n, np.int32), np.ascontiguousarray(needed, np.int32)
        ip = C.POINTER(C.c_int)
        p = lib().gadp_cache_b200(self.h, _f(b), _f(e), nn.ctypes.data_as(ip), slope, nd.ctypes.data_as(ip), len(nd))
        if not p:
            raise RuntimeError(lib().gadp_last_error().decode())
        g = V.RefGrid(p, ref_model)
        g.keep = self
        return g

    def score_docked(self, ref_model, pose_xyz, begin, end, n, slope, cap3, num_tors, e_in):
        b, e = np.ascontiguousarray(begin, np.float32), np.ascontiguousarray(end, np.float32)
        nn = np.ascontiguousarray(n, np.int32)
        xyz = np.ascontiguousarray(pose_xyz, np.float32); out = np.array(e_in, np.float32)
        cap = np.ascontiguousarray(cap3, np.float32)
        if lib().gadp_score_docked(self.h, ref_model.p, _f(xyz), len(out), _f(b), _f(e), nn.ctypes.data_as(C.POINTER(C.c_int)), slope,
                                   _f(cap), num_tors, _f(out)):
            raise RuntimeError(lib().gadp_last_error().decode())
        return out

    def refine_structure(self, ref_model, confs, begin, end, n, cap3, maxiters, accurate=False, early_term=False):
        """b200::refine_structure_b200 on all conformations at once -> (e [k] (max_fl = never inside), confs [k, 7+T])"""
        b, e_ = np.ascontiguousarray(begin, np.float32), np.ascontiguousarray(end, np.float32)
        nn, cap = np.ascontiguousarray(n, np.int32), np.ascontiguousarray(cap3, np.float32)
        x = np.array(confs, np.float32); e = np.zeros(len(x), np.float32)
        if lib().gadp_refine_structure(self.h, ref_model.p, _f(x), len(x), _f(b), _f(e_), nn.ctypes.data_as(C.POINTER(C.c_int)), _f(cap),
                                       maxiters, int(accurate), int(early_term), _f(e)):
            raise RuntimeError(lib().gadp_last_error().decode())
        return e, x

    def parallel_mc(self, ref_model, seed, c1, c2, num_tasks, num_steps, maxiters, state_conf, num_saved_mins=50, min_rmsd=1.0,
                    hunt_cap=(10, 10, 10)):
        """b200::parallel_mc_b200::operator() (needs cache_b200 built on this handle; the model must hold state_conf)
        -> (e [k], confs [k, 7+T]) of the merged container"""
        a, b = np.ascontiguousarray(c1, np.float32), np.ascontiguousarray(c2, np.float32)
        hc, st = np.ascontiguousarray(hunt_cap, np.float32), np.ascontiguousarray(state_conf, np.float32)
        cap = num_saved_mins
        e = np.zeros(cap, np.float32); x = np.zeros((cap, len(st)), np.float32); n = (C.c_int * 1)()
        if lib().gadp_parallel_mc(self.h, ref_model.p, seed, _f(a), _f(b), num_tasks, num_steps, maxiters, num_saved_mins, min_rmsd, _f(hc),
                                  _f(st), len(st), cap, _f(e), _f(x), n):
            raise RuntimeError(lib().gadp_last_error().decode())
        return e[:n[0]].copy(), x[:n[0]].copy()
