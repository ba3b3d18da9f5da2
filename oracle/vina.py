"""ctypes front-end to oracle/vina_ref.c (CPU oracle of the Vina scoring rows; test infrastructure only)."""
import ctypes as C
import numpy as np
from . import gridmaker as _gm

_fp, _ip, _vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p
_ready = False


def lib():
    global _ready
    L = _gm.lib()
    if not _ready:
        L.gvo_eval_terms.argtypes = [_fp, C.c_int, C.c_int, C.c_float]; L.gvo_eval_terms.restype = C.c_float
        L.gvo_prec_create.argtypes = [_fp, C.c_float]; L.gvo_prec_create.restype = _vp
        L.gvo_prec_free.argtypes = [_vp]
        L.gvo_prec_n.argtypes = [_vp]
        L.gvo_prec_table.argtypes = [_vp, C.c_int, C.c_int, _fp, _fp, _fp]
        L.gvo_prec_eval_fast.argtypes = [_vp, C.c_int, C.c_int, C.c_float]; L.gvo_prec_eval_fast.restype = C.c_float
        L.gvo_prec_eval_deriv.argtypes = [_vp, C.c_int, C.c_int, C.c_float, _fp, _fp]
        L.gvo_exact_eval.argtypes = [_vp, C.c_int, C.c_int, C.c_float]; L.gvo_exact_eval.restype = C.c_float
        L.gvo_cache_populate.argtypes = [_vp, _fp, _fp, _ip, C.c_int, _fp, _ip, C.c_int, _fp]
        L.gvo_grid_evaluate.argtypes = [_fp, _fp, _fp, _ip, _fp, C.c_float, C.c_float, _fp]; L.gvo_grid_evaluate.restype = C.c_float
        L.gvo_cache_eval.argtypes = [C.POINTER(_fp), _fp, _fp, _ip, C.c_int, _fp, _ip, C.c_float, C.c_float, _fp]
        L.gvo_cache_eval.restype = C.c_float
        L.gvo_naive_exact.argtypes = [_vp, C.c_int, _fp, _ip, C.c_int, _fp, _ip, C.c_float]; L.gvo_naive_exact.restype = C.c_float
        L.gvo_noncache_eval.argtypes = [_vp, C.c_int, _fp, _ip, C.c_int, _fp, _ip, C.c_float, C.c_float, _fp, _fp]
        L.gvo_noncache_eval.restype = C.c_float
        L.gvo_num_tors_div.argtypes = [_vp, C.c_float, C.c_float]; L.gvo_num_tors_div.restype = C.c_float
        _ready = True
    return L


def _f(a): return a.ctypes.data_as(_fp)
def _i(a): return a.ctypes.data_as(_ip)


class VinaOracle:
    def __init__(self, weights6=None, factor=32.0):
        w = None if weights6 is None else np.ascontiguousarray(weights6, np.float32)
        self.p = lib().gvo_prec_create(None if w is None else _f(w), factor)
        self.n = lib().gvo_prec_n(self.p)

    def __del__(self):
        try:
            lib().gvo_prec_free(self.p)
        except Exception:
            pass

    def splines(self, factor=10.0):
        L = lib()
        L.gvo_splines_create.argtypes = [_vp, C.c_float]; L.gvo_splines_create.restype = _vp
        L.gvo_splines_n.argtypes = [_vp]
        L.gvo_splines_table.argtypes = [_vp, C.c_int, C.c_int, _fp]
        L.gvo_splines_eval_deriv.argtypes = [_vp, C.c_int, C.c_int, C.c_float, _fp, _fp]
        if not hasattr(self, "_sp"):
            self._sp = L.gvo_splines_create(self.p, factor)
        return self._sp

    def spline_table(self, t1, t2):
        sp = self.splines()
        out = np.empty((lib().gvo_splines_n(sp), 4), np.float32)
        lib().gvo_splines_table(sp, t1, t2, _f(out))
        return out

    def spline_eval_deriv(self, t1, t2, r2):
        e, d = C.c_float(), C.c_float()
        lib().gvo_splines_eval_deriv(self.splines(), t1, t2, r2, C.byref(e), C.byref(d))
        return e.value, d.value

    def table(self, t1, t2):
        a, b, c = (np.empty(self.n, np.float32) for _ in range(3))
        lib().gvo_prec_table(self.p, t1, t2, _f(a), _f(b), _f(c))
        return a, b, c

    def eval_fast(self, t1, t2, r2): return lib().gvo_prec_eval_fast(self.p, t1, t2, r2)

    def eval_deriv(self, t1, t2, r2):
        e, d = C.c_float(), C.c_float()
        lib().gvo_prec_eval_deriv(self.p, t1, t2, r2, C.byref(e), C.byref(d))
        return e.value, d.value

    def exact(self, t1, t2, r2): return lib().gvo_exact_eval(self.p, t1, t2, r2)

    def cache_populate(self, begin, end, n, rec_xyz, rec_types, t2):
        begin, end = np.ascontiguousarray(begin, np.float32), np.ascontiguousarray(end, np.float32)
        n = np.ascontiguousarray(n, np.int32)
        rx, rt = np.ascontiguousarray(rec_xyz, np.float32), np.ascontiguousarray(rec_types, np.int32)
        out = np.empty((n[2] + 1, n[1] + 1, n[0] + 1), np.float32)  # x fastest
        lib().gvo_cache_populate(self.p, _f(begin), _f(end), _i(n), len(rt), _f(rx), _i(rt), t2, _f(out))
        return out

    @staticmethod
    def cache_eval(grids, begin, end, n, lig_xyz, lig_types, slope, v, want_deriv=True):
        begin, end = np.ascontiguousarray(begin, np.float32), np.ascontiguousarray(end, np.float32)
        n = np.ascontiguousarray(n, np.int32)
        lx, lt = np.ascontiguousarray(lig_xyz, np.float32), np.ascontiguousarray(lig_types, np.int32)
        ptrs = (_fp * 28)()
        keep = []
        for t, g in grids.items():
            g = np.ascontiguousarray(g, np.float32); keep.append(g)
            ptrs[t] = _f(g)
        d = np.zeros((len(lt), 3), np.float32)
        e = lib().gvo_cache_eval(ptrs, _f(begin), _f(end), _i(n), len(lt), _f(lx), _i(lt), slope, v, _f(d) if want_deriv else None)
        return e, d

    def naive_exact(self, rec_xyz, rec_types, lig_xyz, lig_types, v=1000.0):
        rx, rt = np.ascontiguousarray(rec_xyz, np.float32), np.ascontiguousarray(rec_types, np.int32)
        lx, lt = np.ascontiguousarray(lig_xyz, np.float32), np.ascontiguousarray(lig_types, np.int32)
        return lib().gvo_naive_exact(self.p, len(rt), _f(rx), _i(rt), len(lt), _f(lx), _i(lt), v)

    def noncache_eval(self, rec_xyz, rec_types, lig_xyz, lig_types, begin, end, slope=1e3, v=1000.0):
        """non_cache::eval (lib/non_cache.cpp:52-83): the docking branch's intermolecular energy (table terms, box clamp + penalty)"""
        rx, rt = np.ascontiguousarray(rec_xyz, np.float32), np.ascontiguousarray(rec_types, np.int32)
        lx, lt = np.ascontiguousarray(lig_xyz, np.float32), np.ascontiguousarray(lig_types, np.int32)
        b, e = np.ascontiguousarray(begin, np.float32), np.ascontiguousarray(end, np.float32)
        return lib().gvo_noncache_eval(self.p, len(rt), _f(rx), _i(rt), len(lt), _f(lx), _i(lt), v, slope, _f(b), _f(e))

    def num_tors_div(self, e, num_tors): return lib().gvo_num_tors_div(self.p, e, num_tors)
