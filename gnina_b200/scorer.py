"""Host-side mirror of the reference's CNN scorer interface over the C ABI.

`CNNScorer` follows `CNNTorchScorer<isCUDA>` (gninasrc/lib/cnn_torch_scorer.{h,cpp}) / `DLScorer`
(gninasrc/lib/dl_scorer.h:23-66): same model-name expansion, same four outputs (score, affinity, loss, variance),
same error behaviour ("Invalid model name: X" for unknown names), `fresh_copy()` for another thread.  The batch
entry points are what a batching front-end inside gnina's ligand loop calls; `score()` with one
pose is the drop-in for `DLScorer::score`.
"""
import ctypes as C
import os
import numpy as np
from . import capi
from .model_blob import WEIGHTS_DIR

PRECISION_FP32 = 0
PRECISION_FP16_TC = 1


class usage_error(ValueError):
    """reference: usage_error (gninasrc/lib/common.h) — bad model name / file"""


# the reference's built-in model table (gninasrc/lib/torch_models.h = gninasrc/lib/models/*.pt), '.' -> '_'
REFERENCE_MODELS = tuple(sorted(n.replace(".", "_") for n in """
all_default_to_default_1.3_1 all_default_to_default_1.3_2 all_default_to_default_1.3_3 crossdock_default2018
crossdock_default2018_1.3 crossdock_default2018_1.3_1 crossdock_default2018_1.3_2 crossdock_default2018_1.3_3
crossdock_default2018_1.3_4 crossdock_default2018_1 crossdock_default2018_2 crossdock_default2018_3 crossdock_default2018_4
crossdock_default2018_KD_1 crossdock_default2018_KD_2 crossdock_default2018_KD_3 crossdock_default2018_KD_4
crossdock_default2018_KD_5 default2017 dense dense_1.3 dense_1.3_1 dense_1.3_2 dense_1.3_3 dense_1.3_4 dense_1.3_PT_KD
dense_1.3_PT_KD_1 dense_1.3_PT_KD_2 dense_1.3_PT_KD_3 dense_1.3_PT_KD_4 dense_1.3_PT_KD_def2018 dense_1.3_PT_KD_def2018_1
dense_1.3_PT_KD_def2018_2 dense_1.3_PT_KD_def2018_3 dense_1.3_PT_KD_def2018_4 dense_1 dense_2 dense_3 dense_4
general_default2018 general_default2018_1 general_default2018_2 general_default2018_3 general_default2018_4
general_default2018_KD_1 general_default2018_KD_2 general_default2018_KD_3 general_default2018_KD_4 general_default2018_KD_5
redock_default2018 redock_default2018_1.3 redock_default2018_1.3_1 redock_default2018_1.3_2 redock_default2018_1.3_3
redock_default2018_1.3_4 redock_default2018_1 redock_default2018_2 redock_default2018_3 redock_default2018_4
redock_default2018_KD_1 redock_default2018_KD_2 redock_default2018_KD_3 redock_default2018_KD_4 redock_default2018_KD_5
""".split()))


def builtin_models():
    """names of the packaged built-in models (a subset of REFERENCE_MODELS until every blob is extracted with
    tools/extract_models.py), like builtin_torch_models() (gninasrc/lib/torch_models.h)"""
    return sorted(f[:-4] for f in os.listdir(WEIGHTS_DIR) if f.endswith(".gbw"))


def expand_model_names(names, check=True):
    """cnn_torch_scorer.cpp:28-62: default ensemble, 'fast', 'default1.0', '<prefix>_ensemble'.  An ensemble expands over
    the REFERENCE's model table, never over "whatever is packaged": a missing member is an error, not a smaller
    ensemble with different scores."""
    names = [n.replace(".", "_") if n not in ("default1.0",) else n for n in names]
    if len(names) == 0:
        names = ["dense_1_3", "dense_1_3_PT_KD_3", "crossdock_default2018_KD_4"]
    elif len(names) == 1:
        if names[0] == "fast":
            names = ["all_default_to_default_1_3_1"]
        elif names[0] == "default1.0":
            names = ["dense", "general_default2018_3", "dense_3", "crossdock_default2018", "redock_default2018_2"]
    out = []
    for n in names:
        if n.endswith("_ensemble"):
            prefix = n[: -len("_ensemble")]
            out += [a for a in REFERENCE_MODELS if a.startswith(prefix)]   # std::map order = sorted
        else:
            out.append(n)
    if not check:
        return out
    avail = set(builtin_models())
    missing = [n for n in out if n in REFERENCE_MODELS and n not in avail]
    if missing:
        raise usage_error("built-in model(s) not packaged in gnina_b200/weights: %s (run tools/extract_models.py on the "
                         "reference's gninasrc/lib/models)" % ", ".join(missing))
    return out


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class CNNScorer:
    def __init__(self, cnn_model_names=(), cnn_models=(), device=0, precision=None, max_batch=None, _clone_of=None):
        L = capi.lib()
        self.device = device
        self._h = C.c_void_p()
        self._models = []
        self._n_rec = 0
        self._parent = _clone_of                       # a clone shares the parent's models: model_info() asks the parent
        if _clone_of is not None:
            capi.check(L.gb_cnn_clone(_clone_of._h, C.byref(self._h)))
            self.model_names = list(_clone_of.model_names)
            self._n_rec = _clone_of._n_rec          # the clone carries the receptor
            return
        rc = L.gb_initialize_cuda(device)
        if rc != 0:
            raise capi.GbError(4, "no usable CUDA device %d (cuda error %d); gnina_b200 has no CPU fallback" % (device, rc))
        self.model_names = expand_model_names(list(cnn_model_names)) if not cnn_models or cnn_model_names else []
        paths = []
        for n in self.model_names:
            p = os.path.join(WEIGHTS_DIR, n + ".gbw")
            if not os.path.exists(p):
                raise usage_error("Invalid model name: " + n)          # cnn_torch_scorer.cpp:70-72
            paths.append(p)
        for f in cnn_models:                                           # external model files, :84-89
            if not os.path.exists(f):
                raise usage_error("Could not open file " + f)
            paths.append(f)
            self.model_names.append(os.path.basename(f))
        for p in paths:
            m = C.c_void_p()
            capi.check(L.gb_model_load(p.encode(), device, C.byref(m)))
            self._models.append(m)
        arr = (C.c_void_p * len(self._models))(*[m.value for m in self._models])
        capi.check(L.gb_cnn_create(arr, len(self._models), device, C.byref(self._h)))
        if precision is not None:
            self.set_option("precision", precision)
        if max_batch is not None:
            self.set_option("max_batch", max_batch)

    # -- DLScorer surface ------------------------------------------------------------------------------
    def initialized(self):
        return capi.lib().gb_cnn_num_models(self._h) > 0

    def has_affinity(self):
        return True

    def fresh_copy(self):
        return CNNScorer(_clone_of=self)

    def set_option(self, key, value):
        capi.check(capi.lib().gb_cnn_set_option(self._h, key.encode(), float(value)))

    def get_option(self, key):
        return capi.lib().gb_cnn_get_option(self._h, key.encode())

    def model_info(self, i=0):
        if not self._models and self._parent is not None:
            return self._parent.model_info(i)
        info = capi.ModelInfo()
        capi.check(capi.lib().gb_model_get_info(self._models[i], C.byref(info)))
        return info

    def set_receptor(self, xyz, smina_types):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(smina_types, np.int32)
        assert len(xyz) == len(t)
        capi.check(capi.lib().gb_cnn_set_receptor(self._h, _fp(xyz), _ip(t), len(t)))
        self._n_rec = len(t)

    @staticmethod
    def _poses(lig_xyz, lig_types, pose_offsets, centers):
        xyz = np.ascontiguousarray(lig_xyz, np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(lig_types, np.int32)
        off = np.ascontiguousarray(pose_offsets, np.int32)
        assert len(xyz) == len(t) and off[-1] == len(t)
        c = None if centers is None else np.ascontiguousarray(centers, np.float32).reshape(-1, 3)
        return xyz, t, off, c

    def score_batch(self, lig_xyz, lig_types, pose_offsets, centers=None):
        """-> (score, affinity, loss, variance) float32 arrays, one entry per pose (ensemble means)."""
        xyz, t, off, c = self._poses(lig_xyz, lig_types, pose_offsets, centers)
        n = len(off) - 1
        out = [np.empty(n, np.float32) for _ in range(4)]
        capi.check(capi.lib().gb_cnn_score_batch(self._h, _fp(xyz), _ip(t), _ip(off), n, _fp(c), *[_fp(o) for o in out]))
        return tuple(out)

    def score_batch_models(self, lig_xyz, lig_types, pose_offsets, centers=None):
        """-> (pose, affinity, loss), each [n_models * R, n_poses] (TorchModel::forward outputs; R = max(1, cnn_rotation),
        row = model * R + rotation)."""
        xyz, t, off, c = self._poses(lig_xyz, lig_types, pose_offsets, centers)
        n, m = len(off) - 1, len(self.model_names) * max(1, int(self.get_option("cnn_rotation")))
        out = [np.empty((m, n), np.float32) for _ in range(3)]
        capi.check(capi.lib().gb_cnn_score_batch_models(self._h, _fp(xyz), _ip(t), _ip(off), n, _fp(c),
                                                        *[_fp(o) for o in out]))
        return tuple(out)

    def rotation(self, r, pose):
        """row-major 3x3 matrix of rotation index r (0 = identity) for batch pose index `pose` (gb_cnn_get_rotation)"""
        m = np.zeros(9, np.float32)
        capi.check(capi.lib().gb_cnn_get_rotation(self._h, int(r), int(pose), _fp(m)))
        return m.reshape(3, 3)

    def score_grad_batch(self, lig_xyz, lig_types, pose_offsets, centers=None, receptor=False):
        """score(m, compute_gradient=True) in batch form -> (score, affinity, loss, variance, dloss/dlig_xyz [n_atoms,3]);
        receptor=True (single pose only) appends getReceptorGradient: dloss/drec_xyz [n_receptor_atoms, 3]."""
        xyz, t, off, c = self._poses(lig_xyz, lig_types, pose_offsets, centers)
        n = len(off) - 1
        out = [np.empty(n, np.float32) for _ in range(4)]
        grad = np.zeros((len(t), 3), np.float32)
        rgrad = np.zeros((self._n_rec, 3), np.float32) if receptor else None
        capi.check(capi.lib().gb_cnn_score_grad(self._h, _fp(xyz), _ip(t), _ip(off), n, _fp(c), *[_fp(o) for o in out],
                                                _fp(grad), _fp(rgrad)))
        return (*out, grad, rgrad) if receptor else (*out, grad)

    def score(self, lig_xyz, lig_types, center=None):
        """DLScorer::score(model&, false, aff, loss, var) for one pose -> (score, affinity, loss, variance)."""
        n = len(lig_types)
        s, a, l, v = self.score_batch(lig_xyz, lig_types, [0, n], None if center is None else [center])
        return float(s[0]), float(a[0]), float(l[0]), float(v[0])

    # -- split form (device-resident measurement) ------------------------------------------------------
    def stage(self, lig_xyz, lig_types, pose_offsets, centers=None):
        xyz, t, off, c = self._poses(lig_xyz, lig_types, pose_offsets, centers)
        self._staged_n = len(off) - 1
        capi.check(capi.lib().gb_cnn_stage_poses(self._h, _fp(xyz), _ip(t), _ip(off), self._staged_n, _fp(c)))

    def run_staged(self):
        capi.check(capi.lib().gb_cnn_run_staged(self._h))

    def fetch(self):
        out = [np.empty(self._staged_n, np.float32) for _ in range(4)]
        capi.check(capi.lib().gb_cnn_fetch(self._h, *[_fp(o) for o in out]))
        return tuple(out)

    def fetch_device(self, device_ptr):
        """results [4][n_staged] -> caller-owned DEVICE buffer (raw pointer, >= 16 n bytes): no host bounce (gather via NCCL)"""
        capi.check(capi.lib().gb_cnn_fetch_device(self._h, C.cast(C.c_void_p(int(device_ptr)), C.POINTER(C.c_float))))

    def stream_ptr(self):
        return capi.lib().gb_cnn_stream(self._h)

    def kernel_launches(self):
        return capi.lib().gb_cnn_kernel_launches(self._h)

    def profile(self):
        """-> {kernel class: (total_ms, launches)} accumulated since the last reset (option "profile" must be 1)."""
        out, i = {}, 0
        name = C.create_string_buffer(96)
        ms, cnt = C.c_double(), C.c_int64()
        while capi.lib().gb_cnn_profile_read(self._h, i, name, 96, C.byref(ms), C.byref(cnt)) == 0:
            out[name.value.decode()] = (ms.value, cnt.value)
            i += 1
        return out

    def profile_reset(self):
        capi.check(capi.lib().gb_cnn_profile_reset(self._h))

    def debug_read(self, name):
        """test-only: raw fp16 view of an internal fast-path buffer of the last pass"""
        n = C.c_size_t()
        capi.check(capi.lib().gb_cnn_debug_read(self._h, name.encode(), None, 0, C.byref(n)))
        buf = np.empty(n.value // 2, np.float16)
        capi.check(capi.lib().gb_cnn_debug_read(self._h, name.encode(), buf.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return buf

    def voxelize(self, lig_xyz, lig_types, pose_offsets, centers=None, model_index=0):
        xyz, t, off, c = self._poses(lig_xyz, lig_types, pose_offsets, centers)
        info = self.model_info(model_index)
        n, N = len(off) - 1, info.grid_points
        out = np.empty((n, info.n_rec_channels + info.n_lig_channels, N, N, N), np.float32)
        capi.check(capi.lib().gb_cnn_voxelize(self._h, model_index, _fp(xyz), _ip(t), _ip(off), n, _fp(c), _fp(out)))
        return out

    def type_atoms(self, smina_types, is_ligand, model_index=0):
        t = np.ascontiguousarray(smina_types, np.int32)
        ch = np.empty(len(t), np.int32)
        rad = np.empty(len(t), np.float32)
        capi.check(capi.lib().gb_model_type_atoms(self._models[model_index], int(is_ligand), _ip(t), len(t), _ip(ch),
                                                  _fp(rad)))
        return ch, rad

    def close(self):
        L = capi.lib()
        if self._h:
            L.gb_cnn_destroy(self._h)
            self._h = C.c_void_p()
        for m in self._models:
            L.gb_model_release(m)
        self._models = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
