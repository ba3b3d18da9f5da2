"""ctypes binding of libgnina_b200.so — the same C ABI a C++/cgo/JNI host binds (include/gnina_b200.h).
Fails loudly when the CUDA library is missing or no device is visible: there is NO CPU fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgnina_b200.so")
_lib = None

SYMBOLS = ["gb_last_error", "gb_version", "gb_initialize_cuda", "gb_device_count", "gb_model_load", "gb_model_load_mem",
           "gb_model_get_info", "gb_model_release", "gb_model_type_atoms", "gb_cnn_create", "gb_cnn_clone",
           "gb_cnn_destroy", "gb_cnn_num_models", "gb_cnn_set_option", "gb_cnn_get_option", "gb_cnn_set_receptor",
           "gb_cnn_score_batch", "gb_cnn_score_batch_models", "gb_cnn_get_rotation", "gb_cnn_score_grad", "gb_cnn_stage_poses", "gb_cnn_run_staged",
           "gb_cnn_fetch", "gb_cnn_fetch_device", "gb_cnn_profile_read", "gb_cnn_profile_reset", "gb_cnn_debug_read", "gb_cnn_stream", "gb_cnn_kernel_launches", "gb_cnn_voxelize", "gb_vina_create", "gb_vina_destroy", "gb_vina_table_size", "gb_vina_prec_table",
           "gb_vina_set_receptor", "gb_vina_cache_build", "gb_vina_cache_read", "gb_vina_cache_eval", "gb_vina_score_exact", "gb_vina_score_noncache", "gb_vina_minimize", "gb_vina_refine_minimize", "gb_vina_set_ligand", "gb_vina_eval_deriv",
           "gb_vina_bfgs", "gb_vina_mc", "gb_vina_mc_traced", "gb_vina_eval_deriv_noncache", "gb_vina_refine", "gb_vina_noncache_atoms", "gb_vina_merge_outputs", "gb_vina_spline_size", "gb_vina_spline_table", "gb_vina_set_precalc"]


class LigandTopology(C.Structure):
    _fields_ = [("n_atoms", C.c_int32), ("n_segments", C.c_int32), ("n_pairs", C.c_int32),
                ("local_xyz", C.POINTER(C.c_float)), ("smina_type", C.POINTER(C.c_int32)),
                ("seg_parent", C.POINTER(C.c_int32)), ("seg_atom_begin", C.POINTER(C.c_int32)),
                ("seg_atom_end", C.POINTER(C.c_int32)), ("seg_rel_origin", C.POINTER(C.c_float)),
                ("seg_rel_axis", C.POINTER(C.c_float)), ("pair_a", C.POINTER(C.c_int32)), ("pair_b", C.POINTER(C.c_int32)),
                ("gyration_radius", C.c_float)]


class McParams(C.Structure):
    _fields_ = [("num_steps", C.c_int32), ("maxiters", C.c_int32), ("num_saved_mins", C.c_int32), ("temperature", C.c_float),
                ("mutation_amplitude", C.c_float), ("min_rmsd", C.c_float), ("hunt_cap", C.c_float * 3)]


class MinimizationParams(C.Structure):
    _fields_ = [("maxiters", C.c_int32), ("accurate_line_search", C.c_int32), ("early_term", C.c_int32)]


class GbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gnina_b200 error %d: %s" % (code, msg))
        self.code = code


class ModelInfo(C.Structure):
    _fields_ = [("arch", C.c_int32), ("n_rec_channels", C.c_int32), ("n_lig_channels", C.c_int32),
                ("grid_points", C.c_int32), ("resolution", C.c_float), ("dimension", C.c_float),
                ("radius_scaling", C.c_float), ("apply_logistic_loss", C.c_int32), ("skip_softmax", C.c_int32),
                ("name", C.c_char * 64)]


# 32 hardware work queues instead of 8 (see prefer_many_hw_queues, gb_internal.h): must be in the environment before
# the process initialises CUDA, whoever does it first (this library or torch)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("gnina_b200: %s is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p
    L.gb_last_error.restype = C.c_char_p
    L.gb_version.restype = C.c_char_p
    L.gb_initialize_cuda.argtypes = [C.c_int]
    L.gb_model_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.gb_model_load_mem.argtypes = [vp, C.c_size_t, C.c_int, C.POINTER(vp)]
    L.gb_model_get_info.argtypes = [vp, C.POINTER(ModelInfo)]
    L.gb_model_release.argtypes = [vp]
    L.gb_model_release.restype = None
    L.gb_model_type_atoms.argtypes = [vp, C.c_int, ip, C.c_int, ip, fp]
    L.gb_cnn_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(vp)]
    L.gb_cnn_clone.argtypes = [vp, C.POINTER(vp)]
    L.gb_cnn_destroy.argtypes = [vp]
    L.gb_cnn_destroy.restype = None
    L.gb_cnn_num_models.argtypes = [vp]
    L.gb_cnn_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.gb_cnn_get_option.argtypes = [vp, C.c_char_p]
    L.gb_cnn_get_option.restype = C.c_double
    L.gb_cnn_set_receptor.argtypes = [vp, fp, ip, C.c_int]
    L.gb_cnn_score_batch.argtypes = [vp, fp, ip, ip, C.c_int, fp, fp, fp, fp, fp]
    L.gb_cnn_score_batch_models.argtypes = [vp, fp, ip, ip, C.c_int, fp, fp, fp, fp]
    L.gb_cnn_score_grad.argtypes = [vp, fp, ip, ip, C.c_int, fp, fp, fp, fp, fp, fp, fp]
    L.gb_cnn_get_rotation.argtypes = [vp, C.c_int, C.c_int, fp]
    L.gb_cnn_stage_poses.argtypes = [vp, fp, ip, ip, C.c_int, fp]
    L.gb_cnn_run_staged.argtypes = [vp]
    L.gb_cnn_fetch.argtypes = [vp, fp, fp, fp, fp]
    L.gb_cnn_fetch_device.argtypes = [vp, fp]
    L.gb_cnn_profile_read.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.gb_cnn_profile_reset.argtypes = [vp]
    L.gb_cnn_debug_read.argtypes = [vp, C.c_char_p, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.gb_cnn_stream.argtypes = [vp]
    L.gb_cnn_stream.restype = vp
    L.gb_cnn_kernel_launches.argtypes = [vp]
    L.gb_cnn_kernel_launches.restype = C.c_int64
    L.gb_cnn_voxelize.argtypes = [vp, C.c_int, fp, ip, ip, C.c_int, fp, fp]
    L.gb_vina_create.argtypes = [C.c_int, fp, C.c_float, C.POINTER(vp)]
    L.gb_vina_destroy.argtypes = [vp]
    L.gb_vina_destroy.restype = None
    L.gb_vina_table_size.argtypes = [vp]
    L.gb_vina_prec_table.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp]
    L.gb_vina_set_receptor.argtypes = [vp, fp, ip, C.c_int]
    L.gb_vina_cache_build.argtypes = [vp, fp, fp, ip, ip, C.c_int]
    L.gb_vina_cache_read.argtypes = [vp, C.c_int, fp]
    L.gb_vina_cache_eval.argtypes = [vp, fp, ip, ip, C.c_int, C.c_float, C.c_float, fp, fp]
    L.gb_vina_score_exact.argtypes = [vp, fp, ip, ip, C.c_int, fp, C.c_float, fp, fp]
    L.gb_vina_minimize.argtypes = [vp, fp, C.c_int, C.POINTER(MinimizationParams), fp, C.c_float, fp, fp, ip]
    L.gb_vina_refine_minimize.argtypes = [vp, fp, C.c_int, C.POINTER(MinimizationParams), fp, fp, fp, fp, ip, ip]
    L.gb_vina_score_noncache.argtypes = [vp, fp, ip, ip, C.c_int, fp, C.c_float, C.c_float, fp, fp, fp, fp]
    up = C.POINTER(C.c_uint32)
    L.gb_vina_spline_size.argtypes = [vp]
    L.gb_vina_spline_table.argtypes = [vp, C.c_int, C.c_int, fp]
    L.gb_vina_set_precalc.argtypes = [vp, C.c_int]
    L.gb_vina_set_ligand.argtypes = [vp, C.POINTER(LigandTopology)]
    L.gb_vina_eval_deriv.argtypes = [vp, fp, C.c_int, fp, C.c_float, fp, fp, fp]
    L.gb_vina_bfgs.argtypes = [vp, fp, C.c_int, C.c_int, fp, C.c_float, fp, fp, ip]
    L.gb_vina_mc.argtypes = [vp, C.POINTER(McParams), fp, fp, up, C.c_int, C.c_float, fp, fp, ip]
    L.gb_vina_mc_traced.argtypes = [vp, C.POINTER(McParams), fp, fp, up, C.c_int, C.c_float, fp, fp, ip, fp]
    L.gb_vina_eval_deriv_noncache.argtypes = [vp, fp, C.c_int, fp, C.c_float, fp, fp, fp, fp]
    L.gb_vina_refine.argtypes = [vp, fp, C.c_int, C.c_int, fp, fp, fp, fp, ip, ip]
    L.gb_vina_noncache_atoms.argtypes = [vp, fp, ip, C.c_int, fp, fp, C.c_float, fp, fp]
    L.gb_vina_merge_outputs.argtypes = [fp, fp, ip, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, ip, ip]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise GbError(rc, lib().gb_last_error().decode(errors="replace"))
