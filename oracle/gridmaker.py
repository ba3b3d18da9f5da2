"""ctypes front-end to oracle/gridmaker_ref.c (CPU oracle; test infrastructure only).

Follows gninasrc/lib/torch_model.cpp:120-181 (make_coordset, centre choice, rec+lig merge, GridMaker::forward)
and :203 (GridMaker::backward).  See the header of gridmaker_ref.c for the libmolgrid provenance and pins.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.gbo_parse_typemap.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        L.gbo_parse_typemap.restype = C.c_int
        L.gbo_type_atoms.argtypes = [ip, C.c_int, C.POINTER(C.c_int), C.c_int, ip, fp]
        L.gbo_center.argtypes = [fp, C.c_int, fp]
        L.gbo_grid_npts.argtypes = [C.c_float, C.c_float]
        L.gbo_grid_npts.restype = C.c_int
        L.gbo_grid_forward.argtypes = [fp, C.c_float, C.c_float, C.c_float, C.c_int, fp, ip, fp, C.c_int, fp]
        L.gbo_grid_backward.argtypes = [fp, C.c_float, C.c_float, C.c_float, C.c_int, fp, ip, fp, C.c_int, fp, fp]
        L.gbo_grid_forward_batch.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, fp, ip, fp, fp, ip, fp, ip,
                                             C.c_int, fp, C.c_int, fp, C.c_int]
        L.gbo_smina_radius.argtypes = [C.c_int]
        L.gbo_smina_radius.restype = C.c_float
        L.gbo_smina_name.argtypes = [C.c_int]
        L.gbo_smina_name.restype = C.c_char_p
        _LIB = L
    return _LIB


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def parse_typemap(text):
    """-> (n_channels, int[28] smina type -> channel or -1)."""
    t2c = (C.c_int * 28)()
    n = lib().gbo_parse_typemap(text.encode(), t2c)
    if n < 0:
        raise ValueError("unknown smina type name in map")
    return n, np.array(list(t2c), dtype=np.int32)


def type_atoms(smina_types, t2c, channel_offset=0):
    st = np.ascontiguousarray(smina_types, dtype=np.int32)
    ch = np.empty(len(st), np.int32)
    rad = np.empty(len(st), np.float32)
    t2c_c = (C.c_int * 28)(*[int(x) for x in t2c])
    lib().gbo_type_atoms(_i(st), len(st), t2c_c, channel_offset, _i(ch), _f(rad))
    return ch, rad


def center_of(xyz):
    xyz = np.ascontiguousarray(xyz, np.float32)
    c = np.empty(3, np.float32)
    lib().gbo_center(_f(xyz), len(xyz), _f(c))
    return c


def grid_forward(center, xyz, channel, radius, n_channels, resolution=0.5, dimension=23.5, radius_scale=1.0):
    xyz = np.ascontiguousarray(xyz, np.float32)
    channel = np.ascontiguousarray(channel, np.int32)
    radius = np.ascontiguousarray(radius, np.float32)
    center = np.ascontiguousarray(center, np.float32)
    N = lib().gbo_grid_npts(resolution, dimension)
    out = np.empty((n_channels, N, N, N), np.float32)
    lib().gbo_grid_forward(_f(center), resolution, dimension, radius_scale, len(xyz), _f(xyz), _i(channel),
                           _f(radius), n_channels, _f(out))
    return out


def grid_backward(center, xyz, channel, radius, gridgrad, resolution=0.5, dimension=23.5, radius_scale=1.0):
    xyz = np.ascontiguousarray(xyz, np.float32)
    channel = np.ascontiguousarray(channel, np.int32)
    radius = np.ascontiguousarray(radius, np.float32)
    center = np.ascontiguousarray(center, np.float32)
    gridgrad = np.ascontiguousarray(gridgrad, np.float32)
    out = np.zeros((len(xyz), 3), np.float32)
    lib().gbo_grid_backward(_f(center), resolution, dimension, radius_scale, len(xyz), _f(xyz), _i(channel),
                            _f(radius), gridgrad.shape[0], _f(gridgrad), _f(out))
    return out


def grid_forward_batch(rec_xyz, rec_channel, rec_radius, lig_xyz, lig_channel, lig_radius, pose_offsets,
                       n_channels, centers=None, resolution=0.5, dimension=23.5, radius_scale=1.0, n_threads=1):
    """Faithful per-pose voxelisation (receptor re-gridded for every pose), poses striped over threads."""
    rec_xyz = np.ascontiguousarray(rec_xyz, np.float32)
    lig_xyz = np.ascontiguousarray(lig_xyz, np.float32)
    rec_channel = np.ascontiguousarray(rec_channel, np.int32)
    lig_channel = np.ascontiguousarray(lig_channel, np.int32)
    rec_radius = np.ascontiguousarray(rec_radius, np.float32)
    lig_radius = np.ascontiguousarray(lig_radius, np.float32)
    pose_offsets = np.ascontiguousarray(pose_offsets, np.int32)
    n_poses = len(pose_offsets) - 1
    N = lib().gbo_grid_npts(resolution, dimension)
    out = np.empty((n_poses, n_channels, N, N, N), np.float32)
    cptr = None
    if centers is not None:
        centers = np.ascontiguousarray(centers, np.float32)
        cptr = _f(centers)
    lib().gbo_grid_forward_batch(resolution, dimension, radius_scale, len(rec_xyz), _f(rec_xyz), _i(rec_channel),
                                 _f(rec_radius), _f(lig_xyz), _i(lig_channel), _f(lig_radius), _i(pose_offsets),
                                 n_poses, cptr, n_channels, _f(out), n_threads)
    return out
