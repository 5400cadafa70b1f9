"""ctypes binding of the f4 part of oracle/libsurf_oracle.so (pose_oracle.c: RANSAC plane fit and pose estimate,
reference SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:2085-2188 and :1655-1851).

TEST INFRASTRUCTURE ONLY -- the checker, never the thing measured or shipped.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import surf_oracle as _so


class Rand(C.Structure):
    """glibc rand() restated (TYPE_3); `Rand(seed)` is srand(seed), the reference never seeds (= 1)."""
    _fields_ = [("r", C.c_int32 * 34), ("f", C.c_int), ("b", C.c_int)]

    def __init__(self, seed: int = 1):
        super().__init__()
        _lib().orc_srand(C.byref(self), seed)

    def next(self) -> int:
        return _lib().orc_rand_next(C.byref(self))


_ready = False


def _lib():
    global _ready
    L = _so.lib()
    if not _ready:
        vp, i32, dp = C.c_void_p, C.c_int, C.POINTER(C.c_double)
        L.orc_srand.argtypes = [vp, C.c_uint32]
        L.orc_srand.restype = None
        L.orc_rand_next.argtypes = [vp]
        L.orc_rand_next.restype = i32
        L.orc_ransac_plane.argtypes = [vp, vp, i32, C.c_float, i32, vp, vp, vp, dp, dp]
        L.orc_ransac_plane.restype = i32
        L.orc_match_points.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
        L.orc_match_points.restype = i32
        L.orc_estimate_pose.argtypes = [vp, vp, vp, i32, vp, i32, vp, vp, dp]
        L.orc_estimate_pose.restype = i32
        L.orc_mean_plane_distance.argtypes = [vp, i32, i32, C.c_size_t, i32, vp, vp]
        L.orc_mean_plane_distance.restype = C.c_double
        _ready = True
    return L


def ransac_plane(rng: Rand, points, margin=1.0, iterations=1000):
    """-> (rc, inliers, plane_idx, best_error, n_outliers) as ransac() leaves them."""
    P = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
    n = len(P)
    inl = np.zeros(n + 3, np.int32)
    n_inl = C.c_int32(0)
    plane = np.full(3, -1, np.int32)
    err, nout = C.c_double(0), C.c_double(0)
    rc = _lib().orc_ransac_plane(C.byref(rng), P.ctypes.data, n, margin, iterations, inl.ctypes.data, C.byref(n_inl),
                                 plane.ctypes.data, C.byref(err), C.byref(nout))
    return rc, inl[:n_inl.value].copy(), plane, err.value, int(nout.value)


def match_points(matches, query_kps, train_kps, train_offsets, train_transforms, image_to_probe):
    m = np.ascontiguousarray(matches)
    qk, tk = np.ascontiguousarray(query_kps), np.ascontiguousarray(train_kps)
    off = np.ascontiguousarray(train_offsets, np.int32)
    tt = np.ascontiguousarray(train_transforms, np.float32)
    ip = np.ascontiguousarray(image_to_probe, np.float64)
    tp = np.zeros((len(m), 3), np.float64)
    qp = np.zeros((len(m), 3), np.float64)
    n = _lib().orc_match_points(m.ctypes.data, len(m), qk.ctypes.data, tk.ctypes.data, off.ctypes.data, tt.ctypes.data,
                                ip.ctypes.data, tp.ctypes.data, qp.ctypes.data)
    return tp[:n].copy(), qp[:n].copy()


def estimate_pose(rng: Rand, query_points, train_points, inliers, plane_idx):
    Q = np.ascontiguousarray(query_points, np.float64).reshape(-1, 3)
    T = np.ascontiguousarray(train_points, np.float64).reshape(-1, 3)
    inl = np.ascontiguousarray(inliers, np.int32)
    pl = np.ascontiguousarray(plane_idx, np.int32)
    est = np.zeros(16, np.float64)
    mmd = C.c_double(0)
    _lib().orc_estimate_pose(C.byref(rng), Q.ctypes.data, T.ctypes.data, len(Q), inl.ctypes.data, len(inl), pl.ctypes.data,
                             est.ctypes.data, C.byref(mmd))
    return est.reshape(4, 4), mmd.value


def mean_plane_distance(mask, ground_truth, estimate, step=4):
    m = np.ascontiguousarray(mask, np.uint8)
    g = np.ascontiguousarray(ground_truth, np.float64)
    e = np.ascontiguousarray(estimate, np.float64)
    return _lib().orc_mean_plane_distance(m.ctypes.data, m.shape[0], m.shape[1], m.shape[1], step, g.ctypes.data, e.ctypes.data)
