"""Host-side mirror of the reference's plane fit / pose estimate calls:
vtkSlicerSurfFeaturesLogic::ransac (...Logic.cxx:2085-2188) and the arithmetic of updateMatchNodeRansac
(:1655-1851), over the C ABI (surf_rand_*, surf_match_points, surf_ransac_plane, surf_estimate_pose,
surf_mean_plane_distance).  The RANSAC hypotheses are scored on the GPU; there is no CPU path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .engine import SURF, SurfError, load_library


class Rand:
    """rand() of the reference's platform (glibc TYPE_3), as an explicit state.  The reference never calls
    srand(), so its stream is Rand(1)."""

    def __init__(self, seed: int = 1):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.surf_rand_create(seed, C.byref(h))
        if rc:
            raise SurfError(rc, "surf_rand_create failed")
        self._h = h

    def next(self) -> int:
        return self._lib.surf_rand_next(self._h)

    def close(self):
        if self._h:
            self._lib.surf_rand_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def match_points(matches, query_kps, train_kps, train_offsets, train_transforms, image_to_probe):
    """trainPoints / queryPoints (:1672-1703) of the matches that survived the Hough vote (imgIdx != -1)."""
    lib = load_library()
    m = np.ascontiguousarray(matches)
    qk, tk = np.ascontiguousarray(query_kps), np.ascontiguousarray(train_kps)
    off = np.ascontiguousarray(train_offsets, np.int32)
    tt = np.ascontiguousarray(train_transforms, np.float32).reshape(-1, 16)
    ip = np.ascontiguousarray(image_to_probe, np.float64)
    tp = np.zeros((len(m), 3), np.float64)
    qp = np.zeros((len(m), 3), np.float64)
    n = C.c_int32(0)
    rc = lib.surf_match_points(m.ctypes.data, len(m), qk.ctypes.data, tk.ctypes.data, off.ctypes.data, len(tt), tt.ctypes.data,
                               ip.ctypes.data, tp.ctypes.data, qp.ctypes.data, C.byref(n))
    if rc:
        raise SurfError(rc, "surf_match_points: bad argument")
    return tp[:n.value].copy(), qp[:n.value].copy()


def ransac_plane(engine: SURF, rng: Rand, points, margin: float = 1.0, iterations: int = 1000):
    """ransac(points, inliersIdx, planePointsIdx) -> (inliers, plane_idx, best_error, n_outliers)."""
    P = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
    n = len(P)
    inl = np.zeros(n + 3, np.int32)
    n_inl = C.c_int32(0)
    plane = np.full(3, -1, np.int32)
    err, nout = C.c_double(0), C.c_int32(0)
    engine._check(engine._lib.surf_ransac_plane(engine._h, rng._h, P.ctypes.data, n, margin, iterations, inl.ctypes.data,
                                                C.byref(n_inl), plane.ctypes.data, C.byref(err), C.byref(nout)))
    return inl[:n_inl.value].copy(), plane, err.value, nout.value


def estimate_pose(rng: Rand, query_points, train_points, inliers, plane_idx):
    """-> (estimate 4x4, mean match distance) of updateMatchNodeRansac."""
    lib = load_library()
    Q = np.ascontiguousarray(query_points, np.float64).reshape(-1, 3)
    T = np.ascontiguousarray(train_points, np.float64).reshape(-1, 3)
    inl = np.ascontiguousarray(inliers, np.int32)
    pl = np.ascontiguousarray(plane_idx, np.int32)
    est = np.zeros(16, np.float64)
    mmd = C.c_double(0)
    rc = lib.surf_estimate_pose(rng._h, Q.ctypes.data, T.ctypes.data, len(Q), inl.ctypes.data, len(inl), pl.ctypes.data,
                                est.ctypes.data, C.byref(mmd))
    if rc:
        raise SurfError(rc, "surf_estimate_pose: bad argument")
    return est.reshape(4, 4), mmd.value


def mean_plane_distance(mask, ground_truth, estimate, step: int = 4) -> float:
    lib = load_library()
    m = np.ascontiguousarray(mask, np.uint8)
    g = np.ascontiguousarray(ground_truth, np.float64)
    e = np.ascontiguousarray(estimate, np.float64)
    out = C.c_double(0)
    rc = lib.surf_mean_plane_distance(m.ctypes.data, m.shape[0], m.shape[1], m.shape[1], step, g.ctypes.data, e.ctypes.data,
                                      C.byref(out))
    if rc:
        raise SurfError(rc, "surf_mean_plane_distance: bad argument")
    return out.value
