"""ctypes binding of oracle/libsurf_oracle.so (the C restatement of OpenCV 2.4.5 SURF).

TEST INFRASTRUCTURE ONLY -- the checker, never the thing measured or shipped.
Restates the calls made at SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:1381-1390 and :635-643.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsurf_oracle.so")

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)
DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
assert KEYPOINT_DTYPE.itemsize == 28 and DMATCH_DTYPE.itemsize == 16


class Params(C.Structure):
    _fields_ = [("hessian_threshold", C.c_double), ("n_octaves", C.c_int32), ("n_octave_layers", C.c_int32),
                ("extended", C.c_int32), ("upright", C.c_int32)]


def params(hessianThreshold=100.0, nOctaves=4, nOctaveLayers=3, extended=False, upright=False) -> Params:
    return Params(float(hessianThreshold), int(nOctaves), int(nOctaveLayers), int(bool(extended)), int(bool(upright)))


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, no FMA contraction)."""
    srcs = [os.path.join(_HERE, f) for f in ("surf_oracle.c", "surf_oracle.h", "pose_oracle.c", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.check_call(["make", "-s", "-C", _HERE, "libsurf_oracle.so"], env=env)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, pp = C.c_void_p, C.c_int, C.POINTER(Params)
        L.orc_integral.argtypes = [vp, i32, i32, i32, vp, i32]
        L.orc_integral.restype = None
        L.orc_hessian_layer.argtypes = [vp, i32, i32, i32, i32, vp, vp]
        L.orc_hessian_layer.restype = i32
        L.orc_gaussian_kernel.argtypes = [i32, C.c_double, vp]
        L.orc_gaussian_kernel.restype = None
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_resize_area_21.argtypes = [vp, i32, vp]
        L.orc_resize_area_21.restype = i32
        L.orc_raw_keypoints.argtypes = [vp, i32, i32, i32, vp, i32, pp, vp, i32]
        L.orc_raw_keypoints.restype = i32
        L.orc_describe.argtypes = [vp, i32, i32, i32, pp, vp, i32, vp, vp]
        L.orc_describe.restype = None
        L.orc_detect.argtypes = [vp, i32, i32, i32, vp, i32, pp, vp, i32]
        L.orc_detect.restype = i32
        L.orc_compute.argtypes = [vp, i32, i32, i32, pp, vp, i32, vp]
        L.orc_compute.restype = i32
        L.orc_detect_and_compute.argtypes = [vp, i32, i32, i32, vp, i32, pp, vp, i32, vp]
        L.orc_detect_and_compute.restype = i32
        L.orc_detect_and_compute_batch.argtypes = [vp, i32, i32, i32, pp, vp, i32, vp, vp, i32]
        L.orc_detect_and_compute_batch.restype = C.c_long
        L.orc_knn_match.argtypes = [vp, i32, vp, i32, i32, i32, vp, i32]
        L.orc_knn_match.restype = None
        L.orc_hough_transform.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32]
        L.orc_correspond.argtypes = [vp, i32, vp, vp, vp, vp, i32, C.c_float, i32, i32, vp, vp, i32]
        L.orc_correspond.restype = None
        L.orc_mask_centroid.argtypes = [vp, i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
        L.orc_mask_centroid.restype = None
        L.orc_max_threads.argtypes = []
        L.orc_max_threads.restype = i32
        L.orc_set_switches.argtypes = [i32, i32, vp]
        L.orc_set_switches.restype = None
        _lib = L
    return _lib


class switches:
    """Context manager for the named reading choices of SURVEY Appendix C (process-wide in the C oracle; all off = the
    2.4.5 reading the CUDA path implements).  angle_pre244: kp.angle = fastAtan2(sum_y, sum_x), upright 90 (C-2, the
    convention before 2.4.4); libm_sincos: this platform's sinf / cosf for the dominant angle instead of the correctly
    rounded values; desc_gauss: the 20 taps of the descriptor's Gaussian, e.g. cv2.getGaussianKernel(20, 3.3, cv2.CV_32F)
    of the installed OpenCV (C-4)."""

    def __init__(self, angle_pre244=False, libm_sincos=False, desc_gauss=None):
        self.a, self.l = int(bool(angle_pre244)), int(bool(libm_sincos))
        self.g = None if desc_gauss is None else np.ascontiguousarray(np.asarray(desc_gauss, np.float32).reshape(-1))
        assert self.g is None or self.g.size == 20

    def __enter__(self):
        lib().orc_set_switches(self.a, self.l, _ptr(self.g))
        return self

    def __exit__(self, *exc):
        lib().orc_set_switches(0, 0, None)
        return False


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data


def integral(img, clamp01=False):
    img = _img(img)
    H, W = img.shape
    out = np.empty((H + 1, W + 1), np.int32)
    lib().orc_integral(_ptr(img), H, W, W, _ptr(out), int(clamp01))
    return out


def hessian_layer(sum_, size, step):
    """det, trace of one pyramid layer, (H//step, W//step) fp32, zero outside the written region."""
    sum_ = np.ascontiguousarray(sum_, np.int32)
    H, W = sum_.shape[0] - 1, sum_.shape[1] - 1
    det = np.zeros((H // step, W // step), np.float32)
    tr = np.zeros_like(det)
    ok = lib().orc_hessian_layer(_ptr(sum_), H, W, size, step, _ptr(det), _ptr(tr))
    return det, tr, bool(ok)


def gaussian_kernel(n, sigma):
    out = np.empty(n, np.float32)
    lib().orc_gaussian_kernel(n, float(sigma), _ptr(out))
    return out


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(float(np.float32(y)), float(np.float32(x))))


def resize_area_21(win):
    win = _img(win)
    assert win.shape[0] == win.shape[1]
    out = np.empty((21, 21), np.uint8)
    rc = lib().orc_resize_area_21(_ptr(win), win.shape[0], _ptr(out))
    if rc:
        raise ValueError("empty window")
    return out


def raw_keypoints(img, p: Params, mask=None, cap=1 << 18):
    img = _img(img)
    H, W = img.shape
    m = None if mask is None else _img(mask)
    kps = np.zeros(cap, KEYPOINT_DTYPE)
    n = lib().orc_raw_keypoints(_ptr(img), H, W, W, _ptr(m), W, C.byref(p), _ptr(kps), cap)
    assert n <= cap
    return kps[:n].copy()


def describe(img, p: Params, kps, want_desc=True, want_patches=False):
    """Orientation (+descriptor) for given keypoints WITHOUT removing deleted ones (size=-1 marks them)."""
    img = _img(img)
    H, W = img.shape
    kps = np.array(kps, dtype=KEYPOINT_DTYPE, copy=True)
    n = len(kps)
    D = 128 if p.extended else 64
    desc = np.zeros((n, D), np.float32) if want_desc else None
    patches = np.zeros((n, 21, 21), np.uint8) if want_patches else None
    lib().orc_describe(_ptr(img), H, W, W, C.byref(p), _ptr(kps), n, _ptr(desc), _ptr(patches))
    return kps, desc, patches


def detect(img, p: Params, mask=None, cap=1 << 18):
    img = _img(img)
    H, W = img.shape
    m = None if mask is None else _img(mask)
    kps = np.zeros(cap, KEYPOINT_DTYPE)
    n = lib().orc_detect(_ptr(img), H, W, W, _ptr(m), W, C.byref(p), _ptr(kps), cap)
    assert 0 <= n <= cap
    return kps[:n].copy()


def compute(img, p: Params, kps):
    img = _img(img)
    H, W = img.shape
    kps = np.array(kps, dtype=KEYPOINT_DTYPE, copy=True)
    D = 128 if p.extended else 64
    desc = np.zeros((max(len(kps), 1), D), np.float32)
    n = lib().orc_compute(_ptr(img), H, W, W, C.byref(p), _ptr(kps), len(kps), _ptr(desc))
    return kps[:n].copy(), desc[:n].copy()


def detect_and_compute(img, p: Params, mask=None, cap=1 << 18):
    img = _img(img)
    H, W = img.shape
    m = None if mask is None else _img(mask)
    D = 128 if p.extended else 64
    kps = np.zeros(cap, KEYPOINT_DTYPE)
    desc = np.zeros((cap, D), np.float32)
    n = lib().orc_detect_and_compute(_ptr(img), H, W, W, _ptr(m), W, C.byref(p), _ptr(kps), cap, _ptr(desc))
    assert 0 <= n <= cap
    return kps[:n].copy(), desc[:n].copy()


def detect_and_compute_batch(imgs, p: Params, cap=1 << 15, n_threads=0):
    imgs = np.ascontiguousarray(imgs, np.uint8)
    B, H, W = imgs.shape
    D = 128 if p.extended else 64
    kps = np.zeros((B, cap), KEYPOINT_DTYPE)
    desc = np.zeros((B, cap, D), np.float32)
    counts = np.zeros(B, np.int32)
    lib().orc_detect_and_compute_batch(_ptr(imgs), B, H, W, C.byref(p), _ptr(kps), cap, _ptr(desc), _ptr(counts),
                                       int(n_threads))
    assert (counts >= 0).all(), "oracle capacity exceeded"
    return [(kps[b, :counts[b]].copy(), desc[b, :counts[b]].copy()) for b in range(B)]


def knn_match(q, t, k=2, n_threads=0):
    q = np.ascontiguousarray(q, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    assert q.ndim == 2 and t.ndim == 2 and q.shape[1] == t.shape[1]
    out = np.zeros((q.shape[0], k), DMATCH_DTYPE)
    lib().orc_knn_match(_ptr(q), q.shape[0], _ptr(t), t.shape[0], q.shape[1], k, _ptr(out), int(n_threads))
    return out


def max_threads():
    return int(lib().orc_max_threads())


def db_knn(query, train_list, k=2, n_threads=0):
    """DescriptorMatcher::add(train_list) + knnMatch(query, k): exact L2 over the concatenation, then
    (imgIdx, trainIdx within the image); ties -> lower image, lower row (BFMatcher's scan order)."""
    sizes = np.array([len(t) for t in train_list], np.int64)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    d = np.asarray(query).shape[1]
    allt = np.concatenate([np.asarray(t, np.float32).reshape(-1, d) for t in train_list], axis=0) if len(train_list) else \
        np.zeros((0, d), np.float32)
    m = knn_match(query, allt, k, n_threads)
    valid = m["trainIdx"] >= 0
    img = np.searchsorted(starts, m["trainIdx"], side="right") - 1
    m["imgIdx"] = np.where(valid, img, -1)
    m["trainIdx"] = np.where(valid, m["trainIdx"] - starts[np.clip(img, 0, max(len(sizes) - 1, 0))], -1)
    return m


def hough_transform(query_kps, train_kps_list, matches, refx, refy, libm_float=False):
    """houghTransform(): returns (vote count of the winning booth, matches with rejected entries set to -1)."""
    qk = np.ascontiguousarray(query_kps, KEYPOINT_DTYPE)
    off = np.concatenate([[0], np.cumsum([len(t) for t in train_kps_list])]).astype(np.int32)
    tk = np.ascontiguousarray(np.concatenate([np.asarray(t, KEYPOINT_DTYPE) for t in train_kps_list])) \
        if len(train_kps_list) else np.zeros(0, KEYPOINT_DTYPE)
    m = np.ascontiguousarray(matches, DMATCH_DTYPE).copy()
    c = lib().orc_hough_transform(_ptr(qk), _ptr(tk), _ptr(off), _ptr(m), len(m), int(refx), int(refy), int(libm_float))
    return int(c), m


def correspond(query_kps, query_desc, train_kps_list, train_desc_list, bogus_desc_list, refx, refy, bogus_ratio=0.95,
               libm_float=False, n_threads=0):
    """computeNextInterSliceCorrespondence for one query image.  Returns (after_hough matches, result[4])."""
    qk = np.ascontiguousarray(query_kps, KEYPOINT_DTYPE)
    mt = np.ascontiguousarray(db_knn(query_desc, train_desc_list, 3, n_threads)[:, 0])
    mb = np.ascontiguousarray(db_knn(query_desc, bogus_desc_list, 1, n_threads)[:, 0])
    off = np.concatenate([[0], np.cumsum([len(t) for t in train_kps_list])]).astype(np.int32)
    tk = np.ascontiguousarray(np.concatenate([np.asarray(t, KEYPOINT_DTYPE) for t in train_kps_list]))
    out = np.zeros(max(len(qk), 1), DMATCH_DTYPE)
    res = np.zeros(4, np.int32)
    lib().orc_correspond(_ptr(qk), len(qk), _ptr(mt), _ptr(mb), _ptr(tk), _ptr(off), len(train_kps_list),
                         float(bogus_ratio), int(refx), int(refy), _ptr(out), _ptr(res), int(libm_float))
    return out[:res[0]].copy(), res


def mask_centroid(mask):
    m = np.ascontiguousarray(mask, np.uint8)
    x, y = C.c_int(0), C.c_int(0)
    lib().orc_mask_centroid(_ptr(m), m.shape[0], m.shape[1], m.strides[0], C.byref(x), C.byref(y))
    return x.value, y.value
