"""Host-side mirror of the OpenCV interface the reference calls, over the C ABI of libsurf_b200.so.

`SURF` mirrors cv::SURF / cv2.Feature2D (detect, compute, detectAndCompute, descriptorSize, the five
public fields) as used at SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:1384-1389; `BFMatcher`
mirrors DescriptorMatcher::knnMatch as used at :640 (and clear/add at :1402-1403).

Everything here is plumbing: every numerical step runs in the CUDA library.  There is no CPU path --
if the library or an sm_100 device is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsurf_b200.so")

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)  # cv::KeyPoint, 28 B
DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])  # cv::DMatch

SURF_OK = 0
ERR_NAMES = {-1: "BAD_ARG", -2: "UNSUPPORTED", -3: "CAPACITY", -4: "CUDA", -5: "NO_DEVICE", -6: "NCCL"}
MEM_HOST, MEM_DEVICE = 0, 1

# every symbol include/surf_b200.h declares
ABI_SYMBOLS = [
    "surf_create", "surf_destroy", "surf_set_params", "surf_get_params", "surf_last_error", "surf_required_capacity",
    "surf_descriptor_size", "surf_get_timings", "surf_stream", "surf_detect", "surf_compute", "surf_detect_and_compute",
    "surf_detect_and_compute_batch", "surf_submit_batch", "surf_prefetch_batch", "surf_lookahead_batch", "surf_collect_batch", "surf_device_results", "surf_sync",
    "surf_integral", "surf_hessian_layer", "surf_raw_keypoints", "surf_describe_keypoints", "surf_match_knn",
    "surf_match_ratio", "surf_merge_topk", "surf_version", "surf_match_prev", "surf_stream_reset",
    "surf_device_matches", "surf_set_match_mode", "surf_match_redo_count", "surf_set_async_outputs",
    "surf_set_graphs", "surf_graph_replays",
    "surf_rand_create", "surf_rand_destroy", "surf_rand_next", "surf_match_points", "surf_ransac_plane",
    "surf_estimate_pose", "surf_mean_plane_distance",
    "surf_wait_outputs", "surf_db_create", "surf_db_destroy", "surf_db_clear", "surf_db_add", "surf_db_size",
    "surf_db_knn", "surf_correspond", "surf_mha_open", "surf_mha_close", "surf_mha_last_error", "surf_mha_transforms",
    "surf_mha_validity", "surf_mha_frame_name", "surf_mha_read", "surf_crop_rect", "surf_mask_centroid", "surf_sweep_mha",
    "surf_mark", "surf_mark_elapsed", "surf_match_prev_time", "surf_comm_unique_id", "surf_comm_init", "surf_comm_destroy",
    "surf_comm_info", "surf_shard_range", "surf_sweep_sharded", "surf_db_knn_sharded", "surf_comm_last_times",
]


class SurfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"surf_b200 error {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("hessian_threshold", C.c_double), ("n_octaves", C.c_int32), ("n_octave_layers", C.c_int32),
                ("extended", C.c_int32), ("upright", C.c_int32)]


class Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d", "integral", "hessian", "nms", "sort", "describe", "pack", "d2h", "total")]
    _fields_ += [("launches", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class VoteParams(C.Structure):
    _fields_ = [("bogus_ratio", C.c_float), ("ref_x", C.c_int32), ("ref_y", C.c_int32)]


class SweepStats(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("exchange_ms", C.c_float), ("assemble_ms", C.c_float),
                ("bytes_sent", C.c_int64), ("bytes_received", C.c_int64), ("chunks", C.c_int32), ("launches", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


VOTE_RESULT_DTYPE = np.dtype([("n_matches", "<i4"), ("hough_count", "<i4"), ("best_image", "<i4"),
                              ("best_count", "<i4")])

_lib = None


def load_library():
    """Loads libsurf_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SURF_B200_LIB", LIB_PATH)  # development knob: A/B builds of the same library
    if not os.path.exists(path):
        raise SurfError(-5, f"{path} is missing: run `python -m surffeatures_b200.build` (there is no CPU path)")
    L = C.CDLL(path)
    vp, i32, sz, lng = C.c_void_p, C.c_int, C.c_size_t, C.c_long
    pp = C.POINTER(Params)
    L.surf_create.argtypes = [pp, i32, i32, i32, i32, i32, C.POINTER(vp)]
    L.surf_destroy.argtypes = [vp]
    L.surf_destroy.restype = None
    L.surf_set_params.argtypes = [vp, pp]
    L.surf_get_params.argtypes = [vp, pp]
    L.surf_last_error.argtypes = [vp]
    L.surf_last_error.restype = C.c_char_p
    L.surf_required_capacity.argtypes = [vp]
    L.surf_required_capacity.restype = lng
    L.surf_descriptor_size.argtypes = [vp]
    L.surf_get_timings.argtypes = [vp, C.POINTER(Timings)]
    L.surf_stream.argtypes = [vp]
    L.surf_stream.restype = vp
    L.surf_detect.argtypes = [vp, vp, i32, i32, sz, vp, sz, vp, i32, C.POINTER(i32)]
    L.surf_compute.argtypes = [vp, vp, i32, i32, sz, vp, i32, vp, C.POINTER(i32)]
    L.surf_detect_and_compute.argtypes = [vp, vp, i32, i32, sz, vp, sz, vp, vp, i32, C.POINTER(i32)]
    L.surf_detect_and_compute_batch.argtypes = [vp, vp, i32, i32, i32, i32, sz, sz, vp, sz, vp, vp, i32, lng, vp]
    L.surf_submit_batch.argtypes = [vp, vp, i32, i32, i32, i32, sz, sz, vp, sz, i32]
    L.surf_collect_batch.argtypes = [vp, vp, vp, i32, lng, vp]
    L.surf_prefetch_batch.argtypes = [vp, vp, i32, i32, i32, sz, sz]
    L.surf_lookahead_batch.argtypes = [vp, vp, i32, i32, i32, i32, sz, sz, vp, sz]
    L.surf_rand_create.argtypes = [C.c_uint32, C.POINTER(vp)]
    L.surf_rand_destroy.argtypes = [vp]
    L.surf_rand_destroy.restype = None
    L.surf_rand_next.argtypes = [vp]
    L.surf_match_points.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    L.surf_ransac_plane.argtypes = [vp, vp, vp, i32, C.c_float, i32, vp, vp, vp, vp, vp]
    L.surf_estimate_pose.argtypes = [vp, vp, vp, i32, vp, i32, vp, vp, vp]
    L.surf_mean_plane_distance.argtypes = [vp, i32, i32, sz, i32, vp, vp, vp]
    L.surf_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.surf_sync.argtypes = [vp]
    L.surf_integral.argtypes = [vp, vp, i32, i32, sz, i32, vp]
    L.surf_hessian_layer.argtypes = [vp, vp, i32, i32, sz, i32, i32, vp, C.POINTER(i32), C.POINTER(i32)]
    L.surf_raw_keypoints.argtypes = [vp, vp, i32, i32, sz, vp, sz, vp, i32, C.POINTER(i32)]
    L.surf_describe_keypoints.argtypes = [vp, vp, i32, i32, sz, vp, i32, vp, vp]
    L.surf_match_knn.argtypes = [vp, vp, i32, vp, i32, i32, i32, vp, i32]
    L.surf_match_ratio.argtypes = [vp, vp, i32, vp, i32, i32, C.c_float, vp, vp, i32]
    L.surf_merge_topk.argtypes = [vp, vp, i32, i32, i32, vp]
    L.surf_match_prev.argtypes = [vp, C.c_float, vp, vp, i32, lng]
    L.surf_stream_reset.argtypes = [vp]
    L.surf_device_matches.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.surf_set_match_mode.argtypes = [vp, i32]
    L.surf_match_redo_count.argtypes = [vp, C.POINTER(i32)]
    L.surf_set_async_outputs.argtypes = [vp, i32]
    L.surf_set_graphs.argtypes = [vp, i32, i32]
    L.surf_graph_replays.argtypes = [vp, C.POINTER(C.c_long)]
    L.surf_wait_outputs.argtypes = [vp]
    L.surf_db_create.argtypes = [vp, i32, lng, i32, C.POINTER(vp)]
    L.surf_db_destroy.argtypes = [vp]
    L.surf_db_destroy.restype = None
    L.surf_db_clear.argtypes = [vp]
    L.surf_db_add.argtypes = [vp, vp, vp, vp, i32, i32]
    L.surf_db_size.argtypes = [vp, C.POINTER(lng), C.POINTER(i32)]
    L.surf_db_knn.argtypes = [vp, vp, i32, i32, vp, i32]
    L.surf_correspond.argtypes = [vp, vp, vp, vp, vp, i32, C.POINTER(VoteParams), vp, vp, i32]
    L.surf_mark.argtypes = [vp, i32]
    L.surf_mark_elapsed.argtypes = [vp, i32, i32, C.POINTER(C.c_float)]
    L.surf_match_prev_time.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i32)]
    L.surf_comm_unique_id.argtypes = [vp]
    L.surf_comm_init.argtypes = [vp, vp, i32, i32, C.POINTER(vp)]
    L.surf_comm_destroy.argtypes = [vp]
    L.surf_comm_destroy.restype = None
    L.surf_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.surf_shard_range.argtypes = [lng, i32, i32, C.POINTER(lng), C.POINTER(lng)]
    L.surf_sweep_sharded.argtypes = [vp, vp, vp, i32, lng, i32, i32, sz, sz, vp, sz, i32, i32, vp, vp, i32, lng, vp,
                                     C.POINTER(SweepStats)]
    L.surf_db_knn_sharded.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32]
    L.surf_comm_last_times.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.surf_version.argtypes = []
    L.surf_version.restype = C.c_char_p
    for name in ABI_SYMBOLS:
        getattr(L, name)  # raises AttributeError if the library does not export a declared symbol
    _lib = L
    return L


def _u8(a, what="image"):
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise SurfError(-1, f"{what} must be a 2-D uint8 array (single channel, depth 8U)")
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data


class SURF:
    """cv::SURF(hessianThreshold, nOctaves=4, nOctaveLayers=2, extended=true, upright=false).

    Defaults follow the modern xfeatures2d::SURF::create(100, 4, 3, False, False), which are also the
    BASELINE config-1 values; pass the reference's (minHessian, 4, 2, ...) explicitly to mirror
    vtkSlicerSurfFeaturesLogic.cxx:1384.
    """

    def __init__(self, hessianThreshold: float = 100.0, nOctaves: int = 4, nOctaveLayers: int = 3,
                 extended: bool = False, upright: bool = False, *, device: int = 0, max_width: int = 1024,
                 max_height: int = 1024, max_batch: int = 1, max_keypoints: int = 32768):
        self._lib = load_library()
        self._h = C.c_void_p()
        self._p = Params(float(hessianThreshold), int(nOctaves), int(nOctaveLayers), int(bool(extended)),
                         int(bool(upright)))
        self.device, self.max_width, self.max_height = device, max_width, max_height
        self.max_batch, self.max_keypoints = max_batch, max_keypoints
        rc = self._lib.surf_create(C.byref(self._p), device, max_width, max_height, max_batch, max_keypoints,
                                   C.byref(self._h))
        if rc != SURF_OK:
            self._h = C.c_void_p()
            raise SurfError(rc, "surf_create failed (an sm_100 GPU is required; there is no CPU fallback)")

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.surf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != SURF_OK:
            raise SurfError(rc, self._lib.surf_last_error(self._h).decode())

    # ---- the five public fields of cv::SURF ----
    def _set(self, **kw):
        p = Params(self._p.hessian_threshold, self._p.n_octaves, self._p.n_octave_layers, self._p.extended, self._p.upright)
        for k, v in kw.items():
            setattr(p, k, v)
        self._check(self._lib.surf_set_params(self._h, C.byref(p)))
        self._p = p

    hessianThreshold = property(lambda s: s._p.hessian_threshold, lambda s, v: s._set(hessian_threshold=float(v)))
    nOctaves = property(lambda s: s._p.n_octaves, lambda s, v: s._set(n_octaves=int(v)))
    nOctaveLayers = property(lambda s: s._p.n_octave_layers, lambda s, v: s._set(n_octave_layers=int(v)))
    extended = property(lambda s: bool(s._p.extended), lambda s, v: s._set(extended=int(bool(v))))
    upright = property(lambda s: bool(s._p.upright), lambda s, v: s._set(upright=int(bool(v))))

    def descriptorSize(self) -> int:
        return int(self._lib.surf_descriptor_size(self._h))

    def descriptorType(self) -> int:
        return 5  # CV_32F

    def timings(self) -> dict:
        t = Timings()
        self._check(self._lib.surf_get_timings(self._h, C.byref(t)))
        return t.as_dict()

    def required_capacity(self) -> int:
        return int(self._lib.surf_required_capacity(self._h))

    @property
    def handle(self):
        return self._h

    @property
    def cuda_stream(self) -> int:
        return int(self._lib.surf_stream(self._h) or 0)

    # ---- cv::Feature2D ----
    def detect(self, image, mask=None) -> np.ndarray:
        """FeatureDetector::detect: oriented, sorted keypoints as a KEYPOINT_DTYPE array."""
        kps, _ = self._dac(image, mask, want_desc=False)
        return kps

    def detectAndCompute(self, image, mask=None):
        """SURF::operator()(img, mask, keypoints, descriptors): (keypoints, N x D float32)."""
        return self._dac(image, mask, want_desc=True)

    def _dac(self, image, mask, want_desc):
        image = np.asarray(image)
        D = self.descriptorSize()
        if image.size == 0:
            return np.zeros(0, KEYPOINT_DTYPE), (np.zeros((0, D), np.float32) if want_desc else None)
        img = _u8(image)
        m = None if mask is None else _u8(mask, "mask")
        if m is not None and m.shape != img.shape:
            raise SurfError(-1, "mask must have the image's size")
        cap = self.max_keypoints
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, D), np.float32) if want_desc else None
        n = C.c_int(0)
        H, W = img.shape
        rc = self._lib.surf_detect_and_compute(self._h, _ptr(img), W, H, img.strides[0], _ptr(m),
                                               0 if m is None else m.strides[0], _ptr(kps), _ptr(desc), cap, C.byref(n))
        self._check(rc)
        return kps[:n.value].copy(), (desc[:n.value].copy() if want_desc else None)

    def compute(self, image, keypoints):
        """DescriptorExtractor::compute: returns (surviving keypoints, descriptors)."""
        D = self.descriptorSize()
        kps = np.array(keypoints, dtype=KEYPOINT_DTYPE, copy=True).ravel()
        image = np.asarray(image)
        if image.size == 0 or len(kps) == 0:
            return np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, D), np.float32)
        img = _u8(image)
        H, W = img.shape
        desc = np.zeros((len(kps), D), np.float32)
        n = C.c_int(0)
        self._check(self._lib.surf_compute(self._h, _ptr(img), W, H, img.strides[0], _ptr(kps), len(kps), _ptr(desc),
                                           C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    # ---- batches (the frame loop of computeNext, vtkSlicerSurfFeaturesLogic.cxx:1392-1406) ----
    def detectAndComputeBatch(self, images, mask=None, want_descriptors=True, capacity: Optional[int] = None):
        """images: (B, H, W) uint8.  Returns a list of (keypoints, descriptors) per frame."""
        kps, desc, off = self.detectAndComputeBatchPacked(images, mask, want_descriptors, capacity)
        return [(kps[off[b]:off[b + 1]], None if desc is None else desc[off[b]:off[b + 1]]) for b in range(len(off) - 1)]

    def detectAndComputeBatchPacked(self, images, mask=None, want_descriptors=True, capacity: Optional[int] = None):
        images = np.ascontiguousarray(images, dtype=np.uint8)
        if images.ndim != 3:
            raise SurfError(-1, "images must be (B, H, W) uint8")
        B, H, W = images.shape
        m = None if mask is None else _u8(mask, "mask")
        if m is not None and m.shape != (H, W):
            raise SurfError(-1, "mask must have the frames' size")
        D = self.descriptorSize()
        cap = int(capacity) if capacity else B * self.max_keypoints
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, D), np.float32) if want_descriptors else None
        off = np.zeros(B + 1, np.int32)
        rc = self._lib.surf_detect_and_compute_batch(self._h, _ptr(images), MEM_HOST, B, W, H, W, W * H, _ptr(m),
                                                     0 if m is None else m.strides[0], _ptr(kps), _ptr(desc), MEM_HOST,
                                                     cap, _ptr(off))
        self._check(rc)
        n = int(off[B])
        return kps[:n], (None if desc is None else desc[:n]), off

    # ---- raw pointer batch API (device- or pinned-memory buffers owned by the caller) ----
    def submit(self, images_ptr: int, in_mem: int, B: int, W: int, H: int, pitch: int, frame_stride: int,
               want_descriptors: bool = True, mask_ptr: int = 0, mask_pitch: int = 0):
        self._check(self._lib.surf_submit_batch(self._h, images_ptr, in_mem, B, W, H, pitch, frame_stride,
                                                mask_ptr or None, mask_pitch, int(want_descriptors)))

    def prefetch(self, images_ptr: int, B: int, W: int, H: int, pitch: int, frame_stride: int):
        """Starts the upload of the NEXT batch's host frames while the submitted batch computes; the following
        submit() of the same pointer and geometry uses them (surf_prefetch_batch)."""
        self._check(self._lib.surf_prefetch_batch(self._h, images_ptr, B, W, H, pitch, frame_stride))

    def lookahead(self, images_ptr: int, in_mem: int, B: int, W: int, H: int, pitch: int, frame_stride: int,
                  mask_ptr: int = 0, mask_pitch: int = 0):
        """Uploads the NEXT batch (host frames) and runs its stages 1-3 beside the submitted batch's descriptor stage;
        the following submit() with the same arguments continues from there (surf_lookahead_batch)."""
        self._check(self._lib.surf_lookahead_batch(self._h, images_ptr, in_mem, B, W, H, pitch, frame_stride,
                                                   mask_ptr or None, mask_pitch))

    def collect(self, kps_ptr: int, desc_ptr: int, out_mem: int, capacity: int, offsets: np.ndarray):
        self._check(self._lib.surf_collect_batch(self._h, kps_ptr or None, desc_ptr or None, out_mem, capacity,
                                                 _ptr(offsets)))

    def device_results(self):
        k, d, o = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self._lib.surf_device_results(self._h, C.byref(k), C.byref(d), C.byref(o)))
        return k.value or 0, d.value or 0, o.value or 0

    def sync(self):
        self._check(self._lib.surf_sync(self._h))

    def set_async_outputs(self, enable: bool):
        """collect()/match_prev_to() return before their device->host copies finish; see wait_outputs()."""
        self._check(self._lib.surf_set_async_outputs(self._h, int(bool(enable))))

    def wait_outputs(self):
        self._check(self._lib.surf_wait_outputs(self._h))

    # ---- stage-level exports (parity of the per-stage tolerances) ----
    def integral(self, image, clamp01=False) -> np.ndarray:
        img = _u8(image)
        H, W = img.shape
        out = np.empty((H + 1, W + 1), np.int32)
        self._check(self._lib.surf_integral(self._h, _ptr(img), W, H, img.strides[0], int(clamp01), _ptr(out)))
        return out

    def hessian_layer(self, image, octave: int, layer: int) -> np.ndarray:
        img = _u8(image)
        H, W = img.shape
        out = np.empty((H >> octave, W >> octave), np.float32)
        r, c = C.c_int(0), C.c_int(0)
        self._check(self._lib.surf_hessian_layer(self._h, _ptr(img), W, H, img.strides[0], octave, layer, _ptr(out),
                                                 C.byref(r), C.byref(c)))
        assert (r.value, c.value) == out.shape
        return out

    def raw_keypoints(self, image, mask=None) -> np.ndarray:
        img = _u8(image)
        m = None if mask is None else _u8(mask, "mask")
        H, W = img.shape
        kps = np.zeros(self.max_keypoints, KEYPOINT_DTYPE)
        n = C.c_int(0)
        self._check(self._lib.surf_raw_keypoints(self._h, _ptr(img), W, H, img.strides[0], _ptr(m),
                                                 0 if m is None else m.strides[0], _ptr(kps), len(kps), C.byref(n)))
        return kps[:n.value].copy()

    def describe_keypoints(self, image, keypoints, want_desc=True, want_patches=False):
        """Orientation (+descriptor, +21x21 patch) of given keypoints, deleted ones kept with size=-1."""
        img = _u8(image)
        H, W = img.shape
        kps = np.array(keypoints, dtype=KEYPOINT_DTYPE, copy=True).ravel()
        D = self.descriptorSize()
        desc = np.zeros((len(kps), D), np.float32) if want_desc else None
        patches = np.zeros((len(kps), 21, 21), np.uint8) if want_patches else None
        self._check(self._lib.surf_describe_keypoints(self._h, _ptr(img), W, H, img.strides[0], _ptr(kps), len(kps),
                                                      _ptr(desc), _ptr(patches)))
        return kps, desc, patches

    # ---- stage 5 on this engine's stream ----
    def match_knn(self, query, train, k=2) -> np.ndarray:
        q = np.ascontiguousarray(query, np.float32)
        t = np.ascontiguousarray(train, np.float32)
        if q.ndim != 2 or t.ndim != 2 or (len(t) and q.shape[1] != t.shape[1]):
            raise SurfError(-1, "query/train must be N x D float32 with equal D")
        out = np.zeros((q.shape[0], k), DMATCH_DTYPE)
        self._check(self._lib.surf_match_knn(self._h, _ptr(q), q.shape[0], _ptr(t), t.shape[0], q.shape[1], k, _ptr(out),
                                             MEM_HOST))
        return out

    def match_ratio(self, query, train, ratio=0.8):
        q = np.ascontiguousarray(query, np.float32)
        t = np.ascontiguousarray(train, np.float32)
        out = np.zeros((q.shape[0], 2), DMATCH_DTYPE)
        good = np.zeros(q.shape[0], np.uint8)
        self._check(self._lib.surf_match_ratio(self._h, _ptr(q), q.shape[0], _ptr(t), t.shape[0], q.shape[1],
                                               float(ratio), _ptr(out), _ptr(good), MEM_HOST))
        return out, good.astype(bool)

    def match_ratio_device(self, q_ptr: int, nq: int, t_ptr: int, nt: int, dim: int, ratio: float, out_ptr: int,
                           good_ptr: int):
        self._check(self._lib.surf_match_ratio(self._h, q_ptr, nq, t_ptr, nt, dim, float(ratio), out_ptr, good_ptr,
                                               MEM_DEVICE))

    def set_match_mode(self, exact_only: bool):
        """False (default): tensor-core candidate search + exact rescoring; True: exact CUDA-core kernel."""
        self._check(self._lib.surf_set_match_mode(self._h, int(bool(exact_only))))

    def set_graphs(self, enable: bool, max_batch: int = 0):
        """CUDA-graph replay of the launch sequence of small batches (<= max_batch frames; 0 keeps the current limit)."""
        self._check(self._lib.surf_set_graphs(self._h, int(bool(enable)), int(max_batch)))

    def graph_replays(self) -> int:
        n = C.c_long(0)
        self._check(self._lib.surf_graph_replays(self._h, C.byref(n)))
        return n.value

    def match_redo_count(self) -> int:
        n = C.c_int(0)
        self._check(self._lib.surf_match_redo_count(self._h, C.byref(n)))
        return n.value

    def stream_reset(self):
        self._check(self._lib.surf_stream_reset(self._h))

    def match_prev(self, ratio=0.8, n_total: Optional[int] = None):
        """After detectAndComputeBatch*/collect: matches every frame of that batch to its predecessor.
        Returns ((N, 2) DMATCH_DTYPE, (N,) bool) packed like the batch's keypoints."""
        n = int(n_total)
        out = np.zeros((n, 2), DMATCH_DTYPE)
        good = np.zeros(n, np.uint8)
        self._check(self._lib.surf_match_prev(self._h, float(ratio), _ptr(out), _ptr(good), MEM_HOST, n))
        return out, good.astype(bool)

    def match_prev_device(self, ratio=0.8):
        """Same, results left on the device (device_matches())."""
        self._check(self._lib.surf_match_prev(self._h, float(ratio), None, None, MEM_DEVICE, 0))

    def match_prev_to(self, ratio, out_ptr: int, good_ptr: int, out_mem: int, capacity: int):
        self._check(self._lib.surf_match_prev(self._h, float(ratio), out_ptr or None, good_ptr or None, out_mem, capacity))

    def device_matches(self):
        m, g = C.c_void_p(), C.c_void_p()
        self._check(self._lib.surf_device_matches(self._h, C.byref(m), C.byref(g)))
        return m.value or 0, g.value or 0

    def merge_topk_device(self, in_ptr: int, n_ranks: int, nq: int, k: int, out_ptr: int):
        self._check(self._lib.surf_merge_topk(self._h, in_ptr, n_ranks, nq, k, out_ptr))

    # ---- timing marks on the engine's stream ----
    def mark(self, slot: int):
        self._check(self._lib.surf_mark(self._h, slot))

    def mark_elapsed(self, a: int, b: int) -> float:
        ms = C.c_float(0)
        self._check(self._lib.surf_mark_elapsed(self._h, a, b, C.byref(ms)))
        return float(ms.value)

    def match_prev_time(self):
        """(device ms, kernel launches) of the last match_prev* call; waits for it."""
        ms, n = C.c_float(0), C.c_int(0)
        self._check(self._lib.surf_match_prev_time(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    # ---- frame-sharded sweep (BASELINE config 4) ----
    def sweep_sharded(self, comm, frames_local: np.ndarray, n_frames_total: int, mask=None, want_descriptors=True,
                      dst: int = 0, capacity: Optional[int] = None):
        """computeNext over a sweep sharded by frame (surf_sweep_sharded): `frames_local` is this rank's contiguous
        block (shard_range(n_frames_total, rank, world)), host memory.  Returns (kps, desc, offsets[int64], stats)
        on rank `dst` and (None, None, None, stats) elsewhere.  comm=None: one GPU."""
        fr = np.ascontiguousarray(frames_local, dtype=np.uint8)
        if fr.ndim != 3:
            raise SurfError(-1, "frames_local must be (n, H, W) uint8")
        n, H, W = fr.shape
        rank, world = (comm.rank, comm.world) if comm is not None else (0, 1)
        m = None if mask is None else _u8(mask, "mask")
        if m is not None and m.shape != (H, W):
            raise SurfError(-1, "mask must have the frames' size")
        D = self.descriptorSize()
        st = SweepStats()
        kps = desc = off = None
        cap = 0
        if rank == dst:
            cap = int(capacity) if capacity else n_frames_total * self.max_keypoints
            kps = np.zeros(cap, KEYPOINT_DTYPE)
            desc = np.zeros((cap, D), np.float32) if want_descriptors else None
            off = np.zeros(n_frames_total + 1, np.int64)
        self._check(self._lib.surf_sweep_sharded(
            self._h, None if comm is None else comm._h, _ptr(fr) if n else None, MEM_HOST, n_frames_total, W, H, W, W * H,
            _ptr(m), 0 if m is None else m.strides[0], int(want_descriptors), dst, _ptr(kps), _ptr(desc), MEM_HOST, cap,
            _ptr(off), C.byref(st)))
        if rank != dst:
            return None, None, None, st.as_dict()
        nk = int(off[-1])
        return kps[:nk], (None if desc is None else desc[:nk]), off, st.as_dict()

    def sweep_sharded_device(self, comm, images_ptr: int, n_frames_total: int, W: int, H: int, pitch: int,
                             frame_stride: int, dst: int, kps_ptr: int, desc_ptr: int, capacity: int,
                             offsets: Optional[np.ndarray], want_descriptors: bool = True):
        """Raw-pointer form: device-resident frames in, device-resident gathered results out (on `dst`)."""
        st = SweepStats()
        self._check(self._lib.surf_sweep_sharded(
            self._h, None if comm is None else comm._h, images_ptr or None, MEM_DEVICE, n_frames_total, W, H, pitch,
            frame_stride, None, 0, int(want_descriptors), dst, kps_ptr or None, desc_ptr or None, MEM_DEVICE, capacity,
            _ptr(offsets), C.byref(st)))
        return st.as_dict()


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block [lo, hi) of rank `rank` (surf_shard_range): the first n_items % world ranks get one more."""
    lo, hi = C.c_long(0), C.c_long(0)
    rc = load_library().surf_shard_range(n_items, rank, world, C.byref(lo), C.byref(hi))
    if rc != SURF_OK:
        raise SurfError(rc, "bad shard_range arguments")
    return lo.value, hi.value


class Comm:
    """NCCL communicator of the engine's device (surf_comm_*): one process per GPU.  The 128-byte id comes from
    Comm.unique_id() on one rank and reaches the others by any transport (bench.py: torch.distributed broadcast)."""

    def __init__(self, engine: SURF, unique_id: bytes, rank: int, world: int):
        if len(unique_id) != 128:
            raise SurfError(-1, "unique_id must be 128 bytes")
        self._engine, self._lib = engine, engine._lib
        self.rank, self.world = rank, world
        h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        engine._check(self._lib.surf_comm_init(engine._h, buf, rank, world, C.byref(h)))
        self._h = h

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = load_library().surf_comm_unique_id(buf)
        if rc != SURF_OK:
            raise SurfError(rc, "ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
        return buf.raw

    @classmethod
    def from_torch(cls, engine: SURF):
        """Rendezvous over an initialised torch.distributed process group (plumbing only)."""
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        t = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            t = torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0)
        return cls(engine, bytes(t.cpu().numpy().tobytes()), rank, world)

    def info(self):
        r, n, v = C.c_int(0), C.c_int(0), C.c_int(0)
        self._engine._check(self._lib.surf_comm_info(self._h, C.byref(r), C.byref(n), C.byref(v)))
        return {"rank": r.value, "world": n.value, "nccl_version": v.value}

    def last_times(self):
        a, b = C.c_float(0), C.c_float(0)
        self._engine._check(self._lib.surf_comm_last_times(self._h, C.byref(a), C.byref(b)))
        return {"total_ms": float(a.value), "nccl_ms": float(b.value)}

    def close(self):
        if getattr(self, "_h", None) and getattr(self._engine, "_h", None):
            self._lib.surf_comm_destroy(self._h)
        self._h = None

    __del__ = close


class DescriptorDB:
    """Device-resident train database (surf_db_*): what cv::DescriptorMatcher::add(vector<Mat>) builds at
    vtkSlicerSurfFeaturesLogic.cxx:1402-1403.  Rows are prepared for the tensor-core matcher once per add()."""

    def __init__(self, engine: SURF, dim: int = 64, max_rows: int = 1 << 20, max_images: int = 4096):
        self._engine = engine
        self._lib = engine._lib
        self.dim = dim
        h = C.c_void_p()
        engine._check(self._lib.surf_db_create(engine._h, dim, max_rows, max_images, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(self._engine, "_h", None):
            self._lib.surf_db_destroy(self._h)
        self._h = None

    __del__ = close

    def clear(self):
        self._engine._check(self._lib.surf_db_clear(self._h))

    def add(self, descriptors: Sequence[np.ndarray], keypoints: Optional[Sequence[np.ndarray]] = None):
        """One entry per train image (imgIdx = position in the add() order)."""
        descs = [np.ascontiguousarray(d, np.float32).reshape(-1, self.dim) for d in descriptors]
        if not descs:
            return
        off = np.concatenate([[0], np.cumsum([len(d) for d in descs])]).astype(np.int32)
        allrows = np.ascontiguousarray(np.concatenate(descs, axis=0))
        kp = None
        if keypoints is not None:
            kp = np.ascontiguousarray(np.concatenate([np.asarray(k, KEYPOINT_DTYPE) for k in keypoints]))
            if len(kp) != len(allrows):
                raise SurfError(-1, "keypoints and descriptors must have the same number of rows")
        self._engine._check(self._lib.surf_db_add(self._h, _ptr(allrows), _ptr(kp), _ptr(off), len(descs), MEM_HOST))

    def add_device(self, desc_ptr: int, kps_ptr: int, offsets: np.ndarray):
        """Rows already on the device (e.g. SURF.device_results() of a collected batch)."""
        off = np.ascontiguousarray(offsets, np.int32)
        self._engine._check(self._lib.surf_db_add(self._h, desc_ptr, kps_ptr or None, _ptr(off), len(off) - 1, MEM_DEVICE))

    def size(self):
        rows, images = C.c_long(0), C.c_int(0)
        self._engine._check(self._lib.surf_db_size(self._h, C.byref(rows), C.byref(images)))
        return rows.value, images.value

    def knnMatch(self, queryDescriptors, k=2) -> np.ndarray:
        q = np.ascontiguousarray(queryDescriptors, np.float32)
        out = np.zeros((q.shape[0], k), DMATCH_DTYPE)
        self._engine._check(self._lib.surf_db_knn(self._h, _ptr(q), q.shape[0], k, _ptr(out), MEM_HOST))
        return out

    def knnMatchSharded(self, comm, queryDescriptors, k=2, query_root: int = -1) -> np.ndarray:
        """knnMatch against a database whose rows are sharded over the ranks of `comm` (this object holds this
        rank's block).  trainIdx is the global row; every rank gets the merged result (surf_db_knn_sharded)."""
        q = np.ascontiguousarray(queryDescriptors, np.float32)
        out = np.zeros((q.shape[0], k), DMATCH_DTYPE)
        self._engine._check(self._lib.surf_db_knn_sharded(self._h, None if comm is None else comm._h, _ptr(q), q.shape[0], k,
                                                          query_root, _ptr(out), MEM_HOST))
        return out

    def knn_sharded_device(self, comm, q_ptr: int, nq: int, k: int, out_ptr: int, query_root: int = -1):
        self._engine._check(self._lib.surf_db_knn_sharded(self._h, None if comm is None else comm._h, q_ptr, nq, k,
                                                          query_root, out_ptr, MEM_DEVICE))


def correspond(train: DescriptorDB, bogus: DescriptorDB, query_descriptors: Sequence[np.ndarray],
               query_keypoints: Sequence[np.ndarray], ref_x: int, ref_y: int, bogus_ratio: float = 0.95):
    """computeNextInterSliceCorrespondence (vtkSlicerSurfFeaturesLogic.cxx:1434-1571) for all query images.

    Returns (after_hough, results): after_hough[i] is the DMATCH array of query image i after the bogus filter
    (entries the Hough vote rejected carry -1 indices, as afterHoughMatches does); results is a
    VOTE_RESULT_DTYPE array (n_matches, hough_count, best_image, best_count)."""
    eng = train._engine
    descs = [np.ascontiguousarray(d, np.float32).reshape(-1, train.dim) for d in query_descriptors]
    off = np.concatenate([[0], np.cumsum([len(d) for d in descs])]).astype(np.int32)
    n = int(off[-1])
    allq = np.ascontiguousarray(np.concatenate(descs, axis=0)) if n else np.zeros((0, train.dim), np.float32)
    kp = np.ascontiguousarray(np.concatenate([np.asarray(k, KEYPOINT_DTYPE) for k in query_keypoints])) if n else \
        np.zeros(0, KEYPOINT_DTYPE)
    matches = np.zeros(max(n, 1), DMATCH_DTYPE)
    results = np.zeros(len(descs), VOTE_RESULT_DTYPE)
    vp = VoteParams(float(bogus_ratio), int(ref_x), int(ref_y))
    eng._check(eng._lib.surf_correspond(train._h, bogus._h, _ptr(allq), _ptr(kp), _ptr(off), len(descs), C.byref(vp),
                                        _ptr(matches), _ptr(results), MEM_HOST))
    return [matches[off[i]:off[i] + results["n_matches"][i]].copy() for i in range(len(descs))], results


class BFMatcher:
    """cv::BFMatcher(NORM_L2): add / clear / knnMatch with DMatch records (imgIdx bookkeeping as
    DescriptorMatcher::add(vector<Mat>) does; vtkSlicerSurfFeaturesLogic.cxx:1402-1403, :640).
    The added train images live in a device-resident DescriptorDB."""

    def __init__(self, engine: Optional[SURF] = None, *, max_rows: int = 1 << 18, max_images: int = 1024, **engine_kw):
        self._engine = engine or SURF(**engine_kw)
        self._train: list[np.ndarray] = []
        self._db: Optional[DescriptorDB] = None
        self._cap = (max_rows, max_images)

    def add(self, descriptors: Sequence[np.ndarray]):
        new = [np.ascontiguousarray(d, np.float32) for d in descriptors]
        if not new:
            return
        if self._db is None:
            self._db = DescriptorDB(self._engine, new[0].shape[1], *self._cap)
        self._db.add(new)
        self._train.extend(new)

    def clear(self):
        self._train = []
        if self._db is not None:
            self._db.clear()

    def getTrainDescriptors(self):
        return list(self._train)

    def knnMatch(self, queryDescriptors, trainDescriptors=None, k=2) -> np.ndarray:
        """Returns an (Nq, k) DMATCH_DTYPE array, ascending distance, ties -> lower imgIdx, lower trainIdx."""
        if trainDescriptors is not None:
            return self._engine.match_knn(queryDescriptors, trainDescriptors, k)
        if not self._train:
            return np.zeros((len(queryDescriptors), 0), DMATCH_DTYPE)
        return self._db.knnMatch(queryDescriptors, k)

    def ratioMatch(self, queryDescriptors, trainDescriptors, ratio=0.8):
        return self._engine.match_ratio(queryDescriptors, trainDescriptors, ratio)


def to_cv2_keypoints(kps: np.ndarray):
    """KEYPOINT_DTYPE array -> tuple of cv2.KeyPoint (needs cv2; a consumer-side convenience only)."""
    import cv2

    return tuple(cv2.KeyPoint(float(k["x"]), float(k["y"]), float(k["size"]), float(k["angle"]), float(k["response"]),
                              int(k["octave"]), int(k["class_id"])) for k in kps)
