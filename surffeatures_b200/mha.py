"""MHA sequence ingest -- Python mirror of the surf_mha_* / surf_sweep_mha C ABI.

Reference: SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx read_mha :598-632 (with its helpers :444-595),
cropData :745-752, computeCentroid :839-860, readAndComputeFeaturesOnMhaFile :1337-1379.
`write_mha` produces files of the same layout from synthetic frames (there is no data set offline)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from .engine import KEYPOINT_DTYPE, SURF, SurfError, DescriptorDB, load_library, _ptr

SKIP_INVALID_BYTES = 1
REFERENCE_CROP_RATIOS = (np.float32(1 / 5.), np.float32(3. / 3.9), np.float32(1 / 6.), np.float32(3. / 4.))  # :1008-1009


class MhaInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("width", "height", "frames_in_file", "first_frame", "last_frame", "n_frames",
                                          "n_transforms")]


def _lib():
    L = load_library()
    if not getattr(L, "_mha_ready", False):
        vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
        L.surf_mha_open.argtypes = [C.c_char_p, i32, i32, i32, C.POINTER(vp), C.POINTER(MhaInfo)]
        L.surf_mha_close.argtypes = [vp]
        L.surf_mha_close.restype = None
        L.surf_mha_last_error.argtypes = [vp]
        L.surf_mha_last_error.restype = C.c_char_p
        L.surf_mha_transforms.argtypes = [vp, vp, i32]
        L.surf_mha_validity.argtypes = [vp, vp, i32]
        L.surf_mha_frame_name.argtypes = [vp, i32, C.c_char_p, i32]
        L.surf_mha_read.argtypes = [vp, i32, i32, vp, sz]
        L.surf_crop_rect.argtypes = [i32, i32, C.POINTER(C.c_float * 4), C.POINTER(C.c_int32 * 4)]
        L.surf_mask_centroid.argtypes = [vp, i32, i32, sz, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.surf_sweep_mha.argtypes = [vp, vp, C.POINTER(C.c_float * 4), vp, sz, vp, vp, vp, vp, C.c_long, C.POINTER(C.c_int32)]
        L._mha_ready = True
    return L


def crop_rect(width: int, height: int, crop_ratios: Sequence[float]):
    """cropData's cv::Rect: (x, y, w, h)."""
    r = (C.c_float * 4)(*[float(np.float32(v)) for v in crop_ratios])
    out = (C.c_int32 * 4)()
    rc = _lib().surf_crop_rect(width, height, C.byref(r), C.byref(out))
    if rc:
        raise SurfError(rc, "crop ratios do not give a rectangle inside the frame")
    return tuple(out)


def mask_centroid(mask: np.ndarray):
    m = np.asarray(mask, np.uint8)
    if m.strides[1] != 1:
        m = np.ascontiguousarray(m)
    x, y = C.c_int32(0), C.c_int32(0)
    rc = _lib().surf_mask_centroid(m.ctypes.data, m.shape[1], m.shape[0], m.strides[0], C.byref(x), C.byref(y))
    if rc:
        raise SurfError(rc, "bad mask")
    return x.value, y.value


class MhaSequence:
    """read_mha(filename, images, filenames, transforms, transformsValidity, firstFrame, lastFrame)."""

    def __init__(self, path: str, first_frame: int = -1, last_frame: int = -1, flags: int = 0):
        self._lib = _lib()
        self._h = C.c_void_p()
        self.info = MhaInfo()
        rc = self._lib.surf_mha_open(str(path).encode(), first_frame, last_frame, flags, C.byref(self._h), C.byref(self.info))
        if rc:
            self._h = C.c_void_p()
            raise SurfError(rc, f"cannot read {path} as an MHA sequence")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.surf_mha_close(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __len__(self):
        return self.info.n_frames

    @property
    def shape(self):
        return self.info.height, self.info.width

    def transforms(self) -> np.ndarray:
        t = np.zeros((self.info.n_transforms, 12), np.float32)
        if len(t):
            rc = self._lib.surf_mha_transforms(self._h, t.ctypes.data, len(t))
            assert rc == 0, rc
        return t

    def validity(self) -> np.ndarray:
        n = self.info.last_frame - self.info.first_frame + 1
        v = np.zeros(n, np.uint8)
        rc = self._lib.surf_mha_validity(self._h, v.ctypes.data, n)
        assert rc == 0, rc
        return v.astype(bool)

    def names(self):
        buf = C.create_string_buffer(4096)
        out = []
        for i in range(self.info.n_transforms):
            rc = self._lib.surf_mha_frame_name(self._h, i, buf, 4096)
            assert rc == 0, rc
            out.append(buf.value.decode())
        return out

    def read(self, start: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = len(self) - start if count is None else count
        out = np.zeros((count, self.info.height, self.info.width), np.uint8)
        if count:
            rc = self._lib.surf_mha_read(self._h, start, count, out.ctypes.data, out.strides[0])
            if rc:
                raise SurfError(rc, self._lib.surf_mha_last_error(self._h).decode())
        return out


def sweep(engine: SURF, seq: MhaSequence, crop_ratios: Optional[Sequence[float]] = REFERENCE_CROP_RATIOS,
          mask: Optional[np.ndarray] = None, db: Optional[DescriptorDB] = None, want_host: bool = True):
    """readAndComputeFeaturesOnMhaFile + the computeNext loop: every kept frame cropped, detected (with the mask's
    same rectangle) and described.  Returns (keypoints, descriptors, offsets) packed in frame order (None, None,
    offsets when want_host is False); appends every frame to `db` on the device when given."""
    L = _lib()
    n = len(seq)
    D = engine.descriptorSize()
    cap = max(n * engine.max_keypoints, 1)
    kps = np.zeros(cap, KEYPOINT_DTYPE) if want_host else None
    desc = np.zeros((cap, D), np.float32) if want_host else None
    off = np.zeros(n + 1, np.int32)
    r = None if crop_ratios is None else (C.c_float * 4)(*[float(np.float32(v)) for v in crop_ratios])
    m = None
    if mask is not None:
        m = np.asarray(mask, np.uint8)
        if m.strides[1] != 1:
            m = np.ascontiguousarray(m)
        if m.shape != seq.shape:
            raise SurfError(-1, "the mask must have the size of the (uncropped) frames")
    nf = C.c_int32(0)
    engine._check(L.surf_sweep_mha(engine._h, seq._h, None if r is None else C.byref(r), _ptr(m),
                                   0 if m is None else m.strides[0], None if db is None else db._h, _ptr(kps), _ptr(desc),
                                   _ptr(off), cap, C.byref(nf)))
    total = int(off[n])
    if not want_host:
        return None, None, off
    return kps[:total], desc[:total], off


def write_mha(path: str, frames: np.ndarray, transforms: Optional[np.ndarray] = None,
              status: Optional[Sequence[bool]] = None, field: str = "ProbeToTrackerTransform",
              extra_fields: Sequence[str] = ("ReferenceToTrackerTransform",)):
    """Writes a PLUS-style sequence file: text header (DimSize, per-frame 4x4 transforms + status), then raw frames."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    if transforms is None:
        transforms = np.tile(np.eye(4, dtype=np.float64).ravel(), (n, 1))
        transforms[:, 3] = np.arange(n) * 0.5
    transforms = np.asarray(transforms, np.float64).reshape(n, 16)
    status = [True] * n if status is None else list(status)
    lines = ["ObjectType = Image", "NDims = 3", "AnatomicalOrientation = RAI", "BinaryData = True",
             "BinaryDataByteOrderMSB = False", "CenterOfRotation = 0 0 0", "CompressedData = False",
             f"DimSize = {w} {h} {n}", "ElementSpacing = 1 1 1", "ElementType = MET_UCHAR", "Offset = 0 0 0",
             "TransformMatrix = 1 0 0 0 1 0 0 0 1", "UltrasoundImageOrientation = MF"]
    for i in range(n):
        nums = " ".join(f"{v:.6g}" for v in transforms[i])
        lines.append(f"Seq_Frame{i:04d}_{field} = {nums} ")
        lines.append(f"Seq_Frame{i:04d}_{field}Status = {'OK' if status[i] else 'INVALID'}")
        for ef in extra_fields:
            lines.append(f"Seq_Frame{i:04d}_{ef} = 1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1 ")
            lines.append(f"Seq_Frame{i:04d}_{ef}Status = OK")
        lines.append(f"Seq_Frame{i:04d}_Timestamp = {0.05 * i:.4f}")
    lines.append("ElementDataFile = LOCAL")
    with open(path, "wb") as f:
        f.write(("\n".join(lines) + "\n").encode())
        f.write(frames.tobytes())
