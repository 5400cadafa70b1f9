"""CPU suite: the C-ABI library loads, exports every symbol include/surf_b200.h declares, keeps the
cv::KeyPoint / cv::DMatch layouts, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "surf_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(surf_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    import surffeatures_b200 as sb

    assert _declared_symbols() == sorted(sb.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    import surffeatures_b200 as sb

    lib = sb.load_library()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.surf_version()


def test_record_layouts_match_opencv():
    import surffeatures_b200 as sb

    assert sb.KEYPOINT_DTYPE.itemsize == 28
    assert [sb.KEYPOINT_DTYPE.fields[f][1] for f in ("x", "y", "size", "angle", "response", "octave", "class_id")] == \
        [0, 4, 8, 12, 16, 20, 24]
    assert sb.DMATCH_DTYPE.itemsize == 16
    assert [sb.DMATCH_DTYPE.fields[f][1] for f in ("queryIdx", "trainIdx", "imgIdx", "distance")] == [0, 4, 8, 12]
    cv2 = pytest.importorskip("cv2")
    kp = sb.to_cv2_keypoints(np.array([(1.5, 2.5, 15, 90, 123.0, 1, -1)], sb.KEYPOINT_DTYPE))
    assert isinstance(kp[0], cv2.KeyPoint) and kp[0].pt == (1.5, 2.5) and kp[0].octave == 1 and kp[0].class_id == -1


def test_no_cpu_fallback():
    """Without an sm_100 device construction must fail loudly (this test is the CPU box's view)."""
    import torch

    import surffeatures_b200 as sb

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: covered by the gpu suite")
    with pytest.raises(sb.SurfError) as ei:
        sb.SURF(100)
    assert ei.value.code == -5


def test_bad_arguments_are_status_codes_not_crashes():
    import surffeatures_b200 as sb

    lib = sb.load_library()
    h = ctypes.c_void_p()
    assert lib.surf_create(None, 0, 640, 480, 1, 1024, ctypes.byref(h)) == -1
    p = sb.engine.Params(-1.0, 4, 3, 0, 0)
    assert lib.surf_create(ctypes.byref(p), 0, 640, 480, 1, 1024, ctypes.byref(h)) == -1
    p = sb.engine.Params(100.0, 0, 3, 0, 0)
    assert lib.surf_create(ctypes.byref(p), 0, 640, 480, 1, 1024, ctypes.byref(h)) == -1
    p = sb.engine.Params(100.0, 4, 3, 0, 0)
    assert lib.surf_create(ctypes.byref(p), 0, 0, 480, 1, 1024, ctypes.byref(h)) == -1
    # a band of the integral-image kernel holds whole rows in shared memory: frames wider than 7 224 px are refused
    assert lib.surf_create(ctypes.byref(p), 0, 7225, 480, 1, 1024, ctypes.byref(h)) == -2
    assert lib.surf_sync(None) == -1
    assert lib.surf_last_error(None) == b"null engine"


def test_product_never_touches_the_oracle():
    """The shipped package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "surffeatures_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cuh", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "surf_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f
    import subprocess

    out = subprocess.run(["ldd", os.path.join(pkg, "libsurf_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
