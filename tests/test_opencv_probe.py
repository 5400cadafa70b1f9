"""Run-time probe for genuine OpenCV SURF (last row).

The reference's arithmetic is OpenCV 2.4.5 `nonfree` SURF, which cannot be installed offline: cv2 in this image is the
headless non-contrib wheel, so these tests SKIP here and parity stays "unpinned by the reference".  On a box whose cv2 has
`xfeatures2d.SURF_create` (an opencv-contrib build with OPENCV_ENABLE_NONFREE) they compare the oracle -- and, on a GPU
box, the CUDA path -- with genuine SURF at BASELINE.json's tolerances, and say which of the Appendix-C reading choices the
installed OpenCV confirms (newer OpenCV changed some of them after 2.4.5; a mismatch names the switch)."""
import numpy as np
import pytest

from surffeatures_b200 import speckle


def _genuine_surf():
    cv2 = pytest.importorskip("cv2")
    xf = getattr(cv2, "xfeatures2d", None)
    if xf is None or not hasattr(xf, "SURF_create"):
        pytest.skip("cv2.xfeatures2d.SURF_create is not available (non-contrib OpenCV): genuine SURF cannot arbitrate here")
    try:
        return cv2, xf.SURF_create(hessianThreshold=100, nOctaves=4, nOctaveLayers=3, extended=False, upright=False)
    except cv2.error as ex:   # contrib built without OPENCV_ENABLE_NONFREE
        pytest.skip(f"SURF_create refuses to run: {ex}")


def _paired_fraction(a, b):
    """Fraction of keypoints of b with a partner in a within 0.5 px and 2 % scale (BASELINE.json)."""
    ok = 0
    for k in b:
        d2 = (a["x"] - k["x"]) ** 2 + (a["y"] - k["y"]) ** 2
        ok += bool(((d2 <= 0.25) & (np.abs(a["size"] - k["size"]) <= 0.02 * k["size"])).any())
    return ok / max(len(b), 1)


def _to_struct(cv_kps, dtype):
    out = np.zeros(len(cv_kps), dtype)
    for i, k in enumerate(cv_kps):
        out[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
    return out


def _compare(name, mine_k, mine_d, cv_k, cv_d):
    frac = _paired_fraction(mine_k, cv_k)
    assert frac >= 0.99, f"{name}: only {frac:.4f} of genuine SURF's keypoints are paired within 0.5 px / 2 % scale"
    # descriptors of identically placed keypoints
    key = {(round(float(k["x"]), 3), round(float(k["y"]), 3), float(k["size"])): i for i, k in enumerate(mine_k)}
    errs, ang = [], []
    for j, k in enumerate(cv_k):
        i = key.get((round(float(k["x"]), 3), round(float(k["y"]), 3), float(k["size"])))
        if i is not None:
            errs.append(float(np.linalg.norm(mine_d[i] - cv_d[j])))
            ang.append(abs(float(mine_k["angle"][i]) - float(k["angle"])))
    assert len(errs) >= 0.9 * len(cv_k), f"{name}: only {len(errs)} of {len(cv_k)} keypoints coincide exactly"
    switches = {
        "C-2 angle convention (fastAtan2(-sum_y, sum_x), upright = 270)": float(np.quantile(ang, 0.99)) < 0.05,
        "C-4 closed-form even Gaussian / C-5 INTER_AREA / C-7 fp64 Haar accumulation (descriptor L2 <= 1e-4)":
            float(np.quantile(errs, 0.99)) <= 1e-4,
    }
    failed = [s for s, ok in switches.items() if not ok]
    assert not failed, f"{name}: the installed OpenCV disagrees with the reading choice(s): {failed}"


def test_oracle_against_genuine_opencv_surf(oracle):
    cv2, surf = _genuine_surf()
    img = speckle.speckle_frame(480, 640, 0)
    ck, cd = surf.detectAndCompute(img, None)
    p = oracle.params(100, 4, 3, False, False)
    ok, od = oracle.detect_and_compute(img, p)
    try:
        _compare("oracle", ok, od, _to_struct(ck, ok.dtype), cd)
    except AssertionError as first:
        # the installed OpenCV is not 2.4.5: let it arbitrate between the Appendix-C readings (oracle switches)
        g_cv = cv2.getGaussianKernel(20, 3.3, cv2.CV_32F).reshape(-1)
        verdicts = []
        for angle_pre244 in (False, True):
            for libm in (False, True):
                for gauss in (None, g_cv):
                    if not angle_pre244 and not libm and gauss is None:
                        continue
                    name = (f"C-2 angle_pre244={angle_pre244}, libm_sincos={libm}, "
                            f"C-4 desc_gauss={'cv2.getGaussianKernel' if gauss is not None else 'closed form'}")
                    with oracle.switches(angle_pre244=angle_pre244, libm_sincos=libm, desc_gauss=gauss):
                        k2, d2 = oracle.detect_and_compute(img, p)
                    try:
                        _compare("oracle", k2, d2, _to_struct(ck, k2.dtype), cd)
                        verdicts.append(name)
                    except AssertionError:
                        pass
        raise AssertionError(f"{first}\n  readings of Appendix C the installed OpenCV {cv2.__version__} agrees with: "
                             f"{verdicts or 'none of the named switches'}") from first


@pytest.mark.gpu
def test_gpu_against_genuine_opencv_surf():
    cv2, surf = _genuine_surf()
    from surffeatures_b200 import SURF

    img = speckle.speckle_frame(480, 640, 0)
    ck, cd = surf.detectAndCompute(img, None)
    e = SURF(100, 4, 3, False, False, max_width=640, max_height=480, max_batch=1, max_keypoints=16384)
    try:
        k, d = e.detectAndCompute(img)
    finally:
        e.close()
    _compare("gpu", k, d, _to_struct(ck, k.dtype), cd)
