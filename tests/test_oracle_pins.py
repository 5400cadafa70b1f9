"""CPU suite: pins the oracle's sub-steps against cv2 4.13 (the only genuine OpenCV code available
offline), against an independent numpy restatement, and against the committed golden fixtures.

The reference repository holds no tests or golden vectors for this path
(SurfFeatures/Testing/Cxx/CMakeLists.txt:4-6,16), so these pins are what anchors the oracle.
"""
import json
import os

import numpy as np
import pytest

from surffeatures_b200 import speckle

cv2 = pytest.importorskip("cv2")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------- cv2 sub-step pins
@pytest.mark.parametrize("shape", [(480, 640), (1, 1), (37, 501), (1024, 1024)])
def test_integral_equals_cv2(oracle, shape):
    img = np.random.default_rng(shape[1]).integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(oracle.integral(img), cv2.integral(img))


def test_mask_integral_is_integral_of_min_mask_1(oracle):
    m = (np.random.default_rng(2).integers(0, 4, (60, 80)) * 85).astype(np.uint8)
    assert np.array_equal(oracle.integral(m, clamp01=True), cv2.integral(np.minimum(m, 1)))


def test_fast_atan2_bit_exact_vs_cv2(oracle):
    rng = np.random.default_rng(1)
    xy = rng.standard_normal((4000, 2)).astype(np.float32)
    xy[:50] *= np.float32(1e-6)
    special = [(0, 0), (0, 1), (1, 0), (0, -1), (-1, 0), (-1e-9, 1), (1, 1), (-1, -1), (3, -3)]
    for y, x in list(map(tuple, xy)) + special:
        assert oracle.fast_atan2(y, x) == cv2.fastAtan2(float(np.float32(y)), float(np.float32(x)))
    assert oracle.fast_atan2(-1e-9, 1.0) == 360.0


def test_gaussian_kernels(oracle):
    # odd length: the closed form; cv2 4.13 agrees to the last ulp or two
    g13 = oracle.gaussian_kernel(13, 2.5)
    assert np.abs(g13 - cv2.getGaussianKernel(13, 2.5, cv2.CV_32F).ravel()).max() < 3e-8
    x = np.arange(13) - 6.0
    ref = np.exp(-x * x / (2 * 2.5 ** 2))
    assert np.abs(g13 - ref / ref.sum()).max() < 3e-8
    assert abs(float(g13.sum()) - 1) < 1e-6
    # even length (descriptor weights): 2.4.5 closed form, NOT the 4.x centre-tap quirk (SURVEY C-4)
    g20 = oracle.gaussian_kernel(20, 3.3)
    x = np.arange(20) - 9.5
    ref = np.exp(-x * x / (2 * 3.3 ** 2))
    assert np.abs(g20 - ref / ref.sum()).max() < 3e-8
    assert np.abs(g20 - cv2.getGaussianKernel(20, 3.3, cv2.CV_32F).ravel()).max() > 5e-4


@pytest.mark.parametrize("n", list(range(21, 131)) + [147, 168, 184, 201, 235, 302, 336, 352, 470, 604, 739])
def test_resize_area_bit_exact_vs_cv2(oracle, n):
    w = np.random.default_rng(n).integers(0, 256, (n, n), dtype=np.uint8)
    assert np.array_equal(oracle.resize_area_21(w), cv2.resize(w, (21, 21), interpolation=cv2.INTER_AREA))


@pytest.mark.parametrize("n", range(1, 21))
def test_resize_area_enlarging_bit_exact_vs_cv2(oracle, n):
    """Windows narrower than 21 px (keypoint size < 7.9 handed to compute()): INTER_AREA runs cv::resize's bilinear
    fixed-point code with the area-mode coefficients.  Several seeds, plus flat and extreme inputs."""
    rng = np.random.default_rng(1000 + n)
    wins = [rng.integers(0, 256, (n, n), dtype=np.uint8) for _ in range(8)]
    wins += [np.zeros((n, n), np.uint8), np.full((n, n), 255, np.uint8), (np.indices((n, n)).sum(0) % 2 * 255).astype(np.uint8)]
    for w in wins:
        assert np.array_equal(oracle.resize_area_21(w), cv2.resize(w, (21, 21), interpolation=cv2.INTER_AREA))


def test_knn_match_vs_cv2_bfmatcher(oracle):
    rng = np.random.default_rng(0)
    t = rng.standard_normal((400, 64)).astype(np.float32)
    q = rng.standard_normal((300, 64)).astype(np.float32)
    t[100] = t[3]  # duplicate rows: ties go to the lower trainIdx
    q[0] = t[3]
    m = oracle.knn_match(q, t, 2)
    ref = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, k=2)
    ridx = np.array([[d.trainIdx for d in row] for row in ref])
    rdist = np.array([[d.distance for d in row] for row in ref], np.float32)
    assert np.array_equal(m["trainIdx"], ridx)
    assert np.allclose(m["distance"], rdist, rtol=2e-6, atol=1e-7)
    assert list(m["trainIdx"][0]) == [3, 100]


# ---------------------------------------------------------------- independent numpy restatement
def _np_box(S, y1, x1, y2, x2, step, ni, nj):
    """Box sums over [y1,y2) x [x1,x2) for every filter origin (i*step, j*step)."""
    ys = np.arange(ni)[:, None] * step
    xs = np.arange(nj)[None, :] * step
    return S[ys + y1, xs + x1] + S[ys + y2, xs + x2] - S[ys + y2, xs + x1] - S[ys + y1, xs + x2]


def _np_hessian(S, size, step):
    """Independent vectorised restatement of SURVEY A3.2/A3.3 (fp32 products, fp64 sums)."""
    H, W = S.shape[0] - 1, S.shape[1] - 1
    det = np.zeros((H // step, W // step), np.float32)
    if size > H or size > W:
        return det
    ratio = np.float32(size) / np.float32(9)
    r = lambda c: int(np.rint(np.float64(ratio * np.float32(c))))
    ni, nj, m = 1 + (H - size) // step, 1 + (W - size) // step, (size // 2) // step

    def resp(boxes):
        acc = np.zeros((ni, nj), np.float64)
        for x1, y1, x2, y2, w in boxes:
            X1, Y1, X2, Y2 = r(x1), r(y1), r(x2), r(y2)
            wt = np.float32(w) / (np.float32(X2 - X1) * np.float32(Y2 - Y1))
            acc += (_np_box(S, Y1, X1, Y2, X2, step, ni, nj).astype(np.float32) * wt).astype(np.float64)
        return acc.astype(np.float32)

    dx = resp([(0, 2, 3, 7, 1), (3, 2, 6, 7, -2), (6, 2, 9, 7, 1)])
    dy = resp([(2, 0, 7, 3, 1), (2, 3, 7, 6, -2), (2, 6, 7, 9, 1)])
    dxy = resp([(1, 1, 4, 4, 1), (5, 1, 8, 4, -1), (1, 5, 4, 8, -1), (5, 5, 8, 8, 1)])
    det[m:m + ni, m:m + nj] = dx * dy - (np.float32(0.81) * dxy) * dxy
    return det


@pytest.mark.parametrize("octave,layer", [(0, 0), (0, 4), (1, 2), (2, 1), (3, 3)])
def test_hessian_layer_vs_numpy_restatement(oracle, octave, layer):
    img = speckle.speckle_frame(240, 320, 3, blobs=4)
    S = oracle.integral(img)
    size, step = (9 + 6 * layer) << octave, 1 << octave
    det, _, _ = oracle.hessian_layer(S, size, step)
    assert np.array_equal(det, _np_hessian(S.astype(np.int64), size, step))


def test_maxima_are_strict_and_above_threshold(oracle):
    """Every raw keypoint comes from a strict 3x3x3 maximum above the threshold (recomputed here
    with numpy on the oracle's own layers), and the count of such maxima bounds the keypoints."""
    img = speckle.speckle_frame(240, 320, 3, blobs=4)
    S = oracle.integral(img)
    p = oracle.params(300, 2, 3)
    raw = oracle.raw_keypoints(img, p)
    n_max = 0
    for o in range(2):
        step = 1 << o
        sizes = [(9 + 6 * l) << o for l in range(5)]
        dets = [oracle.hessian_layer(S, s, step)[0] for s in sizes]
        for l in (1, 2, 3):
            margin = (sizes[l + 1] // 2) // step + 1
            rows, cols = dets[l].shape
            c = dets[l][margin:rows - margin, margin:cols - margin]
            ok = c > np.float32(300)
            for k in (l - 1, l, l + 1):
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        if k == l and dy == 0 and dx == 0:
                            continue
                        ok &= c > dets[k][margin + dy:rows - margin + dy, margin + dx:cols - margin + dx]
            n_max += int(ok.sum())
    assert 0 < len(raw) <= n_max
    assert (raw["response"] > 300).all()
    # sorted by the reference comparator's first key
    assert (np.diff(raw["response"]) <= 0).all()


def test_blob_is_found_where_it_is(oracle):
    """SURVEY Appendix B sanity: a sigma=3.2 Gaussian blob at (150,100) gives the top maximum there."""
    yy, xx = np.mgrid[0:240, 0:320].astype(np.float64)
    img = (255 * np.exp(-((xx - 150) ** 2 + (yy - 100) ** 2) / (2 * 3.2 ** 2))).astype(np.uint8)
    raw = oracle.raw_keypoints(img, oracle.params(100, 4, 3))
    assert len(raw) >= 1
    assert abs(raw[0]["x"] - 150) < 0.05 and abs(raw[0]["y"] - 100) < 0.05 and 15 <= raw[0]["size"] <= 27


def test_descriptor_properties(oracle, frame640):
    p = oracle.params(100, 4, 3)
    k, d = oracle.detect_and_compute(frame640, p)
    assert np.allclose(np.linalg.norm(d, axis=1), 1, atol=1e-5)
    assert ((k["angle"] >= 0) & (k["angle"] <= 360)).all()
    # 64-d and 128-d describe the same keypoints; upright fixes the angle at 270
    k2, d2 = oracle.detect_and_compute(frame640, oracle.params(100, 4, 3, True, True))
    assert d2.shape[1] == 128 and (k2["angle"] == 270).all() and len(k2) == len(k)
    # detect() then compute() equals detectAndCompute() (the reference calls them separately)
    kd = oracle.detect(frame640, p)
    kc, dc = oracle.compute(frame640, p, kd)
    assert kc.tobytes() == k.tobytes() and np.array_equal(dc, d)


def test_rotation_covariance(oracle):
    """Rotating the image by 90 degrees rotates keypoints and leaves descriptors (nearly) unchanged."""
    img = speckle.speckle_frame(256, 256, 12, blobs=6)
    p = oracle.params(400, 3, 3)
    k0, d0 = oracle.detect_and_compute(img, p)
    k1, d1 = oracle.detect_and_compute(np.ascontiguousarray(np.rot90(img)), p)
    assert abs(len(k0) - len(k1)) <= 0.05 * len(k0)
    m = oracle.knn_match(d0[:200], d1, 1)
    assert (m["distance"][:, 0] < 0.25).mean() > 0.8


# ---------------------------------------------------------------- golden fixtures
def test_speckle_generator_is_pinned():
    meta = json.load(open(os.path.join(GOLDEN, "golden.json")))
    for name, spec in meta["frames"].items():
        f = _golden_frame(spec)
        assert speckle.sha256(f) == spec["sha256"], name


def _golden_frame(spec):
    if spec["kind"] == "frame":
        return speckle.speckle_frame(spec["h"], spec["w"], spec["seed"], blobs=spec.get("blobs", 0))
    return speckle.speckle_stream(spec["n"], spec["h"], spec["w"], seed=spec["seed"])[spec["index"]]


def test_oracle_reproduces_golden_vectors(oracle):
    meta = json.load(open(os.path.join(GOLDEN, "golden.json")))
    z = np.load(os.path.join(GOLDEN, "golden_surf.npz"))
    for name, spec in meta["frames"].items():
        img = _golden_frame(spec)
        p = oracle.params(**spec["params"])
        k, d = oracle.detect_and_compute(img, p)
        assert int(oracle.integral(img).astype(np.int64).sum()) == spec["integral_checksum"]
        assert k.tobytes() == z[name + "_kp"].tobytes(), name
        assert np.array_equal(d, z[name + "_desc"]), name


def _np_refine(N, step, ds, x, y, size):
    """interpolateKeypoint: central differences, Matx33f::solve by the closed-form 3x3 Cramer rule in fp32
    (OpenCV's Matx_FastSolveOp<float, 3, 1>), written here a second time in numpy scalars."""
    f = np.float32
    N0, N1, N2 = N
    b = [f(-(f(N1[5] - N1[3]) / f(2))), f(-(f(N1[7] - N1[1]) / f(2))), f(-(f(N2[4] - N0[4]) / f(2)))]
    dxx = f(f(N1[3] - f(f(2) * N1[4])) + N1[5])
    dxy = f(f(f(f(N1[8] - N1[6]) - N1[2]) + N1[0]) / f(4))
    dxs = f(f(f(f(N2[5] - N2[3]) - N0[5]) + N0[3]) / f(4))
    dyy = f(f(N1[1] - f(f(2) * N1[4])) + N1[7])
    dys = f(f(f(f(N2[7] - N2[1]) - N0[7]) + N0[1]) / f(4))
    dss = f(f(N0[4] - f(f(2) * N1[4])) + N2[4])
    a = [[dxx, dxy, dxs], [dxy, dyy, dys], [dxs, dys, dss]]

    def m2(p, q, r, s):
        return f(f(p * q) - f(r * s))

    det = f(f(f(a[0][0] * m2(a[1][1], a[2][2], a[2][1], a[1][2])) - f(a[0][1] * m2(a[1][0], a[2][2], a[2][0], a[1][2])))
            + f(a[0][2] * m2(a[1][0], a[2][1], a[2][0], a[1][1])))
    if det == 0:
        return None
    d = f(f(1) / det)
    x0 = f(d * f(f(f(b[0] * m2(a[1][1], a[2][2], a[1][2], a[2][1])) - f(a[0][1] * m2(b[1], a[2][2], a[1][2], b[2])))
                 + f(a[0][2] * m2(b[1], a[2][1], a[1][1], b[2]))))
    x1 = f(d * f(f(f(a[0][0] * m2(b[1], a[2][2], a[1][2], b[2])) - f(b[0] * m2(a[1][0], a[2][2], a[1][2], a[2][0])))
                 + f(a[0][2] * m2(a[1][0], b[2], b[1], a[2][0]))))
    x2 = f(d * f(f(f(a[0][0] * m2(a[1][1], b[2], b[1], a[2][1])) - f(a[0][1] * m2(a[1][0], b[2], b[1], a[2][0])))
                 + f(b[0] * m2(a[1][0], a[2][1], a[1][1], a[2][0]))))
    if (x0 == 0 and x1 == 0 and x2 == 0) or abs(x0) > 1 or abs(x1) > 1 or abs(x2) > 1:
        return None
    return (f(x + f(x0 * f(step))), f(y + f(x1 * f(step))), f(np.rint(f(size + f(x2 * f(ds))))))


def test_raw_keypoints_equal_numpy_restatement_of_stage3(oracle):
    """Maxima search, sub-pixel / scale interpolation and the final sort, restated in numpy on the oracle's own layers
    (which test_hessian_layer_vs_numpy_restatement pins): the raw keypoints must come out byte for byte."""
    img = speckle.speckle_frame(200, 260, 13, blobs=3)
    S = oracle.integral(img)
    thr, O, L = 250.0, 3, 3
    raw = oracle.raw_keypoints(img, oracle.params(thr, O, L))
    mine = []
    for o in range(O):
        step = 1 << o
        sizes = [(9 + 6 * l) << o for l in range(L + 2)]
        layers = [oracle.hessian_layer(S, s, step) for s in sizes]
        dets = [t[0] for t in layers]
        traces = [t[1] for t in layers]
        for l in range(1, L + 1):
            size = sizes[l]
            if sizes[l + 1] > min(img.shape):       # a layer that does not fit leaves nothing to compare with
                continue
            margin = (sizes[l + 1] // 2) // step + 1
            rows, cols = dets[l].shape
            for i in range(margin, rows - margin):
                for j in range(margin, cols - margin):
                    v = dets[l][i, j]
                    if not v > np.float32(thr):
                        continue
                    N = [dets[k][i - 1:i + 2, j - 1:j + 2].reshape(-1) for k in (l - 1, l, l + 1)]
                    if not all(v > N[k][q] for k in range(3) for q in range(9) if not (k == 1 and q == 4)):
                        continue
                    si, sj = step * (i - (size // 2) // step), step * (j - (size // 2) // step)
                    x0, y0 = np.float32(sj + (size - 1) * 0.5), np.float32(si + (size - 1) * 0.5)
                    r = _np_refine(N, step, size - sizes[l - 1], x0, y0, np.float32(size))
                    if r is None:
                        continue
                    tr = traces[l][i, j]
                    mine.append((float(r[0]), float(r[1]), float(r[2]), -1.0, float(v), o, int(tr > 0) - int(tr < 0)))
    mine.sort(key=lambda k: (-k[4], -k[2], -k[5], -k[1], k[0]))
    got = np.array(mine, dtype=oracle.KEYPOINT_DTYPE)
    assert len(got) == len(raw) > 50
    assert got.tobytes() == raw.tobytes()


# ---------------------------------------------------------------- Appendix-C reading choices as named switches
def test_oracle_switches_change_what_they_name_and_nothing_else():
    """every doubtful upstream detail is a named run-time switch of the oracle so that a genuine
    OpenCV (tests/test_opencv_probe.py) can arbitrate.  All off = the 2.4.5 reading the CUDA path implements."""
    import cv2

    from oracle import surf_oracle as so
    from surffeatures_b200 import speckle

    img = speckle.speckle_frame(240, 320, 11)
    p = so.params(100, 4, 3, False, False)
    kp0, d0 = so.detect_and_compute(img, p)
    assert len(kp0) > 300
    # C-2: the convention before 2.4.4 reports the mirrored angle; the sampled window is the same physical one
    with so.switches(angle_pre244=True):
        kp1, d1 = so.detect_and_compute(img, p)
        kpu, _ = so.detect_and_compute(img, so.params(100, 4, 3, False, True))
    assert np.array_equal(kp0["x"], kp1["x"]) and np.array_equal(kp0["size"], kp1["size"])
    mirrored = np.mod(360.0 - kp1["angle"].astype(np.float64), 360.0)
    assert np.abs(((kp0["angle"] - mirrored + 180.0) % 360.0) - 180.0).max() < 1e-3
    e1 = np.linalg.norm(d0 - d1, axis=1)      # sin / cos of the mirrored angle: last-bit effects (a tap rounds the other way)
    assert np.median(e1) == 0 and np.quantile(e1, 0.95) < 1e-3
    assert (kpu["angle"] == 90.0).all()
    # C-4: the installed OpenCV's even-length Gaussian differs from the 2.4.5 closed form on the centre taps
    g_cv = cv2.getGaussianKernel(20, 3.3, cv2.CV_32F).reshape(-1)
    with so.switches(desc_gauss=g_cv):
        kp2, d2 = so.detect_and_compute(img, p)
    assert np.array_equal(kp0["angle"], kp2["angle"])                       # orientation does not use that kernel
    e = np.linalg.norm(d0 - d2, axis=1)
    assert (g_cv != so.gaussian_kernel(20, 3.3)).any() and 0 < np.median(e) < 0.05
    # libm sinf / cosf for the dominant angle: the platform-dependent last bit
    with so.switches(libm_sincos=True):
        kp3, d3 = so.detect_and_compute(img, p)
    assert np.array_equal(kp0["angle"], kp3["angle"])
    e3 = np.linalg.norm(d0 - d3, axis=1)
    assert np.median(e3) == 0 and np.quantile(e3, 0.95) < 1e-3
    # back to the defaults
    kp4, d4 = so.detect_and_compute(img, p)
    assert kp4.tobytes() == kp0.tobytes() and d4.tobytes() == d0.tobytes()
