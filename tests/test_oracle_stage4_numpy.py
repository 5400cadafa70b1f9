"""CPU pin of the oracle's stage 4 (orientation + descriptor) against an independent restatement
written in numpy on top of GENUINE OpenCV sub-steps: cv2.integral, cv2.fastAtan2 and cv2.resize(INTER_AREA) do the work
the C oracle restates by hand.  Only the glue (sample lattice, window walk, sums) and the 2.4.5 Gaussian taps (cv2 4.13
rounds them one ulp differently, see test_oracle_pins.py::test_gaussian_kernels) are written twice.  Pure-Python loops:
a handful of keypoints per case."""
import math

import cv2
import numpy as np
import pytest

from surffeatures_b200 import speckle

F = np.float32


def _gauss_245(n, sigma):
    """getGaussianKernel(n, sigma, CV_32F) as OpenCV 2.4.5 computes it: fp32 taps summed in fp64, scaled, rounded once more."""
    x = np.arange(n, dtype=np.float64) - (n - 1) * 0.5
    t = np.exp(-0.5 / (sigma * sigma) * x * x).astype(F)
    return (t.astype(np.float64) * (1.0 / t.astype(np.float64).sum())).astype(F)


def _cv_round(v):
    return int(np.rint(v))  # half to even, as cvRound (lrint)


def _haar(S, x, y, g):
    """x / y Haar responses of side g at integral-image position (x, y); (int box sum) * w as fp32, summed in fp64."""
    h = g // 2
    wa, wn = F(1) / (F(h) * F(g)), F(-1) / (F(h) * F(g))

    def box(x0, y0, x1, y1):
        return int(S[y + y0, x + x0]) + int(S[y + y1, x + x1]) - int(S[y + y0, x + x1]) - int(S[y + y1, x + x0])

    left, right = box(0, 0, h, g), box(h, 0, g, g)
    top, bottom = box(0, 0, g, h), box(0, h, g, g)
    vx = F(float(F(F(left) * wn)) + float(F(F(right) * wa)))
    vy = F(float(F(F(top) * wa)) + float(F(F(bottom) * wn)))
    return vx, vy


def _orientation(img, S, cx, cy, size):
    H, W = img.shape
    s = F(F(size) * F(1.2)) / F(9.0)
    g = 2 * _cv_round(F(2) * s)
    if H + 1 < g or W + 1 < g:
        return None
    G = _gauss_245(13, 2.5)
    hg = F(g - 1) / F(2)
    X, Y, A = [], [], []
    for i in range(-6, 7):
        for j in range(-6, 7):
            if i * i + j * j > 36:
                continue
            x = _cv_round(F(F(cx) + F(F(i) * s)) - hg)
            y = _cv_round(F(F(cy) + F(F(j) * s)) - hg)
            if y < 0 or y >= (H + 1) - g or x < 0 or x >= (W + 1) - g:
                continue
            vx, vy = _haar(S, x, y, g)
            w = F(G[i + 6] * G[j + 6])
            X.append(F(vx * w))
            Y.append(F(vy * w))
            A.append(_cv_round(F(cv2.fastAtan2(float(Y[-1]), float(X[-1])))))
    if not X:
        return None
    best, bx, by = F(0), F(0), F(0)
    for i in range(0, 360, 5):
        sx, sy = F(0), F(0)
        for x, y, a in zip(X, Y, A):
            d = abs(a - i)
            if d < 30 or d > 330:
                sx, sy = F(sx + x), F(sy + y)
        mod = F(F(sx * sx) + F(sy * sy))
        if mod > best:
            best, bx, by = mod, sx, sy
    return F(cv2.fastAtan2(float(-by), float(bx)))


def _window(img, cx, cy, size, angle):
    H, W = img.shape
    s = F(F(size) * F(1.2)) / F(9.0)
    n = int(F(21) * s)
    th = F(F(angle) * F(math.pi / 180))
    sd, cd = F(-math.sin(float(th))), F(math.cos(float(th)))  # correctly rounded, as the oracle fixes it
    off = F(-F(n - 1) / F(2))
    sx = F(F(F(cx) + F(off * cd)) + F(off * sd))
    sy = F(F(F(cy) - F(off * sd)) + F(off * cd))
    win = np.zeros((n, n), np.uint8)
    for i in range(n):
        px, py = float(sx), float(sy)
        for j in range(n):
            ix, iy = math.floor(px), math.floor(py)
            if 0 <= ix < W - 1 and 0 <= iy < H - 1:
                a, b = F(px - ix), F(py - iy)
                na, nb = F(F(1) - a), F(F(1) - b)
                p = img[iy:iy + 2, ix:ix + 2].astype(F)
                v = F(F(F(F(p[0, 0] * na) * nb) + F(F(p[0, 1] * a) * nb)) + F(F(p[1, 0] * na) * b))
                v = F(v + F(F(p[1, 1] * a) * b))
                win[i, j] = _cv_round(v)
            else:
                x = min(max(_cv_round(px), 0), W - 1)
                y = min(max(_cv_round(py), 0), H - 1)
                win[i, j] = img[y, x]
            px, py = px + float(cd), py - float(sd)
        sx, sy = F(sx + sd), F(sy + cd)
    return win


def _descriptor(patch, extended):
    gd = _gauss_245(20, 3.3)                                     # 2.4.5 closed form (SURVEY C-4)
    P = patch.astype(np.int32)
    vec = []
    for I in range(4):
        for J in range(4):
            acc = [F(0)] * (8 if extended else 4)
            for y in range(5 * I, 5 * I + 5):
                for xx in range(5 * J, 5 * J + 5):
                    w = F(gd[y] * gd[xx])
                    tx = F(F(P[y, xx + 1] - P[y, xx] + P[y + 1, xx + 1] - P[y + 1, xx]) * w)
                    ty = F(F(P[y + 1, xx] - P[y, xx] + P[y + 1, xx + 1] - P[y, xx + 1]) * w)
                    if extended:
                        k = 0 if ty >= 0 else 2
                        acc[k], acc[k + 1] = F(acc[k] + tx), F(acc[k + 1] + abs(tx))
                        k = 4 if tx >= 0 else 6
                        acc[k], acc[k + 1] = F(acc[k] + ty), F(acc[k + 1] + abs(ty))
                    else:
                        acc = [F(acc[0] + tx), F(acc[1] + ty), F(acc[2] + abs(tx)), F(acc[3] + abs(ty))]
            vec += acc
    vec = np.array(vec, F)
    mag = float(sum(float(F(v * v)) for v in vec))
    return (vec * F(1.0 / (math.sqrt(mag) + np.finfo(np.float64).eps))).astype(F)


@pytest.mark.parametrize("extended", [False, True])
def test_stage4_equals_numpy_cv2_restatement(oracle, extended):
    img = speckle.speckle_frame(160, 200, 4)
    p = oracle.params(600, 3, 2, extended, False)
    raw = oracle.raw_keypoints(img, p)
    assert len(raw) >= 12
    pick = raw[np.linspace(0, len(raw) - 1, 12).astype(int)]   # strong and weak, small and large
    kps, desc, patches = oracle.describe(img, p, pick, want_desc=True, want_patches=True)
    S = cv2.integral(img)
    checked = 0
    for k, d, pt in zip(kps, desc, patches):
        if k["size"] <= 0:
            continue
        ang = _orientation(img, S, k["x"], k["y"], k["size"])
        assert ang is not None and ang == k["angle"]
        win = _window(img, k["x"], k["y"], k["size"], ang)
        n = win.shape[0]
        mine = cv2.resize(win, (21, 21), interpolation=cv2.INTER_AREA)   # genuine OpenCV INTER_AREA
        assert np.array_equal(mine, pt), (n, int(np.abs(mine.astype(int) - pt).max()))
        mv = _descriptor(mine, extended)
        assert np.array_equal(mv, d), float(np.abs(mv - d).max())
        checked += 1
    assert checked >= 10


def test_stage4_near_border_takes_the_clamped_taps(oracle):
    """A keypoint whose window leaves the frame: nearest-pixel taps with clamping on both sides of the restatement."""
    img = speckle.speckle_frame(96, 128, 9)
    p = oracle.params(100, 4, 3, False, False)
    kp = np.zeros(2, oracle.KEYPOINT_DTYPE)
    kp["x"], kp["y"], kp["size"], kp["angle"] = [14.3, 120.2], [11.6, 85.1], [15, 21], [-1, -1]
    kps, desc, patches = oracle.describe(img, p, kp, want_desc=True, want_patches=True)
    S = cv2.integral(img)
    for k, d, pt in zip(kps, desc, patches):
        assert k["size"] > 0
        ang = _orientation(img, S, k["x"], k["y"], k["size"])
        assert ang == k["angle"]
        mine = cv2.resize(_window(img, k["x"], k["y"], k["size"], ang), (21, 21), interpolation=cv2.INTER_AREA)
        assert np.array_equal(mine, pt)
        assert np.array_equal(_descriptor(mine, False), d)
