"""CPU suite for row f4 (plane fit + pose estimate): pins of the oracle restatement (oracle/pose_oracle.c) and of the
host-side functions of the product that need no GPU.

The reference (SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:2085-2188, :1655-1851) has no tests or fixtures for
this code, and vnl / VTK are absent: what CAN be pinned is pinned here -- rand() against this machine's real glibc,
the plane fit against a literal Python transcription and against known answers."""
import ctypes
import math

import numpy as np
import pytest

from oracle import pose_oracle as po


def _libc():
    return ctypes.CDLL("libc.so.6")


@pytest.mark.parametrize("seed", [1, 0, 42, 2147483647, 2147483653, 4294967295])
def test_rand_restatement_equals_glibc(seed):
    libc = _libc()
    libc.srand(ctypes.c_uint(seed))
    want = [libc.rand() for _ in range(3000)]
    r = po.Rand(seed)
    assert [r.next() for _ in range(3000)] == want
    # the product's generator (host code of libsurf_b200.so, no GPU involved)
    from surffeatures_b200.pose import Rand

    pr = Rand(seed)
    assert [pr.next() for _ in range(3000)] == want


def test_unseeded_rand_is_seed_1():
    # "If no seed value is provided, rand() is automatically seeded with a value of 1" (rand(3)); a fresh process
    # is needed to observe it, so check the documented equivalence through the restatement only
    a, b = po.Rand(1), po.Rand(0)
    assert [a.next() for _ in range(10)] == [b.next() for _ in range(10)]
    assert po.Rand(1).next() == 1804289383  # the well-known first value of glibc rand()


def _py_normalize(a):
    t = 0.0
    for x in a:
        t += x * x
    if t != 0:
        t = 1.0 / math.sqrt(t)
        a = [t * x for x in a]
    return a


def _py_dot(a, b):
    ip = 0.0
    for x, y in zip(a, b):
        ip += x * y
    return ip


def _py_ransac(points, threshold, iterations, rng):
    """Literal transcription of ransac() :2085-2188 (Python floats are IEEE doubles)."""
    n = len(points)
    best_n, best_err = float(n), 1.7976931348623157e308
    inliers, plane = [], []
    thr = np.float32(threshold)
    for _ in range(iterations):
        idx = [0, 0, 0]
        while idx[0] == idx[1] or idx[1] == idx[2] or idx[0] == idx[2]:
            idx = [rng.next() % n, rng.next() % n, rng.next() % n]
        p1, p2, p3 = (list(points[i]) for i in idx)
        v1 = _py_normalize([a - b for a, b in zip(p1, p2)])
        v2 = _py_normalize([a - b for a, b in zip(p1, p3)])
        if math.sqrt(_py_dot(v1, v1)) < 0.99 or math.sqrt(_py_dot(v2, v2)) < 0.99:
            continue
        if abs(_py_dot(v1, v2)) > 0.99:
            continue
        N = _py_normalize([v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]])
        D = -_py_dot(p1, N)
        n_out, err, tmp = 0.0, 0.0, []
        for j in range(n):
            if j in idx:
                continue
            dist = np.float32(abs(_py_dot(N, list(points[j])) + D))
            if dist > thr:
                n_out += 1
            else:
                err += float(dist)
                tmp.append(j)
        if n_out < best_n or (n_out == best_n and err < best_err):
            best_n, best_err, plane, inliers = n_out, err, list(idx), tmp
    return inliers + plane, plane, best_err, int(best_n)


@pytest.mark.parametrize("n,seed", [(4, 1), (12, 1), (60, 7)])
def test_ransac_oracle_equals_literal_transcription(n, seed):
    g = np.random.default_rng(n)
    pts = g.normal(size=(n, 3)) * [30, 30, 1.5] + [10, -5, 40]
    pts[::5] += g.normal(size=pts[::5].shape) * 10          # outliers
    rc, inl, plane, err, nout = po.ransac_plane(po.Rand(seed), pts, margin=1.0, iterations=60)
    want = _py_ransac(pts, 1.0, 60, po.Rand(seed))
    assert rc == 0
    assert inl.tolist() == want[0] and plane.tolist() == want[1] and err == want[2] and nout == want[3]


def test_ransac_known_answer_plane():
    g = np.random.default_rng(3)
    xy = g.uniform(-50, 50, size=(80, 2))
    z = 0.5 * xy[:, 0] - 0.25 * xy[:, 1] + 2.0
    pts = np.column_stack([xy, z])
    pts[:8, 2] += 25.0                                       # eight gross outliers
    rc, inl, plane, err, nout = po.ransac_plane(po.Rand(1), pts, margin=1.0, iterations=1000)
    assert rc == 0 and nout == 8
    assert sorted(inl.tolist()) == list(range(8, 80)) and err < 1e-9
    assert all(p >= 8 for p in plane)


def test_ransac_edge_cases():
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], float)
    rc, inl, plane, err, nout = po.ransac_plane(po.Rand(1), pts)
    assert rc == 0 and inl.tolist() == [0, 1, 2] and plane.tolist() == [0, 1, 2]
    rc, inl, plane, _, _ = po.ransac_plane(po.Rand(1), pts[:2])
    assert rc == 1 and len(inl) == 0
    # all points coincide: every hypothesis is skipped (the reference would then index an empty vector)
    rc, inl, _, _, _ = po.ransac_plane(po.Rand(1), np.ones((6, 3)), iterations=50)
    assert rc == 2


def _scene(n=40, seed=5):
    """Matches of one query frame against 3 train frames whose tracker transforms are known."""
    g = np.random.default_rng(seed)
    from oracle.surf_oracle import DMATCH_DTYPE, KEYPOINT_DTYPE

    tk = np.zeros(3 * 50, KEYPOINT_DTYPE)
    tk["x"], tk["y"] = g.uniform(0, 360, 150), g.uniform(0, 280, 150)
    qk = np.zeros(60, KEYPOINT_DTYPE)
    qk["x"], qk["y"] = g.uniform(0, 360, 60), g.uniform(0, 280, 60)
    m = np.zeros(n, DMATCH_DTYPE)
    m["queryIdx"] = g.permutation(60)[:n]                   # distinct: two inliers at one location make acos(0/0)
    m["trainIdx"] = g.integers(0, 50, n)
    m["imgIdx"] = g.integers(0, 3, n)
    m["imgIdx"][::7] = -1                                    # rejected by the Hough vote
    T = np.tile(np.eye(4, dtype=np.float32), (3, 1, 1))
    for i in range(3):
        a = 0.02 * i
        T[i, :3, :3] = [[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]]
        T[i, :3, 3] = [5 * i, -3 * i, 0.4 * i]
    I2P = np.diag([0.1, 0.1, 0.1, 1.0])
    I2P[:3, 3] = [-18, 2, 0.5]
    return m, qk, tk, np.array([0, 50, 100, 150], np.int32), T.reshape(3, 16), I2P


def test_product_host_functions_equal_oracle():
    """surf_match_points / surf_estimate_pose / surf_mean_plane_distance are host arithmetic of the product: they
    must reproduce the oracle bit for bit on this machine (same statements, same libm)."""
    from surffeatures_b200 import pose

    m, qk, tk, off, T, I2P = _scene()
    tp, qp = pose.match_points(m, qk, tk, off, T, I2P)
    otp, oqp = po.match_points(m, qk, tk, off, T, I2P)
    assert len(tp) == int((m["imgIdx"] != -1).sum())
    assert np.array_equal(tp, otp) and np.array_equal(qp, oqp)
    # plane + inliers from the oracle; the pose from both, with the generator in the same state
    r = po.Rand(1)
    rc, inl, plane, _, _ = po.ransac_plane(r, otp, margin=1.0, iterations=200)
    assert rc == 0
    draws = 0
    probe = po.Rand(1)
    while [probe.r[i] for i in range(31)] != [r.r[i] for i in range(31)] or probe.f != r.f:
        probe.next()
        draws += 1
    pr = pose.Rand(1)
    for _ in range(draws):
        pr.next()
    est, mmd = pose.estimate_pose(pr, qp, tp, inl, plane)
    oest, ommd = po.estimate_pose(r, oqp, otp, inl, plane)
    assert np.isfinite(oest).all()
    assert np.array_equal(est, oest) and mmd == ommd
    assert pr.next() == r.next()                             # both consumed the same number of draws
    assert np.allclose(np.linalg.norm(est[:3, 0]), 0.10763) and np.allclose(np.linalg.norm(est[:3, 2]), 0.10646)
    mask = np.zeros((280, 364), np.uint8)
    mask[40:200, 60:300] = 255
    gt = np.eye(4)
    gt[:3, 3] = [1, 2, 3]
    assert pose.mean_plane_distance(mask, gt, est) == po.mean_plane_distance(mask, gt, oest)
