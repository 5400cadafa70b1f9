"""CPU suite: pins the oracle's restatement of the "next" rows f2 / f3 -- DescriptorMatcher::add / knnMatch
over several train images, the bogus filter, houghTransform and the per-image vote
(vtkSlicerSurfFeaturesLogic.cxx:73-345, :1402-1403, :1434-1571, :839-860).

Pins available here: genuine cv2.BFMatcher.add(list)+knnMatch for the database semantics, and an independent,
literal pure-Python transcription of the reference's Hough code (float32 via numpy scalars, this machine's
float libm) for the vote.  The reference has no tests or vectors for this path: PARITY UNPINNED upstream."""
import math

import numpy as np
import pytest

KP = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"),
               ("class_id", "<i4")])
f32 = np.float32


def _unit_rows(rng, n, dim=64):
    t = np.abs(rng.standard_normal((n, dim))).astype(np.float32)
    return t / np.linalg.norm(t, axis=1, keepdims=True)


_CV2_SCRIPT = r"""
import json, sys
import numpy as np
import cv2
d = np.load(sys.argv[1])
mats = [d[k] for k in sorted(d.files) if k.startswith("m")]
q = d["q"]
cv2.setNumThreads(1)


def self_consistent(ref):
    # every (imgIdx, trainIdx, distance) cv2 reports must describe a row at that distance
    for i, row in enumerate(ref):
        for m in row:
            if not (0 <= m.imgIdx < len(mats) and 0 <= m.trainIdx < len(mats[m.imgIdx])):
                return False
            true = float(np.sqrt(((q[i].astype(np.float64) - mats[m.imgIdx][m.trainIdx]) ** 2).sum()))
            if abs(true - m.distance) > 1e-5:
                return False
    return True


for attempt in range(20):
    bf = cv2.BFMatcher(cv2.NORM_L2)
    bf.add(mats)
    ref = bf.knnMatch(q, k=3)
    if self_consistent(ref):
        print(json.dumps([[(m.imgIdx, m.trainIdx, float(m.distance)) for m in row] for row in ref]))
        break
else:
    print(json.dumps(None))
"""


def test_db_knn_equals_cv2_bfmatcher_add(oracle, tmp_path):
    """Genuine cv2.BFMatcher.add(list) + knnMatch.  cv2 runs in a fresh interpreter: with torch imported first in
    the same process (as other tests of this suite do) cv2 4.13's multi-image knnMatch returns wrong neighbours
    here -- a library clash that has nothing to do with either side of this comparison.  cv2 4.13's multi-image knnMatch
    is also not repeatable on its own: about every other identical call reports, for one or two of the 200 queries,
    the right distance with an (imgIdx, trainIdx) that is not the row at that distance (measured here, one thread or
    eight).  The script therefore accepts a cv2 answer only if it is self-consistent -- every reported index is the
    row at the reported distance -- and asks again otherwise.  Train images with fewer than k rows are left out of this
    case: cv2 4.13 fills the missing neighbours of such an image from uninitialised memory (query 1 of a 1-row image came
    back as distance 0.0 / row 0 in every call here); the oracle's handling of short images is covered below."""
    pytest.importorskip("cv2")
    import json
    import subprocess
    import sys

    rng = np.random.default_rng(0)
    mats = [_unit_rows(rng, n) for n in (120, 3, 333, 64)]  # every image >= k rows: see the docstring
    q = _unit_rows(rng, 200)
    npz = tmp_path / "case.npz"
    np.savez(npz, q=q, **{f"m{i}": m for i, m in enumerate(mats)})
    out = subprocess.run([sys.executable, "-c", _CV2_SCRIPT, str(npz)], capture_output=True, text=True, check=True).stdout
    ref = json.loads(out.strip().splitlines()[-1])
    if ref is None:
        pytest.skip("cv2's multi-image knnMatch contradicted itself in 20 of 20 calls on this machine")
    got = oracle.db_knn(q, mats, 3)
    for i, row in enumerate(ref):
        assert [m[0] for m in row] == list(got["imgIdx"][i])
        assert [m[1] for m in row] == list(got["trainIdx"][i])
        assert np.allclose([m[2] for m in row], got["distance"][i], rtol=1e-6)


def test_db_knn_ties_follow_add_order(oracle):
    rng = np.random.default_rng(1)
    a, b = _unit_rows(rng, 10), _unit_rows(rng, 10)
    b[3] = a[4]
    m = oracle.db_knn(a[4:5], [a, b], 2)
    assert list(zip(m["imgIdx"][0], m["trainIdx"][0])) == [(0, 4), (1, 3)]


def test_db_knn_with_images_shorter_than_k(oracle):
    """Images with fewer than k rows (the case cv2 4.13 cannot arbitrate): the database is the concatenation of its
    images, so the k nearest rows are those of a stable fp64 argsort over all rows, whichever image they sit in."""
    rng = np.random.default_rng(2)
    mats = [_unit_rows(rng, n) for n in (120, 1, 333, 2, 64)]
    q = _unit_rows(rng, 200)
    got = oracle.db_knn(q, mats, 3)
    allm = np.concatenate(mats).astype(np.float64)
    d = ((q[:, None, :].astype(np.float64) - allm[None]) ** 2).sum(-1)
    idx = np.argsort(d, axis=1, kind="stable")[:, :3]
    off = np.cumsum([0] + [len(m) for m in mats])
    img = np.searchsorted(off, idx, side="right") - 1
    assert np.array_equal(got["imgIdx"], img)
    assert np.array_equal(got["trainIdx"], idx - off[img])
    assert np.allclose(got["distance"], np.sqrt(np.take_along_axis(d, idx, 1)), rtol=1e-6)


# ---- literal transcription of the reference's Hough code (every variable a numpy float32 / python float = double)
def _angle_radians(a):
    return f32((f32(2.0) * math.pi * float(a)) / f32(180.0))


def _ref_booth(q, t, refx, refy):
    key_row, key_col, key_ori, key_scl = f32(q["y"]), f32(q["x"]), _angle_radians(q["angle"]), f32(q["size"])
    diff_row, diff_col = f32(f32(refy) - key_row), f32(f32(refx) - key_col)
    dist_sqr = f32(f32(diff_col * diff_col) + f32(diff_row * diff_row))
    dist_to_booth = f32(np.sqrt(dist_sqr))
    angle_to_booth = f32(np.arctan2(diff_row, diff_col))      # float32 in, float32 libm out
    angle_diff = f32(f32(0) - key_ori)
    scale_diff = f32(f32(10) / key_scl)
    f_row, f_col, f_ori, f_scl = f32(t["y"]), f32(t["x"]), _angle_radians(t["angle"]), f32(t["size"])
    ori_diff = f32(f_ori - key_ori)
    dist = f32(dist_to_booth * f32(f_scl / key_scl))
    angle = f32(angle_to_booth + ori_diff)
    row = f32(f32(np.sin(angle) * dist) + f_row)
    col = f32(f32(np.cos(angle) * dist) + f_col)
    return row, col, f32(f_ori + angle_diff), f32(scale_diff * f_scl)


def _ref_compatible(a, b):
    a_row, a_col, a_ori, a_scl = a
    b_row, b_col, b_ori, b_scl = b
    a_ori, b_ori = _angle_radians(a_ori), _angle_radians(b_ori)
    log_diff = f32(abs(f32(np.log(a_scl)) - f32(np.log(b_scl))))
    dc, dr = f32(b_col - a_col), f32(b_row - a_row)
    dist = f32(np.sqrt(f32(f32(dc * dc) + f32(dr * dr))))
    od = f32(b_ori - a_ori) if b_ori > a_ori else f32(a_ori - b_ori)
    if float(od) > 2 * math.pi - float(od):
        od = f32(2 * math.pi - float(od))
    return bool(dist < f32(f32(30.0) / f32(43.0)) * a_scl and log_diff < f32(np.log(f32(1.5)))
                and od < f32(f32(20.0) * math.pi / f32(180.0)))


def _ref_hough(qk, tk_list, matches, refx, refy):
    m = matches.copy()
    kg = [_ref_booth(qk[r["queryIdx"]], tk_list[r["imgIdx"]][r["trainIdx"]], refx, refy)
          for r in m if r["trainIdx"] >= 0 and r["imgIdx"] >= 0]
    best, best_count = -1, 0
    for i, a in enumerate(kg):
        c = sum(_ref_compatible(a, b) for b in kg)
        if c > best_count:
            best, best_count = i, c
    for j, b in enumerate(kg):
        if best >= 0 and not _ref_compatible(kg[best], b):
            m["imgIdx"][j] = m["queryIdx"][j] = m["trainIdx"][j] = -1
    return best_count, m


def _pose_case(rng, n=120, n_out=40):
    tk = np.zeros(n, KP)
    tk["x"], tk["y"] = rng.uniform(20, 600, n), rng.uniform(20, 440, n)
    tk["size"], tk["angle"] = rng.uniform(9, 40, n), rng.uniform(0, 360, n)
    qk = tk.copy()
    qk["x"] += 9.5
    qk["y"] += 4.25
    qk["x"][-n_out:] = rng.uniform(20, 600, n_out)
    qk["angle"][-n_out:] = rng.uniform(0, 360, n_out)
    from oracle.surf_oracle import DMATCH_DTYPE

    m = np.zeros(n, DMATCH_DTYPE)
    m["queryIdx"] = np.arange(n)
    m["imgIdx"] = (np.arange(n) >= 50).astype(np.int32)
    m["trainIdx"] = np.where(np.arange(n) >= 50, np.arange(n) - 50, np.arange(n))
    m["distance"] = 0.1
    return qk, [tk[:50], tk[50:]], m


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_hough_transform_equals_literal_transcription(oracle, seed):
    qk, tks, m = _pose_case(np.random.default_rng(seed))
    want_count, want = _ref_hough(qk, tks, m, 320, 240)
    # numpy's float32 sin/cos/arctan2/log are this machine's float libm: compare with the libm_float switch on
    got_count, got = oracle.hough_transform(qk, tks, m, 320, 240, libm_float=True)
    assert got_count == want_count
    assert got.tobytes() == want.tobytes()
    kept = got["queryIdx"] >= 0
    assert kept[:80].mean() > 0.9 and kept[80:].mean() < 0.15


def test_correctly_rounded_and_libm_float_models_agree(oracle):
    """The default arithmetic model rounds double results once; this machine's float libm may differ in the last
    ulp of a poll booth, which only matters when a pair sits exactly on a threshold."""
    agree = 0
    for seed in range(20):
        qk, tks, m = _pose_case(np.random.default_rng(100 + seed))
        c0, m0 = oracle.hough_transform(qk, tks, m, 300, 200, libm_float=False)
        c1, m1 = oracle.hough_transform(qk, tks, m, 300, 200, libm_float=True)
        agree += int(c0 == c1 and m0.tobytes() == m1.tobytes())
    assert agree >= 19


def test_hough_edge_cases(oracle):
    from oracle.surf_oracle import DMATCH_DTYPE

    qk, tks, m = _pose_case(np.random.default_rng(5))
    c, out = oracle.hough_transform(qk, tks, m[:0], 1, 1)
    assert c == 0 and len(out) == 0
    one = m[:1].copy()
    c, out = oracle.hough_transform(qk, tks, one, 1, 1)
    assert c == 1 and out.tobytes() == one.tobytes()          # a lone booth votes for itself
    bad = np.zeros(3, DMATCH_DTYPE)
    bad["trainIdx"] = bad["imgIdx"] = -1
    c, out = oracle.hough_transform(qk, tks, bad, 1, 1)
    assert c == 0 and out.tobytes() == bad.tobytes()


def test_mask_centroid(oracle):
    m = np.zeros((40, 60), np.uint8)
    m[10:20, 30:50] = 255
    assert oracle.mask_centroid(m) == (39, 14)                # (int) truncation of 39.5, 14.5
    m[0, 0] = 1
    x, y = oracle.mask_centroid(m)
    assert (x, y) == (int(np.float32(39.5 * 200 / 201)), int(np.float32(14.5 * 200 / 201)))


def test_correspond_pipeline_on_synthetic_descriptors(oracle):
    rng = np.random.default_rng(7)
    qk, tks, _ = _pose_case(rng)
    td = [_unit_rows(rng, 50), _unit_rows(rng, 70)]
    qd = np.concatenate(td) + 0.01 * rng.standard_normal((120, 64)).astype(np.float32)
    bogus = [_unit_rows(rng, 30)]
    out, res = oracle.correspond(qk, qd, tks, td, bogus, 320, 240)
    assert res[0] == len(out) == 120                          # every query has a far better train than bogus match
    assert res[1] >= 70 and res[2] in (0, 1) and res[3] <= res[1]
    # a bogus database that contains the queries themselves vetoes everything (ratio = d/0 = inf or nan>0.95 false)
    out2, res2 = oracle.correspond(qk, qd, tks, td, [np.concatenate(td)], 320, 240)
    assert res2[0] < 5
