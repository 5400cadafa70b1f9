"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Tolerances are BASELINE.json's: integral bit-exact; Hessian responses within 1e-5 relative; >= 99 % of
keypoints paired within 0.5 px and 2 % scale; descriptors within 1e-4 L2.  Most stages are in fact
checked for exact equality, because the kernels reproduce the reference's operation order.
"""
import numpy as np
import pytest

from surffeatures_b200 import speckle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from surffeatures_b200 import SURF

    e = SURF(100, 4, 3, False, False, max_width=1024, max_height=1024, max_batch=8, max_keypoints=32768)
    yield e
    e.close()


def _pair_stats(a, b):
    """Fraction of keypoints of `b` (oracle) with a partner in `a` within 0.5 px and 2 % scale."""
    if len(b) == 0:
        return 1.0 if len(a) == 0 else 0.0
    if len(a) == 0:
        return 0.0
    ok = 0
    ax, ay, asz = a["x"], a["y"], a["size"]
    for k in b:
        d2 = (ax - k["x"]) ** 2 + (ay - k["y"]) ** 2
        near = (d2 <= 0.25) & (np.abs(asz - k["size"]) <= 0.02 * k["size"])
        ok += bool(near.any())
    return ok / len(b)


# ---------------------------------------------------------------- stage 1
@pytest.mark.parametrize("shape", [(480, 640), (17, 33), (1, 1), (479, 641), (1024, 1024), (100, 4), (33, 129)])
def test_integral_bit_exact(eng, oracle, shape):
    rng = np.random.default_rng(shape[0] * 7919 + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(eng.integral(img), oracle.integral(img))


def test_integral_saturated_and_mask(eng, oracle):
    img = np.full((1024, 1024), 255, np.uint8)
    got = eng.integral(img)
    assert got[-1, -1] == 255 * 1024 * 1024
    assert np.array_equal(got, oracle.integral(img))
    m = np.random.default_rng(3).integers(0, 3, (480, 640), dtype=np.uint8) * 100
    assert np.array_equal(eng.integral(m, clamp01=True), oracle.integral(m, clamp01=True))


def test_integral_strided_input(eng, oracle):
    big = np.random.default_rng(5).integers(0, 256, (300, 700), dtype=np.uint8)
    view = big[10:250, 3:643]  # pitch 700, unaligned start
    assert np.array_equal(eng.integral(view), oracle.integral(np.ascontiguousarray(view)))


@pytest.mark.parametrize("shape", [(70, 1400), (37, 2049), (9, 4100), (3, 1299), (5, 7224)])
def test_integral_wide_frames_take_the_short_band_path(oracle, shape):
    """Frames wider than ~1 100 px do not fit 32-row bands in shared memory: the band height shrinks (R = 28 ... 4), the
    row groups are no longer eight rows and the generic vertical-sum loop runs; many bands per frame also exercise the
    flag / carry step of k_integral_band.  Batched, too: the bands of several frames are in flight together."""
    from surffeatures_b200 import SURF

    H, W = shape
    rng = np.random.default_rng(H * 31 + W)
    e = SURF(100, 1, 2, max_width=W, max_height=H, max_batch=3, max_keypoints=256)
    try:
        for rep in range(3):   # epoch flags: nothing is cleared between launches
            img = rng.integers(0, 256, shape, dtype=np.uint8)
            assert np.array_equal(e.integral(img), oracle.integral(img))
        m = rng.integers(0, 2, shape, dtype=np.uint8) * 255
        assert np.array_equal(e.integral(m, clamp01=True), oracle.integral(m, clamp01=True))
    finally:
        e.close()


# ---------------------------------------------------------------- stage 2
def test_hessian_layers(eng, oracle, frame640):
    S = oracle.integral(frame640)
    worst = 0.0
    exact = total = 0
    for o in range(4):
        for l in range(5):
            size, step = (9 + 6 * l) << o, 1 << o
            want, _, _ = oracle.hessian_layer(S, size, step)
            got = eng.hessian_layer(frame640, o, l)
            assert got.shape == want.shape
            scale = np.maximum(np.abs(want), 1.0)
            worst = max(worst, float((np.abs(got - want) / scale).max()))
            exact += int((got == want).sum())
            total += want.size
    assert worst <= 1e-5, worst
    assert exact == total, f"{total - exact} of {total} Hessian samples not bit-identical"


def test_hessian_skips_oversize_layers(eng, oracle):
    img = speckle.speckle_frame(240, 320, 11)
    S = oracle.integral(img)
    want, _, ok = oracle.hessian_layer(S, 264, 8)
    assert not ok
    got = eng.hessian_layer(img, 3, 4)
    assert np.array_equal(got, want) and not got.any()


# ---------------------------------------------------------------- stage 3
@pytest.mark.parametrize("thr,layers", [(100, 3), (400, 2), (1000, 3)])
def test_raw_keypoints_identical(eng, oracle, frame640, thr, layers):
    eng.hessianThreshold, eng.nOctaveLayers = thr, layers
    try:
        got = eng.raw_keypoints(frame640)
    finally:
        eng.hessianThreshold, eng.nOctaveLayers = 100, 3
    want = oracle.raw_keypoints(frame640, oracle.params(thr, 4, layers))
    assert len(got) == len(want) > 500
    for f in ("x", "y", "size", "response", "octave", "class_id", "angle"):
        assert np.array_equal(got[f], want[f]), f


def test_raw_keypoints_blobs_all_octaves(eng, oracle):
    img = speckle.speckle_frame(480, 640, 21, blobs=8)
    got = eng.raw_keypoints(img)
    want = oracle.raw_keypoints(img, oracle.params(100, 4, 3))
    assert set(np.unique(want["octave"])) >= {0, 1, 2}
    assert got.tobytes() == want.tobytes()


def test_mask_semantics(eng, oracle, frame640):
    mask = np.zeros_like(frame640)
    mask[100:400, 150:500] = 255
    got = eng.raw_keypoints(frame640, mask)
    want = oracle.raw_keypoints(frame640, oracle.params(100, 4, 3), mask=mask)
    assert 0 < len(want) < len(oracle.raw_keypoints(frame640, oracle.params(100, 4, 3)))
    assert got.tobytes() == want.tobytes()


# ---------------------------------------------------------------- stage 4
def test_orientation_patches_descriptors_given_keypoints(eng, oracle, frame640):
    p = oracle.params(100, 4, 3)
    raw = oracle.raw_keypoints(frame640, p)
    okp, odesc, opatch = oracle.describe(frame640, p, raw, want_patches=True)
    gkp, gdesc, gpatch = eng.describe_keypoints(frame640, raw, want_patches=True)
    assert np.array_equal(gkp["size"], okp["size"])
    live = okp["size"] > 0
    assert np.array_equal(gkp["angle"][live], okp["angle"][live])
    same_patch = (gpatch[live] == opatch[live]).all(axis=(1, 2))
    assert same_patch.mean() >= 0.999, f"{(~same_patch).sum()} patches differ"
    err = np.linalg.norm(gdesc[live] - odesc[live], axis=1)
    assert err.max() <= 1e-4, float(err.max())
    assert (err == 0).mean() >= 0.99


def test_axis_aligned_directions_take_the_fp64_tap_path(eng, oracle):
    """Keypoints whose direction is within 0.11 degrees of an axis have a sine or cosine that is not a multiple of
    2^-32: their windows are tapped with fp64 coordinates instead of the 32.32 fixed-point ones.  Collected from several
    frames (about 0.3 % of the keypoints), both small and large windows; patches must be identical, byte for byte."""
    p = oracle.params(100, 4, 3)
    picked = 0
    for seed in (1, 2, 3, 4):
        img = speckle.speckle_frame(480, 640, 40 + seed, blobs=6)
        raw = oracle.raw_keypoints(img, p)
        okp, odesc, opatch = oracle.describe(img, p, raw, want_patches=True)
        th = np.deg2rad(okp["angle"].astype(np.float64))
        near_axis = (okp["size"] > 0) & (np.minimum(np.abs(np.cos(th)), np.abs(np.sin(th))) < 2.0 ** -9)
        if not near_axis.any():
            continue
        sel = raw[near_axis]
        gkp, gdesc, gpatch = eng.describe_keypoints(img, sel, want_patches=True)
        assert np.array_equal(gkp["angle"], okp["angle"][near_axis])
        assert np.array_equal(gpatch, opatch[near_axis])
        assert np.abs(gdesc - odesc[near_axis]).max() <= 1e-6
        picked += int(near_axis.sum())
    assert picked >= 8, picked


def test_big_keypoints_unstaged_window_path(eng, oracle):
    img = speckle.speckle_frame(480, 640, 21, blobs=8)
    p = oracle.params(100, 4, 3)
    raw = oracle.raw_keypoints(img, p)
    big = raw[raw["size"] > 56]  # window side > 128: computed on demand, not staged
    assert len(big) > 0
    okp, odesc, opatch = oracle.describe(img, p, big, want_patches=True)
    gkp, gdesc, gpatch = eng.describe_keypoints(img, big, want_patches=True)
    assert np.array_equal(gkp["angle"], okp["angle"])
    assert np.array_equal(gpatch, opatch)
    assert np.abs(gdesc - odesc).max() <= 1e-6


@pytest.mark.parametrize("upright", [False, True])
def test_big_windows_every_resize_path(oracle, upright):
    """Windows wider than 128 pixels are streamed: sub-windows of taps -> row sums of whole INTER_AREA cells -> column
    accumulators.  Given keypoints whose window sides cover the integer scales (multiples of 21), fractional scales up to
    the largest window (739), cell groups of every width, and centres in the frame's interior, on its border and in its
    corners (clamped taps); patches byte for byte, angles bit for bit."""
    from surffeatures_b200 import KEYPOINT_DTYPE, SURF

    img = speckle.speckle_frame(480, 640, 77, blobs=8)
    sides = [129, 130, 146, 147, 148, 160, 168, 189, 191, 192, 193, 210, 231, 255, 288, 289, 300, 336, 384, 385, 420, 441,
             500, 577, 640, 700, 739]
    centres = [(320.3, 240.6), (30.2, 40.9), (600.7, 441.1), (12.5, 300.0), (333.3, 470.2)]
    kps = np.zeros(len(sides) * len(centres), KEYPOINT_DTYPE)
    i = 0
    for n in sides:
        for (x, y) in centres:
            kps[i] = (x, y, (n + 0.4) * 9.0 / (21 * 1.2), -1.0, 500.0, 2, 1)
            i += 1
    got_sides = (np.float32(21) * (kps["size"] * np.float32(1.2) / np.float32(9.0))).astype(np.int32)
    assert set(sides) <= set(got_sides.tolist())          # the sizes hit the intended window sides, 147 = 7 * 21 etc.
    p = oracle.params(100, 4, 3, False, upright)
    okp, odesc, opatch = oracle.describe(img, p, kps, want_patches=True)
    e = SURF(100, 4, 3, False, upright, max_width=640, max_height=480, max_keypoints=4096)
    try:
        gkp, gdesc, gpatch = e.describe_keypoints(img, kps, want_patches=True)
    finally:
        e.close()
    alive = okp["size"] > 0
    assert alive.sum() >= 0.8 * len(kps)
    assert np.array_equal(gkp["size"] > 0, alive)
    assert np.array_equal(gkp["angle"][alive], okp["angle"][alive])
    assert np.array_equal(gpatch[alive], opatch[alive])
    assert np.abs(gdesc[alive] - odesc[alive]).max() <= 1e-6


@pytest.mark.parametrize("extended,upright", [(False, False), (True, False), (False, True), (True, True)])
def test_detect_and_compute_variants(oracle, frame640, extended, upright):
    from surffeatures_b200 import SURF

    e = SURF(100, 4, 3, extended, upright, max_width=640, max_height=480, max_keypoints=16384)
    try:
        kps, desc = e.detectAndCompute(frame640)
    finally:
        e.close()
    okps, odesc = oracle.detect_and_compute(frame640, oracle.params(100, 4, 3, extended, upright))
    assert desc.shape == odesc.shape == (len(okps), 128 if extended else 64)
    assert _pair_stats(kps, okps) >= 0.99
    assert np.array_equal(kps["x"], okps["x"]) and np.array_equal(kps["y"], okps["y"])
    assert np.array_equal(kps["angle"], okps["angle"])
    err = np.linalg.norm(desc - odesc, axis=1)
    assert err.max() <= 1e-4, float(err.max())
    assert np.allclose(np.linalg.norm(desc, axis=1), 1.0, atol=1e-5)


def test_reference_call_sequence_detect_then_compute(oracle, frame640):
    """The reference's own sequence (vtkSlicerSurfFeaturesLogic.cxx:1384-1389): detect with
    (minHessian, 4, 2) and a mask, then compute with a default-constructed extractor."""
    from surffeatures_b200 import SURF

    mask = np.zeros_like(frame640)
    mask[40:440, 60:600] = 1
    det = SURF(400, 4, 2, max_width=640, max_height=480)
    ext = SURF(100, 4, 2, extended=True, max_width=640, max_height=480)
    try:
        kps = det.detect(frame640, mask)
        kps2, desc = ext.compute(frame640, kps)
    finally:
        det.close()
        ext.close()
    okps = oracle.detect(frame640, oracle.params(400, 4, 2), mask=mask)
    okps2, odesc = oracle.compute(frame640, oracle.params(100, 4, 2, True, False), okps)
    assert kps.tobytes() == okps.tobytes()
    assert kps2.tobytes() == okps2.tobytes()
    assert np.linalg.norm(desc - odesc, axis=1).max() <= 1e-4


def test_batch_equals_single_and_is_ordered(eng, oracle):
    frames = speckle.speckle_stream(5, 240, 320, seed=1)
    res = eng.detectAndComputeBatch(frames)
    assert len(res) == 5
    for b in range(5):
        k1, d1 = eng.detectAndCompute(frames[b])
        assert res[b][0].tobytes() == k1.tobytes()
        assert np.array_equal(res[b][1], d1)
    ok, od = oracle.detect_and_compute(frames[3], oracle.params(100, 4, 3))
    assert np.array_equal(res[3][0]["x"], ok["x"])
    assert np.linalg.norm(res[3][1] - od, axis=1).max() <= 1e-4


def test_full_size_1024_extended_upright(oracle):
    """BASELINE config 3 shape (1024x1024, 128-d, upright) on one frame against the oracle."""
    from surffeatures_b200 import SURF

    img = speckle.speckle_frame(1024, 1024, 100)
    e = SURF(100, 4, 3, True, True, max_width=1024, max_height=1024, max_keypoints=32768)
    try:
        kps, desc = e.detectAndCompute(img)
    finally:
        e.close()
    okps, odesc = oracle.detect_and_compute(img, oracle.params(100, 4, 3, True, True))
    assert len(kps) == len(okps) > 10000
    assert _pair_stats(kps[:2000], okps[:2000]) >= 0.99
    assert (kps["angle"] == 270.0).all()
    assert np.linalg.norm(desc - odesc, axis=1).max() <= 1e-4


# ---------------------------------------------------------------- edge cases
def test_edge_cases(eng, oracle):
    from surffeatures_b200 import SurfError

    # empty image -> empty outputs, not an error
    k, d = eng.detectAndCompute(np.zeros((0, 0), np.uint8))
    assert len(k) == 0 and d.shape == (0, 64)
    # flat image -> no keypoints
    k, d = eng.detectAndCompute(np.full((64, 64), 128, np.uint8))
    assert len(k) == 0 and d.shape == (0, 64)
    # image smaller than the smallest filter
    k, d = eng.detectAndCompute(np.random.default_rng(0).integers(0, 255, (8, 8), dtype=np.uint8))
    assert len(k) == 0
    # compute with no keypoints
    k, d = eng.compute(np.zeros((32, 32), np.uint8), np.zeros(0, k.dtype))
    assert len(k) == 0 and d.shape == (0, 64)
    # frame larger than the workspace is refused
    with pytest.raises(SurfError):
        eng.detectAndCompute(np.zeros((2000, 2000), np.uint8))
    # multi-channel / wrong depth is refused
    with pytest.raises(SurfError):
        eng.detectAndCompute(np.zeros((32, 32, 3), np.uint8))


def test_small_image_keypoints_near_border(eng, oracle):
    img = speckle.speckle_frame(72, 96, 5, blobs=3)
    p = oracle.params(50, 4, 3)
    eng.hessianThreshold = 50
    try:
        k, d = eng.detectAndCompute(img)
    finally:
        eng.hessianThreshold = 100
    ok, od = oracle.detect_and_compute(img, p)
    assert k.tobytes() == ok.tobytes()
    if len(ok):
        assert np.linalg.norm(d - od, axis=1).max() <= 1e-4


def test_provided_keypoints_deletion_and_filter(eng, oracle, frame640):
    p = oracle.params(100, 4, 3)
    kps = oracle.raw_keypoints(frame640, p)[:50].copy()
    kps["size"][3] = 0.0        # dropped by runByKeypointSize
    kps["size"][7] = 4000.0     # Haar wavelet larger than the image: deleted by the invoker
    kps["x"][9], kps["y"][9] = -500.0, -500.0  # no orientation sample in bounds: deleted
    gk, gd = eng.compute(frame640, kps)
    ok, od = oracle.compute(frame640, p, kps)
    assert len(gk) == len(ok) == 47
    assert gk.tobytes() == ok.tobytes()
    assert np.linalg.norm(gd - od, axis=1).max() <= 1e-4


@pytest.mark.parametrize("upright", [False, True])
def test_small_provided_keypoints_enlarging_resize(oracle, frame640, upright):
    """DescriptorExtractor::compute() on keypoints smaller than any the detector emits (size < 7.9: window of 1..20
    pixels, enlarged to 21 x 21 by cv::resize's bilinear fixed-point code).  Sizes sweep every window side 1..20,
    positions include the frame border; the tiniest (Haar side 0, window 0) follow the oracle's reading of upstream
    (angle 0 / deleted)."""
    from surffeatures_b200 import SURF

    p = oracle.params(100, 4, 3, False, upright)
    rng = np.random.default_rng(5)
    base = oracle.raw_keypoints(frame640, oracle.params(100, 4, 3))[:400].copy()
    kps = np.concatenate([base, base])
    n_small = len(kps)
    # window side n = (int)(21 * size * 1.2 / 9): sizes 0.2 .. 7.9 cover n = 0 .. 20 several times over
    kps["size"] = np.linspace(0.2, 7.85, n_small).astype(np.float32)[rng.permutation(n_small)]
    kps["x"][:40] = rng.uniform(-2, 6, 40).astype(np.float32)          # left border
    kps["y"][40:80] = rng.uniform(474, 482, 40).astype(np.float32)     # bottom border
    kps["x"][80:120] = rng.uniform(634, 642, 40).astype(np.float32)    # right border
    eng = SURF(100, 4, 3, False, upright, max_width=640, max_height=480, max_batch=1, max_keypoints=8192)
    try:
        okp, odesc, opatch = oracle.describe(frame640, p, kps, want_patches=True)
        gkp, gdesc, gpatch = eng.describe_keypoints(frame640, kps, want_patches=True)
    finally:
        eng.close()
    assert np.array_equal(gkp["size"], okp["size"])
    live = okp["size"] > 0
    side = (21 * (kps["size"] * np.float32(1.2) / np.float32(9.0))).astype(np.int32)
    assert set(range(1, 21)) <= set(side[live].tolist()), "every enlarging window side must be exercised"
    assert (~live).sum() > 0 and (side[~live] < 1).any()
    assert np.array_equal(gkp["angle"][live], okp["angle"][live])
    same_patch = (gpatch[live] == opatch[live]).all(axis=(1, 2))
    assert same_patch.mean() >= 0.999, f"{(~same_patch).sum()} patches differ"
    err = np.linalg.norm(gdesc[live] - odesc[live], axis=1)
    assert np.nanmax(err) <= 1e-4, float(np.nanmax(err))
    assert np.array_equal(np.isnan(gdesc[live]), np.isnan(odesc[live]))


def test_capacity_overflow_is_reported(frame640):
    from surffeatures_b200 import SURF, SurfError

    e = SURF(100, 4, 3, max_width=640, max_height=480, max_keypoints=128)
    try:
        with pytest.raises(SurfError) as ei:
            e.detectAndCompute(frame640)
        assert ei.value.code == -3 and e.required_capacity() > 128
    finally:
        e.close()


# ---------------------------------------------------------------- stage 5
@pytest.mark.parametrize("nq,nt,dim,k", [(300, 400, 64, 2), (1000, 4243, 64, 3), (257, 129, 128, 2), (5, 3, 64, 4),
                                         (64, 1, 64, 1)])
def test_knn_match_equals_oracle(eng, oracle, nq, nt, dim, k):
    rng = np.random.default_rng(nq + nt)
    t = np.abs(rng.standard_normal((nt, dim))).astype(np.float32)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    q = t[rng.integers(0, nt, nq)] + 0.05 * rng.standard_normal((nq, dim)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    got = eng.match_knn(q, t, k)
    want = oracle.knn_match(q, t, k)
    assert np.array_equal(got["trainIdx"], want["trainIdx"])
    assert np.array_equal(got["queryIdx"], want["queryIdx"])
    fin = want["trainIdx"] >= 0
    assert np.allclose(got["distance"][fin], want["distance"][fin], rtol=2e-6, atol=1e-7)
    assert (got["trainIdx"][~fin] == -1).all()


def test_knn_match_ties_go_to_lower_index(eng):
    rng = np.random.default_rng(9)
    t = rng.standard_normal((200, 64)).astype(np.float32)
    t[150] = t[20]
    t[77] = t[20]
    q = t[[20, 20, 5]].copy()
    m = eng.match_knn(q, t, 3)
    assert list(m["trainIdx"][0]) == [20, 77, 150]
    assert list(m["trainIdx"][2][:1]) == [5]


def test_frame_to_frame_ratio_match(eng, oracle):
    """BASELINE config 2 step: descriptors of frame t matched to frame t-1, ratio test 0.8."""
    frames = speckle.speckle_stream(2, 480, 640, seed=1, fresh=0.0)
    (k0, d0), (k1, d1) = eng.detectAndComputeBatch(frames)
    m, good = eng.match_ratio(d1, d0, 0.8)
    om = oracle.knn_match(d1, d0, 2)
    assert np.array_equal(m["trainIdx"], om["trainIdx"])
    ogood = om["distance"][:, 0] < np.float32(0.8) * om["distance"][:, 1]
    assert (good == ogood).mean() >= 0.999
    # matched keypoints move by about the stream's (0.7, 0.3) px shift
    dx = k1["x"][good] - k0["x"][m["trainIdx"][good, 0]]
    assert good.sum() > 500 and abs(np.median(dx) + 0.7) < 0.2


def test_bfmatcher_add_imgidx(eng, oracle):
    from surffeatures_b200 import BFMatcher

    rng = np.random.default_rng(4)
    mats = [rng.standard_normal((n, 64)).astype(np.float32) for n in (30, 50, 20)]
    q = rng.standard_normal((40, 64)).astype(np.float32)
    bf = BFMatcher(eng)
    bf.add(mats)
    m = bf.knnMatch(q, k=3)
    allt = np.concatenate(mats)
    om = oracle.knn_match(q, allt, 3)
    starts = np.array([0, 30, 80])
    img = np.searchsorted(starts, om["trainIdx"], side="right") - 1
    assert np.array_equal(m["imgIdx"], img)
    assert np.array_equal(m["trainIdx"], om["trainIdx"] - starts[img])
    bf.clear()
    assert bf.knnMatch(q, k=3).shape == (40, 0)


# ---------------------------------------------------------------- committed golden fixtures
def test_gpu_reproduces_golden_vectors():
    import json
    import os

    from surffeatures_b200 import SURF

    gdir = os.path.join(os.path.dirname(__file__), "golden")
    meta = json.load(open(os.path.join(gdir, "golden.json")))
    z = np.load(os.path.join(gdir, "golden_surf.npz"))
    for name, spec in meta["frames"].items():
        if spec["kind"] == "frame":
            img = speckle.speckle_frame(spec["h"], spec["w"], spec["seed"], blobs=spec.get("blobs", 0))
        else:
            img = speckle.speckle_stream(spec["n"], spec["h"], spec["w"], seed=spec["seed"])[spec["index"]]
        assert speckle.sha256(img) == spec["sha256"]
        e = SURF(**spec["params"], max_width=spec["w"], max_height=spec["h"], max_keypoints=4096)
        try:
            k, d = e.detectAndCompute(img)
            assert int(e.integral(img).astype(np.int64).sum()) == spec["integral_checksum"]
        finally:
            e.close()
        gk, gd = z[name + "_kp"], z[name + "_desc"]
        assert len(k) == len(gk) == spec["n_keypoints"], name
        for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(k[f], gk[f]), (name, f)
        assert np.linalg.norm(d - gd, axis=1).max() <= 1e-4, name


def test_config1_frame_counts_match_golden(eng):
    import json
    import os

    meta = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))["config1"]
    img = speckle.speckle_frame(480, 640, 0)
    assert speckle.sha256(img) == meta["sha256"]
    k, d = eng.detectAndCompute(img)
    assert len(k) == meta["n_keypoints"]
    assert np.bincount(k["octave"], minlength=4).tolist() == meta["octave_hist"]
    assert abs(float(d.astype(np.float64).sum()) - meta["desc_checksum"]) < 1e-2


# ---------------------------------------------------------------- stage 5 on the tensor cores
def _unit_rows(rng, n, dim):
    t = np.abs(rng.standard_normal((n, dim))).astype(np.float32)
    return t / np.linalg.norm(t, axis=1, keepdims=True)


@pytest.mark.parametrize("nq,nt,dim,k", [(4243, 4100, 64, 2), (300, 9000, 64, 3), (513, 777, 128, 2), (40, 300, 128, 4)])
def test_tensor_core_matcher_equals_exact_and_rarely_redoes(eng, oracle, nq, nt, dim, k):
    rng = np.random.default_rng(nq * 31 + nt)
    t = _unit_rows(rng, nt, dim)
    q = t[rng.integers(0, nt, nq)] + 0.05 * rng.standard_normal((nq, dim)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    eng.set_match_mode(False)
    got = eng.match_knn(q, t, k)
    redo = eng.match_redo_count()
    eng.set_match_mode(True)
    try:
        exact = eng.match_knn(q, t, k)
    finally:
        eng.set_match_mode(False)
    assert got.tobytes() == exact.tobytes()       # bit-identical records, incl. distances
    want = oracle.knn_match(q, t, k)
    assert np.array_equal(got["trainIdx"], want["trainIdx"])
    # the tensor-core candidates must be good enough to prove almost every query
    # (k <= 2: three-entry key lists behind a quad tournament -- a neighbour that shares its quad of train rows with a
    # better one is redone: 3 / nt of the queries on top of the 0.1-0.2 % whose third neighbour lies within the margin)
    assert redo <= max(4, nq // 100), f"{redo} of {nq} queries fell back to the exact kernel"


def test_tensor_core_matcher_near_ties_fall_back_correctly(eng, oracle):
    """Train rows that differ by 1e-7 cannot be separated by the BF16 split: the margin test must send
    those queries to the exact kernel, and the answer must still be the exact one."""
    rng = np.random.default_rng(77)
    base = _unit_rows(rng, 10, 64)
    # 24 distinct near-copies of every base row inside ONE 256-row train tile (12 per 128-column half):
    # more near-ties than the 8 candidates a half-tile keeps, so the candidate set cannot be proven
    t = np.concatenate([base + np.float32(2e-7) * rng.standard_normal(base.shape).astype(np.float32) for _ in range(24)])
    q = base[np.arange(200) % 10] + np.float32(1e-3) * rng.standard_normal((200, 64)).astype(np.float32)
    got = eng.match_knn(q, t, 2)
    assert eng.match_redo_count() == 200
    eng.set_match_mode(True)
    try:
        exact = eng.match_knn(q, t, 2)
    finally:
        eng.set_match_mode(False)
    assert got.tobytes() == exact.tobytes()
    # against the CPU oracle the near-copies are closer together than fp32 summation-order noise, so
    # only the distances (not which of the 24 copies wins) are comparable
    want = oracle.knn_match(q, t, 2)
    assert np.allclose(got["distance"], want["distance"], rtol=1e-4, atol=1e-7)
    assert np.array_equal(got["trainIdx"] % 10, want["trainIdx"] % 10)


def test_match_prev_stream_equals_pairwise_exact(oracle):
    from surffeatures_b200 import SURF

    frames = speckle.speckle_stream(7, 240, 320, seed=3, fresh=0.0)
    e = SURF(100, 4, 3, max_width=320, max_height=240, max_batch=4, max_keypoints=4096)
    try:
        prev = None
        for lo, hi in ((0, 4), (4, 7)):   # two batches: the second one's first frame matches across the call
            kps, desc, off = e.detectAndComputeBatchPacked(frames[lo:hi])
            m, good = e.match_prev(0.8, n_total=int(off[-1]))
            assert e.match_redo_count() <= max(4, int(off[-1]) // 100)   # key lists + quad tournament: <= 1 % unproven at ~1 100 rows
            for b in range(hi - lo):
                d = desc[off[b]:off[b + 1]]
                mm, gg = m[off[b]:off[b + 1]], good[off[b]:off[b + 1]]
                if prev is None:
                    assert (mm["trainIdx"] == -1).all() and not gg.any()
                else:
                    om = oracle.knn_match(d, prev, 2)
                    assert np.array_equal(mm["trainIdx"], om["trainIdx"]), (lo, b)
                    assert np.array_equal(mm["queryIdx"][:, 0], np.arange(len(d)))
                    og = om["distance"][:, 0] < np.float32(0.8) * om["distance"][:, 1]
                    assert (gg == og).mean() > 0.999
                prev = d
        e.stream_reset()
        kps, desc, off = e.detectAndComputeBatchPacked(frames[:2])
        m, good = e.match_prev(0.8, n_total=int(off[-1]))
        assert (m["trainIdx"][:off[1]] == -1).all()      # forgotten predecessor after the reset
        assert (m["trainIdx"][off[1]:off[2], 0] >= 0).all()
    finally:
        e.close()


def test_async_output_mode_gives_identical_results():
    """Copies on the second stream (overlapping the next batch) must deliver the same bytes."""
    import torch

    from surffeatures_b200 import KEYPOINT_DTYPE, MEM_HOST, SURF

    frames = speckle.speckle_stream(6, 240, 320, seed=9, fresh=0.0)
    e = SURF(100, 4, 3, max_width=320, max_height=240, max_batch=2, max_keypoints=4096)
    try:
        ref = []
        for i in range(3):
            k, d, off = e.detectAndComputeBatchPacked(frames[2 * i:2 * i + 2])
            m, g = e.match_prev(0.8, n_total=int(off[-1]))
            ref.append((k.copy(), d.copy(), off.copy(), m.copy(), g.copy()))
        e.stream_reset()
        e.set_async_outputs(True)
        cap = 2 * 4096
        pinned = torch.from_numpy(frames).pin_memory()
        bufs = [dict(k=torch.zeros((cap, 7), dtype=torch.float32).pin_memory(), d=torch.zeros((cap, 64)).pin_memory(),
                     m=torch.zeros((cap * 2, 4), dtype=torch.int32).pin_memory(),
                     g=torch.zeros((cap,), dtype=torch.uint8).pin_memory()) for _ in range(2)]
        offs = [np.zeros(3, np.int32) for _ in range(3)]
        got = []
        for i in range(3):
            hb = bufs[i & 1]
            e.submit(pinned[2 * i:2 * i + 2].data_ptr(), MEM_HOST, 2, 320, 240, 320, 320 * 240, True)
            e.wait_outputs()
            if i >= 1:   # the previous step's buffers are complete now
                pb, po = bufs[(i - 1) & 1], offs[i - 1]
                n = int(po[-1])
                got.append((pb["k"].numpy()[:n].copy().view(np.uint8).reshape(-1), pb["d"].numpy()[:n].copy(),
                            pb["m"].numpy()[:2 * n].copy(), pb["g"].numpy()[:n].copy()))
            e.collect(hb["k"].data_ptr(), hb["d"].data_ptr(), MEM_HOST, cap, offs[i])
            e.match_prev_to(0.8, hb["m"].data_ptr(), hb["g"].data_ptr(), MEM_HOST, cap)
        e.wait_outputs()
        n = int(offs[2][-1])
        pb = bufs[0]
        got.append((pb["k"].numpy()[:n].copy().view(np.uint8).reshape(-1), pb["d"].numpy()[:n].copy(),
                    pb["m"].numpy()[:2 * n].copy(), pb["g"].numpy()[:n].copy()))
        for i in range(3):
            k, d, off, m, g = ref[i]
            assert np.array_equal(offs[i], off)
            assert got[i][0].tobytes() == k.tobytes()
            assert np.array_equal(got[i][1], d)
            assert got[i][2].tobytes() == m.tobytes()
            assert np.array_equal(got[i][3].astype(bool), g)
    finally:
        e.close()


def test_stream_loop_with_lookahead_async_outputs_and_matching():
    """The loop bench.py times end to end -- submit, look-ahead of the next batch, asynchronous result copies and the
    frame-to-frame matcher -- must deliver the bytes of the plain synchronous calls."""
    import torch

    from surffeatures_b200 import MEM_HOST, SURF

    frames = speckle.speckle_stream(8, 240, 320, seed=29, fresh=0.0)
    e = SURF(100, 4, 3, max_width=320, max_height=240, max_batch=2, max_keypoints=4096)
    try:
        ref = []
        for i in range(4):
            k, d, off = e.detectAndComputeBatchPacked(frames[2 * i:2 * i + 2])
            m, g = e.match_prev(0.8, n_total=int(off[-1]))
            ref.append((k.copy(), d.copy(), off.copy(), m.copy(), g.copy()))
        e.stream_reset()
        e.set_async_outputs(True)
        cap = 2 * 4096
        pinned = torch.from_numpy(frames).pin_memory()
        bufs = [dict(k=torch.zeros((cap, 7), dtype=torch.float32).pin_memory(), d=torch.zeros((cap, 64)).pin_memory(),
                     m=torch.zeros((cap * 2, 4), dtype=torch.int32).pin_memory(),
                     g=torch.zeros((cap,), dtype=torch.uint8).pin_memory()) for _ in range(2)]
        offs = [np.zeros(3, np.int32) for _ in range(4)]
        got = []

        def snapshot(i):
            pb, n = bufs[i & 1], int(offs[i][-1])
            got.append((pb["k"].numpy()[:n].copy().view(np.uint8).reshape(-1), pb["d"].numpy()[:n].copy(),
                        pb["m"].numpy()[:2 * n].copy(), pb["g"].numpy()[:n].copy()))

        for i in range(4):
            hb = bufs[i & 1]
            e.submit(pinned[2 * i:2 * i + 2].data_ptr(), MEM_HOST, 2, 320, 240, 320, 320 * 240, True)
            if i + 1 < 4:
                e.lookahead(pinned[2 * i + 2:2 * i + 4].data_ptr(), MEM_HOST, 2, 320, 240, 320, 320 * 240)
            e.wait_outputs()
            if i >= 1:
                snapshot(i - 1)
            e.collect(hb["k"].data_ptr(), hb["d"].data_ptr(), MEM_HOST, cap, offs[i])
            e.match_prev_to(0.8, hb["m"].data_ptr(), hb["g"].data_ptr(), MEM_HOST, cap)
        e.wait_outputs()
        snapshot(3)
        for i in range(4):
            k, d, off, m, g = ref[i]
            assert np.array_equal(offs[i], off)
            assert got[i][0].tobytes() == k.tobytes()
            assert np.array_equal(got[i][1], d)
            assert got[i][2].tobytes() == m.tobytes()
            assert np.array_equal(got[i][3].astype(bool), g)
    finally:
        e.close()


@pytest.mark.parametrize("extended,upright,shape", [(True, False, (200, 264)), (False, True, (251, 333)), (True, True, (240, 320))])
def test_lookahead_variants_and_pending_lookahead_at_destroy(extended, upright, shape):
    """Look-ahead with the other descriptor variants and odd frame sizes (device frames, no TMA alignment for 333 px
    rows), a batch smaller than max_batch, and an engine destroyed while a look-ahead is still pending."""
    import torch

    from surffeatures_b200 import KEYPOINT_DTYPE, MEM_DEVICE, MEM_HOST, SURF

    H, W = shape
    frames = speckle.speckle_stream(7, H, W, seed=31, fresh=0.0)
    D = 128 if extended else 64
    e = SURF(150, 3, 2, extended, upright, max_width=W, max_height=H, max_batch=3, max_keypoints=4096)
    try:
        batches = [frames[0:3], frames[3:6], frames[6:7]]          # the last batch has one frame
        ref = [e.detectAndComputeBatchPacked(b) for b in batches]
        ref = [(k.copy(), d.copy(), o.copy()) for k, d, o in ref]
        dev = [torch.from_numpy(np.ascontiguousarray(b)).cuda() for b in batches]
        torch.cuda.synchronize()
        cap = 3 * 4096
        k = torch.zeros((cap, 7), dtype=torch.float32).pin_memory()
        d = torch.zeros((cap, D), dtype=torch.float32).pin_memory()
        for i, b in enumerate(batches):
            nb = len(b)
            off = np.zeros(nb + 1, np.int32)
            e.submit(dev[i].data_ptr(), MEM_DEVICE, nb, W, H, W, W * H, True)
            if i + 1 < len(batches):
                e.lookahead(dev[i + 1].data_ptr(), MEM_DEVICE, len(batches[i + 1]), W, H, W, W * H)
            e.collect(k.data_ptr(), d.data_ptr(), MEM_HOST, cap, off)
            rk, rd, ro = ref[i]
            n = int(off[-1])
            assert np.array_equal(off, ro)
            assert k.numpy()[:n].copy().view(KEYPOINT_DTYPE).reshape(-1).tobytes() == rk.tobytes()
            assert np.array_equal(d.numpy()[:n], rd)
        # leave a look-ahead pending and destroy the engine: close() must drain it
        e.submit(dev[0].data_ptr(), MEM_DEVICE, 3, W, H, W, W * H, True)
        e.lookahead(dev[1].data_ptr(), MEM_DEVICE, 3, W, H, W, W * H)
    finally:
        e.close()


def test_device_frames_with_unaligned_pitch_take_the_fallback_path(oracle):
    """Frames handed over as device pointers with a pitch that is not a multiple of 16 cannot use TMA:
    the 4-byte cp.async tile path must give the same bytes."""
    import torch

    from surffeatures_b200 import KEYPOINT_DTYPE, MEM_DEVICE, MEM_HOST, SURF

    frames = speckle.speckle_stream(2, 200, 300, seed=11, fresh=0.0)
    pitch = 308                                         # multiple of 4, not of 16
    buf = torch.zeros((2, 200, pitch), dtype=torch.uint8, device="cuda")
    buf[:, :, :300] = torch.from_numpy(frames).cuda()
    e = SURF(100, 4, 3, max_width=300, max_height=200, max_batch=2, max_keypoints=8192)
    try:
        ref = e.detectAndComputeBatch(frames)
        cap = 2 * 8192
        k = torch.zeros((cap, 7), dtype=torch.float32).pin_memory()
        d = torch.zeros((cap, 64), dtype=torch.float32).pin_memory()
        off = np.zeros(3, np.int32)
        torch.cuda.synchronize()
        e.submit(buf.data_ptr(), MEM_DEVICE, 2, 300, 200, pitch, 200 * pitch, True)
        e.collect(k.data_ptr(), d.data_ptr(), MEM_HOST, cap, off)
        for b in range(2):
            kk = k.numpy()[off[b]:off[b + 1]].copy().view(KEYPOINT_DTYPE).reshape(-1)
            assert kk.tobytes() == ref[b][0].tobytes()
            assert np.array_equal(d.numpy()[off[b]:off[b + 1]], ref[b][1])
        ok, od = oracle.detect_and_compute(frames[1], oracle.params(100, 4, 3))
        assert len(ok) == off[2] - off[1]
    finally:
        e.close()


def test_prefetched_upload_gives_identical_results():
    """surf_prefetch_batch uploads batch i+1 on its own stream while batch i computes; the results must be the
    bytes of the plain path, and a submit of OTHER frames must discard the prefetched ones."""
    import torch

    from surffeatures_b200 import KEYPOINT_DTYPE, MEM_HOST, SURF

    frames = speckle.speckle_stream(8, 240, 320, seed=21, fresh=0.0)
    e = SURF(100, 4, 3, max_width=320, max_height=240, max_batch=2, max_keypoints=4096)
    try:
        ref = [e.detectAndComputeBatchPacked(frames[2 * i:2 * i + 2]) for i in range(4)]
        ref = [(k.copy(), d.copy(), o.copy()) for k, d, o in ref]
        pinned = torch.from_numpy(frames).pin_memory()
        cap = 2 * 4096
        k = torch.zeros((cap, 7), dtype=torch.float32).pin_memory()
        d = torch.zeros((cap, 64), dtype=torch.float32).pin_memory()
        off = np.zeros(3, np.int32)

        def batch_ptr(i):
            return pinned[2 * i:2 * i + 2].data_ptr()

        # order 0, 1, 2 with every next batch prefetched; then prefetch 0 but submit 3 (must discard)
        plan = [(0, 1), (1, 2), (2, 0), (3, None)]
        for cur, nxt in plan:
            e.submit(batch_ptr(cur), MEM_HOST, 2, 320, 240, 320, 320 * 240, True)
            if nxt is not None:
                e.prefetch(batch_ptr(nxt), 2, 320, 240, 320, 320 * 240)
            e.collect(k.data_ptr(), d.data_ptr(), MEM_HOST, cap, off)
            rk, rd, ro = ref[cur]
            n = int(off[-1])
            assert np.array_equal(off, ro)
            assert k.numpy()[:n].copy().view(KEYPOINT_DTYPE).reshape(-1).tobytes() == rk.tobytes()
            assert np.array_equal(d.numpy()[:n], rd)
    finally:
        e.close()


@pytest.mark.parametrize("in_mem_name,use_mask", [("host", False), ("device", False), ("host", True), ("device", True)])
def test_lookahead_detect_gives_identical_results(in_mem_name, use_mask):
    """surf_lookahead_batch runs stages 1-3 of batch i+1 on a second stream and buffer set while batch i is in its
    descriptor stage; results must be the bytes of the plain path, over several batches, and a submit of OTHER frames
    must discard the look-ahead."""
    import torch

    from surffeatures_b200 import KEYPOINT_DTYPE, MEM_DEVICE, MEM_HOST, SURF

    frames = speckle.speckle_stream(10, 240, 320, seed=23, fresh=0.0)
    mask = np.zeros((240, 320), np.uint8)
    mask[30:200, 40:300] = 255
    e = SURF(100, 4, 3, max_width=320, max_height=240, max_batch=2, max_keypoints=4096)
    try:
        ref = [e.detectAndComputeBatchPacked(frames[2 * i:2 * i + 2], mask=mask if use_mask else None) for i in range(5)]
        ref = [(k.copy(), d.copy(), o.copy()) for k, d, o in ref]
        if in_mem_name == "host":
            store, mstore, mem = torch.from_numpy(frames).pin_memory(), torch.from_numpy(mask).pin_memory(), MEM_HOST
        else:
            store, mstore, mem = torch.from_numpy(frames).cuda(), torch.from_numpy(mask).cuda(), MEM_DEVICE
        torch.cuda.synchronize()
        mptr, mpitch = (mstore.data_ptr(), 320) if use_mask else (0, 0)
        cap = 2 * 4096
        k = torch.zeros((cap, 7), dtype=torch.float32).pin_memory()
        d = torch.zeros((cap, 64), dtype=torch.float32).pin_memory()
        off = np.zeros(3, np.int32)

        def ptr(i):
            return store[2 * i:2 * i + 2].data_ptr()

        # 0 -> 1 -> 2 -> 3 each looked ahead; then look ahead at 0 but submit 4 (discard); then 1 without look-ahead
        plan = [(0, 1), (1, 2), (2, 3), (3, 0), (4, None), (1, None)]
        for cur, nxt in plan:
            e.submit(ptr(cur), mem, 2, 320, 240, 320, 320 * 240, True, mptr, mpitch)
            if nxt is not None:
                e.lookahead(ptr(nxt), mem, 2, 320, 240, 320, 320 * 240, mptr, mpitch)
            e.collect(k.data_ptr(), d.data_ptr(), MEM_HOST, cap, off)
            rk, rd, ro = ref[cur]
            n = int(off[-1])
            assert np.array_equal(off, ro), (cur, off, ro)
            assert k.numpy()[:n].copy().view(KEYPOINT_DTYPE).reshape(-1).tobytes() == rk.tobytes()
            assert np.array_equal(d.numpy()[:n], rd)
            t = e.timings()
            assert t["hessian"] > 0 and t["describe"] > 0
        # single-frame calls after a pending look-ahead still work (it is discarded, the scratch is ordered)
        e.submit(ptr(0), mem, 2, 320, 240, 320, 320 * 240, True, mptr, mpitch)
        e.lookahead(ptr(1), mem, 2, 320, 240, 320, 320 * 240, mptr, mpitch)
        e.collect(k.data_ptr(), d.data_ptr(), MEM_HOST, cap, off)
        kk, dd = e.detectAndCompute(frames[5], mask=mask if use_mask else None)
        k5, d5, o5 = ref[2]
        assert kk.tobytes() == k5[o5[1]:o5[2]].tobytes() and np.array_equal(dd, d5[o5[1]:o5[2]])
    finally:
        e.close()


@pytest.mark.parametrize("shape,thr,octaves,layers", [((251, 333), 100, 4, 3), ((333, 251), 200, 3, 1),
                                                      ((480, 640), 300, 4, 4), ((200, 1000), 100, 2, 2),
                                                      ((640, 480), 100, 5, 2)])
def test_odd_sizes_and_layer_counts(oracle, shape, thr, octaves, layers):
    """Frame sizes that are not multiples of the tile / alignment granules and layer counts that take the
    generic (run-time table) TMA kernel or the non-fused kernels."""
    from surffeatures_b200 import SURF

    img = speckle.speckle_frame(shape[0], shape[1], seed=shape[0] + layers, blobs=5)
    e = SURF(thr, octaves, layers, max_width=shape[1], max_height=shape[0], max_keypoints=32768)
    try:
        raw = e.raw_keypoints(img)
        k, d = e.detectAndCompute(img)
    finally:
        e.close()
    p = oracle.params(thr, octaves, layers)
    assert raw.tobytes() == oracle.raw_keypoints(img, p).tobytes()
    ok, od = oracle.detect_and_compute(img, p)
    assert len(k) == len(ok) > 50
    for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(k[f], ok[f]), f
    assert np.linalg.norm(d - od, axis=1).max() <= 1e-4


def test_config3_batched_1024_extended_upright(oracle):
    """BASELINE config 3 shape (1024 x 1024, extended + upright, one batched call), at a batch the test box holds
    in seconds: 8 distinct frames, each submitted twice.  Repeats must be byte-identical (frames are independent),
    two frames are checked against the oracle, every angle is 270 and every descriptor has 128 unit-norm values."""
    from surffeatures_b200 import SURF

    frames = speckle.speckle_batch(8, 1024, 1024, 100)
    batch = np.concatenate([frames, frames])
    e = SURF(100, 4, 3, True, True, max_width=1024, max_height=1024, max_batch=16, max_keypoints=32768)
    try:
        kps, desc, off = e.detectAndComputeBatchPacked(batch)
    finally:
        e.close()
    assert len(off) == 17 and desc.shape[1] == 128
    for b in range(8):
        a0, a1, b0, b1 = off[b], off[b + 1], off[b + 8], off[b + 9]
        assert kps[a0:a1].tobytes() == kps[b0:b1].tobytes() and desc[a0:a1].tobytes() == desc[b0:b1].tobytes()
    assert (kps["angle"] == 270.0).all()
    assert np.allclose(np.linalg.norm(desc, axis=1), 1.0, atol=1e-5)
    assert (off[1:] - off[:-1]).min() > 10000                       # ~17 k keypoints per 1024^2 speckle frame
    P = oracle.params(100, 4, 3, True, True)
    for b in (0, 7):
        ok, od = oracle.detect_and_compute(frames[b], P)
        assert kps[off[b]:off[b + 1]].tobytes() == ok.tobytes()
        assert np.linalg.norm(desc[off[b]:off[b + 1]] - od, axis=1).max() <= 1e-4
