"""GPU parity of the "next" rows f2 / f3: the multi-image train database behind
DescriptorMatcher::add / knnMatch (vtkSlicerSurfFeaturesLogic.cxx:1402-1403, :635-643) and the inter-slice
correspondence step (:1434-1571: k=3 vs train, k=1 vs bogus, bogus filter, Hough vote, per-image votes).
Everything goes through the C ABI (surf_db_*, surf_correspond); the oracle is the checker."""
import numpy as np
import pytest

from surffeatures_b200 import speckle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from surffeatures_b200 import SURF

    e = SURF(400, 4, 2, True, False, device=0, max_width=640, max_height=480, max_batch=8, max_keypoints=8192)
    yield e
    e.close()


def _unit_rows(rng, n, dim):
    t = np.abs(rng.standard_normal((n, dim))).astype(np.float32)
    t /= np.maximum(np.linalg.norm(t, axis=1, keepdims=True), 1e-12)
    return t


@pytest.mark.parametrize("dim,k,sizes", [(64, 3, (300, 0, 517, 41)), (128, 1, (1000, 129)), (64, 2, (5,)),
                                         (64, 4, (2500, 2500, 1300))])
def test_db_knn_equals_oracle_and_exact_mode(eng, oracle, dim, k, sizes):
    from surffeatures_b200.engine import DescriptorDB

    rng = np.random.default_rng(sum(sizes) + dim)
    mats = [_unit_rows(rng, n, dim) for n in sizes]
    allt = np.concatenate(mats)
    q = allt[rng.integers(0, len(allt), 700)] + 0.05 * rng.standard_normal((700, dim)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    db = DescriptorDB(eng, dim, max_rows=8192, max_images=16)
    try:
        db.add(mats[:1])          # two add() calls: the second appends at a row that is not a multiple of 8
        db.add(mats[1:])
        assert db.size() == (len(allt), len(mats))
        got = db.knnMatch(q, k)
        want = oracle.db_knn(q, mats, k)
        assert np.array_equal(got["imgIdx"], want["imgIdx"])
        assert np.array_equal(got["trainIdx"], want["trainIdx"])
        fin = want["trainIdx"] >= 0
        assert np.allclose(got["distance"][fin], want["distance"][fin], rtol=2e-6, atol=1e-7)
        eng.set_match_mode(True)
        try:
            exact = db.knnMatch(q, k)
        finally:
            eng.set_match_mode(False)
        assert got.tobytes() == exact.tobytes()     # tensor-core path == exact CUDA-core path, bit for bit
        # clear() empties the matcher like DescriptorMatcher::clear(); re-adding starts at image 0 again
        db.clear()
        assert db.size() == (0, 0)
        db.add(mats[-1:])
        again = db.knnMatch(q[:50], 1)
        assert np.array_equal(again["trainIdx"], oracle.db_knn(q[:50], mats[-1:], 1)["trainIdx"])
        assert (again["imgIdx"][again["trainIdx"] >= 0] == 0).all()
    finally:
        db.close()


def test_db_ties_prefer_lower_image_then_lower_row(eng):
    from surffeatures_b200.engine import DescriptorDB

    rng = np.random.default_rng(11)
    a, b = _unit_rows(rng, 40, 64), _unit_rows(rng, 30, 64)
    b[7] = a[12]
    b[9] = a[12]
    db = DescriptorDB(eng, 64, max_rows=1024, max_images=4)
    try:
        db.add([a, b])
        m = db.knnMatch(a[12:13], 3)
        assert list(zip(m["imgIdx"][0], m["trainIdx"][0])) == [(0, 12), (1, 7), (1, 9)]
    finally:
        db.close()


def test_db_capacity_and_argument_errors(eng):
    from surffeatures_b200 import SurfError
    from surffeatures_b200.engine import DescriptorDB

    rng = np.random.default_rng(1)
    db = DescriptorDB(eng, 64, max_rows=100, max_images=2)
    try:
        db.add([_unit_rows(rng, 60, 64)])
        with pytest.raises(SurfError) as ei:
            db.add([_unit_rows(rng, 60, 64)])
        assert ei.value.code == -3 and db.size() == (60, 1)      # nothing was added
        db.add([_unit_rows(rng, 10, 64)])
        with pytest.raises(SurfError):
            db.add([_unit_rows(rng, 1, 64)])                     # image capacity
        with pytest.raises(SurfError):
            db.knnMatch(_unit_rows(rng, 4, 64), 5)               # k > 4
    finally:
        db.close()


def _sweeps(eng):
    """train / query sweeps of one correlated speckle stream and an unrelated bogus sweep, described on the GPU
    with the reference's effective parameters (detect thr 400 / 4 / 2 then compute, :1384-1389)."""
    stream = speckle.speckle_stream(12, 480, 640, seed=5, fresh=0.0)
    bogus = speckle.speckle_batch(3, 480, 640, 900)
    train = eng.detectAndComputeBatch(stream[0:8])
    query = eng.detectAndComputeBatch(stream[[2, 5, 11]])
    bog = eng.detectAndComputeBatch(bogus)
    return train, query, bog


def test_correspond_equals_oracle(eng, oracle):
    from surffeatures_b200.engine import DescriptorDB, correspond

    train, query, bog = _sweeps(eng)
    D = train[0][1].shape[1]
    tdb = DescriptorDB(eng, D, max_rows=1 << 16, max_images=64)
    bdb = DescriptorDB(eng, D, max_rows=1 << 16, max_images=64)
    try:
        tdb.add([d for _, d in train], [k for k, _ in train])
        bdb.add([d for _, d in bog], [k for k, _ in bog])
        refx, refy = 301, 222
        got, res = correspond(tdb, bdb, [d for _, d in query], [k for k, _ in query], refx, refy, 0.95)
        for i, (qk, qd) in enumerate(query):
            want, wres = oracle.correspond(qk, qd, [k for k, _ in train], [d for _, d in train], [d for _, d in bog],
                                           refx, refy, 0.95)
            assert tuple(res[i]) == tuple(wres), i
            for f in ("queryIdx", "trainIdx", "imgIdx"):
                assert np.array_equal(got[i][f], want[f]), (i, f)
            assert np.allclose(got[i]["distance"], want["distance"], rtol=2e-6)
        # query frames 2 and 5 are train frames: their best train image is themselves, with many surviving votes
        assert res["best_image"][0] == 2 and res["best_image"][1] == 5
        assert res["best_count"][0] > 100 and res["hough_count"][0] >= res["best_count"][0]
        # frame 11 is not in the train sweep; its nearest is the last train frame or none
        assert res["n_matches"][2] <= len(query[2][0])
    finally:
        tdb.close()
        bdb.close()


def test_hough_vote_synthetic_pose_cluster(eng, oracle):
    """Matches that agree on one similarity transform survive; scrambled ones are rejected -- and the GPU keeps
    exactly the entries the oracle keeps."""
    from surffeatures_b200 import KEYPOINT_DTYPE
    from surffeatures_b200.engine import DescriptorDB, correspond

    rng = np.random.default_rng(3)
    n = 300
    tk = np.zeros(n, KEYPOINT_DTYPE)
    tk["x"], tk["y"] = rng.uniform(20, 600, n), rng.uniform(20, 440, n)
    tk["size"], tk["angle"] = rng.uniform(9, 40, n), rng.uniform(0, 360, n)
    qk = tk.copy()
    qk["x"] += 12.5                                  # consistent translation for the first 200
    qk["y"] -= 7.25
    qk["x"][200:] = rng.uniform(20, 600, 100)        # outliers
    qk["angle"][200:] = rng.uniform(0, 360, 100)
    td = _unit_rows(rng, n, 64)
    qd = td + 0.01 * rng.standard_normal((n, 64)).astype(np.float32)
    bd = _unit_rows(rng, 50, 64)
    tdb = DescriptorDB(eng, 64, max_rows=4096, max_images=8)
    bdb = DescriptorDB(eng, 64, max_rows=4096, max_images=8)
    try:
        tdb.add([td[:100], td[100:]], [tk[:100], tk[100:]])
        bdb.add([bd], [np.zeros(50, KEYPOINT_DTYPE)])
        got, res = correspond(tdb, bdb, [qd], [qk], 320, 240)
        want, wres = oracle.correspond(qk, qd, [tk[:100], tk[100:]], [td[:100], td[100:]], [bd], 320, 240)
        assert tuple(res[0]) == tuple(wres)
        assert np.array_equal(got[0]["trainIdx"], want["trainIdx"]) and np.array_equal(got[0]["imgIdx"], want["imgIdx"])
        kept = got[0]["queryIdx"] >= 0
        assert kept[:200].mean() > 0.9 and kept[200:].mean() < 0.1
    finally:
        tdb.close()
        bdb.close()


def test_correspond_rejects_empty_databases(eng):
    from surffeatures_b200 import KEYPOINT_DTYPE, SurfError
    from surffeatures_b200.engine import DescriptorDB, correspond

    tdb = DescriptorDB(eng, 64, max_rows=256, max_images=4)
    bdb = DescriptorDB(eng, 64, max_rows=256, max_images=4)
    try:
        with pytest.raises(SurfError):
            correspond(tdb, bdb, [np.zeros((3, 64), np.float32)], [np.zeros(3, KEYPOINT_DTYPE)], 0, 0)
    finally:
        tdb.close()
        bdb.close()


@pytest.mark.parametrize("dim,k", [(64, 2), (128, 3)])
def test_db_many_exact_duplicates_take_the_grouped_redo_path(eng, oracle, dim, k):
    """More tied rows inside one train tile than the candidate lists hold: the tensor-core proof fails for every
    query and the exact redo (groups of queries sharing each train row) must still return the lowest-index rows,
    like BFMatcher."""
    from surffeatures_b200.engine import DescriptorDB

    rng = np.random.default_rng(5 + dim)
    base = _unit_rows(rng, 50, dim)
    rows = np.repeat(base, 100, axis=0)                              # descriptor j = rows 100 j .. 100 j + 99
    mats = [rows[:1234], rows[1234:3000], rows[3000:]]
    q = base[np.arange(400) % 50]
    db = DescriptorDB(eng, dim, max_rows=1 << 14, max_images=32)
    try:
        db.add(mats)
        got = db.knnMatch(q, k)
        assert eng.match_redo_count() == 400                         # every query was redone, in groups
        want = oracle.db_knn(q, mats, k)
        assert np.array_equal(got["imgIdx"], want["imgIdx"]) and np.array_equal(got["trainIdx"], want["trainIdx"])
        assert (got["distance"] == 0).all()
        few = db.knnMatch(q[:7], k)                                  # fewer entries than CTAs: one query per CTA
        assert eng.match_redo_count() == 7 and few.tobytes() == got[:7].tobytes()
    finally:
        db.close()


def test_config5_full_size_database(eng, oracle):
    """BASELINE config 5 at its full size on one GPU: 1 048 576 x 64 unit rows (added as 8 images of 131 072 rows,
    the per-GPU shards of the multi-GPU layout), 4 096 queries = database rows + noise.  Size-independent
    properties: every query's nearest row is its source row, the records are ascending, and a sample of queries
    equals the oracle's exact scan of the whole database."""
    from surffeatures_b200.engine import DescriptorDB

    rng = np.random.default_rng(3)
    n, shard = 1 << 20, 1 << 17
    rows = np.abs(rng.standard_normal((n, 64), dtype=np.float32))
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    src = rng.integers(0, n, 4096)
    q = rows[src] + 0.05 * rng.standard_normal((4096, 64), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    db = DescriptorDB(eng, 64, max_rows=n, max_images=8)
    try:
        db.add([rows[i * shard:(i + 1) * shard] for i in range(8)])
        assert db.size() == (n, 8)
        m = db.knnMatch(q, 2)
        assert eng.match_redo_count() <= 8
        g = m["imgIdx"].astype(np.int64) * shard + m["trainIdx"]
        assert (g[:, 0] == src).all()
        assert (m["distance"][:, 0] <= m["distance"][:, 1]).all()
        good = m["distance"][:, 0] < np.float32(0.8) * m["distance"][:, 1]
        assert good.mean() > 0.99                                   # the ratio test has its true positives
        sel = np.arange(0, 4096, 128)
        want = oracle.knn_match(q[sel], rows, 2)
        assert np.array_equal(g[sel], want["trainIdx"])
        assert np.allclose(m["distance"][sel], want["distance"], rtol=2e-6, atol=1e-7)
    finally:
        db.close()
