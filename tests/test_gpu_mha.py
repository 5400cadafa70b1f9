"""GPU parity of row f1: whole MHA files through surf_sweep_mha (crop + mask + batches + database fill) against the
oracle on the frames the reference's reader would produce (vtkSlicerSurfFeaturesLogic.cxx:1337-1406), and the
surfcmd flow computeBogus / computeTrain / computeQuery / correspondences (Logic/surfcmd.cxx:3-41) end to end."""
import numpy as np
import pytest

from oracle import mha_oracle as mo
from surffeatures_b200 import mha, speckle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from surffeatures_b200 import SURF

    # the reference's effective parameters: detect (minHessian=400, 4, 2), compute with the default extractor
    e = SURF(400, 4, 2, True, False, device=0, max_width=640, max_height=480, max_batch=4, max_keypoints=8192)
    yield e
    e.close()


def _sector_mask(h=480, w=640):
    yy, xx = np.mgrid[0:h, 0:w]
    return ((((xx - w / 2) ** 2 + (yy + 40) ** 2) < (h * 0.98) ** 2) & (np.abs(xx - w / 2) < (yy + 60) * 0.9)).astype(np.uint8) * 255


def test_sweep_equals_oracle_on_reference_reader_frames(eng, oracle, tmp_path):
    frames = speckle.speckle_stream(11, 480, 640, seed=21)
    status = [True] * 11
    status[4] = False                       # an INVALID frame: later frames shift, as the reference reads them
    p = str(tmp_path / "train.mha")
    mha.write_mha(p, frames, status=status)
    mask = _sector_mask()
    seq = mha.MhaSequence(p, 1, 9)
    try:
        kps, desc, off = mha.sweep(eng, seq, mha.REFERENCE_CROP_RATIOS, mask)
    finally:
        seq.close()
    ref = mo.read_mha(p, 1, 9)
    assert len(off) - 1 == len(ref["images"]) == 8
    cmask = mo.crop(mask, mha.REFERENCE_CROP_RATIOS)
    P = oracle.params(400, 4, 2, True, False)
    for i, img in enumerate(ref["images"]):
        crop = mo.crop(img, mha.REFERENCE_CROP_RATIOS)
        assert crop.shape == (280, 364)
        okps = oracle.detect(crop, oracle.params(400, 4, 2), mask=cmask)
        okps, odesc = oracle.compute(crop, oracle.params(100, 4, 2, True, False), okps)      # the two-call sequence
        k, d = kps[off[i]:off[i + 1]], desc[off[i]:off[i + 1]]
        assert k.tobytes() == okps.tobytes(), i
        assert np.linalg.norm(d - odesc, axis=1).max() <= 1e-4
        k1, d1 = oracle.detect_and_compute(crop, P, mask=cmask)                               # == one pass
        assert k1.tobytes() == okps.tobytes()
    assert off[-1] > 8 * 100


def test_sweep_without_crop_or_mask_and_partial_batches(eng, oracle, tmp_path):
    frames = speckle.speckle_stream(6, 240, 320, seed=22)
    p = str(tmp_path / "q.mha")
    mha.write_mha(p, frames)
    seq = mha.MhaSequence(p)
    try:
        kps, desc, off = mha.sweep(eng, seq, None, None)           # 6 frames, batches of 4 + 2
    finally:
        seq.close()
    P = oracle.params(400, 4, 2, True, False)
    for i in range(6):
        ok, od = oracle.detect_and_compute(frames[i], P)
        assert kps[off[i]:off[i + 1]].tobytes() == ok.tobytes()
        assert np.linalg.norm(desc[off[i]:off[i + 1]] - od, axis=1).max() <= 1e-4


def test_surfcmd_flow_equals_oracle(eng, oracle, tmp_path):
    """computeBogus, computeTrain, computeQuery (sweeps that fill the matchers on the device), then the inter-slice
    correspondences of every query frame."""
    from surffeatures_b200.engine import DescriptorDB, correspond

    stream = speckle.speckle_stream(10, 480, 640, seed=23, fresh=0.0)
    files = {}
    for name, fr in (("train", stream[:7]), ("query", stream[[1, 4, 9]]), ("bogus", speckle.speckle_batch(2, 480, 640, 700))):
        files[name] = str(tmp_path / f"{name}.mha")
        mha.write_mha(files[name], fr)
    mask = _sector_mask()
    cmask = mo.crop(mask, mha.REFERENCE_CROP_RATIOS)
    refx, refy = mha.mask_centroid(cmask)
    assert (refx, refy) == oracle.mask_centroid(cmask)
    D = eng.descriptorSize()
    tdb, bdb = DescriptorDB(eng, D, 1 << 16, 64), DescriptorDB(eng, D, 1 << 16, 64)
    try:
        host = {}
        for name, db in (("bogus", bdb), ("train", tdb), ("query", None)):
            seq = mha.MhaSequence(files[name])
            try:
                host[name] = mha.sweep(eng, seq, mha.REFERENCE_CROP_RATIOS, mask, db=db)
            finally:
                seq.close()
        per = {n: [(k[o[i]:o[i + 1]], d[o[i]:o[i + 1]]) for i in range(len(o) - 1)] for n, (k, d, o) in host.items()}
        assert tdb.size() == (int(host["train"][2][-1]), 7) and bdb.size()[1] == 2
        got, res = correspond(tdb, bdb, [d for _, d in per["query"]], [k for k, _ in per["query"]], refx, refy)
        for i, (qk, qd) in enumerate(per["query"]):
            want, wres = oracle.correspond(qk, qd, [k for k, _ in per["train"]], [d for _, d in per["train"]],
                                           [d for _, d in per["bogus"]], refx, refy)
            assert tuple(res[i]) == tuple(wres), i
            for f in ("queryIdx", "trainIdx", "imgIdx"):
                assert np.array_equal(got[i][f], want[f]), (i, f)
        assert list(res["best_image"][:2]) == [1, 4]
    finally:
        tdb.close()
        bdb.close()
