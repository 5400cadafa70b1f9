"""CPU suite for row f1: the MHA reader / crop / centroid host code of libsurf_b200.so against
the plain-Python restatement of the reference's read_mha (oracle/mha_oracle.py;
vtkSlicerSurfFeaturesLogic.cxx:444-632, :745-752, :839-860).  No GPU needed: these C-ABI calls are host-only."""
import os

import numpy as np
import pytest

from oracle import mha_oracle as mo
from surffeatures_b200 import SurfError, mha


def _frames(n, h=40, w=56, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (n, h, w), dtype=np.uint8)


def _same(seq, ref):
    got = seq.read()
    assert len(got) == len(ref["images"]) == len(seq)
    for a, b in zip(got, ref["images"]):
        assert np.array_equal(a, b)
    assert np.array_equal(seq.transforms(), ref["transforms"])
    assert seq.names() == ref["names"]
    assert np.array_equal(seq.validity(), ref["validity"])
    assert (seq.info.first_frame, seq.info.last_frame) == (ref["first"], ref["last"])


@pytest.mark.parametrize("first,last", [(-1, -1), (0, 3), (2, 100), (5, 2), (100, 100), (3, 3)])
def test_reader_equals_reference_restatement_all_valid(tmp_path, first, last):
    fr = _frames(9)
    t = np.random.default_rng(1).normal(size=(9, 16)) * 50
    p = str(tmp_path / "seq.mha")
    mha.write_mha(p, fr, t)
    with_ref = mo.read_mha(p, first, last)
    s = mha.MhaSequence(p, first, last)
    try:
        _same(s, with_ref)
        lo, hi = with_ref["first"], with_ref["last"]
        assert np.array_equal(s.read(), fr[lo:hi + 1])
        # the transform is parsed with atof into fp32, first 12 of the 16 numbers
        assert np.allclose(s.transforms(), t[lo:hi + 1, :12].astype(np.float32), rtol=1e-5, atol=1e-6)
    finally:
        s.close()


def test_invalid_frames_shift_later_frames_like_the_reference(tmp_path):
    """readImages_mha neither reads nor skips an INVALID frame inside the range: later frames come from earlier bytes."""
    fr = _frames(8, seed=2)
    status = [True, False, True, True, False, True, True, True]
    p = str(tmp_path / "seq.mha")
    mha.write_mha(p, fr, status=status)
    s = mha.MhaSequence(p)
    try:
        _same(s, mo.read_mha(p))
        got = s.read()
        assert len(got) == 6
        assert np.array_equal(got[0], fr[0]) and np.array_equal(got[1], fr[1])      # frame "2" holds frame 1's bytes
        assert np.array_equal(got[5], fr[5])                                        # two invalid frames: shift by 2
    finally:
        s.close()
    fixed = mha.MhaSequence(p, flags=mha.SKIP_INVALID_BYTES)
    try:
        _same(fixed, mo.read_mha(p, skip_invalid_bytes=True))
        assert np.array_equal(fixed.read(), fr[np.array(status)])
    finally:
        fixed.close()


def test_both_tracker_fields_are_collected(tmp_path):
    """A file carrying Probe... and Ultrasound... transforms yields two entries per frame in the reference's
    vectors (both strstr tests match); the kept ones are then chosen by frame index, as read_mha does."""
    fr = _frames(4, seed=3)
    p = str(tmp_path / "seq.mha")
    mha.write_mha(p, fr, extra_fields=("UltrasoundToTrackerTransform",))
    ref = mo.read_mha(p)
    s = mha.MhaSequence(p)
    try:
        _same(s, ref)
        assert len(ref["names"]) == 4 and "Frame0000_Ultrasound" in ref["names"][1]
    finally:
        s.close()


def test_partial_reads_and_windows(tmp_path):
    fr = _frames(10, seed=4)
    p = str(tmp_path / "seq.mha")
    mha.write_mha(p, fr)
    s = mha.MhaSequence(p, 2, 8)
    try:
        assert np.array_equal(s.read(3, 2), fr[5:7])
        assert s.read(7, 0).shape == (0, 40, 56)
        with pytest.raises(SurfError):
            s.read(6, 2)
    finally:
        s.close()


def test_malformed_files_are_errors(tmp_path):
    p = str(tmp_path / "nodims.mha")
    open(p, "wb").write(b"ObjectType = Image\nElementDataFile = LOCAL\n" + bytes(100))
    with pytest.raises(SurfError):
        mha.MhaSequence(p)
    p2 = str(tmp_path / "nostatus.mha")
    open(p2, "wb").write(b"DimSize = 4 4 2\nElementDataFile = LOCAL\n" + bytes(32))
    with pytest.raises(SurfError):          # the reference indexes transformsValidity[i] for every frame in range
        mha.MhaSequence(p2)
    with pytest.raises(SurfError):
        mha.MhaSequence(str(tmp_path / "missing.mha"))
    # an empty line ends the header scan: fields after it are not seen (std::getline loop `if(str.empty()) break`)
    p3 = str(tmp_path / "blank.mha")
    fr = _frames(2, 4, 4)
    mha.write_mha(p3, fr)
    raw = open(p3, "rb").read().replace(b"Seq_Frame0001_ProbeToTrackerTransform =", b"\nSeq_Frame0001_ProbeToTrackerTransform =")
    open(p3, "wb").write(raw)
    with pytest.raises(SurfError):
        mha.MhaSequence(p3)                 # only one status was seen for two frames


@pytest.mark.parametrize("size", [(640, 480), (820, 616), (512, 512), (641, 479), (33, 17)])
def test_crop_rect_is_fp32_truncation(size):
    w, h = size
    for ratios in (mha.REFERENCE_CROP_RATIOS, (0.0, 1.0, 0.0, 1.0), (0.1, 0.9, 0.3, 0.7), (0.25, 0.75, 0.5, 0.99)):
        assert mha.crop_rect(w, h, ratios) == mo.crop_rect(w, h, ratios)
    assert mha.crop_rect(640, 480, mha.REFERENCE_CROP_RATIOS) == (128, 80, 364, 280)
    with pytest.raises(SurfError):
        mha.crop_rect(w, h, (0.5, 0.5, 0.5, 0.5))      # cropRatiosValid
    with pytest.raises(SurfError):
        mha.crop_rect(w, h, (0.5, 1.6, 0.0, 1.0))      # rectangle leaves the image


def test_mask_centroid_equals_oracle(oracle):
    rng = np.random.default_rng(6)
    for _ in range(5):
        m = (rng.random((280, 364)) > 0.6).astype(np.uint8) * 255
        assert mha.mask_centroid(m) == oracle.mask_centroid(m)
    yy, xx = np.mgrid[0:480, 0:640]
    sector = (((xx - 320) ** 2 + (yy + 40) ** 2) < 470 ** 2).astype(np.uint8)
    view = sector[80:360, 128:492]                      # a strided view, like croppedMask = mask(roi)
    assert mha.mask_centroid(view) == oracle.mask_centroid(np.ascontiguousarray(view))
