#!/usr/bin/env python
"""Small run of the whole path for compute-sanitizer (memcheck / racecheck targets):
  compute-sanitizer --tool memcheck  python tools/sanitize_small.py
  compute-sanitizer --tool racecheck python tools/sanitize_small.py
One 320x240 batch of 2 frames through every kernel family: integral, fused TMA Hessian+maxima (octaves 0-1), k_hessian +
k_nms (octaves 2-3), sort, orientation, the descriptor window kernels (a blob frame makes large windows), patch descriptor,
pack, and the tcgen05 matcher (stream matching and a database query).  Results are checked against the synchronous
single-frame calls so that a silent corruption is an error, not just a sanitizer line."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surffeatures_b200 import SURF, DescriptorDB, speckle  # noqa: E402


def main():
    frames = np.stack([speckle.speckle_frame(240, 320, 7), speckle.speckle_frame(240, 320, 8, blobs=6)])
    e = SURF(100, 4, 3, False, False, max_width=320, max_height=240, max_batch=2, max_keypoints=4096)
    k, d, off = e.detectAndComputeBatchPacked(frames)
    m, g = e.match_prev(0.8, n_total=int(off[-1]))
    for b in range(2):
        k1, d1 = e.detectAndCompute(frames[b])
        assert k1.tobytes() == k[off[b]:off[b + 1]].tobytes() and np.array_equal(d1, d[off[b]:off[b + 1]])
    ex = e.match_knn(d[off[1]:off[2]], d[off[0]:off[1]], 2)
    assert np.array_equal(ex["trainIdx"], m["trainIdx"][off[1]:off[2]])
    db = DescriptorDB(e, 64, max_rows=8192, max_images=4)
    db.add([d[off[0]:off[1]], d[off[1]:off[2]]])
    r = db.knnMatch(d[:200], 3)
    assert (r["trainIdx"][:, 0] == np.arange(200)).all() and (r["imgIdx"][:, 0] == 0).all()
    db.close()
    e.extended, e.upright = True, True
    k2, d2 = e.detectAndCompute(frames[1])
    assert d2.shape[1] == 128 and len(k2) > 0
    e.close()
    print(f"sanitize_small ok: {int(off[-1])} keypoints, {int(g.sum())} ratio-test matches")


if __name__ == "__main__":
    main()
