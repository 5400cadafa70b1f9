#!/usr/bin/env python
"""Generates tests/golden/golden.json + golden_surf.npz from the CPU oracle (small frames only).

The reference holds no golden vectors and genuine OpenCV 2.4.5 nonfree cannot be built offline, so
these fixtures pin (a) the synthetic speckle generator (SHA-256 per frame) and (b) the oracle's own
outputs, so that any later change to either is caught.  GPU parity tests compare the CUDA path
against these fixtures as well.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import surf_oracle as so  # noqa: E402
from surffeatures_b200 import speckle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FRAMES = {
    "speckle_160x120_default": dict(kind="frame", h=120, w=160, seed=0,
                                    params=dict(hessianThreshold=100, nOctaves=4, nOctaveLayers=3, extended=False, upright=False)),
    "blobs_192x144_ext": dict(kind="frame", h=144, w=192, seed=5, blobs=4,
                              params=dict(hessianThreshold=400, nOctaves=4, nOctaveLayers=2, extended=True, upright=False)),
    "stream_160x120_upright": dict(kind="stream", n=3, index=2, h=120, w=160, seed=2,
                                   params=dict(hessianThreshold=100, nOctaves=3, nOctaveLayers=3, extended=False, upright=True)),
}


def frame(spec):
    if spec["kind"] == "frame":
        return speckle.speckle_frame(spec["h"], spec["w"], spec["seed"], blobs=spec.get("blobs", 0))
    return speckle.speckle_stream(spec["n"], spec["h"], spec["w"], seed=spec["seed"])[spec["index"]]


def main():
    arrays, meta = {}, {"frames": {}}
    for name, spec in FRAMES.items():
        img = frame(spec)
        k, d = so.detect_and_compute(img, so.params(**spec["params"]))
        spec = dict(spec)
        spec["sha256"] = speckle.sha256(img)
        spec["integral_checksum"] = int(so.integral(img).astype(np.int64).sum())
        spec["n_keypoints"] = int(len(k))
        meta["frames"][name] = spec
        arrays[name + "_kp"] = k
        arrays[name + "_desc"] = d
        print(name, len(k), spec["sha256"][:12])
    # config-1 frame (640x480, seed 0): too large to commit its descriptors; pin hash and counts
    img = speckle.speckle_frame(480, 640, 0)
    k, d = so.detect_and_compute(img, so.params(100, 4, 3, False, False))
    meta["config1"] = dict(sha256=speckle.sha256(img), n_keypoints=int(len(k)),
                           integral_checksum=int(so.integral(img).astype(np.int64).sum()),
                           desc_checksum=float(np.float64(d.astype(np.float64).sum())),
                           octave_hist=np.bincount(k["octave"], minlength=4).tolist())
    json.dump(meta, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "golden_surf.npz"), **arrays)


if __name__ == "__main__":
    main()
