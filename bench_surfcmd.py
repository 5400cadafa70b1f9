#!/usr/bin/env python
"""bench_surfcmd.py -- measurement of the "next" rows f1-f3: the reference's headless flow
(SurfFeatures/Logic/surfcmd.cxx:3-41) on synthetic PLUS sequence files:

  sweeps          : bogus / train / query .mha files -> crop -> masked detect (minHessian 400, 4 octaves, 2 layers)
                    -> 128-d describe -> device-resident matcher databases       (readAndComputeFeaturesOnMhaFile + computeNext)
  correspondences : every query frame: knn k=3 vs train DB, k=1 vs bogus DB, bogus-ratio filter, Hough vote,
                    per-image votes                                             (computeNextInterSliceCorrespondence)

  python bench_surfcmd.py [--train N] [--query N] [--bogus N] [--steps K] [--warmup W] [--no-cpu]

One JSON line: frames/s through the sweeps (file bytes on the host page cache -> results on the device), query
frames/s through the correspondences, and the CPU oracle timed on a bounded sample of the same work.
bench.py (the BASELINE metric) is unchanged; this is the extra line for the widened rows."""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
W_, H_ = 640, 480


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train", type=int, default=384)
    ap.add_argument("--query", type=int, default=96)
    ap.add_argument("--bogus", type=int, default=96)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench_surfcmd.py needs a CUDA device: the engine has no CPU path")
    from surffeatures_b200 import SURF, mha, speckle
    from surffeatures_b200.engine import DescriptorDB, correspond

    tmp = tempfile.mkdtemp(prefix="surfcmd_")
    stream = speckle.speckle_stream(args.train + 8, H_, W_, seed=11)   # default decorrelation (fresh = 0.02), as bench.py
    qsel = np.linspace(0, args.train + 7, args.query).astype(int)
    sets = {"train": stream[:args.train], "query": stream[qsel], "bogus": speckle.speckle_stream(args.bogus, H_, W_, seed=977)}
    files = {}
    for name, fr in sets.items():
        files[name] = os.path.join(tmp, name + ".mha")
        mha.write_mha(files[name], fr)
    yy, xx = np.mgrid[0:H_, 0:W_]
    mask = (((xx - W_ / 2) ** 2 + (yy + 40) ** 2 < (H_ * 0.98) ** 2) & (np.abs(xx - W_ / 2) < (yy + 60) * 0.9)).astype(np.uint8) * 255
    ratios = mha.REFERENCE_CROP_RATIOS
    x, y, cw, ch = mha.crop_rect(W_, H_, ratios)
    refx, refy = mha.mask_centroid(mask[y:y + ch, x:x + cw])

    eng = SURF(400, 4, 2, True, False, device=0, max_width=cw, max_height=ch, max_batch=args.batch, max_keypoints=4096)
    D = eng.descriptorSize()
    tdb = DescriptorDB(eng, D, max_rows=1 << 20, max_images=max(args.train, 16))
    bdb = DescriptorDB(eng, D, max_rows=1 << 19, max_images=max(args.bogus, 16))
    seqs = {n: mha.MhaSequence(f) for n, f in files.items()}
    n_frames = sum(len(s) for s in seqs.values())
    state = {}

    def sweeps():
        tdb.clear()
        bdb.clear()
        mha.sweep(eng, seqs["bogus"], ratios, mask, db=bdb, want_host=False)
        mha.sweep(eng, seqs["train"], ratios, mask, db=tdb, want_host=False)
        state["query"] = mha.sweep(eng, seqs["query"], ratios, mask)          # query rows come back to the host

    def correspondences():
        k, d, o = state["query"]
        qd = [d[o[i]:o[i + 1]] for i in range(len(o) - 1)]
        qk = [k[o[i]:o[i + 1]] for i in range(len(o) - 1)]
        state["res"] = correspond(tdb, bdb, qd, qk, refx, refy)

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / args.steps

    t_sweep = timed(sweeps)
    t_corr = timed(correspondences)
    rows, imgs = tdb.size()
    brows, _ = bdb.size()
    nq_rows = int(state["query"][2][-1])
    res = state["res"][1]
    line = {
        "metric": "surfcmd flow (rows f1-f3): sweep frames/s (MHA file -> crop -> masked detect+describe -> device DB) and "
                  "correspondence query-frames/s (k=3 train + k=1 bogus + filter + Hough vote)",
        "value": n_frames / t_sweep, "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_sweep, "higher_is_better": True, "data": "synthetic", "dtype": "u8/int32/f32",
        "config": {"workload": f"surfcmd flow: {args.bogus}+{args.train}+{args.query} frames {W_}x{H_} in 3 .mha files (page cache), "
                               f"crop {cw}x{ch} at ({x},{y}), sector mask, minHessian 400 / 4 octaves / 2 layers / 128-d, "
                               f"batches of {args.batch}",
                   "timing": "host wall clock around synchronous C-ABI calls (file reads included)"},
        "sweep": {"frames": n_frames, "frames_per_s": n_frames / t_sweep, "ms": 1e3 * t_sweep,
                  "train_rows": rows, "train_images": imgs, "bogus_rows": brows,
                  "keypoints_per_frame": rows / max(imgs, 1)},
        "correspond": {"query_frames": len(res), "query_rows": nq_rows, "query_frames_per_s": len(res) / t_corr,
                       "ms": 1e3 * t_corr, "distances_per_s": nq_rows * (rows + brows) / t_corr,
                       "mean_hough_votes": float(res["hough_count"].mean()), "redo": eng.match_redo_count(),
                       "best_image_is_source_frame": float((res["best_image"] == np.minimum(qsel, args.train - 1)).mean())},
    }
    if not args.no_cpu:
        from oracle import mha_oracle as mo
        from oracle import surf_oracle as so

        threads = so.max_threads()
        P = so.params(400, 4, 2, True, False)
        cmask = mo.crop(mask, ratios)
        ns = max(threads, 8)
        t0 = time.perf_counter()
        imgs_cpu = mo.read_mha(files["train"], 0, ns - 1)["images"]
        crops = np.stack([mo.crop(i, ratios) for i in imgs_cpu])
        # frame-parallel over all host threads; the batch entry has no mask argument, so the masked detect runs per frame
        feats = [so.detect_and_compute(c, P, mask=cmask) for c in crops[:2]]
        t_one = (time.perf_counter() - t0) / 2
        t0 = time.perf_counter()
        so.detect_and_compute_batch(crops, so.params(400, 4, 2, True, False), cap=4096, n_threads=threads)
        t_batch = time.perf_counter() - t0
        # correspondence of 2 query frames against the GPU-built train set sizes (oracle knn on all host threads)
        k, d, o = state["query"]
        tk, td, to = mha.sweep(eng, seqs["train"], ratios, mask)
        bk, bd, bo = mha.sweep(eng, seqs["bogus"], ratios, mask)
        tl_k = [tk[to[i]:to[i + 1]] for i in range(len(to) - 1)]
        tl_d = [td[to[i]:to[i + 1]] for i in range(len(to) - 1)]
        bl_d = [bd[bo[i]:bo[i + 1]] for i in range(len(bo) - 1)]
        t0 = time.perf_counter()
        nc = 2
        for i in range(nc):
            want, wres = so.correspond(k[o[i]:o[i + 1]], d[o[i]:o[i + 1]], tl_k, tl_d, bl_d, refx, refy, n_threads=threads)
            assert tuple(wres) == tuple(res[i]), (i, wres, res[i])       # the timed GPU result is also the right one
        t_c = (time.perf_counter() - t0) / nc
        line["cpu_baseline"] = {"kind": "port", "cores": threads, "unit": "frames/s",
                                "sweep_frames_per_s": ns / t_batch, "sweep_single_thread_masked_frames_per_s": 1 / t_one,
                                "correspond_query_frames_per_s": 1 / t_c,
                                "sample": f"{ns} train frames (unmasked batch entry, {threads} threads) + 2 masked frames single-thread; "
                                          f"{nc} query frames of correspondences against the full databases"}
    print(json.dumps(line), flush=True)
    for s in seqs.values():
        s.close()
    tdb.close()
    bdb.close()
    os._exit(0)


if __name__ == "__main__":
    main()
