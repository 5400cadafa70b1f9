"""CPU suite: the N>1 plumbing (frame-sharded sweep with gather, row-sharded database match with top-k merge)
over gloo with world_size 2, using the oracle as the stand-in for the per-rank compute."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import surf_oracle as so
    from surffeatures_b200 import sharded, speckle

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p = so.params(200, 3, 3)
    frames = speckle.speckle_stream(5, 96, 128, seed=6, fresh=0.0)   # 5 frames over 2 ranks: 3 + 2

    def compute(fr):
        res = [so.detect_and_compute(f, p) for f in fr]
        k = np.concatenate([r[0] for r in res]) if res else np.zeros(0, so.KEYPOINT_DTYPE)
        d = np.concatenate([r[1] for r in res]) if res else np.zeros((0, 64), np.float32)
        return k, d, np.concatenate([[0], np.cumsum([len(r[0]) for r in res])])

    lo, hi = sharded.shard_range(len(frames), rank, world)
    got = sharded.sweep_sharded(frames[lo:hi], compute, dst=0)
    rng = np.random.default_rng(1)
    db = rng.standard_normal((101, 64)).astype(np.float32)
    db[70] = db[10]                                     # a tie across the two shards
    q = np.concatenate([db[[10, 55, 99]], rng.standard_normal((20, 64)).astype(np.float32)])
    dlo, dhi = sharded.shard_range(len(db), rank, world)
    m = sharded.match_sharded(q, db[dlo:dhi], 3, lambda a, b, k: so.knn_match(a, b, k))
    if rank == 0:
        ref = compute(frames)
        np.savez(os.path.join(out_dir, "r0.npz"), k=got[0], d=got[1], off=got[2], rk=ref[0], rd=ref[1], roff=ref[2],
                 m=m, rm=so.knn_match(q, db, 3))
    else:
        assert got is None
        np.savez(os.path.join(out_dir, "r1.npz"), m=m)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_everything():
    from surffeatures_b200.sharded import shard_range

    for n in (0, 1, 7, 64, 4096):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_merge_topk_numpy_ties_and_padding():
    from surffeatures_b200 import DMATCH_DTYPE
    from surffeatures_b200.sharded import merge_topk_numpy

    c = np.zeros((2, 1, 2), DMATCH_DTYPE)
    c[0, 0] = [(0, 5, 0, 1.0), (0, 9, 0, 2.0)]
    c[1, 0] = [(0, 3, 0, 1.0), (0, -1, -1, np.inf)]
    out = merge_topk_numpy(c, 2)
    assert list(out["trainIdx"][0]) == [3, 5] and list(out["distance"][0]) == [1.0, 1.0]


def test_world_size_2_sweep_gather_and_db_merge(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    z = np.load(tmp_path / "r0.npz")
    assert np.array_equal(z["off"], z["roff"])
    assert z["k"].tobytes() == z["rk"].tobytes()          # frame order preserved, nothing lost
    assert np.array_equal(z["d"], z["rd"])
    assert np.array_equal(z["m"]["trainIdx"], z["rm"]["trainIdx"])   # global indices, tie -> lower global row
    assert np.allclose(z["m"]["distance"], z["rm"]["distance"])
    assert list(z["m"]["trainIdx"][0][:2]) == [10, 70]
    z1 = np.load(tmp_path / "r1.npz")
    assert z1["m"].tobytes() == z["m"].tobytes()           # every rank holds the merged result
