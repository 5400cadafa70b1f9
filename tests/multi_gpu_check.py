"""Multi-GPU check, run under torchrun on N GPUs of one box (not collected by pytest):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
Every rank runs its own engine on its shard (NCCL only for the gather / candidate all-gather); rank 0 compares the
sharded results with a single-GPU run of the whole job."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from surffeatures_b200 import SURF, sharded, speckle  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_frames, H, W = 8 * world + 3, 256, 256           # config-4 shape, small: ragged shards
    frames = speckle.speckle_stream(n_frames, H, W, seed=2, fresh=0.0)
    eng = SURF(100, 4, 3, device=local, max_width=W, max_height=H, max_batch=16, max_keypoints=8192)

    def compute(fr):
        if len(fr) == 0:
            return np.zeros(0, eng.detect(np.zeros((0, 0), np.uint8)).dtype), np.zeros((0, 64), np.float32), np.zeros(1, np.int64)
        ks, ds, offs = [], [], [0]
        for i in range(0, len(fr), 16):
            k, d, off = eng.detectAndComputeBatchPacked(fr[i:i + 16])
            ks.append(k.copy()); ds.append(d.copy()); offs.extend((off[1:] + offs[-1]).tolist())
        return np.concatenate(ks), np.concatenate(ds), np.asarray(offs)

    lo, hi = sharded.shard_range(n_frames, rank, world)
    got = sharded.sweep_sharded(frames[lo:hi], compute, dst=0)
    # row-sharded keyframe database (config-5 shape, small)
    rng = np.random.default_rng(3)
    db = np.abs(rng.standard_normal((20000, 64))).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    q = db[rng.integers(0, len(db), 1000)] + 0.05 * rng.standard_normal((1000, 64)).astype(np.float32)
    dlo, dhi = sharded.shard_range(len(db), rank, world)
    m = sharded.match_sharded(q, db[dlo:dhi], 2, lambda a, b, k: eng.match_knn(a, b, k), merge=sharded.engine_merge(eng))
    ok = True
    if rank == 0:
        ref = compute(frames)
        ok &= np.array_equal(got[2], ref[2]) and got[0].tobytes() == ref[0].tobytes() and np.array_equal(got[1], ref[1])
        mref = eng.match_knn(q, db, 2)
        ok &= np.array_equal(m["trainIdx"], mref["trainIdx"]) and np.allclose(m["distance"], mref["distance"], rtol=1e-6)
        print(f"multi_gpu_check world={world}: sweep {n_frames} frames -> {len(got[0])} keypoints gathered in order; "
              f"db match merged; {'OK' if ok else 'MISMATCH'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
