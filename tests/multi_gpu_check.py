"""Multi-GPU check, run under torchrun on N GPUs of one box (not collected by pytest):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
Every rank runs its own engine on its shard; the exchange steps go through the C ABI (surf_comm_*, surf_sweep_sharded,
surf_db_knn_sharded: NCCL on device buffers).  Rank 0 compares the sharded results with a single-GPU run of the whole
job, byte for byte, and appends one line to gpurun_out/multi_gpu_check.log."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from surffeatures_b200 import SURF, Comm, DescriptorDB, shard_range, speckle  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H = W = 512                                          # config-4 shape
    n_frames = 20 * world + 3                            # ragged shards, several chunks per rank at max_batch 8
    frames = speckle.speckle_stream(n_frames, H, W, seed=2, fresh=0.0)
    eng = SURF(100, 4, 3, device=local, max_width=W, max_height=H, max_batch=8, max_keypoints=8192)
    comm = Comm.from_torch(eng)
    info = comm.info()
    ok = True
    report = {"world": world, "nccl_version": info["nccl_version"]}

    lo, hi = shard_range(n_frames, rank, world)
    for dst in (0, world - 1):
        k, d, off, st = eng.sweep_sharded(comm, frames[lo:hi], n_frames, dst=dst)
        if rank == dst:
            ref_k, ref_d, ref_off = [], [], [0]
            for i in range(0, n_frames, 8):
                rk, rd, roff = eng.detectAndComputeBatchPacked(frames[i:i + 8])
                ref_off += [ref_off[-1] + int(x) for x in roff[1:]]
                ref_k.append(rk.copy()); ref_d.append(rd.copy())
            ref_k, ref_d = np.concatenate(ref_k), np.concatenate(ref_d)
            same = (np.array_equal(off, np.asarray(ref_off, np.int64)) and k.tobytes() == ref_k.tobytes()
                    and np.array_equal(d, ref_d))
            ok &= bool(same)
            report[f"sweep_dst{dst}"] = {"frames": n_frames, "keypoints": int(off[-1]), "identical_to_single_gpu": bool(same),
                                         "bytes_received": st["bytes_received"], "exchange_ms": st["exchange_ms"],
                                         "total_ms": st["total_ms"], "chunks": st["chunks"]}
        else:
            assert k is None
        dist.barrier()

    # row-sharded database (config-5 shape, small), ties across shards, k = 2 and 3, broadcast queries
    rng = np.random.default_rng(3)
    db_rows = np.abs(rng.standard_normal((50001, 64))).astype(np.float32)
    db_rows /= np.linalg.norm(db_rows, axis=1, keepdims=True)
    db_rows[40000] = db_rows[10]                          # an exact tie in different shards
    q = db_rows[rng.integers(0, len(db_rows), 1000)] + 0.05 * rng.standard_normal((1000, 64)).astype(np.float32)
    q[0] = db_rows[10]
    dlo, dhi = shard_range(len(db_rows), rank, world)
    db = DescriptorDB(eng, 64, max_rows=dhi - dlo, max_images=1)
    db.add([db_rows[dlo:dhi]])
    for k_ in (2, 3):
        for root in (-1, 0):
            qq = q if (root < 0 or rank == root) else np.zeros_like(q)
            m = db.knnMatchSharded(comm, qq, k_, query_root=root)
            mref = eng.match_knn(q, db_rows, k_)
            same = np.array_equal(m["trainIdx"], mref["trainIdx"]) and np.array_equal(m["distance"], mref["distance"])
            ok &= bool(same)
            if rank == 0:
                report[f"db_k{k_}_root{root}"] = {"identical_to_single_gpu": bool(same), "tie_row": [int(x) for x in m["trainIdx"][0][:2]],
                                                  **comm.last_times()}
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    good = int(flag.item()) == 1
    if rank == 0:
        report["result"] = "OK" if good else "MISMATCH"
        line = json.dumps(report)
        print("multi_gpu_check " + line, flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "multi_gpu_check.log"), "a") as f:
            f.write(line + "\n")
    db.close()
    comm.close()
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if good else 1


if __name__ == "__main__":
    sys.exit(main())
