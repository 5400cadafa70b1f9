"""Multi-GPU use of the path: one process per GPU over torch.distributed.

Frames are independent (computeNext touches only images[index],
SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:1392-1406), so a sweep is sharded by contiguous
frame blocks with NO collective on the data path; the only exchange steps are
  * sweep_sharded : gather of the variable-length keypoint / descriptor blocks to one rank, in frame
                    order (BASELINE config 4), and
  * match_sharded : all-gather of the per-rank top-k candidate records of a row-sharded keyframe
                    database followed by a merge (BASELINE config 5; ties -> lower global row).
On a B200 both exchange steps are product code behind the C ABI (surf_sweep_sharded, surf_db_knn_sharded in
csrc/surf_comm.cu: NCCL on device buffers; Python mirror SURF.sweep_sharded / DescriptorDB.knnMatchSharded).  This
module restates the same protocol over torch.distributed with `compute` / `match` / `merge` as callables, so that the
host-side logic (shard ranges, frame order, global row indices, tie rule) is covered on the CPU over gloo with
world_size 2; it is not on the GPU path.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .engine import DMATCH_DTYPE, KEYPOINT_DTYPE


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; the first n_items % world ranks get one more."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dev(group) -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def sweep_sharded(frames_local: np.ndarray, compute: Callable, group=None, dst: int = 0):
    """Runs `compute(frames_local) -> (kps[KEYPOINT_DTYPE], desc[N,D] float32, offsets[B+1])` on every rank's
    shard and gathers everything to rank `dst` in frame order.  Returns (kps, desc, offsets) on dst, None elsewhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = _dev(group)
    kps, desc, off = compute(frames_local)
    counts = np.diff(np.asarray(off, np.int64)).astype(np.int64)
    D = desc.shape[1] if desc.ndim == 2 and desc.shape[1] else 64
    # per-frame counts of every rank (variable number of frames per rank): gather sizes first
    meta = torch.tensor([len(counts), int(counts.sum()), D], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = [m.cpu().numpy() for m in metas]
    max_frames = max(int(m[0]) for m in metas)
    cpad = torch.zeros(max_frames, dtype=torch.int64, device=dev)
    cpad[:len(counts)] = torch.from_numpy(counts).to(dev)
    call = [torch.zeros_like(cpad) for _ in range(world)]
    dist.all_gather(call, cpad, group=group)
    # payload: keypoints as 7 x 4-byte words, descriptors as float32 rows
    k_t = torch.from_numpy(np.ascontiguousarray(kps).view(np.int32).reshape(-1, 7).copy()).to(dev)
    d_t = torch.from_numpy(np.ascontiguousarray(desc, np.float32)).to(dev)
    if rank != dst:
        if len(kps):
            dist.send(k_t, dst, group=group)
            dist.send(d_t, dst, group=group)
        return None
    all_k, all_d, all_counts = [], [], []
    for r in range(world):
        nf, nk = int(metas[r][0]), int(metas[r][1])
        all_counts.append(call[r][:nf].cpu().numpy())
        if r == dst:
            all_k.append(k_t)
            all_d.append(d_t)
        elif nk:
            kr = torch.empty((nk, 7), dtype=torch.int32, device=dev)
            dr = torch.empty((nk, int(metas[r][2])), dtype=torch.float32, device=dev)   # the SENDER's descriptor length
            dist.recv(kr, r, group=group)
            dist.recv(dr, r, group=group)
            all_k.append(kr)
            all_d.append(dr)
    kk = torch.cat(all_k).cpu().numpy().view(KEYPOINT_DTYPE).reshape(-1) if all_k else np.zeros(0, KEYPOINT_DTYPE)
    dd = torch.cat(all_d).cpu().numpy() if all_d else np.zeros((0, D), np.float32)
    offsets = np.concatenate([[0], np.cumsum(np.concatenate(all_counts))]).astype(np.int64)
    return kk, dd, offsets


def merge_topk_numpy(cands: np.ndarray, k: int) -> np.ndarray:
    """Reference merge of (n_ranks, nq, k) DMATCH records: ascending distance, ties -> lower trainIdx."""
    n_ranks, nq, _ = cands.shape
    out = np.zeros((nq, k), DMATCH_DTYPE)
    flat = cands.transpose(1, 0, 2).reshape(nq, n_ranks * k)
    for q in range(nq):
        c = flat[q]
        c = c[c["trainIdx"] >= 0]
        order = np.lexsort((c["trainIdx"], c["distance"]))[:k]
        out[q, :len(order)] = c[order]
        out[q, len(order):]["trainIdx"] = -1
        out[q, len(order):]["imgIdx"] = -1
        out[q, len(order):]["distance"] = np.inf
        out[q]["queryIdx"] = q
    return out


def match_sharded(query: np.ndarray, db_shard: np.ndarray, k: int, match: Callable, merge: Optional[Callable] = None,
                  group=None) -> np.ndarray:
    """knnMatch against a database whose rows are sharded over the ranks (rank r owns the r-th contiguous block).
    `match(query, db_shard, k)` returns (nq, k) DMATCH records with LOCAL trainIdx; every rank gets the merged
    result with GLOBAL trainIdx."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = _dev(group)
    rows = torch.tensor([len(db_shard)], dtype=torch.int64, device=dev)
    all_rows = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(all_rows, rows, group=group)
    base = int(sum(int(r.item()) for r in all_rows[:rank]))
    local = match(query, db_shard, k).copy()
    local["trainIdx"] = np.where(local["trainIdx"] >= 0, local["trainIdx"] + base, -1)
    t = torch.from_numpy(local.view(np.int32).reshape(len(query), k * 4).copy()).to(dev)
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)          # 16 B x k x nq per rank: the whole exchange step
    stacked = torch.stack(parts)                    # (world, nq, 4k) int32 words on the exchange device
    if merge is not None:
        return merge(stacked, k)
    return merge_topk_numpy(stacked.cpu().numpy().view(DMATCH_DTYPE).reshape(world, len(query), k), k)


def engine_merge(engine) -> Callable:
    """merge callable running surf_merge_topk (CUDA) on the all-gathered candidate records."""

    def _merge(stacked: torch.Tensor, k: int) -> np.ndarray:
        world, nq = stacked.shape[0], stacked.shape[1]
        src = stacked.contiguous()
        out = torch.empty((nq, k * 4), dtype=torch.int32, device=src.device)
        torch.cuda.current_stream().synchronize()
        engine.merge_topk_device(src.data_ptr(), world, nq, k, out.data_ptr())
        engine.sync()
        return out.cpu().numpy().view(DMATCH_DTYPE).reshape(nq, k)

    return _merge
