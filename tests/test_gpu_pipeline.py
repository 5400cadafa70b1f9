"""GPU suite: the batch pipeline around the kernels -- two batches in flight, the loop bench.py times (at the benchmarked
shape), the frame-sharded sweep and the row-sharded database match through the C ABI (single GPU and a one-rank NCCL
communicator), timing marks, and the error paths of surf_match_prev / surf_set_params.
Everything is compared byte for byte with the plain synchronous calls, and sampled frames with the CPU oracle."""
import numpy as np
import pytest

from surffeatures_b200 import speckle

pytestmark = pytest.mark.gpu


def _sync_reference(e, frames, B, ratio=0.8):
    """Synchronous path: one blocking batch call + match_prev per batch."""
    ref = []
    e.stream_reset()
    for i in range(0, len(frames), B):
        k, d, off = e.detectAndComputeBatchPacked(frames[i:i + B])
        m, g = e.match_prev(ratio, n_total=int(off[-1]))
        ref.append((k.copy(), d.copy(), off.copy(), m.copy(), g.copy()))
    return ref


def _pipelined(e, frames, B, W, H, cap, lookahead=True, mem_host=True):
    """The loop of bench.py: submit(i+1) [+ lookahead(i+2)] before collect(i); asynchronous result copies; matcher."""
    import torch

    from surffeatures_b200 import MEM_DEVICE, MEM_HOST

    nb = len(frames) // B
    pinned = torch.from_numpy(frames).pin_memory()
    src = pinned if mem_host else pinned.cuda()
    mem = MEM_HOST if mem_host else MEM_DEVICE
    bufs = [dict(k=torch.zeros((cap, 7), dtype=torch.float32).pin_memory(), d=torch.zeros((cap, 64)).pin_memory(),
                 m=torch.zeros((cap * 2, 4), dtype=torch.int32).pin_memory(),
                 g=torch.zeros((cap,), dtype=torch.uint8).pin_memory()) for _ in range(2)]
    offs = [np.zeros(B + 1, np.int32) for _ in range(nb)]
    got = []

    def snapshot(i):
        pb, n = bufs[i & 1], int(offs[i][-1])
        got.append((pb["k"].numpy()[:n].copy().view(np.uint8).reshape(-1), pb["d"].numpy()[:n].copy(),
                    pb["m"].numpy()[:2 * n].copy(), pb["g"].numpy()[:n].copy()))

    def submit(i):
        e.submit(src[i * B:(i + 1) * B].data_ptr(), mem, B, W, H, W, W * H, True)
        if lookahead and i + 1 < nb:
            e.lookahead(src[(i + 1) * B:(i + 2) * B].data_ptr(), mem, B, W, H, W, W * H)

    e.stream_reset()
    e.set_async_outputs(True)
    try:
        submit(0)
        for i in range(nb):
            hb = bufs[i & 1]
            if i + 1 < nb:
                submit(i + 1)
            e.wait_outputs()
            if i >= 1:
                snapshot(i - 1)
            e.collect(hb["k"].data_ptr(), hb["d"].data_ptr(), MEM_HOST, cap, offs[i])
            e.match_prev_to(0.8, hb["m"].data_ptr(), hb["g"].data_ptr(), MEM_HOST, cap)
        e.wait_outputs()
        snapshot(nb - 1)
    finally:
        e.set_async_outputs(False)
    return offs, got


def _assert_same(ref, offs, got):
    assert len(ref) == len(got)
    for i, (k, d, off, m, g) in enumerate(ref):
        assert np.array_equal(offs[i], off), i
        assert got[i][0].tobytes() == k.tobytes(), i
        assert np.array_equal(got[i][1], d), i
        assert got[i][2].tobytes() == m.tobytes(), i
        assert np.array_equal(got[i][3].astype(bool), g), i


@pytest.mark.parametrize("lookahead,mem_host", [(True, True), (False, True), (True, False)])
def test_two_batches_in_flight_equal_the_synchronous_calls(lookahead, mem_host):
    from surffeatures_b200 import SURF

    W, H, B = 320, 240, 3
    frames = speckle.speckle_stream(5 * B, H, W, seed=41, fresh=0.0)
    e = SURF(100, 4, 3, max_width=W, max_height=H, max_batch=B, max_keypoints=4096)
    try:
        ref = _sync_reference(e, frames, B)
        offs, got = _pipelined(e, frames, B, W, H, B * 4096, lookahead, mem_host)
        _assert_same(ref, offs, got)
    finally:
        e.close()


def test_benchmarked_loop_at_the_benchmarked_shape(oracle):
    """BASELINE config 2 exactly as bench.py runs it: batches of 64 frames 640x480, look-ahead, two batches in flight,
    asynchronous outputs, frame-to-frame matcher -- against the synchronous calls byte for byte, and against the CPU
    oracle for sampled frames (keypoints, descriptors, matches)."""
    from surffeatures_b200 import SURF

    W, H, B, CAP = 640, 480, 64, 8192
    frames = speckle.speckle_stream(3 * B, H, W, seed=1)
    e = SURF(100, 4, 3, max_width=W, max_height=H, max_batch=B, max_keypoints=CAP)
    try:
        ref = _sync_reference(e, frames, B)
        offs, got = _pipelined(e, frames, B, W, H, B * CAP)
        _assert_same(ref, offs, got)
    finally:
        e.close()
    p = oracle.params(100, 4, 3, False, False)
    for batch, b in ((0, 5), (1, 0), (2, 63)):           # incl. the first frame of a batch: matched across the call
        k, d, off, m, g = ref[batch]
        f = batch * B + b
        ok, od = oracle.detect_and_compute(frames[f], p)
        mine_k, mine_d = k[off[b]:off[b + 1]], d[off[b]:off[b + 1]]
        assert len(ok) == len(mine_k)
        assert np.array_equal(mine_k["x"], ok["x"]) and np.array_equal(mine_k["y"], ok["y"])
        assert np.array_equal(mine_k["angle"], ok["angle"])
        assert np.quantile(np.linalg.norm(mine_d - od, axis=1), 0.999) <= 1e-4
        if b > 0:
            prev_d = d[off[b - 1]:off[b]]
        else:                                            # last frame of the previous batch
            pk, pdesc, poff = ref[batch - 1][:3]
            prev_d = pdesc[poff[B - 1]:poff[B]]
        om = oracle.knn_match(mine_d, prev_d, 2)
        mm = m[off[b]:off[b + 1]]
        assert np.array_equal(mm["trainIdx"], om["trainIdx"])


def test_third_submit_without_collect_is_refused_and_collect_takes_the_oldest():
    import torch

    from surffeatures_b200 import MEM_DEVICE, SURF, SurfError

    W, H = 160, 120
    frames = speckle.speckle_stream(3, H, W, seed=5, fresh=0.0)
    d = torch.from_numpy(frames).cuda()
    e = SURF(100, 3, 3, max_width=W, max_height=H, max_batch=1, max_keypoints=2048)
    try:
        counts = [len(e.detectAndCompute(f)[0]) for f in frames]
        e.submit(d[0:1].data_ptr(), MEM_DEVICE, 1, W, H, W, W * H, True)
        e.submit(d[1:2].data_ptr(), MEM_DEVICE, 1, W, H, W, W * H, True)
        with pytest.raises(SurfError) as ei:
            e.submit(d[2:3].data_ptr(), MEM_DEVICE, 1, W, H, W, W * H, True)
        assert ei.value.code == -1
        off = np.zeros(2, np.int32)
        e.collect(0, 0, MEM_DEVICE, 4096, off)
        assert off[1] == counts[0]
        e.submit(d[2:3].data_ptr(), MEM_DEVICE, 1, W, H, W, W * H, True)
        e.collect(0, 0, MEM_DEVICE, 4096, off)
        assert off[1] == counts[1]
        e.collect(0, 0, MEM_DEVICE, 4096, off)
        assert off[1] == counts[2]
        off[:] = 0
        e.collect(0, 0, MEM_DEVICE, 4096, off)            # nothing in flight: the last batch again (two-call sizing)
        assert off[1] == counts[2]
        e.integral(frames[0])                              # a stage-level call abandons the batch bookkeeping
        with pytest.raises(SurfError):
            e.collect(0, 0, MEM_DEVICE, 4096, off)
    finally:
        e.close()


def test_marks_and_match_time():
    from surffeatures_b200 import SURF

    frames = speckle.speckle_stream(4, 240, 320, seed=3, fresh=0.0)
    e = SURF(100, 4, 3, max_width=320, max_height=240, max_batch=2, max_keypoints=4096)
    try:
        assert e.match_prev_time() == (0.0, 0)
        e.mark(0)
        for i in range(2):
            k, d, off = e.detectAndComputeBatchPacked(frames[2 * i:2 * i + 2])
            e.match_prev(0.8, n_total=int(off[-1]))
        e.mark(1)
        ms = e.mark_elapsed(0, 1)
        mt, launches = e.match_prev_time()
        assert 0 < mt < ms and launches >= 4
        t = e.timings()
        assert t["describe"] > 0 and t["hessian"] > 0
    finally:
        e.close()


def test_match_prev_capacity_error_changes_nothing_and_second_call_is_refused():
    from surffeatures_b200 import SURF, SurfError

    frames = speckle.speckle_stream(4, 240, 320, seed=13, fresh=0.0)
    e = SURF(100, 4, 3, max_width=320, max_height=240, max_batch=2, max_keypoints=4096)
    try:
        ref = _sync_reference(e, frames, 2)
        e.stream_reset()
        k, d, off = e.detectAndComputeBatchPacked(frames[0:2])
        e.match_prev(0.8, n_total=int(off[-1]))
        k, d, off = e.detectAndComputeBatchPacked(frames[2:4])
        with pytest.raises(SurfError) as ei:
            e.match_prev(0.8, n_total=int(off[-1]) - 1)          # arrays one record too small
        assert ei.value.code == -3 and e.required_capacity() == int(off[-1])
        m, g = e.match_prev(0.8, n_total=int(off[-1]))           # the retry sees the untouched stream state
        assert m.tobytes() == ref[1][3].tobytes() and np.array_equal(g, ref[1][4])
        with pytest.raises(SurfError) as ei:
            e.match_prev(0.8, n_total=int(off[-1]))
        assert ei.value.code == -1
    finally:
        e.close()


def test_set_params_between_batches_forgets_the_predecessor_frame():
    from surffeatures_b200 import SURF

    frames = speckle.speckle_stream(4, 240, 320, seed=17, fresh=0.0)
    e = SURF(100, 4, 3, max_width=320, max_height=240, max_batch=2, max_keypoints=4096)
    try:
        k, d, off = e.detectAndComputeBatchPacked(frames[0:2])
        e.match_prev(0.8, n_total=int(off[-1]))
        e.extended = True                                        # 64 -> 128: the stored predecessor rows are 64-d
        k, d, off = e.detectAndComputeBatchPacked(frames[2:4])
        assert d.shape[1] == 128
        m, g = e.match_prev(0.8, n_total=int(off[-1]))
        assert (m["trainIdx"][:off[1]] == -1).all() and not g[:off[1]].any()   # first frame: no predecessor any more
        assert (m["trainIdx"][off[1]:off[2], 0] >= 0).all()
        e.extended = False
        k2, d2, off2 = e.detectAndComputeBatchPacked(frames[0:2])
        assert d2.shape[1] == 64
    finally:
        e.close()


# ---------------------------------------------------------------- frame-sharded sweep / row-sharded database
def _batchwise(e, frames, B, mask=None, want_desc=True):
    ks, ds, offs = [], [], [0]
    for i in range(0, len(frames), B):
        k, d, off = e.detectAndComputeBatchPacked(frames[i:i + B], mask, want_desc)
        ks.append(k.copy())
        if want_desc:
            ds.append(d.copy())
        offs += [offs[-1] + int(x) for x in off[1:]]
    return np.concatenate(ks), (np.concatenate(ds) if want_desc else None), np.asarray(offs, np.int64)


@pytest.fixture(scope="module")
def sweep_setup():
    from surffeatures_b200 import SURF

    H = W = 512                                           # BASELINE config-4 slice shape
    frames = speckle.speckle_stream(19, H, W, seed=2, fresh=0.0)   # 19 frames at max_batch 8: ragged last batch
    e = SURF(100, 4, 3, max_width=W, max_height=H, max_batch=8, max_keypoints=8192)
    yield e, frames
    e.close()


def test_sweep_on_one_gpu_equals_the_batch_calls(sweep_setup):
    e, frames = sweep_setup
    rk, rd, roff = _batchwise(e, frames, 8)
    k, d, off, st = e.sweep_sharded(None, frames, len(frames))
    assert np.array_equal(off, roff) and k.tobytes() == rk.tobytes() and np.array_equal(d, rd)
    assert st["chunks"] == 3 and st["bytes_received"] == 0 and st["total_ms"] > 0 and st["launches"] > 0
    # detect only, with a mask
    mask = np.zeros((512, 512), np.uint8)
    mask[100:400, 50:300] = 255
    rk, _, roff = _batchwise(e, frames[:9], 8, mask, False)
    k, d, off, _ = e.sweep_sharded(None, frames[:9], 9, mask=mask, want_descriptors=False)
    assert d is None and np.array_equal(off, roff) and k.tobytes() == rk.tobytes()
    # nothing to do
    k, d, off, st = e.sweep_sharded(None, frames[:0], 0)
    assert len(k) == 0 and list(off) == [0]


def test_sweep_capacity_overflow_is_reported(sweep_setup):
    from surffeatures_b200 import SurfError

    e, frames = sweep_setup
    with pytest.raises(SurfError) as ei:
        e.sweep_sharded(None, frames[:4], 4, capacity=1000)
    assert ei.value.code == -3 and e.required_capacity() > 1000
    k, d, off, _ = e.sweep_sharded(None, frames[:4], 4)      # the engine is usable afterwards
    assert len(off) == 5 and off[-1] == len(k)


def test_one_rank_nccl_communicator_runs_the_exchange_code(sweep_setup):
    """ncclCommInitRank with n_ranks = 1: the NCCL calls of both exchange steps (count all-gather, grouped send/recv with
    no peer, record all-gather, broadcast) run on this GPU; results equal the local ones."""
    from surffeatures_b200 import Comm, DescriptorDB

    e, frames = sweep_setup
    comm = Comm(e, Comm.unique_id(), 0, 1)
    try:
        assert comm.info()["world"] == 1 and comm.info()["nccl_version"] > 20000
        rk, rd, roff = _batchwise(e, frames[:11], 8)
        k, d, off, st = e.sweep_sharded(comm, frames[:11], 11)
        assert np.array_equal(off, roff) and k.tobytes() == rk.tobytes() and np.array_equal(d, rd)
        rng = np.random.default_rng(3)
        rows = np.abs(rng.standard_normal((30000, 64))).astype(np.float32)
        rows /= np.linalg.norm(rows, axis=1, keepdims=True)
        rows[20000] = rows[7]
        q = rows[rng.integers(0, len(rows), 500)] + 0.05 * rng.standard_normal((500, 64)).astype(np.float32)
        q[0] = rows[7]
        db = DescriptorDB(e, 64, max_rows=len(rows), max_images=1)
        try:
            db.add([rows])
            for kk in (1, 2, 4):
                ref = e.match_knn(q, rows, kk)
                for root in (-1, 0):
                    m = db.knnMatchSharded(comm, q, kk, query_root=root)
                    assert np.array_equal(m["trainIdx"], ref["trainIdx"]) and np.array_equal(m["distance"], ref["distance"])
                    assert (m["imgIdx"] == 0).all()
            assert list(db.knnMatchSharded(comm, q, 2)["trainIdx"][0]) == [7, 20000]
            t = comm.last_times()
            assert t["total_ms"] > 0 and 0 <= t["nccl_ms"] < t["total_ms"]
            m = db.knnMatchSharded(None, q, 2)                   # comm = NULL: plain database call
            assert np.array_equal(m["trainIdx"], e.match_knn(q, rows, 2)["trainIdx"])
        finally:
            db.close()
    finally:
        comm.close()


def test_results_do_not_depend_on_batch_composition_or_repetition():
    """compute-sanitizer is closed on the GPU pool (profiles/r2/compute_sanitizer_closed.log), so races are hunted the
    other way round: a frame's bytes must not depend on which batch it rides in, on its position in the batch, on what
    else the GPU is doing (look-ahead of another batch, matcher on its own stream), or on how often the call is
    repeated.  Shared-memory races in the fused Hessian tiles, the descriptor CTAs' ring and the atomic work queues
    would show up as a keypoint order or a descriptor bit that differs between runs."""
    import torch

    from surffeatures_b200 import MEM_DEVICE, SURF

    W, H = 320, 240
    frames = np.stack([speckle.speckle_frame(H, W, 50 + i, blobs=4 if i % 3 == 0 else 0) for i in range(12)])
    e = SURF(100, 4, 3, max_width=W, max_height=H, max_batch=12, max_keypoints=4096)
    try:
        single = [e.detectAndCompute(f) for f in frames]
        for rep in range(6):
            order = np.random.default_rng(rep).permutation(12)
            B = (12, 6, 4, 3, 5, 7)[rep]
            for lo in range(0, 12, B):
                idx = order[lo:lo + B]
                d_frames = torch.from_numpy(frames[idx]).cuda()
                nxt = torch.from_numpy(frames[order[:len(idx)]]).cuda()
                e.submit(d_frames.data_ptr(), MEM_DEVICE, len(idx), W, H, W, W * H, True)
                e.lookahead(nxt.data_ptr(), MEM_DEVICE, len(idx), W, H, W, W * H)   # competing work on the look-ahead stream
                cap = len(idx) * 4096
                k = torch.zeros((cap, 7), dtype=torch.float32).pin_memory()
                d = torch.zeros((cap, 64)).pin_memory()
                off = np.zeros(len(idx) + 1, np.int32)
                e.collect(k.data_ptr(), d.data_ptr(), 0, cap, off)
                e.match_prev_device(0.8)
                for j, f in enumerate(idx):
                    sk, sd = single[f]
                    assert off[j + 1] - off[j] == len(sk), (rep, f)
                    assert k.numpy()[off[j]:off[j + 1]].tobytes() == sk.tobytes(), (rep, f)
                    assert np.array_equal(d.numpy()[off[j]:off[j + 1]], sd), (rep, f)
        e.sync()
    finally:
        e.close()


def test_graph_replays_are_byte_identical_to_the_eager_launches():
    """Small batches are served by CUDA-graph replays from the second time a (buffers, size) key comes round
    (surf_set_graphs): same kernels, same arguments, so keypoints, descriptors and stage timings must not change."""
    from surffeatures_b200 import SURF

    W, H = 320, 240
    frames = speckle.speckle_stream(6, H, W, seed=23, fresh=0.0)
    mask = np.zeros((H, W), np.uint8)
    mask[30:200, 40:300] = 255
    e = SURF(100, 4, 3, max_width=W, max_height=H, max_batch=4, max_keypoints=4096)
    try:
        e.set_graphs(False)
        eager = [e.detectAndCompute(f) for f in frames]
        eager_mask = [e.detectAndCompute(f, mask) for f in frames[:2]]
        eager_batch = e.detectAndComputeBatchPacked(frames[:4])
        assert e.graph_replays() == 0
        e.set_graphs(True)
        for rep in range(4):
            for f, (kp, de) in zip(frames, eager):
                k2, d2 = e.detectAndCompute(f)
                assert k2.tobytes() == kp.tobytes() and d2.tobytes() == de.tobytes()
            t = e.timings()
            assert t["total"] > 0 and t["describe"] > 0          # the stage events are recorded by the graph, too
            for f, (kp, de) in zip(frames[:2], eager_mask):
                k2, d2 = e.detectAndCompute(f, mask)
                assert k2.tobytes() == kp.tobytes() and d2.tobytes() == de.tobytes()
            kb, db, ob = e.detectAndComputeBatchPacked(frames[:4])
            assert kb.tobytes() == eager_batch[0].tobytes() and db.tobytes() == eager_batch[1].tobytes()
            assert np.array_equal(ob, eager_batch[2])
        assert e.graph_replays() >= 12, e.graph_replays()         # most of the 4 x 9 submits were replays
        # new parameters drop the graphs: the next calls run (and are captured) with the new thresholds
        e.hessianThreshold = 400.0
        a = e.detectAndCompute(frames[0])
        b = e.detectAndCompute(frames[0])
        c = e.detectAndCompute(frames[0])
        assert len(a[0]) < len(eager[0][0])
        assert a[0].tobytes() == b[0].tobytes() == c[0].tobytes() and a[1].tobytes() == c[1].tobytes()
    finally:
        e.close()
