#!/usr/bin/env python
"""bench.py -- SURF detect+describe frames/sec @640x480 (BASELINE.json metric).

Headline workload = BASELINE config 2: a stream of 640x480 speckle frames, detect + describe every frame and BF L2
ratio-match it to the previous frame.  One step = one batch of BATCH consecutive frames of the stream.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5] [--no-extras]

value   : whole-job frames/s with the frames already resident in HBM (device pointers in/out)
e2e     : the same through the host-buffer API (pinned host frames -> H2D, packed keypoints,
          descriptors and matches -> D2H inside the timed region)
roofline: dominant stage's algorithmic bytes / its CUDA-event time, against MEASURED_PEAKS.json;
          roofline_stages: the same for every stage (integral, Hessian+maxima, describe, matcher), timed in a pass
          where no other stream runs beside the stage
latency : host-in/host-out milliseconds per call at batch 1, 4, 16 (the reference's one-frame-per-tick loop)
configs : BASELINE configs 3 (1024^2, 128-d upright, 256 frames), 4 (sweep of 512^2 slices sharded by frame,
          results gathered to rank 0 by NCCL) and 5 (1M-row database sharded by row, NCCL all-gather of top-2
          records) measured in the same run; under torchrun 4 and 5 use all ranks (strong scaling)
cpu_baseline: the CPU oracle (restatement of OpenCV 2.4.5 SURF + BFMatcher) on the box's host cores

All device times are CUDA events recorded by the engine on its own stream (surf_mark), max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_, H_ = 640, 480
PARAMS = dict(hessianThreshold=100.0, nOctaves=4, nOctaveLayers=3, extended=False, upright=False)
RATIO = 0.8
CAP = 8192          # raw keypoints per frame the workspace holds
LOOKAHEAD = os.environ.get("SURF_BENCH_LOOKAHEAD", "1") != "0"   # surf_lookahead_batch in the throughput loops
E2E_DIAG = os.environ.get("SURF_BENCH_E2E_DIAG", "")   # "no_d2h" / "no_h2d" / "no_desc": diagnosis runs of the end-to-end loop, never a reported number
POOL_BATCHES = 4    # distinct batches cycled through; the per-step working set (integral + det pyramid of 64 frames, ~600 MB) exceeds the 126 MB L2


def _metric():
    return "SURF detect+describe frames/sec @640x480 (1/2/4/8 GPU); HBM GB/s vs peak"


def workload_config(B: int) -> dict:
    """The `config` object -- identical in both arms (--impl ours / reference)."""
    return {"workload": "config2: 640x480 speckle stream, detect+describe (thr=100, 4 octaves, 3 layers, 64-d) "
                        "+ BF L2 k=2 ratio(0.8) match to previous frame",
            "frames_per_step": B, "frame": "640x480 u8", "ratio": RATIO,
            "l2": f"inputs cycle through {POOL_BATCHES} batches (no L2 flush needed): per-step working set "
                  f"(integral images + det pyramid of {B} frames) > 126 MB L2"}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": float(d["hbm_gbs"]), "bf16_burst": float(d.get("bf16_tflops", 1590.0)),
                "bf16_sustained": float(d.get("bf16_tflops_sustained", 1400.0)), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def host_threads() -> int:
    """Threads this process may use: the affinity mask, not OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  NVML (nvidia_ml_py) is sampled from a thread every
    10 ms -- the timed region of a short run is well under a second, less than nvidia-smi needs to start on an
    8-GPU box; `nvidia-smi -lms` (the B200_PROFILING.md recipe) is the fallback when NVML cannot be imported."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index
        self.sm, self.reasons, self.mx, self.stop_flag, self.thread, self.nvml = [], set(), None, False, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber: resolve through the PCI bus id of the torch device
            import torch

            props = torch.cuda.get_device_properties(gpu_index)
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(
                    f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0".encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{props.uuid}".encode())
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _nvml_loop(self):
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                r = int(get_reasons(self.handle))
                for name, b in bits.items():
                    if r & b:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            if not self.sm:
                return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": ["no NVML sample"], "source": "nvml"}
            return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "nvml"}
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi"}


def make_stream(n_frames: int, seed: int, h: int = H_, w: int = W_) -> np.ndarray:
    from surffeatures_b200 import speckle

    return speckle.speckle_stream(n_frames, h, w, seed=seed)


def make_frames_parallel(n: int, h: int, w: int, seed0: int) -> np.ndarray:
    """n independent speckle frames (seeds seed0...), generated on a few host threads (numpy releases the GIL)."""
    from surffeatures_b200 import speckle

    with ThreadPoolExecutor(max_workers=min(8, host_threads())) as ex:
        return np.stack(list(ex.map(lambda i: speckle.speckle_frame(h, w, seed0 + i), range(n))))


def algorithmic_bytes_per_frame(n_kp: float, D: int = 64, L: int = 3, O: int = 4, W: int = W_, H: int = H_):
    """B1..B4 per frame."""
    A = 4 * (W + 1) * (H + 1)
    P = sum((H >> o) * (W >> o) for o in range(O))
    b1 = W * H + A
    b2 = O * A + 4 * (L + 2) * P
    b3 = 4 * (L + 2) * P + 28 * n_kp
    b4 = W * H + A + n_kp * (28 + 4 + 4 * D)
    return {"integral": b1, "hessian": b2, "nms": b3, "hessian_nms_fused": O * A + 28 * n_kp, "describe": b4,
            "total": b1 + b2 + b3 + b4}


# ------------------------------------------------------------------------------------------------
def _cpu_stream_pass(so, frames, p, threads):
    """One pass of the config-2 workload on the CPU oracle: detect+describe every frame of `frames[1:]`, match each to
    its predecessor.  Returns keypoints seen."""
    res = so.detect_and_compute_batch(frames[1:], p, cap=CAP, n_threads=threads)
    prev = so.detect_and_compute(frames[0], p)[1]
    for _, d in res:
        so.knn_match(d, prev, 2, n_threads=threads)
        prev = d
    return sum(len(k) for k, _ in res)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  Genuine OpenCV 2.4.5 nonfree cannot be
    built offline, so this times the oracle port (all host threads, frame-parallel).  Same `config` as the GPU arm;
    each step is a bounded sample of that step's 64 frames so that the whole run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)   # torchrun exports 1; the oracle also gets num_threads explicitly
    from oracle import surf_oracle as so

    B = args.batch
    p = so.params(**PARAMS)
    probe = make_stream(3, seed=1)
    t0 = time.perf_counter()
    so.detect_and_compute_batch(probe[1:], p, cap=CAP, n_threads=min(threads, 2))
    per_frame_1t = (time.perf_counter() - t0) / 2 * min(threads, 2)       # thread-seconds per frame, roughly
    # bound the run: about 150 s of wall clock over warm-up + timed steps
    n_steps = max(1, args.steps + args.warmup)
    est = per_frame_1t / threads * 1.4
    sample = int(max(min(B, 150.0 / (n_steps * est)), min(B, threads)))
    frames = make_stream(sample + 1, seed=1)
    for _ in range(args.warmup):
        _cpu_stream_pass(so, frames, p, threads)
    t0 = time.perf_counter()
    nk = 0
    for _ in range(args.steps):
        nk += _cpu_stream_pass(so, frames, p, threads)
    dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": _metric(), "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps * (B / sample),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32/f32", "data": "synthetic",
        "config": workload_config(B),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} of the step's {B} frames per timed step x {args.steps} steps, detect+describe+"
                                   f"match, oracle port of OpenCV 2.4.5 SURF (genuine nonfree not installable offline), "
                                   f"frame-parallel on {threads} threads; ms_per_step is scaled to {B} frames"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "keypoints_per_frame": nk / (sample * args.steps),
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
def _parse_cpulist(txt: str):
    cpus = []
    for part in txt.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def bind_rank_to_gpu_numa(torch, local: int, world: int):
    """One process per GPU on a shared host: give every rank a private slice of the cores of its GPU's NUMA node, before
    any pinned buffer is allocated (first touch decides where the pages live).  Returns a note for the JSON line."""
    if world == 1:
        return "single process: affinity left alone"
    try:
        props = torch.cuda.get_device_properties(local)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node_cpus = _parse_cpulist(open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read())
        allowed = sorted(set(node_cpus) & set(os.sched_getaffinity(0))) or sorted(os.sched_getaffinity(0))
        # ranks whose GPUs share this CPU list split it evenly (rank order = local rank order)
        n_local = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        sharing = []
        for r in range(n_local):
            pr = torch.cuda.get_device_properties(r)
            f = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/local_cpulist"
            if _parse_cpulist(open(f).read()) == node_cpus:
                sharing.append(r)
        k, n = sharing.index(local), len(sharing)
        per = max(1, len(allowed) // n)
        mine = allowed[k * per:(k + 1) * per] or allowed
        os.sched_setaffinity(0, mine)
        return f"rank bound to cpus {mine[0]}-{mine[-1]} ({len(mine)} of the {len(allowed)} cores local to GPU {bdf}, shared by {n} ranks)"
    except Exception as ex:   # containers without sysfs PCI topology: run unbound
        return f"unbound ({type(ex).__name__})"


class Ctx:
    """Process-wide state of the GPU arm."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.affinity = bind_rank_to_gpu_numa(torch, self.local, self.world)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


class StreamBench:
    """detect+describe(+match to the previous frame) over batches of B frames cycled from a pool (configs 2 and 3)."""

    def __init__(self, ctx: Ctx, eng, pool: np.ndarray, B: int, D: int, cap: int, match: bool):
        from surffeatures_b200 import MEM_DEVICE, MEM_HOST

        torch = ctx.torch
        self.ctx, self.eng, self.B, self.D, self.cap, self.match = ctx, eng, B, D, cap, match
        self.MEM_DEVICE, self.MEM_HOST = MEM_DEVICE, MEM_HOST
        self.n_pool = len(pool) // B
        self.H, self.W = pool.shape[1:]
        self.h_pool = torch.from_numpy(pool).pin_memory()
        self.d_pool = self.h_pool.to(ctx.dev)
        self.cap_total = B * cap
        self.offsets = np.zeros(B + 1, np.int32)
        self.launches = 0
        self.kp_seen = 0
        self.lookahead = LOOKAHEAD
        self.hbuf = None
        self.bytes_io = {"h2d": 0, "d2h": 0}

    def _slice(self, pool, i):
        j = i % self.n_pool
        return pool[j * self.B:(j + 1) * self.B]

    # Two batches in flight: step i submits batch i+1 (which continues from its look-ahead), starts the look-ahead of
    # batch i+2, and only then collects batch i -- the host never keeps the GPU waiting for its own wake-up.
    def _submit(self, i, mem):
        e, B, W, H = self.eng, self.B, self.W, self.H
        pool = self.d_pool if mem == self.MEM_DEVICE else self.h_pool
        e.submit(self._slice(pool, i).data_ptr(), mem, B, W, H, W, W * H, True)
        if self.lookahead:   # stages 1-3 (and, for host frames, the H2D) of the next batch run beside this batch's descriptor stage
            e.lookahead(self._slice(pool, i + 1).data_ptr(), mem, B, W, H, W, W * H)
        elif mem == self.MEM_HOST:
            e.prefetch(self._slice(pool, i + 1).data_ptr(), B, W, H, W, W * H)

    def prime_device(self):
        self._submit(0, self.MEM_DEVICE)

    def prime_e2e(self):
        self._submit(0, self.MEM_HOST)

    def drain(self):
        """Collects (into nowhere) the batch left in flight by a pipelined loop."""
        try:
            self.eng.collect(0, 0, self.MEM_DEVICE, self.cap_total, self.offsets)
        except Exception:
            pass
        self.eng.wait_outputs()
        self.eng.sync()

    def step_device(self, i):
        e, B = self.eng, self.B
        self._submit(i + 1, self.MEM_DEVICE)
        e.collect(0, 0, self.MEM_DEVICE, self.cap_total, self.offsets)   # batch i: waits for its counts; copies nothing
        self.launches += e.timings()["launches"]
        self.kp_seen += int(self.offsets[B])
        if self.match:
            e.match_prev_device(RATIO)       # frame t vs t-1 for the whole batch, on the device, own stream

    def alloc_host_outputs(self):
        """Two sets of pinned output buffers alternate so that the device->host copies of step i overlap the kernels
        of step i+1."""
        torch = self.ctx.torch
        n = self.cap_total
        self.hbuf = [dict(kps=torch.empty((n, 7), dtype=torch.float32).pin_memory(),
                          desc=torch.empty((n, self.D), dtype=torch.float32).pin_memory(),
                          matches=torch.empty((n * 2, 4), dtype=torch.int32).pin_memory() if self.match else None,
                          good=torch.empty((n,), dtype=torch.uint8).pin_memory() if self.match else None) for _ in range(2)]

    def step_e2e(self, i):
        e, B, W, H = self.eng, self.B, self.W, self.H
        hb = self.hbuf[i & 1]
        self._submit(i + 1, self.MEM_DEVICE if E2E_DIAG == "no_h2d" else self.MEM_HOST)
        e.wait_outputs()      # copies of step i-1 (other buffer set) have landed; this set is free again
        if E2E_DIAG == "no_d2h":   # diagnosis only: the end-to-end loop without its result copies
            e.collect(0, 0, self.MEM_DEVICE, self.cap_total, self.offsets)
            if self.match:
                e.match_prev_device(RATIO)
            return
        e.collect(hb["kps"].data_ptr(), 0 if E2E_DIAG == "no_desc" else hb["desc"].data_ptr(), self.MEM_HOST, self.cap_total,
                  self.offsets)
        n = int(self.offsets[B])
        self.kp_seen += n
        d2h = n * (28 + 4 * self.D) + (B + 1) * 4
        if self.match:
            e.match_prev_to(RATIO, hb["matches"].data_ptr(), hb["good"].data_ptr(), self.MEM_HOST, self.cap_total)
            d2h += 2 * n * 16 + n
        self.bytes_io = {"h2d": B * W * H, "d2h": d2h}

    def timed(self, step_fn, steps, warmup, sample_clocks=False):
        e, ctx = self.eng, self.ctx
        e.stream_reset()
        if step_fn == self.step_e2e and E2E_DIAG != "no_h2d":   # batch 0 in flight before step 0
            self.prime_e2e()
        else:
            self.prime_device()
        for i in range(warmup):
            step_fn(i)
        sampler = ClockSampler(ctx.local) if sample_clocks else None
        ctx.barrier()
        if sampler:
            sampler.start()
        self.launches = 0
        self.kp_seen = 0
        n_match = 0
        e.mark(0)
        for i in range(steps):
            step_fn(warmup + i)
            if self.match:
                n_match += 1
        e.wait_outputs()
        e.sync()               # compute stream and the matcher's own stream
        e.mark(1)
        ms = e.mark_elapsed(0, 1)
        if self.match and n_match:
            self.launches += n_match * e.match_prev_time()[1]
        clocks = sampler.stop() if sampler else None
        self.drain()           # the batch submitted by the last step
        ctx.barrier()
        return ctx.max_over_ranks(ms), clocks

    def stage_pass(self, steps, warmup):
        """Every stage timed with nothing else on the GPU: no look-ahead, the matcher drained before the next submit."""
        e, B, W, H = self.eng, self.B, self.W, self.H
        e.stream_reset()
        acc = {k: 0.0 for k in ("integral", "hessian", "sort", "describe", "pack", "match")}
        for i in range(warmup + steps):
            e.sync()
            e.submit(self._slice(self.d_pool, i).data_ptr(), self.MEM_DEVICE, B, W, H, W, W * H, True)
            e.collect(0, 0, self.MEM_DEVICE, self.cap_total, self.offsets)
            t = e.timings()
            mt = 0.0
            if self.match:
                e.match_prev_device(RATIO)
                mt = e.match_prev_time()[0]
            if i >= warmup:
                for k in ("integral", "hessian", "sort", "describe", "pack"):
                    acc[k] += t[k]
                acc["match"] += mt
        e.sync()
        return {k: v / steps for k, v in acc.items()}


def measure_latency(ctx: Ctx, eng, frames: np.ndarray, D: int, reps: int = 30):
    """Host-in / host-out wall-clock per call at small batches (blocking API, pinned buffers): the one-frame-per-tick
    use of computeNext (vtkSlicerSurfFeaturesLogic.cxx:1392-1406)."""
    from surffeatures_b200 import MEM_HOST

    torch = ctx.torch
    out = {}
    H, W = frames.shape[1:]
    for nb in (1, 4, 16):
        if nb > len(frames) or nb > eng.max_batch:
            continue
        h_in = torch.from_numpy(frames[:nb].copy()).pin_memory()
        kps = torch.empty((nb * CAP, 7), dtype=torch.float32).pin_memory()
        desc = torch.empty((nb * CAP, D), dtype=torch.float32).pin_memory()
        off = np.zeros(nb + 1, np.int32)
        def calls(graphs: bool):
            eng.set_graphs(graphs)
            ts = []
            for r in range(reps + 8):
                t0 = time.perf_counter()
                eng.submit(h_in.data_ptr(), MEM_HOST, nb, W, H, W, W * H, True)
                eng.collect(kps.data_ptr(), desc.data_ptr(), MEM_HOST, nb * CAP, off)
                ts.append((time.perf_counter() - t0) * 1e3)
            return np.array(ts[8:])

        ts_eager = calls(False)
        r0 = eng.graph_replays()
        ts = calls(True)   # the default: the launch sequence replayed as a CUDA graph
        out[f"batch{nb}"] = {"ms_per_call_median": float(np.median(ts)), "ms_p10": float(np.quantile(ts, 0.1)),
                             "ms_p90": float(np.quantile(ts, 0.9)), "ms_per_frame": float(np.median(ts) / nb),
                             "frames_per_s": float(nb / np.median(ts) * 1e3), "keypoints": int(off[nb]),
                             "graph_replays": int(eng.graph_replays() - r0), "calls": int(reps + 8),
                             "eager_ms_per_call_median": float(np.median(ts_eager)),
                             "launches_per_call": int(eng.timings()["launches"]),
                             "device_stage_ms": {k: round(float(v), 4) for k, v in eng.timings().items()
                                                 if k in ("h2d", "integral", "hessian", "sort", "describe", "pack", "d2h", "total")}}
    return out


# ------------------------------------------------------------------------------------------------
def run_config3(ctx: Ctx, steps_frames: int = 256):
    """BASELINE config 3: extended 128-d upright SURF on 1024x1024 frames, 256 frames (batches of 32), one GPU per rank."""
    from surffeatures_b200 import SURF

    B, cap, D, S = 32, 32768, 128, 1024
    pool = make_frames_parallel(B, S, S, 100 + 1000 * ctx.rank)      # 32 distinct frames, cycled 8x = 256 frames
    eng = SURF(100.0, 4, 3, True, True, device=ctx.local, max_width=S, max_height=S, max_batch=B, max_keypoints=cap)
    try:
        sb = StreamBench(ctx, eng, pool, B, D, cap, match=False)
        steps = steps_frames // B
        ms, _ = sb.timed(sb.step_device, steps, 2)
        kp = sb.kp_seen / (steps * B)
        sb.alloc_host_outputs()
        eng.set_async_outputs(True)
        ms_e2e, _ = sb.timed(sb.step_e2e, steps, 2)
        eng.set_async_outputs(False)
        return {"workload": "config3: 256 frames 1024x1024, extended (128-d) upright, thr=100, 4 octaves, 3 layers; batches "
                            "of 32 (32 distinct frames cycled)", "frames": steps * B * ctx.world,
                "value": ctx.world * steps * B / (ms * 1e-3), "unit": "frames/s", "ms_per_256_frames": ms * 256 / (steps * B),
                "keypoints_per_frame": kp,
                "e2e": {"value": ctx.world * steps * B / (ms_e2e * 1e-3), "unit": "frames/s",
                        "h2d_bytes_per_step": sb.bytes_io["h2d"], "d2h_bytes_per_step": sb.bytes_io["d2h"]},
                "scaling": "weak (one replica per GPU)"}
    finally:
        eng.close()


def run_config4(ctx: Ctx, comm_factory, n_total: int = 4096, reps: int = 2):
    """BASELINE config 4: a sweep of n_total 512x512 slices sharded by frame over the ranks; every batch's keypoints and
    descriptors go to rank 0 by NCCL (surf_sweep_sharded) while the next batch computes.  Strong scaling."""
    from surffeatures_b200 import SURF, shard_range

    torch = ctx.torch
    S, B, cap, D = 512, 64, 8192, 64
    lo, hi = shard_range(n_total, ctx.rank, ctx.world)
    n_local = hi - lo
    n_pool = 128
    pool = make_stream(n_pool, seed=2 + ctx.rank, h=S, w=S)           # 128 correlated slices, cycled over the shard
    reps_pool = (n_local + n_pool - 1) // n_pool
    d_frames = torch.from_numpy(pool).to(ctx.dev).repeat(reps_pool, 1, 1)[:n_local].contiguous()
    eng = SURF(100.0, 4, 3, False, False, device=ctx.local, max_width=S, max_height=S, max_batch=B, max_keypoints=cap)
    comm = comm_factory(eng) if ctx.world > 1 else None
    try:
        per_frame = 5000                                             # output capacity on rank 0 (~3.9k keypoints / slice)
        cap_rows = n_total * per_frame if ctx.rank == 0 else 0
        d_kps = torch.empty((max(cap_rows, 1), 7), dtype=torch.float32, device=ctx.dev)
        d_desc = torch.empty((max(cap_rows, 1), D), dtype=torch.float32, device=ctx.dev)
        off = np.zeros(n_total + 1, np.int64) if ctx.rank == 0 else None
        best = None
        for r in range(reps + 1):
            ctx.barrier()
            st = eng.sweep_sharded_device(comm, d_frames.data_ptr(), n_total, S, S, S, S * S, 0, d_kps.data_ptr(),
                                          d_desc.data_ptr(), cap_rows, off)
            ms = ctx.max_over_ranks(st["total_ms"])
            if r > 0 and (best is None or ms < best[0]):
                best = (ms, st)
        ms, st = best
        total_kp = int(off[-1]) if ctx.rank == 0 else 0
        res = {"workload": f"config4: sweep of {n_total} slices 512x512 sharded by frame over {ctx.world} GPU(s), thr=100, 4 "
                           f"octaves, 3 layers, 64-d; keypoints + descriptors gathered to rank 0 in frame order "
                           f"(NCCL send/recv per batch, overlapped with the next batch); 128 distinct slices per rank cycled",
               "frames": n_total, "value": n_total / (ms * 1e-3), "unit": "frames/s", "ms_per_sweep": ms,
               "scaling": "strong", "keypoints_gathered": total_kp,
               "nccl": {"bytes_into_rank0": int(st["bytes_received"]) if ctx.rank == 0 else None,
                        "exchange_ms_rank0": st["exchange_ms"], "assemble_ms_rank0": st["assemble_ms"],
                        "chunks_per_rank": st["chunks"],
                        "ingress_gbs": (st["bytes_received"] / (st["exchange_ms"] * 1e-3) / 1e9
                                        if ctx.rank == 0 and st["exchange_ms"] > 0 else None)},
               "data_path": "device-resident frames in, device-resident gathered results out (rank 0)"}
        return res
    finally:
        if comm is not None:
            comm.close()
        eng.close()


def run_config5(ctx: Ctx, comm_factory, n_rows: int = 1 << 20, nq: int = 4096, reps: int = 10):
    """BASELINE config 5: knnMatch(k=2) + ratio test of 4096 queries against a 1M x 64 database sharded by row over the
    ranks; per-rank tensor-core candidates + exact rescoring, NCCL all-gather of the 16 B x 2 x Nq records, merge."""
    from surffeatures_b200 import SURF, DescriptorDB, MEM_DEVICE, shard_range

    torch = ctx.torch
    D, k = 64, 2
    lo, hi = shard_range(n_rows, ctx.rank, ctx.world)
    rng = np.random.default_rng(3 + 17 * ctx.rank)
    rows = np.abs(rng.standard_normal((hi - lo, D), dtype=np.float32))
    rows *= rng.choice(np.array([-1.0, 1.0], np.float32), size=rows.shape, p=[0.35, 0.65])
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    q = rows[rng.integers(0, len(rows), nq)] + 0.05 * rng.standard_normal((nq, D), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    eng = SURF(100.0, 4, 3, False, False, device=ctx.local, max_width=64, max_height=64, max_batch=1, max_keypoints=1024)
    comm = comm_factory(eng) if ctx.world > 1 else None
    db = DescriptorDB(eng, D, max_rows=hi - lo, max_images=1)
    try:
        d_rows = torch.from_numpy(rows).to(ctx.dev)
        db.add_device(d_rows.data_ptr(), 0, np.array([0, hi - lo], np.int32))
        d_q = torch.from_numpy(q).to(ctx.dev)            # rank 0's rows are broadcast to everybody (query_root = 0)
        d_out = torch.empty((nq * k, 4), dtype=torch.int32, device=ctx.dev)
        times, nccl = [], []
        for r in range(reps + 3):
            ctx.barrier()
            eng.mark(2)
            db.knn_sharded_device(comm, d_q.data_ptr(), nq, k, d_out.data_ptr(), 0 if comm is not None else -1)
            eng.mark(3)
            if r >= 3:
                times.append(ctx.max_over_ranks(eng.mark_elapsed(2, 3)))
                nccl.append(comm.last_times()["nccl_ms"] if comm is not None else 0.0)
        ms = float(np.median(times))
        out = d_out.cpu().numpy().view(np.dtype([("q", "<i4"), ("t", "<i4"), ("img", "<i4"), ("d", "<f4")])).reshape(nq, k)
        passed = float(np.mean(out["d"][:, 0] < RATIO * out["d"][:, 1]))
        flops = 2.0 * nq * n_rows * D
        pk = _peaks()
        return {"workload": f"config5: knnMatch(k=2) + ratio(0.8) of {nq} queries against a {n_rows} x 64 database sharded by "
                            f"row over {ctx.world} GPU(s); tcgen05 candidates + exact fp32 rescoring per rank, NCCL all-gather "
                            f"of 16 B x 2 x Nq records per rank, merge (ties -> lower global row)",
                "value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_query_batch": ms, "nccl_ms": float(np.median(nccl)),
                "nccl_bytes_per_rank": 16 * k * nq + 4, "scaling": "strong", "ratio_test_pass_fraction": passed,
                "useful_tflops": flops / (ms * 1e-3) / 1e12,
                "tensor_frac_of_burst_peak": flops / (ms * 1e-3) / 1e12 / (pk["bf16_burst"] * ctx.world),
                "note": "the BF16 hi/lo split issues 3 MMAs per useful product: the tensor pipe does 3x the useful flops"}
    finally:
        db.close()
        if comm is not None:
            comm.close()
        eng.close()


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    ctx = Ctx()
    torch = ctx.torch
    from surffeatures_b200 import SURF, Comm

    if os.environ.get("SURF_B200_LIB") and not args.dev_lib:
        raise SystemExit("SURF_B200_LIB is a development knob (A/B builds); pass --dev-lib to measure another library")
    B, D = args.batch, 64
    steps, warmup = args.steps, max(args.warmup, 3)
    peaks = _peaks()

    pool = make_stream(POOL_BATCHES * B, seed=1 + ctx.rank)   # each rank owns its own shard of the stream (weak scaling)
    eng = SURF(**PARAMS, device=ctx.local, max_width=W_, max_height=H_, max_batch=B, max_keypoints=CAP)
    sb = StreamBench(ctx, eng, pool, B, D, CAP, match=True)

    ms_dev, clocks = sb.timed(sb.step_device, steps, warmup, sample_clocks=True)
    n_launch = sb.launches
    kp_per_frame = sb.kp_seen / (steps * B)
    sb.alloc_host_outputs()
    eng.set_async_outputs(True)
    ms_e2e, clocks_e2e = sb.timed(sb.step_e2e, steps, warmup, sample_clocks=True)
    eng.set_async_outputs(False)
    stage_ms = sb.stage_pass(max(5, steps // 3), 2)
    fps = ctx.world * B * steps / (ms_dev * 1e-3)
    fps_e2e = ctx.world * B * steps / (ms_e2e * 1e-3)

    # ---- roofline: every stage, then the dominant one in the contract's `roofline` object ----
    alg = algorithmic_bytes_per_frame(kp_per_frame, D)
    match_flops = 2.0 * kp_per_frame * kp_per_frame * D * B
    stages = [
        {"stage": "integral", "kernels": "k_integral_band", "bound": "hbm",
         "algorithmic_bytes": alg["integral"] * B, "ms": stage_ms["integral"]},
        {"stage": "hessian+maxima", "kernels": "k_hessian_nms_tma_c<0|1>+k_hessian+k_nms", "bound": "hbm",
         "algorithmic_bytes": alg["hessian_nms_fused"] * B, "algorithmic_bytes_unfused": (alg["hessian"] + alg["nms"]) * B,
         "ms": stage_ms["hessian"]},
        {"stage": "sort", "kernels": "k_sort_chunks+k_merge_pass", "bound": "hbm",
         "algorithmic_bytes": 2 * 28 * kp_per_frame * B, "ms": stage_ms["sort"]},
        {"stage": "describe", "kernels": "k_orient+k_descriptor<0..2>+k_descriptor_big", "bound": "hbm",
         "algorithmic_bytes": alg["describe"] * B, "ms": stage_ms["describe"]},
    ]
    for s in stages:
        s["achieved"] = s["algorithmic_bytes"] / (s["ms"] * 1e-3) / 1e9 if s["ms"] > 0 else None
        s["peak"], s["unit"] = peaks["hbm"], "GB/s"
        s["frac"] = s["achieved"] / peaks["hbm"] if s["achieved"] else None
        if "algorithmic_bytes_unfused" in s and s["ms"] > 0:
            s["frac_unfused_accounting"] = s["algorithmic_bytes_unfused"] / (s["ms"] * 1e-3) / 1e9 / peaks["hbm"]
    stages.append({"stage": "match", "kernels": "k_tc_prep+k_tc_candidates3w+k_tc_rescore+k_tc_redo", "bound": "tensor",
                   "algorithmic_flops": match_flops, "ms": stage_ms["match"],
                   "achieved": match_flops / (stage_ms["match"] * 1e-3) / 1e12 if stage_ms["match"] > 0 else None,
                   "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                   "frac": match_flops / (stage_ms["match"] * 1e-3) / 1e12 / peaks["bf16_burst"] if stage_ms["match"] > 0 else None,
                   "note": "useful flops 2*Nq*Nt*D; the BF16 hi/lo split issues 3 MMAs per useful product"})
    dom = max(stages[:4], key=lambda s: s["ms"])
    # dram__bytes_read+write and smsp__inst_executed of the describe-stage kernels from one `ncu --set full` capture of this
    # very command at batch 64 (profiles/: see NCU_* below)
    traffic = NCU_DESCRIBE_TRAFFIC if (dom["stage"] == "describe" and B == 64) else None
    roofline = {"bound": "hbm", "kernel": dom["kernels"], "achieved": dom["achieved"], "peak": peaks["hbm"],
                "peak_source": peaks["source"] + " hbm_gbs, burst copy peak", "unit": "GB/s", "frac": dom["frac"],
                "traffic": traffic, "traffic_source": NCU_SOURCE if traffic else None,
                "limiter": "instruction issue, not HBM: the window kernels sit at 74-84 % issue-slot utilisation "
                           "(about 6300 exact bilinear taps per keypoint, 56-74 % of their instructions); see "
                           "profiles/r2/describe_stage_v5_sections.txt",
                "issue_slots": ({"warp_inst_per_step": NCU_DESCRIBE_INST,
                                 "achieved_ginst_s": NCU_DESCRIBE_INST / (dom["ms"] * 1e-3) / 1e9,
                                 "peak_ginst_s": 148 * 4 * 1.965,
                                 "frac": NCU_DESCRIBE_INST / (dom["ms"] * 1e-3) / (148 * 4 * 1.965e9)}
                                if traffic else None),
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes"], "ms_per_launch": dom["ms"],
                "whole_path_frac": alg["total"] * B * steps / (ms_dev * 1e-3) / 1e9 / peaks["hbm"]}

    latency = None
    if ctx.rank == 0 and not args.no_extras:
        latency = measure_latency(ctx, eng, pool[:16], D)
    ctx.barrier()

    def make_line(extras, cpu_baseline):
        return {
            "metric": _metric(), "value": fps, "unit": "frames/s", "n_gpus": ctx.world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32/f32", "data": "synthetic",
            "config": workload_config(B),
            "notes": {"frames_per_step_all_gpus": B * ctx.world, "sharding": "frames sharded by rank, no data-path collective "
                      "(config 2 does not shard below a frame: replicas); configs 4 and 5 below use NCCL",
                      "lookahead": bool(sb.lookahead), "two_batches_in_flight": True, "e2e_diag": E2E_DIAG or None,
                      "host_affinity": ctx.affinity,
                      "library": os.environ.get("SURF_B200_LIB", "surffeatures_b200/libsurf_b200.so")},
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": sb.bytes_io["h2d"],
                    "d2h_bytes_per_step": sb.bytes_io["d2h"], "ms_per_step": ms_e2e / steps, "clocks": clocks_e2e},
            "gpu_launches": n_launch, "keypoints_per_frame": kp_per_frame,
            "roofline": roofline, "roofline_stages": stages,
            "stage_ms_note": "stages timed in a separate pass with nothing else on the GPU (no look-ahead, matcher drained "
                             "before the next submit); the headline overlaps stages 1-3 of the next batch and the matcher "
                             "with stage 4",
            "latency": latency, "latency_ms_batch1": latency["batch1"]["ms_per_call_median"] if latency else None,
            "configs": extras, "cpu_baseline": cpu_baseline, "clocks": clocks,
        }

    extras = {}
    if not args.no_extras:
        comm_factory = Comm.from_torch
        # The extras run collectives; a rank that fails inside one would leave the others waiting forever and take the
        # headline down with it.  A watchdog bounds them: when it fires, rank 0 prints the line with what is done so far
        # and every rank leaves at once (the one case in which the process does not shut down in order).
        done = threading.Event()

        def bail():
            if done.is_set():
                return
            if ctx.rank == 0:
                extras.setdefault("error", f"extras did not finish within {args.extras_timeout} s")
                print(json.dumps(make_line(extras, None)), flush=True)
            os._exit(0)

        dog = threading.Timer(args.extras_timeout, bail)
        dog.daemon = True
        dog.start()
        for name, fn in (("config3", lambda: run_config3(ctx)),
                         ("config4", lambda: run_config4(ctx, comm_factory, args.sweep_frames)),
                         ("config5", lambda: run_config5(ctx, comm_factory))):
            if args.only_extra and name != args.only_extra:
                continue
            failed = 0.0
            try:
                extras[name] = fn()
            except Exception as ex:   # an extra must not take the headline down with it (recorded, not hidden)
                extras[name] = {"error": f"{type(ex).__name__}: {ex}"}
                failed = 1.0
            # every rank learns whether any rank failed: the remaining multi-rank extras are skipped by all or by none
            if ctx.max_over_ranks(failed) > 0:
                extras.setdefault(name, {"error": "another rank failed"})
                break
            ctx.barrier()
        done.set()
        dog.cancel()

    cpu_baseline = None
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu:
        cpu_baseline = cpu_baseline_leg(pool)

    if ctx.rank == 0:
        print(json.dumps(make_line(extras, cpu_baseline)), flush=True)
    # orderly shutdown: engine first (its streams and events), then the process group; no os._exit
    eng.close()
    torch.cuda.synchronize()
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    return 0


# from profiles/r2/describe_stage_v11_ncu_summary.txt (ncu --set full of this command, batch 64, the kernels of HEAD): k_orient +
# the four window kernels (the patch -> descriptor step runs inside them): 188.9 MB read + 13.0 MB written, 3.704 G warp
# instructions.  Round 1: 428.9 MB and 4.42 G; first half of round 2: 203.2 MB and 3.93 G.
NCU_DESCRIBE_TRAFFIC, NCU_DESCRIBE_INST = 201.9e6, 3.704e9
NCU_SOURCE = "profiles/r2/describe_stage_v11_ncu_summary.txt"


def cpu_baseline_leg(pool: np.ndarray):
    """The oracle port on the box's host cores, on a bounded sample of the same stream (about 10-30 s of CPU work):
    T-thread frame-parallel passes (median of the repetitions that fit the budget) and a 1-thread row."""
    from oracle import surf_oracle as so

    threads = host_threads()
    p = so.params(**PARAMS)
    ns = max(threads, 8)
    frames = pool[:ns + 1]
    _cpu_stream_pass(so, frames[:3], p, threads)       # warm-up
    reps, t_all = [], time.perf_counter()
    while len(reps) < 20 and (time.perf_counter() - t_all < 15.0 or len(reps) < 3):
        t0 = time.perf_counter()
        _cpu_stream_pass(so, frames, p, threads)
        reps.append(ns / (time.perf_counter() - t0))
    t0 = time.perf_counter()
    _cpu_stream_pass(so, frames[:3], p, 1)
    one = 2 / (time.perf_counter() - t0)
    return {"value": float(np.median(reps)), "unit": "frames/s", "cores": threads, "kind": "port",
            "p10": float(np.quantile(reps, 0.1)), "p90": float(np.quantile(reps, 0.9)), "repetitions": len(reps),
            "single_thread_frames_per_s": one,
            "sample": f"{ns} frames of the same stream per repetition, detect+describe+match, oracle port of OpenCV 2.4.5 "
                      f"SURF (genuine nonfree not installable offline), {threads} threads frame-parallel; 1-thread row on 2 frames"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="headline only: no latency table, no configs 3/4/5")
    ap.add_argument("--only-extra", default="", help="run only this extra (config3|config4|config5)")
    ap.add_argument("--sweep-frames", type=int, default=4096, help="slices of the config-4 sweep")
    ap.add_argument("--extras-timeout", type=float, default=420.0, help="watchdog of the configs 3/4/5 section, seconds")
    ap.add_argument("--dev-lib", action="store_true", help="allow SURF_B200_LIB (development A/B builds)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
