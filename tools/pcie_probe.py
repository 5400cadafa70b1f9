#!/usr/bin/env python
"""Host<->device copy bandwidth of this box (pinned memory, CUDA events): what bounds the end-to-end loop of bench.py,
which moves 19.7 MB up and ~86 MB down per 64-frame step.  Prints one JSON line."""
import json

import torch

dev = torch.device("cuda", 0)
up = torch.empty(64 * 640 * 480, dtype=torch.uint8).pin_memory()
down = torch.empty(86_000_000, dtype=torch.uint8).pin_memory()
d_up = torch.empty_like(up, device=dev)
d_down = torch.empty_like(down, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_up.copy_(up, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        down.copy_(d_down, non_blocking=True)


def both():
    h2d(); d2h()


t_up, t_down, t_both = timed(h2d), timed(d2h), timed(both)
print(json.dumps({"h2d_19.7MB_ms": t_up, "h2d_GBs": up.numel() / t_up / 1e6, "d2h_86MB_ms": t_down,
                  "d2h_GBs": down.numel() / t_down / 1e6, "both_concurrent_ms": t_both,
                  "note": "both = one upload and one download issued together on two streams"}))
