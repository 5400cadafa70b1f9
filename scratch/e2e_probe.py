import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from surffeatures_b200 import SURF, MEM_HOST, MEM_DEVICE, speckle
B, W, H, D, CAP = 64, 640, 480, 64, 8192
pool = speckle.speckle_stream(2 * B, H, W, seed=1)
h_pool = torch.from_numpy(pool).pin_memory(); d_pool = h_pool.cuda()
eng = SURF(100, 4, 3, max_width=W, max_height=H, max_batch=B, max_keypoints=CAP)
st = torch.cuda.ExternalStream(eng.cuda_stream)
cap = B * CAP
hb = [dict(k=torch.empty((cap, 7)).pin_memory(), d=torch.empty((cap, D)).pin_memory(), m=torch.empty((cap * 2, 4), dtype=torch.int32).pin_memory(), g=torch.empty((cap,), dtype=torch.uint8).pin_memory()) for _ in range(2)]
off = np.zeros(B + 1, np.int32)
def run(name, host_in, copy_out, do_match, async_out, steps=8):
    eng.set_async_outputs(async_out); eng.stream_reset()
    def step(i):
        b = hb[i & 1]
        src = (h_pool if host_in else d_pool)[(i % 2) * B:(i % 2 + 1) * B]
        eng.submit(src.data_ptr(), MEM_HOST if host_in else MEM_DEVICE, B, W, H, W, W * H, True)
        if async_out: eng.wait_outputs()
        if copy_out: eng.collect(b['k'].data_ptr(), b['d'].data_ptr(), MEM_HOST, cap, off)
        else: eng.collect(0, 0, MEM_DEVICE, cap, off)
        if do_match:
            if copy_out: eng.match_prev_to(0.8, b['m'].data_ptr(), b['g'].data_ptr(), MEM_HOST, cap)
            else: eng.match_prev_device(0.8)
    for i in range(3): step(i)
    eng.wait_outputs(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps): step(3 + i)
    eng.wait_outputs(); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps * 1e3
    print(f"{name:45s} {dt:7.2f} ms/step  {B / dt * 1e3:8.0f} fps", flush=True)
run("device in, no out, no match", False, False, False, False)
run("device in, no out, match", False, False, True, False)
run("host in, no out, match", True, False, True, False)
run("host in, copy out sync, match", True, True, True, False)
run("host in, copy out async, match", True, True, True, True)
run("device in, copy out async, match", False, True, True, True)
import os; os._exit(0)
