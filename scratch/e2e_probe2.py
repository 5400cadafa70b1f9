import os; os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from surffeatures_b200 import SURF, MEM_HOST, MEM_DEVICE, speckle
B, W, H, D, CAP = 64, 640, 480, 64, 8192
pool = speckle.speckle_stream(2 * B, H, W, seed=1)
h_pool = torch.from_numpy(pool).pin_memory(); d_pool = h_pool.cuda()
eng = SURF(100, 4, 3, max_width=W, max_height=H, max_batch=B, max_keypoints=CAP)
cap = B * CAP
hb = [dict(k=torch.empty((cap, 7)).pin_memory(), d=torch.empty((cap, D)).pin_memory(), m=torch.empty((cap * 2, 4), dtype=torch.int32).pin_memory(), g=torch.empty((cap,), dtype=torch.uint8).pin_memory()) for _ in range(2)]
off = np.zeros(B + 1, np.int32)
eng.set_async_outputs(True); eng.stream_reset()
T = {k: 0.0 for k in ('submit', 'wait', 'collect', 'match')}
def step(i, rec):
    b = hb[i & 1]
    src = d_pool[(i % 2) * B:(i % 2 + 1) * B]
    t0 = time.perf_counter(); eng.submit(src.data_ptr(), MEM_DEVICE, B, W, H, W, W * H, True)
    t1 = time.perf_counter(); eng.wait_outputs()
    t2 = time.perf_counter(); eng.collect(b['k'].data_ptr(), b['d'].data_ptr(), MEM_HOST, cap, off)
    t3 = time.perf_counter(); eng.match_prev_to(0.8, b['m'].data_ptr(), b['g'].data_ptr(), MEM_HOST, cap)
    t4 = time.perf_counter()
    if rec:
        T['submit'] += t1 - t0; T['wait'] += t2 - t1; T['collect'] += t3 - t2; T['match'] += t4 - t3
for i in range(3): step(i, False)
eng.wait_outputs(); torch.cuda.synchronize()
t0 = time.perf_counter()
N = 8
for i in range(N): step(3 + i, True)
eng.wait_outputs(); torch.cuda.synchronize()
print('total ms/step', (time.perf_counter() - t0) / N * 1e3, {k: round(v / N * 1e3, 3) for k, v in T.items()}, eng.timings())
import os; os._exit(0)
