import sys, numpy as np, torch
sys.path.insert(0, '.')
from surffeatures_b200 import SURF, MEM_DEVICE, speckle
B, W, H, CAP = 64, 640, 480, 8192
pool = speckle.speckle_stream(2 * B, H, W, seed=1)
d_pool = torch.from_numpy(pool).cuda()
eng = SURF(100, 4, 3, max_width=W, max_height=H, max_batch=B, max_keypoints=CAP)
off = np.zeros(B + 1, np.int32)
for i in range(2):
    src = d_pool[i * B:(i + 1) * B]
    eng.submit(src.data_ptr(), MEM_DEVICE, B, W, H, W, W * H, True)
    eng.collect(0, 0, MEM_DEVICE, B * CAP, off)
    eng.match_prev_device(0.8)
    print('step', i, 'keypoints', int(off[B]), 'redo', eng.match_redo_count(), flush=True)
import os; os._exit(0)
