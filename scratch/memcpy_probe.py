import ctypes, time, torch
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
rt.cudaStreamCreateWithFlags.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint]
rt.cudaStreamSynchronize.argtypes = [ctypes.c_void_p]
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
d = torch.zeros(80 * 1024 * 1024, dtype=torch.uint8, device='cuda')
h = torch.empty(80 * 1024 * 1024, dtype=torch.uint8).pin_memory()
h2 = ctypes.c_void_p(); print('hostalloc', rt.cudaHostAlloc(ctypes.byref(h2), 80 * 1024 * 1024, 0))
s = ctypes.c_void_p(); rt.cudaStreamCreateWithFlags(ctypes.byref(s), 1)
torch.cuda.synchronize()
for name, dst in (('torch pinned', h.data_ptr()), ('cudaHostAlloc', h2.value)):
    for _ in range(3):
        t0 = time.perf_counter(); r = rt.cudaMemcpyAsync(dst, d.data_ptr(), 80 * 1024 * 1024, 2, s); t1 = time.perf_counter()
        rt.cudaStreamSynchronize(s); t2 = time.perf_counter()
        print(name, 'rc', r, 'enqueue %.3f ms, total %.3f ms' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
