import ctypes, time, torch
rt = ctypes.CDLL("libcudart.so.12")
vp = ctypes.c_void_p
rt.cudaMemcpyAsync.argtypes = [vp, vp, ctypes.c_size_t, ctypes.c_int, vp]
rt.cudaStreamCreateWithFlags.argtypes = [ctypes.POINTER(vp), ctypes.c_uint]
rt.cudaStreamSynchronize.argtypes = [vp]
rt.cudaEventCreate.argtypes = [ctypes.POINTER(vp)]
rt.cudaEventRecord.argtypes = [vp, vp]
rt.cudaEventSynchronize.argtypes = [vp]
N = 80 * 1024 * 1024
d = torch.zeros(N, dtype=torch.uint8, device='cuda')
h = torch.empty(N, dtype=torch.uint8).pin_memory()
a, b = vp(), vp(); rt.cudaStreamCreateWithFlags(ctypes.byref(a), 1); rt.cudaStreamCreateWithFlags(ctypes.byref(b), 1)
ev = vp(); rt.cudaEventCreate(ctypes.byref(ev))
x = torch.zeros(1 << 20, device='cuda')
torch.cuda.synchronize()
for it in range(3):
    t0 = time.perf_counter()
    rt.cudaMemcpyAsync(h.data_ptr(), d.data_ptr(), N, 2, b)
    rt.cudaEventRecord(ev, a); rt.cudaEventSynchronize(ev)
    t1 = time.perf_counter(); rt.cudaStreamSynchronize(b); t2 = time.perf_counter()
    print('event on other stream after D2H enqueue: %.3f ms; copy done %.3f ms' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
# kernel on torch stream concurrently with copy
s = torch.cuda.Stream()
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rt.cudaMemcpyAsync(h.data_ptr(), d.data_ptr(), N, 2, b)
    with torch.cuda.stream(s):
        for _ in range(20): x.mul_(1.0001)
    s.synchronize(); t1 = time.perf_counter(); rt.cudaStreamSynchronize(b); t2 = time.perf_counter()
    print('20 small kernels on another stream: %.3f ms; copy done %.3f ms' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
