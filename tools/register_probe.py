"""How fast is pinning the caller's own arrays (cudaHostRegister) compared with staging them?  C3 sizes: 1024 pageable
arrays of 1e5 float64 up, one fresh 410 MB int32 array down.  (Decides whether the mover should DMA from / to the caller's
memory directly: one pass over host memory instead of three.)"""
import ctypes as C
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
import torch

rt = None
for name in ('libcudart.so.12', 'libcudart.so'):
    try:
        rt = C.CDLL(name); break
    except OSError:
        pass
assert rt is not None
rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
rt.cudaHostUnregister.argtypes = [C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
K, T = 1024, 100000
rng = np.random.default_rng(0)
host = [rng.standard_normal(T) for _ in range(K)]
dev = torch.device('cuda', 0)
torch.zeros(1, device=dev)
dst = torch.empty(K * T, dtype=torch.float64, device=dev)
path = torch.randint(0, 10, (K * T,), dtype=torch.int32, device=dev)
def sync():
    torch.cuda.synchronize(); return time.perf_counter()
st = torch.cuda.current_stream().cuda_stream

for threads in (1, 4, 8):
    pool = ThreadPoolExecutor(threads)
    for rep in range(2):
        t0 = sync()
        rcs = list(pool.map(lambda a: rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0), host))
        t1 = sync()
        off = 0
        for a in host:
            rt.cudaMemcpyAsync(dst.data_ptr() + off, a.ctypes.data, a.nbytes, 1, st); off += a.nbytes
        t2 = sync()
        list(pool.map(lambda a: rt.cudaHostUnregister(a.ctypes.data), host))
        t3 = sync()
    assert not any(rcs), rcs[:4]
    print('upload, %d threads: register %.1f ms, DMA from the arrays %.1f ms (%.1f GB/s), unregister %.1f ms -> %.1f ms in all'
          % (threads, 1e3 * (t1 - t0), 1e3 * (t2 - t1), K * T * 8 / (t2 - t1) / 1e9, 1e3 * (t3 - t2), 1e3 * (t3 - t0)), flush=True)
assert np.array_equal(dst[:T].cpu().numpy(), host[0])

nb = K * T * 4
for threads in (1, 4, 8):
    pool = ThreadPoolExecutor(threads)
    for rep in range(2):
        t0 = sync()
        out = np.empty(K * T, dtype=np.int32)
        piece = -(-nb // (threads * 4096)) * 4096
        parts = [(out.ctypes.data + o, min(piece, nb - o)) for o in range(0, nb, piece)]
        rcs = list(pool.map(lambda p: rt.cudaHostRegister(p[0], p[1], 0), parts))
        t1 = sync()
        rt.cudaMemcpyAsync(out.ctypes.data, path.data_ptr(), nb, 2, st)
        t2 = sync()
        list(pool.map(lambda p: rt.cudaHostUnregister(p[0]), parts))
        t3 = sync()
    print('download, %d threads: register a fresh array %.1f ms (rc %s), DMA into it %.1f ms (%.1f GB/s), unregister %.1f ms -> %.1f ms in all'
          % (threads, 1e3 * (t1 - t0), set(rcs), 1e3 * (t2 - t1), nb / (t2 - t1) / 1e9, 1e3 * (t3 - t2), 1e3 * (t3 - t0)), flush=True)
assert np.array_equal(out, path.cpu().numpy())
