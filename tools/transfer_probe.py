"""Throughput of the estimators' one-off transfers (csrc/transfer.cu) at the C3 size: 1024 host arrays of 1e5 float64 up,
1.024e8 int32 path entries down, for several worker-thread counts; the first call of a process (pinned staging slots,
streams) is timed separately.  Slot size: BHMM_B200_STAGE_KB (read at the first transfer of the process)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bhmm_b200.engine as eng

K, T = 1024, 100000
rng = np.random.default_rng(0)
host = [rng.standard_normal(T) for _ in range(K)]
dev = torch.device('cuda', 0)
torch.cuda.init()
torch.zeros(1, device=dev)
def sync():
    torch.cuda.synchronize(); return time.perf_counter()
dst = torch.empty(K * T, dtype=torch.float64, device=dev)
path = torch.randint(0, 10, (K * T,), dtype=torch.int32, device=dev)
threads0 = os.environ.get('BHMM_B200_TRANSFER_THREADS', 'auto')
t0 = sync()
eng.upload_arrays(dst[:8], [np.zeros(8)])
t1 = sync()
print('stage %s KB, threads %s: first (tiny) transfer of the process %.1f ms' % (os.environ.get('BHMM_B200_STAGE_KB', '2048'), threads0, 1e3 * (t1 - t0)), flush=True)
for th in [int(x) for x in os.environ.get('PROBE_THREADS', '2,4,8,12,16').split(',')]:
    os.environ['BHMM_B200_TRANSFER_THREADS'] = str(th)
    ups, downs = [], []
    for rep in range(3):
        t0 = sync()
        eng.upload_arrays(dst, host)
        t1 = sync()
        out = eng.download_array(path)
        t2 = sync()
        ups.append(t1 - t0); downs.append(t2 - t1)
    print('  threads %2d: upload %.1f ms (%.1f GB/s), download into a fresh array %.1f ms (%.1f GB/s)'
          % (th, 1e3 * min(ups), K * T * 8 / min(ups) / 1e9, 1e3 * min(downs), K * T * 4 / min(downs) / 1e9), flush=True)
assert np.array_equal(dst.cpu().numpy()[:T], host[0]) and np.array_equal(out, path.cpu().numpy())
# for scale: what torch does with the same data
t0 = sync()
off = 0
for a in host[:256]:
    dst[off:off + T].copy_(torch.from_numpy(a)); off += T
t1 = sync()
p = path.cpu().numpy()
t2 = sync()
print('  torch: per-array pageable copy_ %.1f GB/s, .cpu() of the paths %.1f ms (%.1f GB/s)'
      % (256 * T * 8 / (t1 - t0) / 1e9, 1e3 * (t2 - t1), K * T * 4 / (t2 - t1) / 1e9), flush=True)
