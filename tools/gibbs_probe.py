import os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from bench import synth_gaussian_gpu
from bhmm_b200.engine import TrajectoryBatch
for N, K, T in ((10, 1024, 100000), (3, 4096, 100000)):
    dev = torch.device('cuda', 0)
    pi, A, means, sigmas, O = synth_gaussian_gpu(N, K, T, 3, dev)
    b = TrajectoryBatch.from_concatenated(O.reshape(-1), [T] * K, N)
    for _ in range(2): b.gibbs_gaussian(A, pi, means, sigmas, seed=1, sweep=0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(5): b.gibbs_gaussian(A, pi, means, sigmas, seed=1, sweep=k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(os.path.basename(os.environ.get('BHMM_B200_LIB', 'default')), 'N=%d gibbs sweep %.3f ms -> %.2f Gframe*sweeps/s' % (N, ms, K * T / ms / 1e6), flush=True)
    b.close(); del b, O
