"""Throughput of the general-N team kernels at the C4 / C5 shapes (reduced trajectory counts)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhmm_b200.engine import TrajectoryBatch, unpack_stats
from bhmm_b200.util import testsystems as ts

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

# C5 shape: N=32 Gaussian
N, K, T = 32, 64, 100000
pi, A, means, sigmas, O, S = ts.gaussian_observations(N, K, T, seed=5)
b = TrajectoryBatch(list(O), N); b.set_profiling(True)
ms = timeit(lambda: b.estep_gaussian(A, pi, means, sigmas))
print('N=32 gaussian K=%d T=%d: estep %.2f ms -> %.3f Gframe/s, kernels %s, info %s' % (K, T, ms, K*T/ms/1e6, b.kernel_ms(), b.info()), flush=True)
ms = timeit(lambda: b.viterbi_gaussian(A, pi, means, sigmas), reps=1)
print('   viterbi %.2f ms -> %.3f Gframe/s' % (ms, K*T/ms/1e6), flush=True)
b.close()
# C4 shape: N=100 discrete, M=1000
N, M, K, T = 100, 1000, 32, 100000
pi, A, B, O, S = ts.discrete_observations(N, M, K, T, seed=4)
b = TrajectoryBatch(list(O), N); b.set_profiling(True)
ms = timeit(lambda: b.estep_discrete(A, pi, B), reps=2)
print('N=100 discrete M=1000 K=%d T=%d: estep %.2f ms -> %.4f Gframe/s, kernels %s, info %s' % (K, T, ms, K*T/ms/1e6, b.kernel_ms(), b.info()), flush=True)
ms = timeit(lambda: b.viterbi_discrete(A, pi, B), reps=1)
print('   viterbi %.2f ms -> %.4f Gframe/s' % (ms, K*T/ms/1e6), flush=True)
b.close()
# C1 / C2 shapes: N=3
for K, T in ((10, 10000), (100, 10000)):
    pi, A, means, sigmas, O, S = ts.gaussian_observations(3, K, T, seed=1)
    b = TrajectoryBatch(list(O), 3); b.set_profiling(True)
    ms = timeit(lambda: b.estep_gaussian(A, pi, means, sigmas), reps=10)
    msg = timeit(lambda: b.gibbs_gaussian(A, pi, means, sigmas, seed=1, sweep=0), reps=10)
    print('N=3 K=%d T=%d: estep %.3f ms -> %.3f Gframe/s (kernels %s) gibbs %.3f ms -> %.3f Gframe/s info %s' % (K, T, ms, K*T/ms/1e6, b.kernel_ms(), msg, K*T/msg/1e6, b.info()), flush=True)
    b.close()
