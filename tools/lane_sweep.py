"""Kernel-level sweep of the lane family on the C3 shape (N=10, K x 1e5 frames): forward / backward+statistics
kernel times for a list of chunk lengths.  The library variant is chosen with BHMM_B200_LIB."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_gaussian_gpu  # noqa: E402
from bhmm_b200.engine import TrajectoryBatch  # noqa: E402
from bhmm_b200.util import testsystems as ts  # noqa: E402

N = int(os.environ.get('SWEEP_N', 10))
K = int(os.environ.get('SWEEP_K', 1024))
T = int(os.environ.get('SWEEP_T', 100000))
chunks = [int(x) for x in os.environ.get('SWEEP_CHUNKS', '2703,1802,1352').split(',')]
warm = int(os.environ.get('SWEEP_WARM', 320))
dev = torch.device('cuda', 0)
pi, A, means, sigmas, O = synth_gaussian_gpu(N, K, T, 3, dev)
pi0, A0, m0, s0 = ts.perturbed_initial_model(A, means, N)
tag = os.path.basename(os.environ.get('BHMM_B200_LIB', 'default'))
import subprocess
def clocks():
    try:
        return subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,temperature.gpu', '--format=csv,noheader'],
                              capture_output=True, text=True).stdout.strip()
    except Exception:
        return '?'

for chunk in chunks:
    b = TrajectoryBatch.from_concatenated(O.reshape(-1), [T] * K, N, chunk=chunk, warm=warm)
    b.set_profiling(True)
    # parameters near the truth (the regime EM spends its time in)
    b.estep_gaussian(A, pi, means, sigmas)
    f = bw = 0.0
    reps = int(os.environ.get('SWEEP_REPS', 5))
    for _ in range(reps):
        b.estep_gaussian(A, pi, means, sigmas)
        k = b.kernel_ms()
        f += k['forward'] / reps
        bw += k['backward_stats'] / reps
    info = b.info()
    print('%s N=%d K=%d chunk=%d warm=%d chains=%d: fwd %.3f ms bwd %.3f ms sum %.3f ms -> %.2f Gframe/s (kernels) fix=%d/%d [%s]'
          % (tag, N, K, info['chunk'], info['warm'], info['chains'], f, bw, f + bw, K * T / (f + bw) / 1e6,
             info['fixups_fwd'], info['fixups_bwd'], clocks()), flush=True)
    b.close()
    del b
