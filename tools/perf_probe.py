"""Quick throughput probe of the fused E-step / Gibbs sweep (not the bench: no JSON contract)."""
import sys
import time

import numpy as np
import torch

from bhmm_b200.engine import TrajectoryBatch, unpack_stats
from bhmm_b200.util import testsystems as ts


def probe(N, K, T, chunks=(0,), reps=3):
    pi, A, means, sigmas, O, S = ts.gaussian_observations(N, K, T, seed=3)
    pi0, A0, m0, s0 = ts.perturbed_initial_model(A, means, N)
    for chunk in chunks:
        b = TrajectoryBatch(list(O), N, chunk=chunk)
        b.set_profiling(True)
        b.estep_gaussian(A0, pi0, m0, s0)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(reps):
            st = b.estep_gaussian(A0, pi0, m0, s0)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        info = b.info()
        km = b.kernel_ms()
        print('estep[%s] N=%d K=%d T=%d chunk=%d warm=%d chains=%d: %.3f ms (fwd %.3f bwd %.3f)  %.3f Gframe/s  fix=%d/%d worst=%.1e/%.1e' % (
            'lane' if b.uses_lane_kernels else 'team', N, K, T, info['chunk'], info['warm'], info['chains'], ms,
            km['forward'], km['backward_stats'], K * T / ms / 1e6, info['fixups_fwd'],
            info['fixups_bwd'], info['worst_fwd'], info['worst_bwd']), flush=True)
        ev[0].record()
        for s in range(reps):
            b.gibbs_gaussian(A0, pi0, m0, s0, seed=1, sweep=s)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        print('gibbs N=%d K=%d T=%d chunk=%d: %.3f ms  %.3f Gframe/s' % (N, K, T, info['chunk'], ms, K * T / ms / 1e6), flush=True)
        b.close()


if __name__ == '__main__':
    t0 = time.time()
    probe(3, 100, 10000, chunks=(0, 512))
    probe(10, 256, 100000, chunks=(0, 1352, 2704, 5408))
    probe(16, 64, 50000, chunks=(0,))
    probe(32, 16, 20000, chunks=(0,))
    print('total %.1f s' % (time.time() - t0))
