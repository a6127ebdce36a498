"""C5 shape, Viterbi: ONE trajectory of --frames frames (default 1e9), 32 states, on ONE GPU.

    BHMM_B200_PANEL=1 python tools/c5_viterbi.py --frames 1e9        # time-chunked Viterbi (opt-in kernels)
    python tools/c5_viterbi.py --frames 2e7                            # sequential kernel: ~0.5 us per frame

The batch is Viterbi-only (no forward-variable workspace: N + 12 bytes per frame), the observations are generated on the
device in blocks.  Prints one JSON line: frames per second, the plan (chains, chunk, warm-up, fix-up sweeps) and a
checksum of the path: run a SHORT trajectory with and without BHMM_B200_PANEL and compare the checksums of the two lines
-- the paths must be identical.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhmm_b200.engine import TrajectoryBatch  # noqa: E402

BLOCK = 1 << 22


def frames(T, A, means, sigmas, dev, seed=99):
    """A real hidden path would cost a sequential pass; blocks of i.i.d. states with long dwell times (each state repeated
    64 frames) give the Viterbi recursion realistic, well-separated decisions at no sequential cost."""
    out = torch.empty(T, dtype=torch.float64, device=dev)
    mu = torch.as_tensor(means, device=dev)
    sg = torch.as_tensor(sigmas, device=dev)
    for b in range(0, T, BLOCK):
        n = min(BLOCK, T - b)
        g = torch.Generator(device=dev)
        g.manual_seed(seed + b // BLOCK)
        s = torch.randint(0, len(means), ((n + 63) // 64,), generator=g, device=dev).repeat_interleave(64)[:n]
        out[b:b + n] = mu[s] + sg[s] * torch.randn(n, generator=g, device=dev, dtype=torch.float64)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=float, default=1e9)
    ap.add_argument('--nstates', type=int, default=32)
    ap.add_argument('--reps', type=int, default=1)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    N, T = args.nstates, int(args.frames)
    rng = np.random.default_rng(7)
    X = rng.random((N, N)) + 0.2
    X += np.eye(N) * N * 0.5
    A = X / X.sum(axis=1)[:, None]
    pi = np.ones(N) / N
    means, sigmas = np.linspace(-N, N, N) * 1.0, np.full(N, 1.0)
    obs = frames(T, A, means, sigmas, dev)
    batch = TrajectoryBatch.from_concatenated(obs, [T], N, device=dev, viterbi_only=True)
    del obs
    path = batch.viterbi_gaussian(A, pi, means, sigmas)          # warm-up (adapts the warm-up length)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        path = batch.viterbi_gaussian(A, pi, means, sigmas)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.reps
    w = torch.arange(1, 1025, device=dev, dtype=torch.int64)
    checksum = int((path[:T // 1024 * 1024].view(-1, 1024).to(torch.int64) * w).sum().item()) if T >= 1024 else int(path.sum().item())
    print(json.dumps({'workload': 'C5 Viterbi: one trajectory, %d frames, %d states, one GPU' % (T, N),
                      'panel': os.environ.get('BHMM_B200_PANEL', '1'), 'seconds': dt, 'frames_per_s': T / dt,
                      'workspace_GB': batch.workspace_bytes / 1e9, 'info': batch.info(), 'path_checksum': checksum}))
    batch.close()


if __name__ == '__main__':
    main()
