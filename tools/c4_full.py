"""C4 shape at full size on ONE GPU: 100-state discrete HMM (1000 symbols), 4096 trajectories x 1e5 frames, one EM E-step
and one Viterbi pass.  The forward variables (328 GB) do not fit, so the trajectories run in groups on a shared workspace
(engine.SubBatchedTrajectories).  Prints one JSON line.

    python tools/c4_full.py [--trajectories 4096] [--budget-gb 140]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhmm_b200.engine import SubBatchedTrajectories, unpack_stats  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--trajectories', type=int, default=4096)
    ap.add_argument('--frames', type=int, default=100000)
    ap.add_argument('--budget-gb', type=float, default=140.0)
    args = ap.parse_args()
    N, M, K, T = 100, 1000, args.trajectories, args.frames
    dev = torch.device('cuda', 0)
    rng = np.random.default_rng(4)
    X = rng.random((N, N)) + 0.05
    X += np.eye(N) * N * 0.4
    A = X / X.sum(axis=1)[:, None]
    pi = np.ones(N) / N
    B = np.full((N, M), 0.2 / M)
    for i in range(N):
        B[i, 10 * i:10 * i + 10] += 0.08
    B /= B.sum(axis=1)[:, None]
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    rows = K * T
    # metastable-looking symbol stream: the state changes every ~50 frames, symbols from the state's own block
    s = torch.randint(0, N, (rows // 50 + 1,), generator=g, device=dev).repeat_interleave(50)[:rows]
    sym = (10 * s + torch.randint(0, 10, (rows,), generator=g, device=dev)).to(torch.int32)
    del s
    batch = SubBatchedTrajectories.from_concatenated(sym, [T] * K, N, int(args.budget_gb * 1e9), device=dev)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    e[0].record()
    stats, Bnum = batch.estep_discrete(A, pi, B)
    e[1].record()
    path = batch.viterbi_discrete(A, pi, B)
    e[2].record()
    torch.cuda.synchronize()
    st = unpack_stats(stats.cpu().numpy(), N)
    ms_e, ms_v = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    agree = float((path.to(torch.int64) == (sym // 10).to(torch.int64)).double().mean().item())
    print(json.dumps({'workload': 'C4: %d-state discrete HMM, %d symbols, %d trajectories x %d frames, one GPU' % (N, M, K, T),
                      'groups': len(batch.groups), 'workspace_gb': batch.workspace_bytes / 1e9,
                      'estep_ms': ms_e, 'estep_frames_per_s': rows / (ms_e * 1e-3),
                      'viterbi_ms': ms_v, 'viterbi_frames_per_s': rows / (ms_v * 1e-3),
                      'loglik': st['loglik'], 'transitions_counted': float(st['C'].sum()),
                      'Bnum_total': float(Bnum.sum().item()), 'viterbi_state_agreement': agree, 'info': batch.info()}), flush=True)
    batch.close()


if __name__ == '__main__':
    main()
