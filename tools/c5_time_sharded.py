"""C5 shape: ONE long trajectory, 32 states, cut in time across the GPUs of one box.

    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/c5_time_sharded.py --frames 4e8

Every rank regenerates the frames it needs (owned range + halo) from a block-seeded generator, so neighbouring ranks
see identical observations in their overlap; runs `--steps` E-steps (forward, backward + statistics, all-reduce of the
statistics, all-gather and certification of the border hand-overs) and prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhmm_b200.engine import TrajectoryBatch, TimeShardedTrajectories, unpack_stats  # noqa: E402

BLOCK = 1 << 20


def frames(lo, hi, means, sigmas, dev, seed=1234):
    """Observations of frames [lo, hi) of the (conceptually infinite) trajectory: block b = frames [b*BLOCK, (b+1)*BLOCK)
    is drawn with its own seed, state sequence i.i.d. uniform (throughput does not depend on the hidden dynamics)."""
    out = []
    mu = torch.as_tensor(means, device=dev)
    sg = torch.as_tensor(sigmas, device=dev)
    for b in range(lo // BLOCK, (hi - 1) // BLOCK + 1):
        g = torch.Generator(device=dev)
        g.manual_seed(seed + b)
        s = torch.randint(0, len(means), (BLOCK,), generator=g, device=dev)
        x = mu[s] + sg[s] * torch.randn(BLOCK, generator=g, device=dev, dtype=torch.float64)
        a, e = max(lo, b * BLOCK), min(hi, (b + 1) * BLOCK)
        out.append(x[a - b * BLOCK:e - b * BLOCK])
    return torch.cat(out)


MBLOCK = 100000


def frames_markov(lo, hi, pi, A, means, sigmas, dev, seed=4321):
    """Observations of frames [lo, hi) of one long trajectory drawn FROM THE MODEL ("generated on device in
    blocks each started from pi"): block b = frames [b*MBLOCK, (b+1)*MBLOCK) is a Markov chain started from pi with its own
    seed, so that neighbouring ranks see identical observations in their overlap; the chains of all blocks of the range
    advance together (one vectorised inverse-CDF draw per frame)."""
    import numpy as np
    b0, b1 = lo // MBLOCK, (hi - 1) // MBLOCK
    K = b1 - b0 + 1
    N = len(pi)
    cumA = torch.as_tensor(np.cumsum(A, axis=1), device=dev)
    cumA[:, -1] = 1.0
    cpi = torch.as_tensor(np.cumsum(pi), device=dev)
    cpi[-1] = 1.0
    U = torch.empty((MBLOCK, K), dtype=torch.float64, device=dev)
    Z = torch.empty((K, MBLOCK), dtype=torch.float64, device=dev)
    g = torch.Generator(device=dev)
    for k in range(K):
        g.manual_seed(seed + b0 + k)
        U[:, k] = torch.rand(MBLOCK, generator=g, device=dev, dtype=torch.float64)
        Z[k] = torch.randn(MBLOCK, generator=g, device=dev, dtype=torch.float64)
    S = torch.empty((MBLOCK, K), dtype=torch.int64, device=dev)
    s = (U[0][:, None] > cpi[None, :]).sum(dim=1).clamp_(max=N - 1)
    S[0] = s
    for t in range(1, MBLOCK):
        s = (U[t][:, None] > cumA[s]).sum(dim=1).clamp_(max=N - 1)
        S[t] = s
    del U
    S = S.t().contiguous()
    x = (torch.as_tensor(means, device=dev)[S] + torch.as_tensor(sigmas, device=dev)[S] * Z).reshape(-1)
    return x[lo - b0 * MBLOCK:hi - b0 * MBLOCK].contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=float, default=1e8)
    ap.add_argument('--nstates', type=int, default=32)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--halo', type=int, default=0)
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        td.init_process_group('nccl', device_id=dev)
    N, T = args.nstates, int(args.frames)
    rng = np.random.default_rng(7)
    X = rng.random((N, N)) + 0.2
    X += np.eye(N) * N * 0.5
    A = X / X.sum(axis=1)[:, None]
    pi = np.ones(N) / N
    means, sigmas = np.linspace(-N, N, N) * 1.0, np.full(N, 1.0)
    halo = args.halo if args.halo > 0 else 8 * max(128, 48 * N)
    lo, hi = (T * rank) // world, (T * (rank + 1)) // world
    a, b = max(0, lo - halo), min(T, hi + halo)
    piece = frames(a, b, means, sigmas, dev)
    batch = TrajectoryBatch.from_concatenated(piece, [b - a], N, device=dev, own_ranges=[(lo - a, hi - a)])
    del piece
    ranges = [[((T * r) // world, (T * (r + 1)) // world, T)] for r in range(world)]

    def step():
        stats = batch.estep_gaussian(A, pi, means, sigmas).clone()
        borders = torch.from_numpy(np.array([batch.border_handovers(0)])).to(dev)
        if world > 1:
            td.all_reduce(stats)
            gathered = [torch.empty_like(borders) for _ in range(world)]
            td.all_gather(gathered, borders)
        else:
            gathered = [borders]
        worst = TimeShardedTrajectories.certify([g.cpu().numpy() for g in gathered], ranges, 1e-11)
        return stats, worst

    for _ in range(args.warmup):
        stats, worst = step()
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        stats, worst = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    st = unpack_stats(stats.cpu().numpy(), N)
    if rank == 0:
        print(json.dumps({'workload': 'C5: one trajectory of %d frames, %d states, time-sharded' % (T, N), 'n_gpus': world,
                          'ms_per_estep': ms, 'value': T / (ms * 1e-3), 'unit': 'frames*iters/s (E-step only)',
                          'halo': halo, 'worst_border_mismatch': worst, 'loglik': st['loglik'],
                          'transitions_counted': float(st['C'].sum()), 'info': batch.info(),
                          'lane_kernels': bool(batch.uses_lane_kernels)}), flush=True)
    if world > 1:
        td.destroy_process_group()


if __name__ == '__main__':
    main()
