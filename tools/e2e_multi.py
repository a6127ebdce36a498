"""End-to-end fit (MaximumLikelihoodEstimator(host lists).fit(maxit=20), C3 shape per rank) under torchrun for several settings
of the host <-> device mover: worker threads, staging slot size, streaming stores.  Every rank fits its own 1024 x 1e5
trajectories (weak scaling, shard=False, like bench.py's e2e leg); the time of a setting is the maximum over ranks.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/e2e_multi.py
"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as td
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bhmm_b200 import _lib
from bhmm_b200.estimators import MaximumLikelihoodEstimator
from bhmm_b200.hmm import HMM
from bhmm_b200.output_models import GaussianOutputModel
from bhmm_b200.util import testsystems as ts

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    td.init_process_group('nccl', device_id=dev)
N, K, T = 10, int(os.environ.get('E2E_K', 1024)), 100000
pi, A, means, sigmas, O = bench.synth_gaussian_gpu(N, K, T, 3 + rank, dev)
host = O.cpu().numpy()
del O
lst = [host[k] for k in range(K)]
pi0, A0, m0, s0 = ts.perturbed_initial_model(A, means, N)
init = HMM(pi0, A0, GaussianOutputModel(N, means=m0, sigmas=s0))
iters = 20
flag = torch.zeros(1, device=dev)


def barrier():
    if world > 1:
        td.all_reduce(flag)
    torch.cuda.synchronize()


def run(threads, stage_kb, nt):
    os.environ['BHMM_B200_TRANSFER_THREADS'] = str(threads)
    _lib.lib.bhmm_b200_transfer_config(int(stage_kb), int(nt))
    best = None
    for rep in range(2):
        torch.cuda.empty_cache()
        barrier()
        t0 = time.perf_counter()
        est = MaximumLikelihoodEstimator(lst, N, initial_model=init, reversible=False, stationary=False, accuracy=-np.inf,
                                         maxit=iters, shard=False)
        t1 = time.perf_counter()
        est.fit()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        v = torch.tensor([t2 - t0, t1 - t0, est.timings['estep'], est.timings['viterbi']], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(v, op=td.ReduceOp.MAX)
        est._batch.close()
        del est
        v = v.cpu().numpy()
        if best is None or v[0] < best[0]:
            best = v
    if rank == 0:
        print('threads %2d  slot %5d KB  nt %d : fit %.3f s (max over ranks: constructor+upload %.3f, E-steps %.3f, Viterbi+paths %.3f) -> %.2f G frame*iter/s'
              % (threads, stage_kb, nt, best[0], best[1], best[2], best[3], world * K * T * iters / best[0] / 1e9), flush=True)


cores = len(os.sched_getaffinity(0))
per = max(1, cores // max(1, int(os.environ.get('LOCAL_WORLD_SIZE', world))))
if rank == 0:
    print('world %d, %d host cores, %d per rank' % (world, cores, per), flush=True)
default = [(min(8, per), 2048, 1), (min(8, per), 2048, 0), (min(8, per), 512, 1), (min(8, per), 8192, 1),
           (max(1, per // 2), 2048, 1), (min(16, 2 * per), 2048, 1), (min(8, per), 2048, 1)]
if os.environ.get('E2E_CONFIGS'):        # "threads:slot_kb:nt,..."
    default = [tuple(int(x) for x in c.split(':')) for c in os.environ['E2E_CONFIGS'].split(',')]
for threads, stage_kb, nt in default:
    run(threads, stage_kb, nt)
if world > 1:
    td.destroy_process_group()
