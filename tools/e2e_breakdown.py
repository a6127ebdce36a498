"""Where the end-to-end fit of bench.py's C3 line spends its time (one GPU)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhmm_b200.engine import TrajectoryBatch, make_batch, download_array
from bhmm_b200.util import testsystems as ts
from bhmm_b200.estimators import MaximumLikelihoodEstimator
from bhmm_b200.hmm import HMM
from bhmm_b200.output_models import GaussianOutputModel
import bench

N, K, T = 10, 1024, 100000
dev = torch.device('cuda', 0)
pi, A, means, sigmas, O = bench.synth_gaussian_gpu(N, K, T, 3, dev)
host = O.cpu().numpy()
del O
lst = [host[k] for k in range(K)]
pi0, A0, m0, s0 = ts.perturbed_initial_model(A, means, N)
def sync():
    torch.cuda.synchronize(); return time.perf_counter()
from bhmm_b200.engine import SubBatchedTrajectories, upload_arrays
for rep in range(2):
    # the pieces of make_batch, one by one
    torch.cuda.empty_cache()
    a0 = sync()
    groups = SubBatchedTrajectories.plan_groups([T] * K, N, int(0.8 * torch.cuda.mem_get_info()[0]))
    a1 = sync()
    cat = torch.empty(K * T, dtype=torch.float64, device=dev)
    a2 = sync()
    upload_arrays(cat, lst)
    a3 = sync()
    bb = TrajectoryBatch.from_concatenated(cat, [T] * K, N)
    a4 = sync()
    print('rep %d pieces: plan_groups %.1f ms, torch.empty(obs) %.1f ms, upload %.1f ms, from_concatenated (clone + create + workspace) %.1f ms'
          % (rep, 1e3 * (a1 - a0), 1e3 * (a2 - a1), 1e3 * (a3 - a2), 1e3 * (a4 - a3)), flush=True)
    bb.close(); del bb, cat
    torch.cuda.empty_cache()
    t0 = sync()
    b = make_batch(lst, N)
    t1 = sync()
    st = b.estep_gaussian(A0, pi0, m0, s0)
    t2 = sync()
    p = b.viterbi_gaussian(A0, pi0, m0, s0)
    t3 = sync()
    flat = download_array(p)
    t4 = sync()
    paths = list(b.split(flat))
    t5 = sync()
    its = []
    for _ in range(6):
        i0 = sync(); b.estep_gaussian(A0, pi0, m0, s0); its.append(1e3 * (sync() - i0))
    print('rep %d: E-steps 2..7 of the fresh batch: %s ms' % (rep, ' '.join('%.2f' % x for x in its)), flush=True)
    print('rep %d: upload/make_batch %.3f s, first E-step %.3f s, viterbi %.3f s, D2H paths %.3f s, split+copy %.3f s; info %s'
          % (rep, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, b.info()), flush=True)
    b.close(); del b
