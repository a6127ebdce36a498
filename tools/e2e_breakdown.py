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
for rep in range(2):
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
    print('rep %d: upload/make_batch %.3f s, first E-step %.3f s, viterbi %.3f s, D2H paths %.3f s, split+copy %.3f s; info %s'
          % (rep, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, b.info()), flush=True)
    b.close(); del b
