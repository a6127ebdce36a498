"""World-size-2 test of the multi-GPU host logic on CPU (gloo): trajectories are sharded, each rank reduces the
statistics of its shard (here with the CPU oracle standing in for the GPU E-step, test infrastructure only), one
all-reduce of the packed statistics, and every rank derives the same M-step.  Mirrors what
MaximumLikelihoodEstimator does with NCCL on the B200s."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as td
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    td.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from bhmm_b200 import dist
        from bhmm_b200.engine import unpack_stats
        from bhmm_b200.util import testsystems as ts
        from bhmm_b200.util import tmatrix
        from oracle.oracle import Oracle
        N = 3
        pi, A, means, sigmas, O, S = ts.gaussian_observations(N, 7, 400, seed=9)
        obs = [O[k][:400 - 30 * k] for k in range(7)]
        pi0, A0, m0, s0 = ts.perturbed_initial_model(A, means, N)
        lo, hi = dist.shard_bounds([len(o) for o in obs], dist.rank(), dist.world_size())
        assert dist.world_size() == world and dist.rank() == rank
        orc = Oracle('port')
        st = orc.estep_gaussian(obs[lo:hi], A0, pi0, m0, s0) if hi > lo else None
        packed = np.zeros(1 + N + N * N + 3 * N)
        if st is not None:
            wd = sum(g.T.dot(o) for g, o in zip(st['gammas'], obs[lo:hi])) - m0 * st['wsum']
            wdd = sum((g * (o[:, None] - m0) ** 2).sum(axis=0) for g, o in zip(st['gammas'], obs[lo:hi]))
            packed = np.concatenate([[st['loglik']], st['gamma0'], st['C'].ravel(), st['wsum'], wd, wdd])
        t = torch.from_numpy(packed.copy())
        dist.allreduce_sum(t)
        tot = unpack_stats(t.numpy(), N)
        Anew = tmatrix.estimate_P(tot['C'], reversible=False, mincount_connectivity=1e-16)
        out[rank] = dict(lo=lo, hi=hi, loglik=tot['loglik'], A=Anew, C=tot['C'], wsum=tot['wsum'])
    finally:
        td.destroy_process_group()


def test_sharded_estep_allreduce_world2():
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0['lo'] == 0 and r0['hi'] == r1['lo'] and r1['hi'] == 7
    # both ranks hold identical reduced statistics and M-step
    assert r0['loglik'] == r1['loglik']
    assert np.array_equal(r0['A'], r1['A'])
    # ... equal to the single-process result over all trajectories
    from bhmm_b200.util import testsystems as ts
    from oracle.oracle import Oracle
    pi, A, means, sigmas, O, S = ts.gaussian_observations(3, 7, 400, seed=9)
    obs = [O[k][:400 - 30 * k] for k in range(7)]
    pi0, A0, m0, s0 = ts.perturbed_initial_model(A, means, 3)
    ref = Oracle('port').estep_gaussian(obs, A0, pi0, m0, s0)
    assert abs(r0['loglik'] - ref['loglik']) <= 1e-12 * abs(ref['loglik'])
    np.testing.assert_allclose(r0['C'], ref['C'], rtol=1e-12)
    np.testing.assert_allclose(r0['wsum'], ref['wsum'], rtol=1e-12)


def _worker_param_draws(rank, world, port, out):
    """Ranks whose global numpy RNGs are in DIFFERENT states must still draw identical Gibbs parameters."""
    import torch.distributed as td
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    td.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from bhmm_b200 import dist
        from bhmm_b200.estimators.bayesian_sampling import BayesianHMMSampler
        from bhmm_b200.hmm.generic_hmm import HMM
        from bhmm_b200.output_models.gaussian import GaussianOutputModel
        np.random.seed(1000 + 17 * rank)                    # deliberately different per rank
        seed = dist.broadcast_int(np.random.randint(0, 2 ** 31 - 1))
        A = np.array([[0.9, 0.1], [0.2, 0.8]])
        pi = np.array([2.0 / 3, 1.0 / 3])
        st = {'C': np.array([[900, 100], [110, 390]], dtype=np.int64), 'n0': np.array([2, 1], dtype=np.int64),
              'count': np.array([1000.0, 500.0]), 'so': np.array([-1000.0, 750.0]), 'soo': np.array([2000.0, 1700.0])}
        draws = []
        for reversible in (False, True):
            s = BayesianHMMSampler.__new__(BayesianHMMSampler)
            s.reversible, s.stationary, s.nstates, s._output = reversible, False, 2, 'gaussian'
            s._rng = np.random.RandomState(seed)            # what __init__ sets up when the batch is sharded
            s._np_rng = np.random.default_rng(s._rng.randint(0, 2 ** 31 - 1))
            s.prior_C, s.prior_n0, s.transition_matrix_sampling_steps = A.copy(), pi.copy(), 1000
            s.model = HMM(pi, A, GaussianOutputModel(2, means=[-1.0, 1.5], sigmas=[1.0, 1.0]))
            s._updateEmissionProbabilities(st)
            s._updateTransitionMatrix(st)
            draws.append(np.concatenate([s.model.transition_matrix.ravel(), s.model.initial_distribution,
                                         s.model.output_model.means, s.model.output_model.sigmas]))
        out[rank] = dict(seed=seed, draws=np.concatenate(draws))
    finally:
        td.destroy_process_group()


def test_gibbs_parameter_draws_agree_across_ranks_world2():
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_param_draws, args=(world, port, out), nprocs=world, join=True)
    assert out[0]['seed'] == out[1]['seed']
    assert np.array_equal(out[0]['draws'], out[1]['draws'])
    assert np.all(np.isfinite(out[0]['draws']))
