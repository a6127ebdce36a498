"""Parity of the batched, device-resident engine (TrajectoryBatch -> bhmm_b200_batch_* C ABI) against the CPU
oracle and the reference fixtures: fused E-step statistics, EM iterations, batched Viterbi and the Gibbs
hidden-path sweep.  float64 within 1e-10 relative, integer outputs bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-10


@pytest.fixture(scope='module')
def eng():
    import torch
    assert torch.cuda.is_available()
    import bhmm_b200.engine as e
    return e


@pytest.fixture(autouse=True, params=['lane', 'team'])
def family(request, monkeypatch):
    """Every engine test runs twice: with the small-N one-thread-per-chain kernels (where N <= 16 allows) and with
    the general-N team kernels forced (BHMM_B200_FAMILY is read when a batch is created)."""
    if request.param == 'team':
        monkeypatch.setenv('BHMM_B200_FAMILY', 'team')
    else:
        monkeypatch.delenv('BHMM_B200_FAMILY', raising=False)
    return request.param


def oracle_stats_gaussian(oracle, obs, A, pi, means, sigmas):
    st = oracle.estep_gaussian(obs, A, pi, means, sigmas)
    wd = sum(g.T.dot(o) for g, o in zip(st['gammas'], obs)) - means * st['wsum']
    wdd = np.zeros(len(means))
    for g, o in zip(st['gammas'], obs):
        d = o[:, None] - means[None, :]
        wdd += (g * d * d).sum(axis=0)
    return st, wd, wdd


def em_obs(g):
    return [g['obs%d' % k] for k in range(len(g['lengths']))]


@pytest.mark.parametrize('chunk,warm', [(0, 0), (200, 150), (64, 4), (5000, 10)])
def test_estep_gaussian_statistics(eng, oracle_port, golden, family, chunk, warm):
    g = golden('em_gauss3')
    obs = em_obs(g)
    A, pi, means, sigmas = g['A0'], g['pi0'], g['means0'], g['sigmas0']
    batch = eng.TrajectoryBatch(obs, 3, chunk=chunk, warm=warm)
    assert batch.uses_lane_kernels == (family == 'lane')
    st = eng.unpack_stats(batch.estep_gaussian(A, pi, means, sigmas).cpu().numpy(), 3)
    ref, wd, wdd = oracle_stats_gaussian(oracle_port, obs, A, pi, means, sigmas)
    assert abs(st['loglik'] - ref['loglik']) <= RTOL * abs(ref['loglik'])
    np.testing.assert_allclose(st['gamma0'], ref['gamma0'], rtol=RTOL)
    np.testing.assert_allclose(st['C'], ref['C'], rtol=RTOL)
    np.testing.assert_allclose(st['wsum'], ref['wsum'], rtol=RTOL)
    np.testing.assert_allclose(st['wd'], wd, rtol=1e-9, atol=1e-9 * np.abs(wd).max())
    np.testing.assert_allclose(st['wdd'], wdd, rtol=RTOL)
    info = batch.info()
    if chunk == 64:
        assert info['chains'] > 4 and (info['fixups_fwd'] + info['fixups_bwd']) >= 1
    # gamma on request
    import torch
    gam = torch.zeros((batch.rows, 3), dtype=torch.float64, device='cuda')
    batch.estep_gaussian(A, pi, means, sigmas, gamma_out=gam)
    gam = gam.cpu().numpy()
    ref_gamma = np.concatenate(ref['gammas'])
    assert np.max(np.abs(gam - ref_gamma)) <= RTOL
    batch.close()


@pytest.mark.parametrize('device_mstep', [True, False])
def test_em_iterations_match_reference_estimator(eng, golden, device_mstep):
    """6 Baum-Welch iterations from the fixture's initial model reproduce MaximumLikelihoodEstimator.fit
    (log-likelihood history, A, pi, means, sigmas within 1e-10; Viterbi paths identical), with the M-step on the GPU
    (engine.mstep_device) and with the host M-step."""
    from bhmm_b200.estimators import MaximumLikelihoodEstimator
    from bhmm_b200.hmm import HMM
    from bhmm_b200.output_models import GaussianOutputModel
    g = golden('em_gauss3')
    obs = em_obs(g)
    init = HMM(g['pi0'], g['A0'], GaussianOutputModel(3, means=g['means0'], sigmas=g['sigmas0']))
    est = MaximumLikelihoodEstimator(obs, 3, initial_model=init, reversible=False, stationary=False,
                                     accuracy=-np.inf, maxit=6, device_mstep=device_mstep)
    model = est.fit()
    assert est.device_msteps == (6 if device_mstep else 0)
    np.testing.assert_allclose(est.likelihoods, g['likelihoods'], rtol=RTOL)
    np.testing.assert_allclose(model.transition_matrix, g['A'], rtol=RTOL)
    np.testing.assert_allclose(model.initial_distribution, g['pi'], rtol=RTOL, atol=1e-300)
    np.testing.assert_allclose(model.output_model.means, g['means'], rtol=RTOL)
    np.testing.assert_allclose(model.output_model.sigmas, g['sigmas'], rtol=RTOL)
    np.testing.assert_allclose(est.count_matrix, g['count_matrix'], rtol=1e-9)
    np.testing.assert_allclose(est.initial_count, g['initial_count'], rtol=1e-9, atol=1e-300)
    for k in range(len(obs)):
        assert np.array_equal(model.hidden_state_trajectories[k], g['viterbi%d' % k])


@pytest.mark.parametrize('device_mstep', [True, False])
def test_em_discrete_matches_reference_estimator(eng, golden, device_mstep):
    from bhmm_b200.estimators import MaximumLikelihoodEstimator
    from bhmm_b200.hmm import HMM
    from bhmm_b200.output_models import DiscreteOutputModel
    g = golden('em_discrete')
    obs = em_obs(g)
    init = HMM(g['pi0'], g['A0'], DiscreteOutputModel(g['B0']))
    est = MaximumLikelihoodEstimator(obs, 4, initial_model=init, reversible=False, stationary=False,
                                     accuracy=-np.inf, maxit=4, output='discrete', device_mstep=device_mstep)
    model = est.fit()
    assert est.device_msteps == (4 if device_mstep else 0)
    np.testing.assert_allclose(est.likelihoods, g['likelihoods'], rtol=RTOL)
    np.testing.assert_allclose(model.transition_matrix, g['A'], rtol=RTOL)
    np.testing.assert_allclose(model.output_model.output_probabilities, g['B'], rtol=1e-9, atol=1e-300)
    for k in range(len(obs)):
        assert np.array_equal(model.hidden_state_trajectories[k], g['viterbi%d' % k])


def test_device_mstep_flags_cases_for_the_host(eng):
    """engine.mstep_device against the host formulas on random statistics; an empty count or a collapsed sigma sets the
    flags word so that the estimator falls back to the general host code (util/tmatrix.estimate_P, gaussian.py:271-272)."""
    import torch
    rng = np.random.default_rng(5)
    for N in (1, 3, 10, 100):
        Cm = rng.random((N, N)) * 50 + 1e-3
        g0, w = rng.random(N) + 0.1, rng.random(N) * 1e3 + 1.0
        wd = (rng.random(N) - 0.5) * w
        wdd = (rng.random(N) + 0.3) * w + wd * wd / w
        mu = np.linspace(-3, 3, N)
        stats = np.concatenate([[-123.5], g0, Cm.ravel(), w, wd, wdd])
        dev = torch.as_tensor(stats).cuda()
        res = eng.unpack_mstep(eng.mstep_device(dev, N, means_old=mu).cpu().numpy(), N)
        assert res['flags'] == 0 and res['loglik'] == -123.5
        np.testing.assert_allclose(res['A'], Cm / Cm.sum(axis=1)[:, None], rtol=1e-14)
        np.testing.assert_allclose(res['pi'], g0 / g0.sum(), rtol=1e-14)
        np.testing.assert_allclose(res['means'], mu + wd / w, rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(res['sigmas'], np.sqrt(wdd / w - (wd / w) ** 2), rtol=1e-12)
        if N >= 3:
            bad = stats.copy()
            bad[1 + N + 1] = 0.0                                      # C[0, 1] = 0: connectivity is the host's business
            assert eng.unpack_mstep(eng.mstep_device(torch.as_tensor(bad).cuda(), N, means_old=mu).cpu().numpy(), N)['flags'] == 1
            bad = stats.copy()
            bad[1 + N + N * N + 2 * N + 2] = 0.0                    # sum gamma d^2 = 0 for state 2: variance clamps to 0
            assert eng.unpack_mstep(eng.mstep_device(torch.as_tensor(bad).cuda(), N, means_old=mu).cpu().numpy(), N)['flags'] >= 1024
        Bn = torch.as_tensor(rng.random((N, 37)) + 1e-3).cuda()
        np.testing.assert_allclose(eng.mstep_discrete_device(Bn).cpu().numpy(),
                                   Bn.cpu().numpy() / Bn.cpu().numpy().sum(axis=1)[:, None], rtol=1e-14)


@pytest.mark.parametrize('N,K,T', [(2, 3, 50), (10, 6, 3000), (20, 3, 900), (32, 4, 1500), (40, 2, 500), (100, 2, 260)])
def test_estep_random_models(eng, oracle_port, N, K, T):
    rng = np.random.default_rng(N * 7 + K)
    X = rng.random((N, N)) ** 2 + 1e-3
    A = X / X.sum(axis=1)[:, None]
    pi = rng.random(N) + 0.01
    pi /= pi.sum()
    means, sigmas = np.linspace(-5, 5, N), np.linspace(0.5, 2.0, N)
    obs = []
    for k in range(K):
        Tk = T - (T // 9) * k
        s = rng.integers(0, N, size=Tk)
        obs.append(means[s] + sigmas[s] * rng.standard_normal(Tk))
    batch = eng.TrajectoryBatch(obs, N, chunk=max(64, T // 7), warm=0)
    st = eng.unpack_stats(batch.estep_gaussian(A, pi, means, sigmas).cpu().numpy(), N)
    ref, wd, wdd = oracle_stats_gaussian(oracle_port, obs, A, pi, means, sigmas)
    assert abs(st['loglik'] - ref['loglik']) <= RTOL * abs(ref['loglik'])
    np.testing.assert_allclose(st['gamma0'], ref['gamma0'], rtol=RTOL, atol=1e-300)
    np.testing.assert_allclose(st['C'], ref['C'], rtol=1e-9, atol=1e-12 * ref['C'].max())
    np.testing.assert_allclose(st['wsum'], ref['wsum'], rtol=RTOL)
    np.testing.assert_allclose(st['wdd'], wdd, rtol=1e-9)
    # batched Viterbi vs the oracle, trajectory by trajectory
    paths = batch.split(batch.viterbi_gaussian(A, pi, means, sigmas).cpu().numpy())
    mism = 0
    for o, p in zip(obs, paths):
        ref_p = oracle_port.viterbi(A, oracle_port.gaussian_p_obs(o, means, sigmas), pi)
        mism += int(np.sum(ref_p != p))
    # emission values come from CUDA's exp (<= 1-2 ulp from glibc's): paths agree unless a comparison is tied to
    # the last bit, which these seeded cases do not hit
    assert mism == 0
    batch.close()


def test_estep_discrete_statistics(eng, oracle_port):
    rng = np.random.default_rng(12)
    N, M, K, T = 7, 40, 5, 2000
    X = rng.random((N, N)) ** 2 + 1e-2
    A = X / X.sum(axis=1)[:, None]
    pi = np.ones(N) / N
    B = rng.random((N, M)) ** 3 + 1e-4
    B /= B.sum(axis=1)[:, None]
    obs = [rng.integers(0, M, size=T - 11 * k).astype(np.int32) for k in range(K)]
    batch = eng.TrajectoryBatch(obs, N, chunk=300, warm=0)
    stats, Bnum = batch.estep_discrete(A, pi, B)
    st = eng.unpack_stats(stats.cpu().numpy(), N)
    ref = oracle_port.estep_discrete(obs, A, pi, B)
    assert abs(st['loglik'] - ref['loglik']) <= RTOL * abs(ref['loglik'])
    np.testing.assert_allclose(st['C'], ref['C'], rtol=1e-9)
    np.testing.assert_allclose(st['gamma0'], ref['gamma0'], rtol=RTOL)
    np.testing.assert_allclose(Bnum.cpu().numpy(), ref['Bnum'], rtol=1e-9)
    paths = batch.split(batch.viterbi_discrete(A, pi, B).cpu().numpy())
    for o, p in zip(obs, paths):
        assert np.array_equal(p, oracle_port.viterbi(A, oracle_port.discrete_p_obs(o, B), pi))   # exact gather
    batch.close()


def test_gibbs_sweep_reproduces_reference_paths(eng, oracle_port, golden):
    """With the reference's uniforms (glibc stream, re-seeded for every trajectory like
    bayesian_sampling.py:288-290 does on the first sweep) the sampled paths and their integer statistics are
    identical to the fixture."""
    import torch
    g = golden('gibbs_gauss3')
    K = len(g['lengths'])
    obs = [g['obs%d' % k] for k in range(K)]
    batch = eng.TrajectoryBatch(obs, 3, chunk=128, warm=0)
    u_rows = np.concatenate([oracle_port.glibc_uniforms(int(g['seed']), int(L))[::-1] for L in g['lengths']])
    u_dev = torch.from_numpy(np.ascontiguousarray(u_rows)).cuda()
    path, counts, sums, ll = batch.gibbs_gaussian(g['A'], g['pi'], g['means'], g['sigmas'], uniforms=u_dev)
    paths = batch.split(path.cpu().numpy())
    for k in range(K):
        assert np.array_equal(paths[k], g['path%d' % k])
    c = batch.unpack_counts(counts)
    assert np.array_equal(c['C'], g['count_matrix'].astype(np.int64))
    assert np.array_equal(c['n0'], g['count_init'])
    sums = sums.cpu().numpy()
    for i in range(3):
        n = int(g['obs_in_state_n%d' % i])
        assert c['count'][i] == n
        np.testing.assert_allclose(sums[i] / n, float(g['obs_in_state_mean%d' % i]), rtol=1e-11)
    ll_ref = sum(oracle_port.forward(g['A'], oracle_port.gaussian_p_obs(o, g['means'], g['sigmas']), g['pi'])[0] for o in obs)
    assert abs(ll - ll_ref) <= RTOL * abs(ll_ref)
    batch.close()


def test_gibbs_philox_is_distributionally_correct(eng, oracle_port):
    """Device Philox draws: the sampled paths follow P(S | O, model): compare state occupancies and transition
    counts, averaged over sweeps, with the smoothed expectations (sum gamma, Baum-Welch C)."""
    rng = np.random.default_rng(21)
    N, K, T = 3, 8, 4000
    A = np.array([[0.95, 0.03, 0.02], [0.04, 0.9, 0.06], [0.02, 0.05, 0.93]])
    pi = np.array([0.3, 0.3, 0.4])
    means, sigmas = np.array([-2.0, 0.0, 2.0]), np.array([1.0, 1.2, 1.0])
    obs = []
    for k in range(K):
        s = np.zeros(T, dtype=int)
        for t in range(1, T):
            s[t] = rng.choice(N, p=A[s[t - 1]])
        obs.append(means[s] + sigmas[s] * rng.standard_normal(T))
    batch = eng.TrajectoryBatch(obs, N)
    ref = oracle_port.estep_gaussian(obs, A, pi, means, sigmas)
    sweeps = 60
    occ = np.zeros(N)
    Cs = np.zeros((N, N))
    seen = set()
    for s in range(sweeps):
        path, counts, sums, ll = batch.gibbs_gaussian(A, pi, means, sigmas, seed=1234, sweep=s)
        c = batch.unpack_counts(counts)
        occ += c['count']
        Cs += c['C']
        seen.add(hash(path[:2000].cpu().numpy().tobytes()))
    assert len(seen) == sweeps                     # different sweeps give different paths
    occ /= sweeps
    Cs /= sweeps
    np.testing.assert_allclose(occ, ref['wsum'], rtol=0.02)
    np.testing.assert_allclose(Cs, ref['C'], rtol=0.08, atol=3.0)
    p2, c2, _, _ = batch.gibbs_gaussian(A, pi, means, sigmas, seed=1234, sweep=7)
    p2 = p2.cpu().numpy().copy()
    p3, _, _, _ = batch.gibbs_gaussian(A, pi, means, sigmas, seed=1234, sweep=7)
    assert np.array_equal(p2, p3.cpu().numpy())    # same (seed, sweep) -> same paths
    batch.close()


def test_gibbs_philox_chi_square_ten_states(eng, oracle_port):
    """N = 10 (the C3 state count, lane sampler with 4-bit hypothesis words): a chi-square test of the sampled state
    occupancies at individual frames against the smoothed marginals gamma.  For a fixed frame t the draws of different
    sweeps are independent samples of P(s_t | O), so sum_i (n_i - S gamma_ti)^2 / (S gamma_ti) over states with an expected
    count >= 5 is chi-square distributed; summed over many frames it is tested at 5 standard deviations."""
    from bhmm_b200.util import testsystems as ts
    N, K, T = 10, 4, 1500
    pi, A, means, sigmas, O, S = ts.gaussian_observations(N, K, T, seed=31)
    means = np.linspace(-3, 3, N)                  # overlapping states: the posterior is not concentrated on one state
    obs = [means[S[k]] + sigmas[S[k]] * np.random.default_rng(5 + k).standard_normal(T) for k in range(K)]
    batch = eng.TrajectoryBatch(obs, N, chunk=300, warm=0)
    ref = oracle_port.estep_gaussian(obs, A, pi, means, sigmas)
    gamma = np.concatenate(ref['gammas'])
    sweeps = 400
    hist = np.zeros((gamma.shape[0], N))
    rows = np.arange(gamma.shape[0])
    for s in range(sweeps):
        path, counts, sums, ll = batch.gibbs_gaussian(A, pi, means, sigmas, seed=77, sweep=s)
        hist[rows, path.cpu().numpy()] += 1.0
    batch.close()
    frames = rows[::7]                             # every 7th frame: neighbouring frames are strongly correlated
    chi2, dof = 0.0, 0
    for t in frames:
        e = sweeps * gamma[t]
        big = e >= 5.0
        if big.sum() < 2:
            continue
        rest_e, rest_o = e[~big].sum(), hist[t][~big].sum()
        ee, oo = np.append(e[big], rest_e), np.append(hist[t][big], rest_o)
        keep = ee > 1e-9
        chi2 += float(((oo[keep] - ee[keep]) ** 2 / ee[keep]).sum())
        dof += int(keep.sum()) - 1
    assert dof > 500
    assert abs(chi2 - dof) < 5.0 * np.sqrt(2.0 * dof), (chi2, dof)


def test_edge_cases_short_ragged_outliers(eng, oracle_port):
    """Trajectories of 1 and 2 frames next to long ones, an observation whose density underflows for every state
    (outlier rule, outputmodel.py:119-131) and a NaN-free far tail (denormal densities): E-step statistics, Viterbi and
    the Gibbs log-likelihood agree with the oracle."""
    rng = np.random.default_rng(99)
    N = 5
    X = rng.random((N, N)) + 0.05
    A = X / X.sum(axis=1)[:, None]
    pi = rng.random(N) + 0.1
    pi /= pi.sum()
    means, sigmas = np.linspace(-4, 4, N), np.linspace(0.3, 1.2, N)
    lengths = [1, 2, 700, 3, 1500, 64, 65]
    obs = []
    for T in lengths:
        s = rng.integers(0, N, size=T)
        obs.append(means[s] + sigmas[s] * rng.standard_normal(T))
    obs[2][100] = 400.0          # all densities underflow to 0 -> row of ones
    obs[4][700] = -17.0          # far tail: densities of ~1e-300 and denormals
    obs[4][701] = 15.5
    batch = eng.TrajectoryBatch(obs, N, chunk=64, warm=16)
    st = eng.unpack_stats(batch.estep_gaussian(A, pi, means, sigmas).cpu().numpy(), N)
    ref, wd, wdd = oracle_stats_gaussian(oracle_port, obs, A, pi, means, sigmas)
    assert abs(st['loglik'] - ref['loglik']) <= RTOL * abs(ref['loglik'])
    np.testing.assert_allclose(st['gamma0'], ref['gamma0'], rtol=RTOL)
    np.testing.assert_allclose(st['C'], ref['C'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st['wsum'], ref['wsum'], rtol=RTOL)
    np.testing.assert_allclose(st['wdd'], wdd, rtol=1e-9)
    assert abs(st['C'].sum() - sum(T - 1 for T in lengths)) < 1e-8
    paths = batch.split(batch.viterbi_gaussian(A, pi, means, sigmas).cpu().numpy())
    for o, p in zip(obs, paths):
        assert np.array_equal(p, oracle_port.viterbi(A, oracle_port.gaussian_p_obs(o, means, sigmas), pi))
    path, counts, sums, ll = batch.gibbs_gaussian(A, pi, means, sigmas, seed=3, sweep=0)
    assert abs(ll - ref['loglik']) <= RTOL * abs(ref['loglik'])
    c = batch.unpack_counts(counts)
    assert c['count'].sum() == sum(lengths) and c['n0'].sum() == len(lengths)
    assert c['C'].sum() == sum(T - 1 for T in lengths)
    # the counts are those of the returned paths
    ref_stats = oracle_port.path_stats(batch.split(path.cpu().numpy()), obs, N)
    assert np.array_equal(c['C'], ref_stats['C']) and np.array_equal(c['n0'], ref_stats['n0'])
    assert np.array_equal(c['count'], ref_stats['count'])
    np.testing.assert_allclose(sums.cpu().numpy()[:N], ref_stats['so'], rtol=1e-12)
    np.testing.assert_allclose(sums.cpu().numpy()[N:], ref_stats['soo'], rtol=1e-12)
    # ignore_outliers=False: the impossible frame makes the likelihood -inf in the reference too
    st2 = eng.unpack_stats(batch.estep_gaussian(A, pi, means, sigmas, ignore_outliers=False).cpu().numpy(), N)
    ref2 = oracle_port.estep_gaussian(obs, A, pi, means, sigmas, ignore_outliers=False)
    assert st2['loglik'] == -np.inf and ref2['loglik'] == -np.inf
    batch.close()


@pytest.mark.parametrize('N', [1, 16])
def test_lane_family_extremes(eng, oracle_port, N):
    rng = np.random.default_rng(N)
    X = rng.random((N, N)) + 0.02
    A = X / X.sum(axis=1)[:, None]
    pi = np.ones(N) / N
    means, sigmas = np.linspace(-3, 3, N) if N > 1 else np.zeros(1), np.linspace(0.5, 1.0, N)
    obs = [rng.normal(size=T) * 2.0 for T in (900, 333)]
    batch = eng.TrajectoryBatch(obs, N, chunk=200, warm=0)
    st = eng.unpack_stats(batch.estep_gaussian(A, pi, means, sigmas).cpu().numpy(), N)
    ref, wd, wdd = oracle_stats_gaussian(oracle_port, obs, A, pi, means, sigmas)
    assert abs(st['loglik'] - ref['loglik']) <= RTOL * abs(ref['loglik'])
    np.testing.assert_allclose(st['C'], ref['C'], rtol=1e-9, atol=1e-12 * ref['C'].max())
    np.testing.assert_allclose(st['wsum'], ref['wsum'], rtol=RTOL)
    np.testing.assert_allclose(st['wdd'], wdd, rtol=1e-9)
    path, counts, sums, ll = batch.gibbs_gaussian(A, pi, means, sigmas, seed=5, sweep=1)
    ref_stats = oracle_port.path_stats(batch.split(path.cpu().numpy()), obs, N)
    assert np.array_equal(batch.unpack_counts(counts)['C'], ref_stats['C'])
    batch.close()


@pytest.mark.parametrize('case', ['sharp', 'broad', 'mixed', 'denormal_band'])
def test_scaling_and_tail_extremes(eng, oracle_port, case):
    """The lane kernels rescale alpha/beta lazily (powers of two, only when the largest component leaves a band) and
    treat far-tail densities specially (exactly zero below exp(-746), library exp in the denormal band).  Long
    trajectories whose likelihood per frame is far from one in either direction, and observations placed so that every
    density of a frame is denormal, must still give the oracle's statistics."""
    rng = np.random.default_rng(7)
    N = 4
    X = rng.random((N, N)) + 0.05
    A = X / X.sum(axis=1)[:, None]
    pi = np.ones(N) / N
    if case == 'sharp':            # densities up to 4e5 per frame: the scaled vectors grow by 2^18 per step
        means, sigmas = np.array([-3e-6, -1e-6, 1e-6, 3e-6]), np.full(N, 1e-6)
    elif case == 'broad':          # densities of 4e-7 per frame: they shrink by 2^-21 per step
        means, sigmas = np.array([-3e6, -1e6, 1e6, 3e6]), np.full(N, 1e6)
    else:
        means, sigmas = np.array([-3.0, -1.0, 1.0, 3.0]), np.array([1e-3, 0.5, 1.0, 20.0])
    obs = []
    for T in (6000, 2500):
        s = rng.integers(0, N, size=T)
        obs.append(means[s] + sigmas[s] * rng.standard_normal(T))
    if case == 'denormal_band':
        # |o - mu| / sigma of 37.8 .. 38.4 for the broadest state and more for the others: every density of the frame
        # is denormal or zero, none of the frames is an "outlier" (the row is not all zero)
        obs[0][1234] = 3.0 + 20.0 * 38.0
        obs[0][1235] = 3.0 - 20.0 * 38.3
        obs[1][77] = 3.0 + 20.0 * 37.7
    batch = eng.TrajectoryBatch(obs, N, chunk=500, warm=64)
    st = eng.unpack_stats(batch.estep_gaussian(A, pi, means, sigmas).cpu().numpy(), N)
    ref, wd, wdd = oracle_stats_gaussian(oracle_port, obs, A, pi, means, sigmas)
    assert np.isfinite(ref['loglik'])
    if case == 'denormal_band':
        # A denormal density carries only a few significant bits and exp() rounds differently into that range on the
        # CPU and on the GPU (both are within one unit of the last -- denormal -- place), so the three frames' scaling
        # factors can only agree to ~1e-4 each; everything that is a RATIO (gamma, xi) is unaffected.  Measured on a B200
        # (round 2): |d loglik| = 4.6e-4, max |d C| = 2.2e-7, max |d sum gamma| = 2e-11; the bounds below leave a factor
        # of ~10-50 and still separate the semantics (a dropped or fatal frame moves the log-likelihood by hundreds).
        print('denormal_band deviations: loglik %.3e, C %.3e, wsum %.3e'
              % (abs(st['loglik'] - ref['loglik']), np.abs(st['C'] - ref['C']).max(), np.abs(st['wsum'] - ref['wsum']).max()))
        assert abs(st['loglik'] - ref['loglik']) < 5e-3
        np.testing.assert_allclose(st['C'], ref['C'], atol=1e-5)
        np.testing.assert_allclose(st['wsum'], ref['wsum'], atol=1e-8)
    else:
        assert abs(st['loglik'] - ref['loglik']) <= RTOL * abs(ref['loglik'])
        np.testing.assert_allclose(st['gamma0'], ref['gamma0'], rtol=1e-9, atol=1e-300)
        np.testing.assert_allclose(st['C'], ref['C'], rtol=1e-9, atol=1e-9 * ref['C'].max())
        np.testing.assert_allclose(st['wsum'], ref['wsum'], rtol=1e-9, atol=1e-9 * ref['wsum'].max())
        np.testing.assert_allclose(st['wdd'], wdd, rtol=1e-8, atol=1e-9 * np.abs(wdd).max())
    assert abs(st['C'].sum() - sum(len(o) - 1 for o in obs)) < 1e-6
    path, counts, sums, ll = batch.gibbs_gaussian(A, pi, means, sigmas, seed=11, sweep=0)
    assert abs(ll - st['loglik']) <= RTOL * abs(ref['loglik'])
    batch.close()


def test_outlier_rule_with_narrow_states(eng, oracle_port):
    """The reference multiplies exp(-z^2/2) by 1/(sigma sqrt(2 pi)) (_gaussian.c:18-20): once the exponential underflows
    (z > 38.6) the density is exactly zero whatever the constant, and a frame whose densities are ALL zero gets the outlier
    rule (outputmodel.py:119-131).  The lane kernels fold the constant into the exponent's argument; for narrow states
    (constant 399 here) that would keep a denormal density up to z = 38.76 and the frame would not be an outlier.  Frames
    placed in that band must be treated like the reference treats them, on both kernel families."""
    import torch
    rng = np.random.default_rng(17)
    N = 4
    X = rng.random((N, N)) + 0.05
    A = X / X.sum(axis=1)[:, None]
    pi = np.ones(N) / N
    means, sigmas = 1e-3 * np.array([-3.0, -1.0, 1.0, 3.0]), np.full(N, 1e-3)
    obs = []
    for T in (3000, 1200):
        s = rng.integers(0, N, size=T)
        obs.append(means[s] + sigmas[s] * rng.standard_normal(T))
    obs[0][100] = 3e-3 + 38.65e-3       # z = 38.65 for the nearest state: exp(-z^2/2) = 0, exp(log 399 - z^2/2) is denormal
    obs[0][2500] = -3e-3 - 38.72e-3
    obs[1][50] = 3e-3 + 38.70e-3
    for k, t in ((0, 100), (0, 2500), (1, 50)):
        assert not oracle_port.gaussian_p_obs(obs[k][t:t + 1], means, sigmas, ignore_outliers=False).any()     # the reference sees an all-zero row
    batch = eng.TrajectoryBatch(obs, N, chunk=400, warm=64)
    gam = torch.zeros((batch.rows, N), dtype=torch.float64, device='cuda')
    st = eng.unpack_stats(batch.estep_gaussian(A, pi, means, sigmas, gamma_out=gam).cpu().numpy(), N)
    ref, wd, wdd = oracle_stats_gaussian(oracle_port, obs, A, pi, means, sigmas)
    assert np.isfinite(ref['loglik'])
    assert abs(st['loglik'] - ref['loglik']) <= RTOL * abs(ref['loglik'])
    np.testing.assert_allclose(st['C'], ref['C'], rtol=1e-9, atol=1e-9 * ref['C'].max())
    np.testing.assert_allclose(st['wsum'], ref['wsum'], rtol=1e-9)
    assert np.max(np.abs(gam.cpu().numpy() - np.concatenate(ref['gammas']))) <= RTOL
    batch.close()


def test_degenerate_sigma_runs_on_the_team_family(eng, oracle_port, family):
    """sigma below 1e-100 could overflow the lane kernels' lazily scaled band.  The reference has no such limit
    (_gaussian.c:18-20), so the batch switches to the team kernels for that call instead of refusing it."""
    rng = np.random.default_rng(3)
    obs = [np.where(rng.random(60) < 0.5, 0.0, rng.standard_normal(60)), rng.standard_normal(45)]
    A, pi = np.array([[0.6, 0.4], [0.3, 0.7]]), np.array([0.5, 0.5])
    means, sigmas = np.zeros(2), np.array([1e-120, 1.0])
    batch = eng.TrajectoryBatch(obs, 2)
    st = eng.unpack_stats(batch.estep_gaussian(A, pi, means, sigmas).cpu().numpy(), 2)
    ref = oracle_port.estep_gaussian(obs, A, pi, means, sigmas)
    assert not batch.uses_lane_kernels
    assert abs(st['loglik'] - ref['loglik']) <= RTOL * abs(ref['loglik'])
    np.testing.assert_allclose(st['C'], ref['C'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st['wsum'], ref['wsum'], rtol=RTOL, atol=1e-12)
    # and back on the lane kernels (when they are the batch's family) with an ordinary model
    st2 = eng.unpack_stats(batch.estep_gaussian(A, pi, means, np.array([0.8, 1.0])).cpu().numpy(), 2)
    ref2 = oracle_port.estep_gaussian(obs, A, pi, means, np.array([0.8, 1.0]))
    assert batch.uses_lane_kernels == (family == 'lane')
    assert abs(st2['loglik'] - ref2['loglik']) <= RTOL * abs(ref2['loglik'])
    batch.close()


def test_sub_batched_trajectories_match_single_batch(eng, oracle_port):
    """Groups of trajectories run one after the other on a shared, budgeted workspace: statistics add up to those of
    the single batch (and the oracle), Viterbi paths and Gibbs paths for given uniforms are identical."""
    import torch
    rng = np.random.default_rng(21)
    N = 6
    X = rng.random((N, N)) + 0.05
    A = X / X.sum(axis=1)[:, None]
    pi = np.ones(N) / N
    means, sigmas = np.linspace(-5, 5, N), np.linspace(0.6, 1.4, N)
    lengths = [15000, 400, 22000, 9000, 1, 31000, 7000, 13000]
    obs = []
    for T in lengths:
        s = rng.integers(0, N, size=T)
        obs.append(means[s] + sigmas[s] * rng.standard_normal(T))
    one = eng.TrajectoryBatch(obs, N, chunk=256, warm=64)
    budget = int(0.45 * one.workspace_bytes)
    many = eng.SubBatchedTrajectories(obs, N, budget, chunk=256, warm=64)
    assert len(many.groups) >= 3 and many.workspace_bytes <= budget
    assert many.groups[0][0] == 0 and many.groups[-1][1] == len(lengths)
    assert all(a[1] == b[0] for a, b in zip(many.groups, many.groups[1:]))
    s1 = one.estep_gaussian(A, pi, means, sigmas).cpu().numpy()
    s2 = many.estep_gaussian(A, pi, means, sigmas).cpu().numpy()
    np.testing.assert_allclose(s2, s1, rtol=1e-11, atol=1e-11)
    ref, wd, wdd = oracle_stats_gaussian(oracle_port, obs, A, pi, means, sigmas)
    st = eng.unpack_stats(s2, N)
    assert abs(st['loglik'] - ref['loglik']) <= RTOL * abs(ref['loglik'])
    np.testing.assert_allclose(st['C'], ref['C'], rtol=1e-9, atol=1e-12)
    # a second E-step re-borrows the workspace and gives the same answer
    np.testing.assert_allclose(many.estep_gaussian(A, pi, means, sigmas).cpu().numpy(), s2, rtol=1e-12, atol=1e-12)
    p1 = one.viterbi_gaussian(A, pi, means, sigmas).cpu().numpy()
    p2 = many.viterbi_gaussian(A, pi, means, sigmas).cpu().numpy()
    assert np.array_equal(p1, p2)
    u = torch.rand(sum(lengths), dtype=torch.float64, device='cuda')
    g1 = one.gibbs_gaussian(A, pi, means, sigmas, uniforms=u)
    g2 = many.gibbs_gaussian(A, pi, means, sigmas, uniforms=u)
    assert np.array_equal(g1[0].cpu().numpy(), g2[0].cpu().numpy())
    assert np.array_equal(g1[1].cpu().numpy(), g2[1].cpu().numpy())
    np.testing.assert_allclose(g2[2].cpu().numpy(), g1[2].cpu().numpy(), rtol=1e-12)
    assert abs(g1[3] - g2[3]) <= 1e-11 * abs(g1[3])
    # make_batch picks the plain batch when everything fits
    assert isinstance(eng.make_batch(obs, N), eng.TrajectoryBatch)
    assert isinstance(eng.make_batch(obs, N, chunk=256, warm=64, max_workspace_bytes=budget), eng.SubBatchedTrajectories)
    with pytest.raises(MemoryError):
        eng.SubBatchedTrajectories(obs, N, 1000)
    one.close()
    many.close()


@pytest.mark.parametrize('N,world', [(3, 3), (10, 4), (20, 2)])
def test_time_sharded_trajectories(eng, oracle_port, N, world):
    """Trajectories cut in TIME across shards (C5: one trajectory too long for one GPU): every shard owns a frame range
    plus a halo, the statistics of the shards add up to those of the whole trajectories, and the hand-overs at the shard
    borders are certified against the neighbour's value.  All shards run in this one process here."""
    rng = np.random.default_rng(100 + N)
    X = rng.random((N, N)) + 0.3 / N
    A = X / X.sum(axis=1)[:, None]
    pi = np.ones(N) / N
    means, sigmas = np.linspace(-N, N, N), np.linspace(0.7, 1.3, N)
    obs = []
    for T in (30000, 8000, 5):
        s = rng.integers(0, N, size=T)
        obs.append(means[s] + sigmas[s] * rng.standard_normal(T))
    shards = [eng.TimeShardedTrajectories(obs, N, r, world, halo=3000) for r in range(world)]
    assert [g[:2] for g in shards[1].global_ranges][:2] == [(30000 // world, 2 * 30000 // world), (8000 // world, 2 * 8000 // world)]
    local = [s.estep_gaussian_local(A, pi, means, sigmas) for s in shards]
    stats, worst = eng.TimeShardedTrajectories.combine(shards, local)
    assert worst <= 1e-11
    st = eng.unpack_stats(stats.cpu().numpy(), N)
    ref, wd, wdd = oracle_stats_gaussian(oracle_port, obs, A, pi, means, sigmas)
    assert abs(st['loglik'] - ref['loglik']) <= RTOL * abs(ref['loglik'])
    np.testing.assert_allclose(st['gamma0'], ref['gamma0'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st['C'], ref['C'], rtol=1e-9, atol=1e-9 * ref['C'].max())
    np.testing.assert_allclose(st['wsum'], ref['wsum'], rtol=RTOL)
    np.testing.assert_allclose(st['wdd'], wdd, rtol=1e-9)
    assert abs(st['C'].sum() - sum(len(o) - 1 for o in obs)) < 1e-6
    # a shard's E-step batch (frames after its owned range) is not a Viterbi shard
    with pytest.raises(Exception):
        shards[0].batch.viterbi_gaussian(A, pi, means, sigmas)
    # Viterbi ACROSS the shards (N <= 32): chain-parallel maps per shard with certified borders, paths resolved from the last
    # shard to the first with one state per trajectory handed to the left; equal to the sequential paths of the oracle
    if N <= 32 and os.environ.get('BHMM_B200_PANEL', '1') != '0':
        paths, vworst = eng.TimeShardedTrajectories.viterbi_combine(shards, (A, pi, means, sigmas, True))
        assert vworst <= 1e-11
        for o, p in zip(obs, paths):
            want = oracle_port.viterbi(A, oracle_port.gaussian_p_obs(o, means, sigmas), pi)
            assert p.shape == want.shape and np.array_equal(p, want), int(np.sum(p != want))
    for s in shards:
        s.close()
    # a halo that is far too short is detected, not silently accepted
    short = [eng.TimeShardedTrajectories(obs, N, r, world, halo=2, warm=2, chunk=4000) for r in range(world)]
    with pytest.raises(RuntimeError):
        eng.TimeShardedTrajectories.combine(short, [s.estep_gaussian_local(A, pi, means, sigmas) for s in short])
    for s in short:
        s.close()


@pytest.mark.parametrize('N', [4, 12, 20, 32, 40, 100])
def test_exact_scan_for_models_that_do_not_forget(eng, oracle_port, N):
    """North star (4): a nearly reducible transition matrix with uninformative emissions never forgets its start, so the
    warm-up starts of the chains fail their certification and chain-by-chain repairs would walk the trajectory
    sequentially.  The engine falls back to the exact time-parallel start -- transfer operators of all chains and a scan
    over them (scan_kernels.cu) -- in both directions, and the statistics still match the oracle."""
    rng = np.random.default_rng(900 + N)
    eps = 2e-5
    A = (1.0 - eps) * np.eye(N) + eps * np.ones((N, N)) / N
    A /= A.sum(axis=1)[:, None]
    pi = rng.random(N) + 0.5
    pi /= pi.sum()
    means, sigmas = np.linspace(-0.02, 0.02, N), np.ones(N)      # emissions that tell the states apart very slowly
    obs = [rng.standard_normal(T) for T in (9000, 4000, 300)]
    batch = eng.TrajectoryBatch(obs, N, chunk=400, warm=32)
    st = eng.unpack_stats(batch.estep_gaussian(A, pi, means, sigmas).cpu().numpy(), N)
    info = batch.info()
    assert info['chains'] > 20
    assert batch.exact_scans >= 2, (batch.exact_scans, info)    # forward and backward fell back to the scan
    assert info['fixups_fwd'] <= 4 and info['fixups_bwd'] <= 4, info   # ... instead of one repair sweep per chain
    ref, wd, wdd = oracle_stats_gaussian(oracle_port, obs, A, pi, means, sigmas)
    assert abs(st['loglik'] - ref['loglik']) <= RTOL * abs(ref['loglik'])
    np.testing.assert_allclose(st['gamma0'], ref['gamma0'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st['C'], ref['C'], rtol=1e-8, atol=1e-9 * ref['C'].max())
    np.testing.assert_allclose(st['wsum'], ref['wsum'], rtol=1e-9)
    # a model that mixes keeps the certified warm-up starts
    A2 = rng.random((N, N)) + 0.1
    A2 /= A2.sum(axis=1)[:, None]
    m2 = np.linspace(-3, 3, N)
    o2 = [m2[rng.integers(0, N, size=T)] + rng.standard_normal(T) for T in (9000, 4000)]
    b2 = eng.TrajectoryBatch(o2, N, chunk=400, warm=200)
    b2.estep_gaussian(A2, pi, m2, sigmas)
    assert b2.exact_scans == 0
    batch.close()
    b2.close()


def test_viterbi_only_batch(eng, oracle_port):
    """A Viterbi-only batch (no forward-variable workspace: how one C5-sized trajectory fits a GPU for Viterbi) returns the
    same paths as a full batch and refuses the E-step."""
    rng = np.random.default_rng(77)
    N = 5
    X = rng.random((N, N)) + 2.0 * np.eye(N)
    A = X / X.sum(axis=1)[:, None]
    pi = np.ones(N) / N
    means, sigmas = np.linspace(-4, 4, N), np.linspace(0.6, 1.5, N)
    obs = []
    for T in (3000, 1200):
        s = rng.integers(0, N, size=T)
        obs.append(means[s] + sigmas[s] * rng.standard_normal(T))
    full = eng.TrajectoryBatch(obs, N, chunk=400, warm=0)
    lean = eng.TrajectoryBatch(obs, N, chunk=400, warm=0, viterbi_only=True)
    assert lean.workspace_bytes < full.workspace_bytes
    p_full = full.viterbi_gaussian(A, pi, means, sigmas).cpu().numpy()
    p_lean = lean.viterbi_gaussian(A, pi, means, sigmas).cpu().numpy()
    assert np.array_equal(p_full, p_lean)
    for o, p in zip(obs, lean.split(p_lean)):
        assert np.array_equal(p, oracle_port.viterbi(A, oracle_port.gaussian_p_obs(o, means, sigmas), pi))
    with pytest.raises(NotImplementedError):
        lean.estep_gaussian(A, pi, means, sigmas)
    full.close()
    lean.close()
