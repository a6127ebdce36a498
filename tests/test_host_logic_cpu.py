"""Host-side logic that needs no GPU: M-step from sufficient statistics, transition-matrix helpers, trajectory
sharding, synthetic test systems."""
import numpy as np
import pytest

from oracle import oracle as orc


def test_gaussian_mstep_from_shifted_moments_equals_reference_two_pass(oracle_port):
    """GaussianOutputModel.estimate_from_statistics (shifted one-pass moments) == gaussian.py:214-272 (two passes)."""
    from bhmm_b200.output_models import GaussianOutputModel
    rng = np.random.default_rng(0)
    N = 4
    means0, sig0 = np.array([-3.0, -0.5, 0.7, 4.0]), np.array([0.6, 1.0, 0.8, 1.5])
    obs = [rng.normal(size=T) * 3.0 for T in (500, 333)]
    gammas = [rng.dirichlet(np.ones(N), size=len(o)) for o in obs]
    ref_m, ref_s = orc.mstep_gaussian(obs, gammas)
    wsum = sum(g.sum(axis=0) for g in gammas)
    wd = sum((g * (o[:, None] - means0)).sum(axis=0) for g, o in zip(gammas, obs))
    wdd = sum((g * (o[:, None] - means0) ** 2).sum(axis=0) for g, o in zip(gammas, obs))
    om = GaussianOutputModel(N, means=means0, sigmas=sig0)
    om.estimate_from_statistics(wsum, wd, wdd)
    np.testing.assert_allclose(om.means, ref_m, rtol=1e-12)
    np.testing.assert_allclose(om.sigmas, ref_s, rtol=1e-11)


def test_discrete_mstep_and_estimate_P():
    from bhmm_b200.output_models import DiscreteOutputModel
    from bhmm_b200.util import tmatrix
    Bnum = np.array([[1.0, 3.0, 0.0], [2.0, 2.0, 4.0]])
    om = DiscreteOutputModel(np.full((2, 3), 1.0 / 3))
    om.estimate_from_statistics(Bnum)
    np.testing.assert_allclose(om.output_probabilities, orc.mstep_discrete(Bnum))
    C = np.array([[5.0, 1.0, 0.0], [2.0, 6.0, 0.0], [0.0, 0.0, 0.0]])
    P = tmatrix.estimate_P(C, reversible=False, mincount_connectivity=1e-16)
    np.testing.assert_allclose(P[:2, :2], C[:2, :2] / C[:2, :2].sum(axis=1)[:, None])
    assert P[2, 2] == 1.0 and tmatrix.is_transition_matrix(P)
    assert not tmatrix.is_connected(C)
    # reversible estimator (parity-unpinned): check its defining properties
    Cr = np.array([[10.0, 2.0, 1.0], [3.0, 20.0, 4.0], [0.5, 5.0, 8.0]])
    Pr = tmatrix.estimate_P(Cr, reversible=True, maxerr=1e-14)
    assert tmatrix.is_transition_matrix(Pr) and tmatrix.is_reversible(Pr)
    pi = tmatrix.stationary_distribution(Pr)
    np.testing.assert_allclose(pi[:, None] * Pr, (pi[:, None] * Pr).T, atol=1e-10)
    ll = lambda P: np.sum(Cr * np.log(P))
    assert ll(Pr) <= ll(Cr / Cr.sum(axis=1)[:, None]) + 1e-9      # cannot beat the unconstrained MLE
    assert ll(Pr) >= ll(0.5 * (Pr + np.full((3, 3), 1 / 3.0)))    # ... but beats a perturbed reversible-ish matrix


def test_hmm_container_and_path_statistics(golden):
    from bhmm_b200 import HMM, GaussianOutputModel
    g = golden('gibbs_gauss3')
    K = len(g['lengths'])
    hm = HMM(g['pi'], g['A'], GaussianOutputModel(3, means=g['means'], sigmas=g['sigmas']))
    hm.hidden_state_trajectories = [g['path%d' % k] for k in range(K)]
    assert np.array_equal(hm.count_matrix(), g['count_matrix'])
    assert np.array_equal(hm.count_init(), g['count_init'])
    obs = [g['obs%d' % k] for k in range(K)]
    for i in range(3):
        oi = hm.collect_observations_in_state(obs, i)
        assert len(oi) == int(g['obs_in_state_n%d' % i])
        np.testing.assert_allclose(np.mean(oi), float(g['obs_in_state_mean%d' % i]), rtol=1e-13)
    with pytest.raises(AssertionError):
        hm.update(g['pi'], g['A'] * 2.0)


def test_shard_bounds_partition_every_trajectory_once():
    from bhmm_b200 import dist
    rng = np.random.default_rng(1)
    for world in (1, 2, 3, 4, 8):
        for K in (1, 2, 7, 8, 100):
            lengths = rng.integers(1, 1000, size=K)
            cuts = [dist.shard_bounds(lengths, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == K
            for (a, b), (c, d) in zip(cuts[:-1], cuts[1:]):
                assert b == c and a <= b
            if K >= 4 * world:
                frames = [lengths[a:b].sum() for a, b in cuts]
                assert max(frames) <= 2.0 * lengths.sum() / world + lengths.max()
    assert dist.shard_bounds([10] * 8, 1, 2) == (4, 8)


def test_testsystems_recipe():
    from bhmm_b200.util import testsystems as ts
    pi, A, means, sigmas, O, S = ts.gaussian_observations(5, 3, 2000, seed=4)
    assert O.shape == (3, 2000) and S.max() < 5
    np.testing.assert_allclose(A.sum(axis=1), 1.0)
    np.testing.assert_allclose(pi @ A, pi, atol=1e-12)
    np.testing.assert_allclose(means, np.linspace(-5, 5, 5))
    lifetimes = 1.0 / (1.0 - np.diag(A))
    np.testing.assert_allclose(lifetimes, np.exp(np.linspace(np.log(10), np.log(100), 5)), rtol=1e-10)
    pi, A, B, Od, Sd = ts.discrete_observations(6, 40, 2, 500, seed=1)
    assert Od.dtype == np.int32 and Od.max() < 40 and Od.min() >= 0
    pi0, A0, m0, s0 = ts.perturbed_initial_model(A, np.linspace(-5, 5, 6), 6)
    np.testing.assert_allclose(A0.sum(axis=1), 1.0)
    assert not np.allclose(A0, A0.T)       # asymmetric on purpose: selects the non-reversible M-step


def test_time_shard_border_certification_logic():
    """Host side of the time-sharded E-step (engine.TimeShardedTrajectories.certify): the forward vector a shard started
    from is compared with the left neighbour's vector at the same frame, the backward one with the right neighbour's;
    shards that own nothing of a trajectory are skipped; a mismatch above the tolerance raises."""
    import numpy as np
    from bhmm_b200.engine import TimeShardedTrajectories as TS, _rel_mismatch
    assert _rel_mismatch([1.0, 0.0, 2.0], [1.0, 0.0, 2.0]) == 0.0
    assert abs(_rel_mismatch([1.0, 2.0], [1.0, 2.0 * (1 + 1e-9)]) - 1e-9) < 1e-12
    assert _rel_mismatch([np.nan, 1.0], [1.0, 1.0]) == 1.0
    N, world = 3, 3
    rng = np.random.default_rng(0)
    fwd_at = {10: rng.random(N), 20: rng.random(N)}      # exact alpha at the frame before each border
    bwd_at = {10: rng.random(N), 20: rng.random(N)}      # exact beta at the first frame after each border
    ranges = [[(0, 10, 30), (0, 0, 2)], [(10, 20, 30), (0, 1, 2)], [(20, 30, 30), (1, 2, 2)]]
    borders = np.zeros((world, 2, 4, N))
    # trajectory 0: three owners
    borders[0, 0, 1], borders[1, 0, 0] = fwd_at[10], fwd_at[10] * (1 + 1e-14)
    borders[1, 0, 1], borders[2, 0, 0] = fwd_at[20], fwd_at[20]
    borders[0, 0, 2], borders[1, 0, 3] = bwd_at[10], bwd_at[10]
    borders[1, 0, 2], borders[2, 0, 3] = bwd_at[20] * (1 - 2e-14), bwd_at[20]
    # trajectory 1 (2 frames): shard 0 owns nothing, shards 1 and 2 one frame each
    v, w = rng.random(N), rng.random(N)
    borders[1, 1, 1], borders[2, 1, 0] = v, v
    borders[1, 1, 2], borders[2, 1, 3] = w, w
    worst = TS.certify(list(borders), ranges, 1e-11)
    assert 1e-14 < worst < 1e-13
    borders[2, 0, 0] = fwd_at[20] * (1 + 1e-6)
    import pytest
    with pytest.raises(RuntimeError):
        TS.certify(list(borders), ranges, 1e-11)


def test_warm_up_adaptation_policy():
    """bhmm_b200_adapt_warm (capi.cu:adapt_warm): a failed pass lengthens the warm-up, a mismatch at the rounding floor
    shrinks it slowly but not below 1.12 x the remembered need, and under a mixing rate that jitters by 5 % from pass to
    pass the certification fails in well under 1.5 % of the passes while the warm-up stays within 15 % of the ideal."""
    import ctypes as C
    import math
    import numpy as np
    from bhmm_b200 import _lib
    fn = _lib.lib.bhmm_b200_adapt_warm

    def step(cur, need, worst, failed, state, wmin=32, cap=10 ** 6):
        return fn(int(cur), float(need), float(worst), int(failed), wmin, cap, state.ctypes.data_as(C.POINTER(C.c_double)))

    st = np.zeros(2)
    assert step(100, 300.0, 1e-5, True, st) == 368              # 1.2 x need, multiple of 16
    assert st[0] == 300.0
    assert step(368, 0.0, 3e-15, False, st) == 352              # floor: -5 %, but >= 1.12 x 300 = 336
    assert step(352, 0.0, 3e-15, False, st) == 336
    assert step(336, 0.0, 3e-15, False, st) == 336              # held by the remembered need
    st = np.zeros(2)
    assert step(64, 2 * 64 + 32, 0.9, True, st) == 192 and st[0] == 0.0     # unmeasurable mismatch: doubling rule only
    assert step(100, 50.0, 3e-15, False, np.zeros(2), wmin=128) == 128       # explicit floor
    assert step(1000, 5000.0, 1e-3, True, np.zeros(2), cap=2000) == 2000     # capped at the longest trajectory

    def simulate(jitter, iters=3000, need0=500.0, seed=0):
        rng = np.random.default_rng(seed)
        w, fails, wsum = 576, 0, 0
        state = np.zeros(2)
        for _ in range(iters):
            true_need = need0 * (1 + jitter * rng.standard_normal())
            m = max(10.0 ** (-13.0 * w / true_need), 3e-15 * math.exp(0.5 * rng.standard_normal()))
            failed = m > 1e-13
            need = w * math.log(1e-13) / math.log(m) if m < 0.5 else 2 * w + 32
            fails += failed
            wsum += w
            w = step(w, need, m, failed, state)
        return fails / iters, wsum / iters

    for jitter, max_fail in ((0.03, 0.006), (0.05, 0.015)):
        fail, mean_w = simulate(jitter)
        assert fail <= max_fail, (jitter, fail)
        assert 500.0 * 1.05 < mean_w < 500.0 * 1.30, (jitter, mean_w)


def test_reversible_transition_matrix_sampler_properties():
    """util/tmatrix.py:sample_P_reversible replaces msmtools' sampler (absent, parity unpinned): every sample is a
    stochastic matrix in detailed balance, the sample mean approaches the reversible MLE as counts grow, the spread
    shrinks like 1/sqrt(counts), and a disconnected count matrix is refused like in the reference."""
    import numpy as np
    import pytest
    from bhmm_b200.util import tmatrix
    rng = np.random.default_rng(5)
    P = np.array([[0.90, 0.08, 0.02], [0.10, 0.80, 0.10], [0.02, 0.18, 0.80]])
    pi = tmatrix.stationary_distribution(P)
    spreads = []
    for total in (3e3, 3e5):
        C = total * pi[:, None] * P + 1.0 / 3          # expected counts + a 'mixed'-like prior
        mle = tmatrix.transition_matrix_reversible(C)
        samples = np.array([tmatrix.sample_P_reversible(C, nsteps=30, rng=rng) for _ in range(60)])
        for T in samples[::10]:
            assert np.all(T >= 0) and np.allclose(T.sum(axis=1), 1.0)
            mu = tmatrix.stationary_distribution(T)
            np.testing.assert_allclose(mu[:, None] * T, (mu[:, None] * T).T, atol=1e-12)
        np.testing.assert_allclose(samples.mean(axis=0), mle, atol=6.0 / np.sqrt(total) + 1e-3)
        spreads.append(samples.std(axis=0).max())
    assert 0 < spreads[1] < spreads[0] / 4                 # 100 x the counts: about 10 x narrower
    with pytest.raises(NotImplementedError):
        tmatrix.sample_P_reversible(np.array([[5.0, 0.0], [0.0, 7.0]]), rng=rng)


def test_gibbs_transition_matrix_update_reversible_and_not():
    """BayesianHMMSampler._updateTransitionMatrix (bayesian_sampling.py:341-373) on given path statistics, without a
    device: the reversible branch hands posterior counts to the from-scratch sampler and yields a model in detailed
    balance; the non-reversible branch draws row-wise Dirichlets; both leave a valid initial distribution."""
    import numpy as np
    from bhmm_b200.estimators.bayesian_sampling import BayesianHMMSampler
    from bhmm_b200.hmm.generic_hmm import HMM
    from bhmm_b200.output_models.gaussian import GaussianOutputModel
    from bhmm_b200.util import tmatrix
    np.random.seed(3)
    A = np.array([[0.95, 0.04, 0.01], [0.05, 0.90, 0.05], [0.02, 0.08, 0.90]])
    pi = tmatrix.stationary_distribution(A)
    st = {'C': np.round(20000 * pi[:, None] * A).astype(np.int64), 'n0': np.array([3, 1, 0], dtype=np.int64)}
    for reversible in (True, False):
        s = BayesianHMMSampler.__new__(BayesianHMMSampler)
        s.reversible, s.stationary, s.nstates = reversible, False, 3
        s._np_rng = np.random.default_rng(11)
        s._rng = np.random
        s.prior_C, s.prior_n0 = A.copy(), pi.copy()
        s.transition_matrix_sampling_steps = 1000
        s.model = HMM(pi, A, GaussianOutputModel(3, means=[-1.0, 0.0, 1.0], sigmas=[1.0, 1.0, 1.0]))
        s._updateTransitionMatrix(st)
        T, p0 = s.model.transition_matrix, s.model.initial_distribution
        assert tmatrix.is_transition_matrix(T) and not np.array_equal(T, A)
        np.testing.assert_allclose(T, A, atol=0.02)
        assert np.all(p0 >= 0) and abs(p0.sum() - 1.0) < 1e-12
        if reversible:
            assert tmatrix.is_reversible(T)


def test_device_cache_bookkeeping_without_a_device():
    """The buffer-identity cache of bhmm_b200.hidden is host-side bookkeeping: entries are keyed by the
    host array's data pointer, served only while shape and fingerprint still match, dropped when the caller changed the
    array, and bounded in number.  (Device tensors are stand-ins here; the GPU path is tests/test_reference_on_cuda.py.)"""
    from bhmm_b200.hidden import api
    api.set_device_cache(True)
    try:
        a = np.random.default_rng(1).random((50, 3))
        api._remember(a, 'DEV-A', 40)
        assert api._lookup(a, 40, 3) == 'DEV-A' and api._lookup(a, 10, 3) == 'DEV-A'    # a prefix of what was written
        assert api._lookup(a, 45, 3) is None                                            # more rows than were written
        assert api._lookup(a, 40, 4) is None                                            # another state count
        assert api._lookup(a.copy(), 40, 3) is None                                     # equal content, other buffer
        a[39, 2] += 1.0                                                                 # the caller touched the last row
        assert api._lookup(a, 40, 3) is None
        st = api.device_cache_stats()
        assert st['stale'] == 1 and st['entries'] == 0
        keep = [np.zeros((4, 2)) + k for k in range(api._CACHE_MAX + 3)]
        for k, b in enumerate(keep):
            api._remember(b, k, 4)
        assert api.device_cache_stats()['entries'] == api._CACHE_MAX
        assert api._lookup(keep[0], 4, 2) is None and api._lookup(keep[-1], 4, 2) == len(keep) - 1
        f = np.asfortranarray(np.ones((5, 3)))
        api._remember(f, 'X', 5)                                                        # not C-contiguous: never cached
        assert api._lookup(f, 5, 3) is None
    finally:
        api.set_device_cache(False)
    assert api.device_cache_stats()['entries'] == 0 and not api.device_cache_stats()['enabled']


def test_reference_arm_runs_without_the_product_library(tmp_path):
    """`bench.py --impl reference` times the reference's C code in spawned single-threaded workers and must not map
    libbhmm_b200.so (the arm imported bhmm_b200 for its data generator)."""
    import json
    import os
    import subprocess
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, json, io, contextlib\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-traj-per-core', '1', '--workload', 'small']\n"
            "import bench\n"
            "if __name__ == '__main__':\n"
            "    bench.main()\n"
            "    maps = open('/proc/self/maps').read()\n"
            "    print('MAPPED_PRODUCT' if 'libbhmm_b200' in maps else 'CLEAN')\n")
    script = tmp_path / 'run_ref.py'
    script.write_text(code)
    r = subprocess.run([sys.executable, str(script)], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=600, env=dict(os.environ, PYTHONPATH=ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['value'] > 0 and line['cpu_baseline']['cores'] >= 1
    assert line['cpu_baseline']['single_core_as_shipped'] > 0 and line['e2e']['h2d_bytes_per_step'] == 0
    assert 'CLEAN' in r.stdout


def test_reversible_sampler_matches_numerical_integration_at_small_counts():
    """util/tmatrix.sample_P_reversible against its documented target on a 2 x 2 count matrix with SMALL counts (where the
    prior matters: large-count tests cannot tell a flat prior on the simplex from a 1/x-type prior -- round-1 advisor
    finding).  Target: prod T_ij^c_ij under the flat measure on the normalised symmetric weights (x00, x01, x11);
    T01 = x01 / (x00 + x01), T10 = x01 / (x01 + x11).  Reference values by a 600 x 600 grid on the simplex."""
    from bhmm_b200.util import tmatrix
    Cm = np.array([[2.0, 1.0], [1.0, 3.0]])
    g = (np.arange(600) + 0.5) / 600.0
    x00, x01 = np.meshgrid(g, g, indexing='ij')
    x11 = 1.0 - x00 - x01
    ok = x11 > 0
    with np.errstate(invalid='ignore', divide='ignore'):
        t01 = x01 / (x00 + x01)
        t10 = x01 / (x01 + x11)
        dens = np.where(ok, (1 - t01) ** Cm[0, 0] * t01 ** Cm[0, 1] * t10 ** Cm[1, 0] * (1 - t10) ** Cm[1, 1], 0.0)
    dens = np.nan_to_num(dens)
    want01 = float((dens * np.nan_to_num(t01)).sum() / dens.sum())
    want10 = float((dens * np.nan_to_num(t10)).sum() / dens.sum())
    rng = np.random.default_rng(12)
    draws = np.array([tmatrix.sample_P_reversible(Cm, nsteps=12, rng=rng) for _ in range(2500)])
    got01, got10 = draws[:, 0, 1].mean(), draws[:, 1, 0].mean()
    assert abs(got01 - want01) < 0.02 and abs(got10 - want10) < 0.02, ((got01, got10), (want01, want10))
    for T in draws[:50]:                                   # every draw is stochastic and reversible
        assert np.allclose(T.sum(axis=1), 1.0)
        pi = tmatrix.stationary_vector(T)
        assert np.allclose(pi[:, None] * T, (pi[:, None] * T).T)


def test_prefault_touches_pages_without_changing_zeros_and_rejects_null():
    """bhmm_b200_prefault (host code of csrc/transfer.cu): write-touches one byte per page of a fresh allocation; the estimators
    run it on a helper thread (engine.HostBufferInBackground) while the GPU iterates.  No device needed."""
    import ctypes as C
    from bhmm_b200 import _lib
    a = np.zeros(3 * 4096 + 17, dtype=np.uint8)
    assert _lib.lib.bhmm_b200_prefault(C.c_void_p(a.ctypes.data), a.nbytes, 3) == _lib.OK
    assert not a.any()
    assert _lib.lib.bhmm_b200_prefault(C.c_void_p(a.ctypes.data), 0, 0) == _lib.OK
    assert _lib.lib.bhmm_b200_prefault(None, 100, 1) == _lib.ERR_INVALID
    # the background wrapper hands back an array of the requested shape and dtype after its thread has finished
    import bhmm_b200.engine as eng
    h = eng.HostBufferInBackground((2_000_000,), np.int32, threads=2)
    out = h.get()
    assert out.shape == (2_000_000,) and out.dtype == np.int32 and h.get() is out


def test_gibbs_emission_draws_match_the_reference_with_the_same_seed(golden):
    """SURVEY 8f N1, the sampling half: GaussianOutputModel.sample / DiscreteOutputModel.sample of the reference
    (gaussian.py:274-320, discrete.py:217-251; fixture from the reference package, make_golden.py section 9) consume numpy's
    global stream in a fixed order.  The draws here work from the path statistics the GPU returns (count, sum o, sum o^2 /
    symbol histogram) and reproduce the reference's parameters for the same seed: states with 5000, 37, one and no frames."""
    from bhmm_b200.output_models import GaussianOutputModel, DiscreteOutputModel
    g = golden('gibbs_emission_draws')
    obs = [g['obs_in_state%d' % i] for i in range(4)]
    om = GaussianOutputModel(4, means=g['means0'].copy(), sigmas=g['sigmas0'].copy())
    count = np.array([len(o) for o in obs])
    so = np.array([o.sum() for o in obs])
    soo = np.array([(o * o).sum() for o in obs])
    np.random.seed(int(g['seed']))
    om.sample_from_statistics(count, so, soo)
    np.testing.assert_allclose(om.means, g['means'], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(om.sigmas, g['sigmas'], rtol=1e-10)
    assert om.means[3] == g['means0'][3] and om.sigmas[3] == g['sigmas0'][3]       # no frames: untouched
    assert om.sigmas[2] == g['sigmas0'][2]                                          # one frame: mean only
    sym = [g['sym_in_state%d' % i] for i in range(3)]
    dm = DiscreteOutputModel(g['B0'].copy())
    hist = np.array([np.bincount(s, minlength=6) for s in sym])
    np.random.seed(int(g['dseed']))
    dm.sample_from_histogram(hist)
    np.testing.assert_allclose(dm.output_probabilities, g['B'], rtol=1e-12, atol=1e-15)
