"""Parity and timing of the opt-in panel kernels (bhmm_b200/csrc/panel_kernels.cu: N = 32 and 32 < N <= 104) on a B200.

    timeout 900 python tests/panel_check.py            # parity against the oracle, then timing next to the team kernels
    timeout 600 python tests/panel_check.py --quick    # parity only (what tests/test_panel_cuda.py runs)
    timeout 900 python tests/panel_check.py --mode=2   # N = 32 on the 4-warp wide kernels (BHMM_B200_PANEL=2)

The panel family is selected by BHMM_B200_PANEL=1, which the library reads once per process; this script sets it for
itself and times the team kernels in a child process without it.  Every GPU call sits under the caller's `timeout`: the
kernels had never run on hardware when they were written (end of round 1), so treat a hang as a possibility.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CHILD = '--team-child' in sys.argv
TEAM_PARITY = '--team-parity' in sys.argv          # the same parity cases on the team kernels (BHMM_B200_PANEL=0)
MODE = '2' if '--mode=2' in sys.argv else '1'      # 2: N = 32 on the 4-warp wide kernels instead of the one-warp kernels
os.environ['BHMM_B200_PANEL'] = '0' if (CHILD or TEAM_PARITY) else MODE

import numpy as np   # noqa: E402
import torch         # noqa: E402

from bhmm_b200.engine import TrajectoryBatch, unpack_stats   # noqa: E402
from bhmm_b200.util import testsystems as ts                  # noqa: E402

RTOL = 1e-10
failures = []


def check(name, ok, detail=''):
    print('%-58s %s %s' % (name, 'ok' if ok else 'FAIL', detail), flush=True)
    if not ok:
        failures.append(name)


def close(a, b, rtol, atol=0.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return bool(np.all(np.abs(a - b) <= rtol * np.abs(b) + atol)), float(np.max(np.abs(a - b) / (np.abs(b) + 1e-300)))


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def timing_wide():
    """C4 shape (reduced trajectory count): N = 100 discrete, M = 1000."""
    N, M, K, T = 100, 1000, 32, 100000
    pi, A, B, O, S = ts.discrete_observations(N, M, K, T, seed=4)
    b = TrajectoryBatch(list(O), N)
    b.set_profiling(True)
    ms = timeit(lambda: b.estep_discrete(A, pi, B), reps=2)
    st = unpack_stats(b.estep_discrete(A, pi, B)[0].cpu().numpy(), N)
    print('%s N=100 discrete M=1000 K=%d T=%d: E-step %.2f ms -> %.4f G frames/s; kernels %s; info %s; loglik %.10e'
          % ('team ' if CHILD else 'panel', K, T, ms, K * T / ms / 1e6, b.kernel_ms(), b.info(), st['loglik']), flush=True)
    ms = timeit(lambda: b.viterbi_discrete(A, pi, B), reps=1)
    print('      Viterbi %.2f ms -> %.4f G frames/s' % (ms, K * T / ms / 1e6), flush=True)
    b.close()


def parity_wide():
    """32 < N <= 104: the wide kernels (a block of N/8 warps per 8 chains)."""
    from oracle.oracle import Oracle
    orc = Oracle('port')
    rng = np.random.default_rng(43)
    for N in (100, 37, 64, 21):
        X = rng.random((N, N)) ** 2 + 1e-3
        A = X / X.sum(axis=1)[:, None]
        pi = rng.random(N) + 0.01
        pi /= pi.sum()
        means, sigmas = np.linspace(-5, 5, N), np.linspace(0.5, 2.0, N)
        obs = []
        for Tk in (900, 777, 40, 1, 333, 1200):
            s = rng.integers(0, N, size=Tk)
            obs.append(means[s] + sigmas[s] * rng.standard_normal(Tk))
        ref = orc.estep_gaussian(obs, A, pi, means, sigmas)
        wdd = np.zeros(N)
        for g, o in zip(ref['gammas'], obs):
            d = o[:, None] - means[None, :]
            wdd += (g * d * d).sum(axis=0)
        for chunk, warm in ((0, 0), (150, 0)):
            b = TrajectoryBatch(obs, N, chunk=chunk, warm=warm)
            gam = torch.zeros((b.rows, N), dtype=torch.float64, device='cuda')
            st = unpack_stats(b.estep_gaussian(A, pi, means, sigmas, gamma_out=gam).cpu().numpy(), N)
            tag = 'wide N=%d gaussian chunk=%d: ' % (N, chunk)
            check(tag + 'loglik', abs(st['loglik'] - ref['loglik']) <= RTOL * abs(ref['loglik']), '%.12e vs %.12e' % (st['loglik'], ref['loglik']))
            for key, r, rt, at in (('gamma0', ref['gamma0'], RTOL, 1e-300), ('C', ref['C'], 1e-9, 1e-12 * ref['C'].max()),
                                   ('wsum', ref['wsum'], RTOL, 0.0), ('wdd', wdd, 1e-9, 0.0)):
                ok, worst = close(st[key], r, rt, at)
                check(tag + key, ok, 'worst rel %.2e' % worst)
            ok, worst = close(gam.cpu().numpy(), np.vstack(ref['gammas']), 1e-9, 1e-14)
            check(tag + 'gamma rows', ok, 'worst rel %.2e' % worst)
            b.close()
    # discrete, the C4 model family at its own shape ("discrete N=100/M=1000")
    N, M = 100, 1000
    X = rng.random((N, N)) ** 2 + 1e-3
    A = X / X.sum(axis=1)[:, None]
    pi = np.ones(N) / N
    B = rng.random((N, M)) ** 3 + 1e-4
    B /= B.sum(axis=1)[:, None]
    sym = [rng.integers(0, M, size=Tk).astype(np.int32) for Tk in (1500, 801, 60, 1, 2500)]
    sym[1][::7] = M - 1                                    # the last symbol column
    sym[1][3::7] = 0
    b = TrajectoryBatch(sym, N, chunk=200, warm=0)
    stats, Bnum = b.estep_discrete(A, pi, B)
    st = unpack_stats(stats.cpu().numpy(), N)
    rd = orc.estep_discrete(sym, A, pi, B)
    paths = b.split(b.viterbi_discrete(A, pi, B).cpu().numpy())
    check('viterbi N=100 discrete (matrix column in registers): paths bit-exact',
          all(np.array_equal(p_, orc.viterbi(A, orc.discrete_p_obs(o_, B), pi)) for o_, p_ in zip(sym, paths)))
    check('wide N=100 discrete: loglik', abs(st['loglik'] - rd['loglik']) <= RTOL * abs(rd['loglik']))
    for key, got, r in (('C', st['C'], rd['C']), ('gamma0', st['gamma0'], rd['gamma0']), ('Bnum', Bnum.cpu().numpy(), rd['Bnum'])):
        ok, worst = close(got, r, 1e-9, 1e-13)
        check('wide N=100 discrete: ' + key, ok, 'worst rel %.2e' % worst)
    b.close()


def timing_viterbi_c5():
    """C5 shape, reduced: ONE trajectory of 2e7 frames, 32 states (the sequential kernel walks it at ~0.5 us per frame)."""
    N, T = 32, 20000000
    pi, A, means, sigmas, O, S = ts.gaussian_observations(N, 1, T // 10, seed=5)     # the host generator walks frame by frame
    b = TrajectoryBatch([np.tile(O[0], 10)], N)
    ms = timeit(lambda: b.viterbi_gaussian(A, pi, means, sigmas), reps=1)
    print('%s N=32 one trajectory of %d frames: Viterbi %.1f ms -> %.4f G frames/s; info %s'
          % ('team ' if CHILD else 'panel', T, ms, T / ms / 1e6, b.info()), flush=True)
    b.close()


def timing():
    N, K, T = 32, 64, 100000
    pi, A, means, sigmas, O, S = ts.gaussian_observations(N, K, T, seed=5)
    b = TrajectoryBatch(list(O), N)
    b.set_profiling(True)
    ms = timeit(lambda: b.estep_gaussian(A, pi, means, sigmas))
    st = unpack_stats(b.estep_gaussian(A, pi, means, sigmas).cpu().numpy(), N)
    print('%s N=32 gaussian K=%d T=%d: E-step %.2f ms -> %.3f G frames/s; kernels %s; info %s; loglik %.10e'
          % ('team ' if CHILD else 'panel mode ' + MODE, K, T, ms, K * T / ms / 1e6, b.kernel_ms(), b.info(), st['loglik']), flush=True)
    b.close()


def parity():
    from oracle.oracle import Oracle
    import bhmm_b200.hidden as hidden
    orc = Oracle('port')
    N = 32
    rng = np.random.default_rng(41)
    X = rng.random((N, N)) ** 2 + 1e-3
    A = X / X.sum(axis=1)[:, None]
    pi = rng.random(N) + 0.01
    pi /= pi.sum()
    means, sigmas = np.linspace(-5, 5, N), np.linspace(0.5, 2.0, N)
    # ---- Gaussian E-step: ragged trajectories (one shorter than a chunk, one of a single frame), three chunkings
    obs = []
    for Tk in (1500, 1333, 700, 40, 1, 977, 2000, 333, 1200, 64):
        s = rng.integers(0, N, size=Tk)
        obs.append(means[s] + sigmas[s] * rng.standard_normal(Tk))
    ref = orc.estep_gaussian(obs, A, pi, means, sigmas)
    wdd = np.zeros(N)
    wd = np.zeros(N)
    for g, o in zip(ref['gammas'], obs):
        d = o[:, None] - means[None, :]
        wd += (g * d).sum(axis=0)
        wdd += (g * d * d).sum(axis=0)
    for chunk, warm in ((0, 0), (214, 0), (97, 150), (5000, 10)):
        b = TrajectoryBatch(obs, N, chunk=chunk, warm=warm)
        gam = torch.zeros((b.rows, N), dtype=torch.float64, device='cuda')
        st = unpack_stats(b.estep_gaussian(A, pi, means, sigmas, gamma_out=gam).cpu().numpy(), N)
        tag = 'gaussian chunk=%d warm=%d: ' % (chunk, warm)
        check(tag + 'loglik', abs(st['loglik'] - ref['loglik']) <= RTOL * abs(ref['loglik']), '%.12e vs %.12e' % (st['loglik'], ref['loglik']))
        for key, r, rt, at in (('gamma0', ref['gamma0'], RTOL, 1e-300), ('C', ref['C'], 1e-9, 1e-12 * ref['C'].max()),
                               ('wsum', ref['wsum'], RTOL, 0.0), ('wd', wd, 1e-9, 1e-9 * np.abs(wd).max()), ('wdd', wdd, 1e-9, 0.0)):
            ok, worst = close(st[key], r, rt, at)
            check(tag + key, ok, 'worst rel %.2e' % worst)
        ok, worst = close(gam.cpu().numpy(), np.vstack(ref['gammas']), 1e-9, 1e-14)
        check(tag + 'gamma rows', ok, 'worst rel %.2e' % worst)
        print('    info', b.info(), flush=True)
        b.close()
    # ---- outlier rule: an observation 60 sigma away from every state, with and without ignore_outliers
    o2 = [np.concatenate([obs[0][:300], [400.0], obs[0][300:600]])]
    b = TrajectoryBatch(o2, N, chunk=100, warm=0)
    st = unpack_stats(b.estep_gaussian(A, pi, means, sigmas, ignore_outliers=True).cpu().numpy(), N)
    r2 = orc.estep_gaussian(o2, A, pi, means, sigmas, ignore_outliers=True)
    check('outlier frame (ignored): loglik', abs(st['loglik'] - r2['loglik']) <= RTOL * abs(r2['loglik']))
    ok, worst = close(st['C'], r2['C'], 1e-9, 1e-12)
    check('outlier frame (ignored): C', ok, 'worst rel %.2e' % worst)
    b.close()
    # ---- discrete E-step
    M = 50
    B = rng.random((N, M)) ** 3 + 1e-4
    B /= B.sum(axis=1)[:, None]
    sym = [rng.integers(0, M, size=Tk).astype(np.int32) for Tk in (1800, 1234, 600, 17)]
    b = TrajectoryBatch(sym, N, chunk=300, warm=0)
    stats, Bnum = b.estep_discrete(A, pi, B)
    st = unpack_stats(stats.cpu().numpy(), N)
    rd = orc.estep_discrete(sym, A, pi, B)
    check('discrete: loglik', abs(st['loglik'] - rd['loglik']) <= RTOL * abs(rd['loglik']))
    for key, got, r in (('C', st['C'], rd['C']), ('gamma0', st['gamma0'], rd['gamma0']), ('Bnum', Bnum.cpu().numpy(), rd['Bnum'])):
        ok, worst = close(got, r, 1e-9, 1e-13)
        check('discrete: ' + key, ok, 'worst rel %.2e' % worst)
    b.close()
    # ---- literal API: forward over a caller's p_obs table (EM_POBS) and the Gibbs sweep's filter
    pobs = orc.gaussian_p_obs(obs[0], means, sigmas)
    lp, alpha = hidden.forward(A, pobs, pi)
    lp_ref, alpha_ref = orc.forward(A, pobs, pi)
    check('hidden.forward N=32: logprob', abs(lp - lp_ref) <= RTOL * abs(lp_ref))
    check('hidden.forward N=32: alpha', float(np.max(np.abs(alpha - alpha_ref))) <= 1e-10)
    # ---- time-chunked Viterbi: one long trajectory cut into chains, against the sequential oracle
    s = rng.integers(0, N, size=60000)
    long_obs = [means[s] + sigmas[s] * rng.standard_normal(60000)]
    for chunk in (0, 1500):
        b = TrajectoryBatch(long_obs, N, chunk=chunk, warm=0)
        got = b.viterbi_gaussian(A, pi, means, sigmas).cpu().numpy()
        want = orc.viterbi(A, orc.gaussian_p_obs(long_obs[0], means, sigmas), pi)
        check('chunked viterbi N=32, 60000 frames, chunk=%d: path' % chunk, np.array_equal(got, want),
              '%d mismatches; info %s' % (int(np.sum(got != want)), b.info()))
        b.close()
    # ---- literal hidden.viterbi on one long trajectory (the case where the sequential kernel loses to a CPU core)
    N3 = 3
    A3 = np.array([[0.97, 0.02, 0.01], [0.03, 0.94, 0.03], [0.01, 0.04, 0.95]])
    pobs3 = np.ascontiguousarray(rng.random((1000000, N3)) ** 2 + 1e-3)
    pi3 = np.ones(N3) / N3
    t0 = time.time()
    got = hidden.viterbi(A3, pobs3, pi3)
    t1 = time.time()
    want = orc.viterbi(A3, pobs3, pi3)
    t2 = time.time()
    check('literal hidden.viterbi N=3, T=1e6 (chunked): path', np.array_equal(got, want),
          'GPU call %.1f ms (incl. copies), oracle C %.1f ms' % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    b = TrajectoryBatch(obs, N, chunk=214, warm=0)
    path, counts, sums, ll = b.gibbs_gaussian(A, pi, means, sigmas, seed=7, sweep=0)
    check('gibbs sweep N=32: loglik of the filter', abs(ll - ref['loglik']) <= RTOL * abs(ref['loglik']))
    b.close()


if __name__ == '__main__':
    assert torch.cuda.is_available(), 'needs a CUDA device'
    if CHILD:
        timing()
        timing_wide()
        if '--with-sequential-c5' in sys.argv:
            timing_viterbi_c5()                             # ~10 s of a single warp: only on request
        sys.exit(0)
    t0 = time.time()
    parity()
    parity_wide()
    print('parity: %d failure(s) in %.1f s' % (len(failures), time.time() - t0), flush=True)
    if '--quick' not in sys.argv and not failures:
        timing()
        timing_wide()
        timing_viterbi_c5()
        subprocess.run([sys.executable, os.path.abspath(__file__), '--team-child'], timeout=600)
    sys.exit(1 if failures else 0)
