"""Randomised runs of the emulated LIBRARY (engine_emu.so, see engine_emu_driver.py) -- not collected by pytest:

    BHMM_B200_PANEL=1 python tests/emu/fuzz_engine.py [seconds] [seed]

Random state counts (2..48), mixing rates, ragged trajectories, chunk and warm-up lengths through the C ABI: E-step
log-likelihood / transition counts and batched Viterbi paths against the oracle (certification, fix-up sweeps, the chunked
Viterbi with its sequential fallback and the tiled path chase all take part, depending on the draw)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.argv, args = sys.argv[:1], sys.argv[1:]
os.environ.setdefault('BHMM_B200_CHASE_TILED', '1')
import engine_emu_driver as D     # noqa: E402  (loads engine_emu.so)

budget = float(args[0]) if args else 60.0
rng = np.random.default_rng(int(args[1]) if len(args) > 1 else 0)
orc = D.Oracle('port')
t_end = time.time() + budget
case = 0
while time.time() < t_end:
    case += 1
    N = int(rng.integers(2, 49))
    diag = float(rng.choice([1.0, 3.0, 30.0, 200.0]))
    X = rng.random((N, N)) + diag * np.eye(N)
    A = np.ascontiguousarray(X / X.sum(axis=1)[:, None])
    pi = rng.random(N)
    pi /= pi.sum()
    means, sigmas = np.linspace(-5, 5, N), np.linspace(0.5, 2.0, N)
    lengths = [int(rng.integers(1, 200)) for _ in range(int(rng.integers(1, 5)))]
    obs = []
    for T in lengths:
        s = rng.integers(0, N, T)
        obs.append(means[s] + sigmas[s] * rng.standard_normal(T))
    cat = np.ascontiguousarray(np.concatenate(obs))
    chunk, warm = int(rng.integers(8, 60)), int(rng.integers(2, 60))
    b = D.Batch(lengths, N, chunk, warm)
    stats = np.zeros(D.lib.bhmm_b200_stats_len_gaussian(N))
    D.rc_ok(D.lib.bhmm_b200_estep_gaussian(b.h, D.d(cat), D.d(A), D.d(pi), D.d(means), D.d(sigmas), 1, None, D.d(stats), None))
    ref = orc.estep_gaussian(obs, A, pi, means, sigmas)
    st = D.unpack(stats, N)
    e_ll = abs(st['loglik'] - ref['loglik']) / abs(ref['loglik'])
    e_c = float(np.max(np.abs(st['C'] - ref['C']))) / max(1e-300, float(ref['C'].max()))
    path = np.zeros(b.rows, dtype=np.int32)
    D.rc_ok(D.lib.bhmm_b200_viterbi_gaussian(b.h, D.d(cat), D.d(A), D.d(pi), D.d(means), D.d(sigmas), 1,
                                             path.ctypes.data_as(D.C.POINTER(D.C.c_int)), None))
    vit = all(np.array_equal(path[b.offsets[k]:b.offsets[k + 1]], orc.viterbi(A, orc.gaussian_p_obs(o, means, sigmas), pi))
              for k, o in enumerate(obs))
    info = b.info()
    b.close()
    ok = e_ll < 1e-10 and e_c < 1e-9 and vit
    print('case %d: N=%d diag=%g lengths=%s chunk=%d warm=%d chains=%d fix %g/%g: loglik %.1e C %.1e viterbi %s %s'
          % (case, N, diag, lengths, chunk, warm, info['chains'], info['fix_f'], info['fix_b'], e_ll, e_c, vit, 'ok' if ok else 'FAIL'),
          flush=True)
    if not ok:
        sys.exit(1)
print('%d cases ok' % case)
