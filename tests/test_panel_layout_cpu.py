"""Lane-level emulation of bhmm_b200/csrc/panel_kernels.cu (N = 32 on the FP64 tensor pipe, 8 chains per warp).

The panel kernels were written without GPU time left in round 1, so their index algebra -- which lane holds which state,
which element of the transition matrix sits in which B fragment, why the accumulator fragment of one frame IS the A
operand of the next, the shared-memory transposition for the xi product, the layout of the partial-statistics row -- is
checked here on the CPU: every per-lane register of one warp is a numpy array over the 32 lanes, mma.sync.m8n8k4.f64 is
emulated from the PTX fragment layout, and the control flow (chain table, warm-up / exact / trajectory-start modes,
virtual frame T, emitted frame f-1) mirrors the kernels statement by statement.  The emulated warp must reproduce a plain
scaled forward-backward (the algorithm of bhmm/hidden/impl_c/_hidden.c:16-183) on whole trajectories.
"""
import numpy as np

PN, PCH, PSTRIDE = 32, 8, 36
LANE = np.arange(32)
G, Q = LANE >> 2, LANE & 3


def state_of(k, q):
    return 8 * (k >> 1) + 2 * q + (k & 1)


STATE = np.array([[state_of(k, q) for k in range(8)] for q in Q])         # [lane][k]


def dmma(d0, d1, a, b):
    """mma.sync.aligned.m8n8k4.row.col.f64: A[row = lane / 4][k = lane % 4], B[k = lane % 4][n = lane / 4],
    C/D[row = lane / 4][col = 2 (lane % 4) + {0, 1}]."""
    Am = np.zeros((8, 4))
    Bm = np.zeros((4, 8))
    Am[G, Q] = a
    Bm[Q, G] = b
    D = Am.dot(Bm)
    return d0 + D[G, 2 * Q], d1 + D[G, 2 * Q + 1]


def quad_sum(v):
    v = v + v[LANE ^ 1]
    return v + v[LANE ^ 2]


def load8(rows_base, table):
    """table[(row of each lane), state_of(k, q)] -> [lane][k]"""
    return table[rows_base[:, None], STATE]


def gauss(o, mu, sigma):
    return np.exp(-0.5 * ((o - mu) / sigma) ** 2) / (np.sqrt(2 * np.pi) * sigma)


def emulate_forward(chains, warm, exact, obs, A, pi, mu, sigma, alpha, hand_used, hand_end, chain_ll):
    """One warp of k_forward_panel32<EM_GAUSS>; chains = list of (row0, len, t0, T) for the 8 quads (None: no chain)."""
    Bf = np.zeros((8, 4, 32))
    for ks in range(8):
        for nt in range(4):
            Bf[ks, nt] = A[STATE[:, ks], 8 * nt + G]
    have = np.array([chains[g] is not None for g in G])
    c = np.array([g if chains[g] is not None else -1 for g in G])
    ln = np.array([chains[g][1] if chains[g] else 0 for g in G])
    t0 = np.array([chains[g][2] if chains[g] else 0 for g in G])
    trow = np.array([chains[g][0] - chains[g][2] if chains[g] else 0 for g in G])
    tstart = np.where(t0 == 0, 0, (t0 - 1) if exact else np.maximum(0, t0 - warm))
    mode = np.where(t0 == 0, 0, 2 if exact else np.where(tstart == 0, 0, 1))
    npre = np.where(have, t0 - tstart, 0)
    maxpre = npre.max()
    total = maxpre + ln.max()
    tend = t0 + ln
    av = np.zeros((32, 8))
    for l in LANE:
        if have[l] and mode[l] == 2:
            av[l] = hand_end[c[l] - 1, STATE[l]]
    ll = np.zeros(32)
    for s in range(total):
        t = t0 - maxpre + s
        on = have & (t >= tstart) & (t < tend)
        init = on & (t == tstart)
        tc = np.minimum(np.maximum(t, tstart + (mode == 2)), tend - 1)
        o = np.where(have, obs[np.where(have, trow + tc, 0)], 0.0)
        p = gauss(o[:, None], mu[STATE], sigma[STATE])
        d = np.zeros((32, 8))
        for ks in range(8):
            for nt in range(4):
                d[:, 2 * nt], d[:, 2 * nt + 1] = dmma(d[:, 2 * nt], d[:, 2 * nt + 1], av[:, ks], Bf[ks, nt])
        v = np.where(on[:, None], d * p, 0.0)
        for l in LANE[init]:
            v[l] = pi[STATE[l]] * p[l] if mode[l] == 0 else (p[l] if mode[l] == 1 else av[l])
        csum = quad_sum(v.sum(axis=1))
        for l in LANE[on]:
            av[l] = v[l] / csum[l]
            if t[l] >= t0[l]:
                alpha[trow[l] + t[l], STATE[l]] = av[l]
                ll[l] += np.log(csum[l])
                if t[l] == tend[l] - 1:
                    hand_end[c[l], STATE[l]] = av[l]
            elif t[l] == t0[l] - 1:
                hand_used[c[l], STATE[l]] = av[l]
    for l in LANE[have & (Q == 0)]:
        chain_ll[c[l]] = ll[l]


def emulate_backward_stats(chains, warm, exact, obs, A, mu, sigma, alpha, hand_used, hand_end):
    """One warp of k_backward_stats_panel32<EM_GAUSS>; returns the warp's row of partial statistics."""
    Bt = np.zeros((8, 4, 32))
    for ks in range(8):
        for nt in range(4):
            Bt[ks, nt] = A[8 * nt + G, STATE[:, ks]]
    X = np.zeros((4, 4, 2, 32))
    st_g, st_gd, st_gdd = np.zeros((32, 8)), np.zeros((32, 8)), np.zeros((32, 8))
    out = np.zeros(PN * PN + 4 * PN)
    have = np.array([chains[g] is not None for g in G])
    c = np.array([g if chains[g] is not None else -1 for g in G])
    ln = np.array([chains[g][1] if chains[g] else 0 for g in G])
    t0 = np.array([chains[g][2] if chains[g] else 0 for g in G])
    T = np.array([chains[g][3] if chains[g] else 0 for g in G])
    trow = np.array([chains[g][0] - chains[g][2] if chains[g] else 0 for g in G])
    e = t0 + ln
    virt = have & (e >= T)
    mode = np.where(virt, 0, 2 if exact else 0)
    fstart = np.where(virt, T, e if exact else np.minimum(T - 1, e + warm - 1))
    flast = t0 + 1
    npre = np.where(have, fstart - (e - 1), 0)
    maxpre = npre.max()
    total = maxpre + np.where(have, e - flast, 0).max()
    bn = np.full((32, 8), 1.0 / PN)
    for l in LANE:
        if have[l] and mode[l] == 2:
            bn[l] = hand_end[c[l] + 1, STATE[l]]
    Us, Ws = np.zeros(PCH * PSTRIDE), np.zeros(PCH * PSTRIDE)
    for s in range(total):
        f = (e - 1) + maxpre - s
        on = have & (f <= fstart) & (f >= flast)
        isvirt = on & virt & (f == T)
        fc = np.minimum(np.maximum(f, t0), T - 1)
        o = np.where(have, obs[np.where(have, trow + fc, 0)], 0.0)
        fa = np.minimum(np.maximum(f - 1, t0), e - 1)
        al = np.where(have[:, None], load8(np.where(have, trow + fa, 0), alpha), 0.0)
        fn = np.minimum(np.maximum(f - 1, t0), T - 1)                      # raw_next: observation of the emitted frame
        o_next = np.where(have, obs[np.where(have, trow + fn, 0)], 0.0)
        p = gauss(o[:, None], mu[STATE], sigma[STATE])
        w = np.where((on & ~isvirt)[:, None], p * bn, 0.0)
        d = np.zeros((32, 8))
        for ks in range(8):
            for nt in range(4):
                d[:, 2 * nt], d[:, 2 * nt + 1] = dmma(d[:, 2 * nt], d[:, 2 * nt + 1], w[:, ks], Bt[ks, nt])
        d[isvirt] = 1.0
        for l in LANE[on & ~isvirt & (f == e)]:
            hand_used[c[l], STATE[l]] = bn[l]
        emit = on & (f - 1 < e)
        xi = emit & ~isvirt
        gq = np.where(emit[:, None], al * d, 0.0)
        S = quad_sum(gq.sum(axis=1))
        sbn = quad_sum(d.sum(axis=1))
        with np.errstate(divide='ignore', invalid='ignore'):
            rS = 1.0 / S
            u = np.where(xi[:, None], al * rS[:, None], 0.0)
        wz = np.where(xi[:, None], w, 0.0)
        any_xi = xi.any()
        if any_xi:
            for l in LANE:                                                  # store8(Us + g * PSTRIDE, q, u)
                Us[G[l] * PSTRIDE + STATE[l]] = u[l]
                Ws[G[l] * PSTRIDE + STATE[l]] = wz[l]
        for l in LANE[emit]:
            gam = gq[l] * rS[l]
            st_g[l] += gam
            if f[l] - 1 == 0:
                out[PN * PN + STATE[l]] += gam
            dd = o_next[l] - mu[STATE[l]]
            st_gd[l] += gam * dd
            st_gdd[l] += gam * dd * dd
            if f[l] - 1 == t0[l] and t0[l] > 0:
                hand_end[c[l], STATE[l]] = d[l] / sbn[l]
        for l in LANE[on]:
            bn[l] = d[l] / sbn[l]
        if any_xi:
            ua = np.zeros((2, 4, 32))
            wb = np.zeros((2, 4, 32))
            for kk in range(2):
                for t in range(4):
                    ua[kk, t] = Us[(4 * kk + Q) * PSTRIDE + 8 * t + G]
                    wb[kk, t] = Ws[(4 * kk + Q) * PSTRIDE + 8 * t + G]
            for kk in range(2):
                for mt in range(4):
                    for nt in range(4):
                        X[mt, nt, 0], X[mt, nt, 1] = dmma(X[mt, nt, 0], X[mt, nt, 1], ua[kk, mt], wb[kk, nt])
    for mt in range(4):
        for nt in range(4):
            for l in LANE:
                out[(8 * mt + G[l]) * PN + 8 * nt + 2 * Q[l]] = X[mt, nt, 0, l]
                out[(8 * mt + G[l]) * PN + 8 * nt + 2 * Q[l] + 1] = X[mt, nt, 1, l]
    for arr in (st_g, st_gd, st_gdd):
        for off in (4, 8, 16):
            arr += arr[LANE ^ off]
    for l in LANE[G == 0]:
        out[PN * PN + PN + STATE[l]] = st_g[l]
        out[PN * PN + 2 * PN + STATE[l]] = st_gd[l]
        out[PN * PN + 3 * PN + STATE[l]] = st_gdd[l]
    return out


def plain_estep(trajs, A, pi, mu, sigma):
    """Scaled forward-backward with per-frame normalisation (_hidden.c:16-183), whole trajectories."""
    N = len(pi)
    alphas, betas, ll = [], [], 0.0
    C, g0, sg, sgd, sgdd = np.zeros((N, N)), np.zeros(N), np.zeros(N), np.zeros(N), np.zeros(N)
    for o in trajs:
        T = len(o)
        p = gauss(o[:, None], mu[None, :], sigma[None, :])
        al, be = np.zeros((T, N)), np.zeros((T, N))
        v = pi * p[0]
        for t in range(T):
            if t > 0:
                v = al[t - 1].dot(A) * p[t]
            ll += np.log(v.sum())
            al[t] = v / v.sum()
        be[T - 1] = 1.0 / N
        for t in range(T - 2, -1, -1):
            v = A.dot(p[t + 1] * be[t + 1])
            be[t] = v / v.sum()
        gam = al * be
        gam /= gam.sum(axis=1)[:, None]
        for t in range(T - 1):
            x = al[t][:, None] * A * (p[t + 1] * be[t + 1])[None, :]
            C += x / x.sum()
        g0 += gam[0]
        sg += gam.sum(axis=0)
        dd = o[:, None] - mu[None, :]
        sgd += (gam * dd).sum(axis=0)
        sgdd += (gam * dd * dd).sum(axis=0)
        alphas.append(al)
        betas.append(be)
    return dict(alpha=np.vstack(alphas), beta=np.vstack(betas), ll=ll, C=C, g0=g0, sg=sg, sgd=sgd, sgdd=sgdd)


def _model(seed):
    rng = np.random.default_rng(seed)
    A = rng.random((PN, PN)) + 3.0 * np.eye(PN)          # fast mixing: a 60-frame warm-up forgets its start
    A /= A.sum(axis=1)[:, None]
    pi = rng.random(PN)
    pi /= pi.sum()
    mu = np.linspace(-5, 5, PN)
    sigma = np.linspace(0.5, 2.0, PN)
    return rng, A, pi, mu, sigma


def _run(exact):
    rng, A, pi, mu, sigma = _model(7 + exact)
    # three trajectories cut into 8 chains of uneven length; the third is a single chain (starts AND ends a trajectory)
    Ts = [150, 121, 37]
    cuts = [[0, 50, 100, 150], [0, 40, 81, 100, 121], [0, 37]]
    trajs = [mu[rng.integers(0, PN, T)] + 0.7 * rng.standard_normal(T) for T in Ts]
    obs = np.concatenate(trajs)
    chains, row = [], 0
    for T, cut in zip(Ts, cuts):
        for a, b in zip(cut[:-1], cut[1:]):
            chains.append((row + a, b - a, a, T))
        row += T
    assert len(chains) == PCH
    ref = plain_estep(trajs, A, pi, mu, sigma)
    rows = len(obs)
    alpha = np.zeros((rows, PN))
    hu_f, he_f = np.zeros((PCH, PN)), np.zeros((PCH, PN))
    hu_b, he_b = np.zeros((PCH, PN)), np.zeros((PCH, PN))
    chain_ll = np.zeros(PCH)
    warm = 60
    if exact:
        # exact fix-up passes start from the neighbour's recorded hand-over: provide the true ones
        for ci, ch in enumerate(chains):
            he_f[ci] = ref['alpha'][ch[0] + ch[1] - 1]
            he_b[ci] = ref['beta'][ch[0]]
    emulate_forward(chains, warm, exact, obs, A, pi, mu, sigma, alpha, hu_f, he_f, chain_ll)
    np.testing.assert_allclose(alpha, ref['alpha'], rtol=1e-11, atol=1e-300)
    assert abs(chain_ll.sum() - ref['ll']) <= 1e-12 * abs(ref['ll'])
    for ci, ch in enumerate(chains):
        np.testing.assert_allclose(he_f[ci], ref['alpha'][ch[0] + ch[1] - 1], rtol=1e-11)
        if ch[2] > 0:
            np.testing.assert_allclose(hu_f[ci], ref['alpha'][ch[0] - 1], rtol=1e-11)
    out = emulate_backward_stats(chains, warm, exact, obs, A, mu, sigma, ref['alpha'], hu_b, he_b)
    np.testing.assert_allclose(A * out[:PN * PN].reshape(PN, PN), ref['C'], rtol=1e-10)
    np.testing.assert_allclose(out[PN * PN:PN * PN + PN], ref['g0'], rtol=1e-10)
    np.testing.assert_allclose(out[PN * PN + PN:PN * PN + 2 * PN], ref['sg'], rtol=1e-10)
    np.testing.assert_allclose(out[PN * PN + 2 * PN:PN * PN + 3 * PN], ref['sgd'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(out[PN * PN + 3 * PN:], ref['sgdd'], rtol=1e-10)
    assert abs((A * out[:PN * PN].reshape(PN, PN)).sum() - (sum(Ts) - len(Ts))) < 1e-8
    for ci, ch in enumerate(chains):
        if ch[2] > 0:
            np.testing.assert_allclose(he_b[ci], ref['beta'][ch[0]], rtol=1e-10)
        if ch[2] + ch[1] < ch[3]:
            np.testing.assert_allclose(hu_b[ci], ref['beta'][ch[0] + ch[1]], rtol=1e-10)


def test_panel_warp_emulation_warm_up_mode():
    _run(exact=0)


def test_panel_warp_emulation_exact_mode():
    _run(exact=1)


def test_panel_partial_warp_and_operand_banks():
    """A warp with fewer than 8 chains leaves the missing quads idle; the xi operand loads of one half warp hit 16
    distinct 8-byte banks with the padded row stride."""
    rng, A, pi, mu, sigma = _model(3)
    T = 45
    o = mu[rng.integers(0, PN, T)] + 0.5 * rng.standard_normal(T)
    ref = plain_estep([o], A, pi, mu, sigma)
    chains = [(0, 20, 0, T), (20, 25, 20, T)] + [None] * 6
    alpha = np.zeros((T, PN))
    hu, he, cl = np.zeros((PCH, PN)), np.zeros((PCH, PN)), np.zeros(PCH)
    he[0] = ref['alpha'][19]
    he[1] = ref['beta'][20]
    emulate_forward(chains, 0, 1, o, A, pi, mu, sigma, alpha, hu, he, cl)
    np.testing.assert_allclose(alpha, ref['alpha'], rtol=1e-11)
    hb_used, hb_end = np.zeros((PCH, PN)), np.zeros((PCH, PN))
    hb_end[1] = ref['beta'][20]
    out = emulate_backward_stats(chains, 0, 1, o, A, mu, sigma, ref['alpha'], hb_used, hb_end)
    np.testing.assert_allclose(A * out[:PN * PN].reshape(PN, PN), ref['C'], rtol=1e-10)
    for half in (LANE[:16], LANE[16:]):
        for kk in range(2):
            for t in range(4):
                banks = ((4 * kk + Q[half]) * PSTRIDE + 8 * t + G[half]) % 16
                assert len(set(banks.tolist())) == 16
