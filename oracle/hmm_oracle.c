/*
 * oracle/hmm_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the bhmm hidden-Markov-model
 * dynamic-programming path.  It exists so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg can check the CUDA kernels; nothing in the
 * product package (bhmm_b200/) may import, link or call it.
 *
 * Every function states the reference lines whose arithmetic (including the
 * order of floating-point operations, which matters for bit-exact integer
 * outputs) it follows.  Paths are relative to the reference checkout
 * (bhmm/bhmm).  Parity status: PINNED -- tests/test_oracle.py checks each
 * function bit-for-bit against oracle/_ref/libbhmm_ref.so (the reference's own
 * C sources compiled in place, see oracle/Makefile) whenever that library is
 * present, and against tests/golden/ vectors generated from it otherwise.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared  (x86-64 baseline: no FMA, so
 * results round exactly like the reference's setuptools -O2 build).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR_NO_MEM 2   /* same value as _BHMM_ERR_NO_MEM, bhmm/hidden/impl_c/_hidden.h:5 */
#define ORC_ERR_SAMPLE 3   /* the reference calls exit(1) here (_hidden.c:299-304); we report instead */

/* ------------------------------------------------------------------------- */
/* Emission models                                                           */
/* ------------------------------------------------------------------------- */

/* Gaussian density, bhmm/output_models/impl_c/_gaussian.c:5-21:
 * C = 1/(sqrt(2 pi) sigma) recomputed per element, d = (o-mu)/sigma,
 * value = C * exp(-0.5*d*d)  (i.e. ((-0.5*d)*d) left to right). */
static double orc_gauss(double o, double mu, double sigma)
{
    double norm = 1.0 / (sqrt(2.0 * M_PI) * sigma);
    double z = (o - mu) / sigma;
    return norm * exp(-0.5 * z * z);
}

/* pobs[t,i] for a whole trajectory, _gaussian.c:45-70 (time-major, state fastest). */
void orc_gaussian_pobs(const double *obs, const double *mu, const double *sigma,
                       int N, int T, double *pobs)
{
    for (int t = 0; t < T; ++t) {
        double *row = pobs + (size_t)t * N;
        for (int i = 0; i < N; ++i)
            row[i] = orc_gauss(obs[t], mu[i], sigma[i]);
    }
}

/* Discrete emission gather pobs[t,:] = B[:,obs[t]], bhmm/output_models/discrete.py:146-153. */
void orc_discrete_pobs(const int *obs, const double *B, int N, int M, int T, double *pobs)
{
    for (int t = 0; t < T; ++t)
        for (int i = 0; i < N; ++i)
            pobs[(size_t)t * N + i] = B[(size_t)i * M + obs[t]];
}

/* Outlier rule, bhmm/output_models/outputmodel.py:119-131: a row whose sum is
 * exactly 0 becomes all ones.  numpy's sum(axis=1) over a C-contiguous (T,N)
 * array adds the N entries of a row; a sum of non-negative doubles is 0 iff
 * every entry is 0, so the summation order is irrelevant for the test.
 * Returns the number of rows replaced (found_outliers = result > 0). */
int orc_handle_outliers(double *pobs, int N, int T)
{
    int hit = 0;
    for (int t = 0; t < T; ++t) {
        double *row = pobs + (size_t)t * N;
        double s = 0.0;
        for (int i = 0; i < N; ++i) s += row[i];
        if (s == 0.0) {
            for (int i = 0; i < N; ++i) row[i] = 1.0;
            ++hit;
        }
    }
    return hit;
}

/* B-numerator scatter-add, bhmm/output_models/impl_c/_discrete.c:1-32. */
void orc_discrete_update_pout(const int *obs, const double *w, int T, int N, int M, double *pout)
{
    for (int t = 0; t < T; ++t) {
        const double *wr = w + (size_t)t * N;
        int sym = obs[t];
        for (int i = 0; i < N; ++i)
            pout[(size_t)i * M + sym] += wr[i];
    }
}

/* ------------------------------------------------------------------------- */
/* Scaled forward / backward                                                 */
/* ------------------------------------------------------------------------- */

/* bhmm/hidden/impl_c/_hidden.c:16-66.  alpha_0 = pi*p_0, each row divided by
 * its own sum iff the sum is non-zero, logprob = sum_t log(c_t).  The dot
 * product over i is accumulated in increasing i starting from 0.0. */
double orc_forward(double *alpha, const double *A, const double *pobs, const double *pi,
                   int N, int T)
{
    double c = 0.0;
    for (int i = 0; i < N; ++i) {
        alpha[i] = pi[i] * pobs[i];
        c += alpha[i];
    }
    double ll = log(c);
    if (c != 0)
        for (int i = 0; i < N; ++i) alpha[i] /= c;

    for (int t = 1; t < T; ++t) {
        const double *prev = alpha + (size_t)(t - 1) * N;
        double *cur = alpha + (size_t)t * N;
        const double *p = pobs + (size_t)t * N;
        c = 0.0;
        for (int j = 0; j < N; ++j) {
            double dot = 0.0;
            for (int i = 0; i < N; ++i) dot += prev[i] * A[i * N + j];
            cur[j] = dot * p[j];
            c += cur[j];
        }
        if (c != 0)
            for (int j = 0; j < N; ++j) cur[j] /= c;
        ll += log(c);
    }
    return ll;
}

/* _hidden.c:69-110.  beta_{T-1} = 1/N (computed as 1.0 / sum of N ones);
 * beta_t[i] = sum_j (A[i,j]*p_{t+1}[j])*beta_{t+1}[j]; row rescaled by its own
 * sum iff non-zero (independent of the forward scaling). */
void orc_backward(double *beta, const double *A, const double *pobs, int N, int T)
{
    double *last = beta + (size_t)(T - 1) * N;
    double c = 0.0;
    for (int i = 0; i < N; ++i) { last[i] = 1.0; c += last[i]; }
    for (int i = 0; i < N; ++i) last[i] /= c;

    for (int t = T - 2; t >= 0; --t) {
        const double *nxt = beta + (size_t)(t + 1) * N;
        const double *p = pobs + (size_t)(t + 1) * N;
        double *cur = beta + (size_t)t * N;
        c = 0.0;
        for (int i = 0; i < N; ++i) {
            double dot = 0.0;
            for (int j = 0; j < N; ++j) dot += A[i * N + j] * p[j] * nxt[j];
            cur[i] = dot;
            c += dot;
        }
        if (c != 0)
            for (int i = 0; i < N; ++i) cur[i] /= c;
    }
}

/* State probabilities, bhmm/hidden/api.py:133-188 (numpy: gamma = alpha*beta,
 * then divided by the row sum obtained from a BLAS dot with a ones vector).
 * The row-sum order of BLAS is unspecified, so agreement with the reference is
 * to a few ulp, not bit-exact; we add left to right. */
void orc_state_probabilities(double *gamma, const double *alpha, const double *beta, int N, int T)
{
    for (int t = 0; t < T; ++t) {
        size_t o = (size_t)t * N;
        double s = 0.0;
        for (int i = 0; i < N; ++i) { gamma[o + i] = alpha[o + i] * beta[o + i]; s += gamma[o + i]; }
        for (int i = 0; i < N; ++i) gamma[o + i] /= s;
    }
}

/* State counts, bhmm/hidden/api.py:191-211 (np.sum(gamma[0:T], axis=0)). */
void orc_state_counts(double *counts, const double *gamma, int N, int T)
{
    for (int i = 0; i < N; ++i) counts[i] = 0.0;
    for (int t = 0; t < T; ++t)
        for (int i = 0; i < N; ++i) counts[i] += gamma[(size_t)t * N + i];
}

/* Baum-Welch transition counts, _hidden.c:148-183.
 * xi_t[i,j] = ((alpha_t[i]*A[i,j])*p_{t+1}[j])*beta_{t+1}[j]; the N*N entries
 * are summed in row-major order and each entry is divided by that sum. */
int orc_transition_counts(double *C, const double *A, const double *pobs,
                          const double *alpha, const double *beta, int N, int T)
{
    size_t nn = (size_t)N * N;
    for (size_t k = 0; k < nn; ++k) C[k] = 0.0;
    double *xi = (double *)malloc(nn * sizeof(double));
    if (!xi) return ORC_ERR_NO_MEM;
    for (int t = 0; t + 1 < T; ++t) {
        const double *a = alpha + (size_t)t * N;
        const double *b = beta + (size_t)(t + 1) * N;
        const double *p = pobs + (size_t)(t + 1) * N;
        double tot = 0.0;
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                double v = a[i] * A[i * N + j] * p[j] * b[j];
                xi[i * N + j] = v;
                tot += v;
            }
        for (size_t k = 0; k < nn; ++k) C[k] += xi[k] / tot;
    }
    free(xi);
    return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* Viterbi                                                                   */
/* ------------------------------------------------------------------------- */

/* First index of the maximum with a strict '>' test, _hidden.c:186-200. */
static int orc_first_max(const double *v, int n)
{
    int best = 0;
    double m = v[0];
    for (int i = 1; i < n; ++i)
        if (v[i] > m) { m = v[i]; best = i; }
    return best;
}

/* _hidden.c:203-281.  v_0 = (p_0*pi)/sum; for every target state j the
 * candidates are h_i = v_i*A[i,j]; the kept value is (p_t[j]*v[best])*A[best,j]
 * (left to right), the row sum runs over j in increasing order and every entry
 * is divided by it (unconditionally).  Back-pointers, then backtrace. */
int orc_viterbi(int *path, const double *A, const double *pobs, const double *pi, int N, int T)
{
    double *v = (double *)malloc(sizeof(double) * N);
    double *w = (double *)malloc(sizeof(double) * N);
    double *h = (double *)malloc(sizeof(double) * N);
    int *bp = (int *)malloc(sizeof(int) * (size_t)T * N);
    int rc = ORC_OK;
    if (!v || !w || !h || !bp) { rc = ORC_ERR_NO_MEM; goto done; }

    double s = 0.0;
    for (int i = 0; i < N; ++i) { v[i] = pobs[i] * pi[i]; s += v[i]; }
    for (int i = 0; i < N; ++i) v[i] /= s;

    for (int t = 1; t < T; ++t) {
        const double *p = pobs + (size_t)t * N;
        s = 0.0;
        for (int j = 0; j < N; ++j) {
            for (int i = 0; i < N; ++i) h[i] = v[i] * A[i * N + j];
            int best = orc_first_max(h, N);
            bp[(size_t)t * N + j] = best;
            w[j] = p[j] * v[best] * A[best * N + j];
            s += w[j];
        }
        for (int j = 0; j < N; ++j) w[j] /= s;
        double *tmp = v; v = w; w = tmp;
    }
    path[T - 1] = orc_first_max(v, N);
    for (int t = T - 2; t >= 0; --t)
        path[t] = bp[(size_t)(t + 1) * N + path[t + 1]];
done:
    free(v); free(w); free(h); free(bp);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* Forward-filter / backward-sample                                          */
/* ------------------------------------------------------------------------- */

/* glibc srand()/rand() (TYPE_3 additive-feedback generator, degree 31, sep 3)
 * restated from its published algorithm: seed expansion with the Lehmer step
 * 16807*x mod (2^31-1) computed by Schrage's method, 310 discarded outputs,
 * then o_k = (r[k-31] + r[k-3]) >> 1.  Writes the uniforms the reference's
 * _random_choice forms, r = rand()/(RAND_MAX+1.0)  (_hidden.c:285-287), in draw
 * order.  tests/test_oracle.py checks this against the C library on the box. */
void orc_glibc_uniforms(int seed, long n, double *u)
{
    int32_t r[34];
    uint32_t *ring;
    long total = 344 + n;
    ring = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)total);
    if (!ring) return;
    uint32_t s = (uint32_t)seed;
    if (s == 0) s = 1;
    r[0] = (int32_t)s;
    for (int i = 1; i < 31; ++i) {
        long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t)w;
    }
    for (int i = 0; i < 31; ++i) ring[i] = (uint32_t)r[i];
    for (int i = 31; i < 34; ++i) ring[i] = ring[i - 31];
    for (long i = 34; i < total; ++i) ring[i] = ring[i - 31] + ring[i - 3];
    for (long k = 0; k < n; ++k) {
        uint32_t o = ring[344 + k] >> 1;
        u[k] = (double)o / (2147483647.0 + 1.0);
    }
    free(ring);
}

/* Sequential normalisation (_hidden.c:307-319) followed by the inverse-CDF
 * draw of _hidden.c:283-305 (first i whose running sum is >= r). */
static int orc_draw(double *p, int n, double r)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += p[i];
    for (int i = 0; i < n; ++i) p[i] /= s;
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        acc += p[i];
        if (acc >= r) return i;
    }
    return -1;
}

/* _hidden.c:330-378 with the uniforms made explicit: u[0] is consumed for
 * t = T-1, u[1] for t = T-2, ...  psel_i = alpha_t[i]*A[i, s_{t+1}]. */
int orc_sample_path(int *path, const double *alpha, const double *A, const double *u, int N, int T)
{
    double *psel = (double *)malloc(sizeof(double) * N);
    if (!psel) return ORC_ERR_NO_MEM;
    int rc = ORC_OK;
    for (int i = 0; i < N; ++i) psel[i] = alpha[(size_t)(T - 1) * N + i];
    int s = orc_draw(psel, N, u[0]);
    if (s < 0) { rc = ORC_ERR_SAMPLE; goto done; }
    path[T - 1] = s;
    for (int t = T - 2; t >= 0; --t) {
        int nxt = path[t + 1];
        for (int i = 0; i < N; ++i) psel[i] = alpha[(size_t)t * N + i] * A[i * N + nxt];
        s = orc_draw(psel, N, u[T - 1 - t]);
        if (s < 0) { rc = ORC_ERR_SAMPLE; goto done; }
        path[t] = s;
    }
done:
    free(psel);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* Estimator-level composites (what the EM / Gibbs loops do per trajectory)   */
/* ------------------------------------------------------------------------- */

/* One trajectory of MaximumLikelihoodEstimator._forward_backward
 * (bhmm/estimators/maximum_likelihood.py:221-269) for a Gaussian output model:
 * p_obs (+ outlier rule) -> forward -> backward -> gamma -> transition counts,
 * then the sufficient statistics the M-step consumes
 * (maximum_likelihood.py:271-282, output_models/gaussian.py:214-272):
 *   gamma0[i] += gamma[0,i];  C += C_k;  wsum[i] += sum_t gamma[t,i];
 *   wo[i] += sum_t gamma[t,i]*o_t.
 * work must hold 4*T*N doubles.  gamma (T*N) is returned for the second
 * (variance) pass of GaussianOutputModel.estimate.  Returns log-likelihood. */
double orc_estep_gaussian_traj(const double *obs, int T, int N,
                               const double *A, const double *pi,
                               const double *mu, const double *sigma, int ignore_outliers,
                               double *work, double *gamma,
                               double *gamma0, double *C, double *wsum, double *wo)
{
    size_t tn = (size_t)T * N;
    double *pobs = work, *alpha = work + tn, *beta = work + 2 * tn, *Ck = work + 3 * tn;
    orc_gaussian_pobs(obs, mu, sigma, N, T, pobs);
    if (ignore_outliers) orc_handle_outliers(pobs, N, T);
    double ll = orc_forward(alpha, A, pobs, pi, N, T);
    orc_backward(beta, A, pobs, N, T);
    orc_state_probabilities(gamma, alpha, beta, N, T);
    orc_transition_counts(Ck, A, pobs, alpha, beta, N, T);
    for (int i = 0; i < N; ++i) gamma0[i] += gamma[i];
    for (int k = 0; k < N * N; ++k) C[k] += Ck[k];
    for (int t = 0; t < T; ++t)
        for (int i = 0; i < N; ++i) {
            double g = gamma[(size_t)t * N + i];
            wsum[i] += g;
            wo[i] += g * obs[t];
        }
    return ll;
}

/* Second pass of GaussianOutputModel.estimate (gaussian.py:259-270):
 * wvar[i] += sum_t gamma[t,i]*(o_t - mean_new[i])^2. */
void orc_gaussian_var_pass(const double *obs, const double *gamma, int T, int N,
                           const double *mean_new, double *wvar)
{
    for (int t = 0; t < T; ++t)
        for (int i = 0; i < N; ++i) {
            double d = obs[t] - mean_new[i];
            wvar[i] += gamma[(size_t)t * N + i] * (d * d);
        }
}

/* Same composite for a discrete output model (discrete.py:130-215):
 * Bnum (N*M) receives the _update_pout scatter-add of gamma. */
double orc_estep_discrete_traj(const int *obs, int T, int N, int M,
                               const double *A, const double *pi, const double *B,
                               int ignore_outliers, double *work, double *gamma,
                               double *gamma0, double *C, double *Bnum)
{
    size_t tn = (size_t)T * N;
    double *pobs = work, *alpha = work + tn, *beta = work + 2 * tn, *Ck = work + 3 * tn;
    orc_discrete_pobs(obs, B, N, M, T, pobs);
    if (ignore_outliers) orc_handle_outliers(pobs, N, T);
    double ll = orc_forward(alpha, A, pobs, pi, N, T);
    orc_backward(beta, A, pobs, N, T);
    orc_state_probabilities(gamma, alpha, beta, N, T);
    orc_transition_counts(Ck, A, pobs, alpha, beta, N, T);
    for (int i = 0; i < N; ++i) gamma0[i] += gamma[i];
    for (int k = 0; k < N * N; ++k) C[k] += Ck[k];
    orc_discrete_update_pout(obs, gamma, T, N, M, Bnum);
    return ll;
}

/* Path statistics of one Gibbs sweep for one trajectory: lag-1 transition
 * counts of the sampled path (HMM.count_matrix, bhmm/hmm/generic_hmm.py:297-319:
 * msmtools count_matrix with lag 1, sliding), first-state histogram
 * (count_init, :321-334) and, per state, the number of frames and the sums the
 * Gaussian sampler needs (collect_observations_in_state :398-431 followed by
 * np.mean / np.mean((o-mu)^2), gaussian.py:274-320). */
void orc_path_stats(const int *path, const double *obs, int T, int N,
                    long long *Cint, long long *n0, long long *cnt, double *so, double *soo)
{
    n0[path[0]] += 1;
    for (int t = 0; t < T; ++t) {
        int s = path[t];
        cnt[s] += 1;
        so[s] += obs[t];
        soo[s] += obs[t] * obs[t];
        if (t + 1 < T) Cint[(size_t)s * N + path[t + 1]] += 1;
    }
}
