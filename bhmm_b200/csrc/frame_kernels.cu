// bhmm_b200/csrc/frame_kernels.cu -- frame-parallel pieces of the bhmm.hidden surface.
//
// These have no dependence between frames, so every frame (or every (frame,state) pair) is its own
// thread.  They back the literal per-function API; the fused E-step never materialises their outputs.
//   k_gaussian_pobs / k_discrete_pobs   OutputModel.p_obs + _handle_outliers
//                                       (_gaussian.c:45-70, discrete.py:146-153, outputmodel.py:119-131)
//   k_state_probabilities               hidden/api.py:133-188
//   k_colsum_*                          state_counts, hidden/api.py:191-211
//   k_xi_norm / k_xi_accumulate         _compute_transition_counts, _hidden.c:148-183
//   k_update_pout                       _update_pout, output_models/impl_c/_discrete.c:1-32
//   k_mstep_hmm / k_mstep_rows          the M-step from the reduced statistics, on the device:
//                                       maximum_likelihood.py:284-330 with estimate_P -> C / rowsum
//                                       (_tmatrix_disconnected.py:110-115, non-reversible), GaussianOutputModel.estimate
//                                       gaussian.py:214-272, DiscreteOutputModel.estimate discrete.py:214-215
#include "common.cuh"
#include "kernels.h"

namespace {

__global__ void k_gaussian_pobs(const double* __restrict__ obs, const double* __restrict__ mu,
                                const double* __restrict__ sigma, int N, long long rows, int ignore_outliers,
                                double* __restrict__ pobs)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double o = obs[r];
    double* out = pobs + r * N;
    bool anynz = false;
    for (int i = 0; i < N; ++i) {
        const double p = gauss_pdf(o, mu[i], sigma[i]);
        out[i] = p;
        anynz |= (p != 0.0);
    }
    if (ignore_outliers && !anynz)
        for (int i = 0; i < N; ++i) out[i] = 1.0;
}

__global__ void k_discrete_pobs(const int* __restrict__ sym, const double* __restrict__ Bt, int N, int M,
                                long long rows, int ignore_outliers, double* __restrict__ pobs)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double* src = Bt + (long long)sym[r] * N;
    double* out = pobs + r * N;
    bool anynz = false;
    for (int i = 0; i < N; ++i) {
        const double p = src[i];
        out[i] = p;
        anynz |= (p != 0.0);
    }
    if (ignore_outliers && !anynz)
        for (int i = 0; i < N; ++i) out[i] = 1.0;
}

__global__ void k_state_probabilities(const double* __restrict__ alpha, const double* __restrict__ beta, int N,
                                      long long rows, double* __restrict__ gamma)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double* a = alpha + r * N;
    const double* b = beta + r * N;
    double* g = gamma + r * N;
    double s = 0.0;
    for (int i = 0; i < N; ++i) {
        const double v = __dmul_rn(a[i], b[i]);
        g[i] = v;
        s = __dadd_rn(s, v);
    }
    for (int i = 0; i < N; ++i) g[i] = g[i] / s;
}

// column sums, stage 1: block b sums rows [b*R, (b+1)*R) into partial[b][0..N)
constexpr int COLSUM_ROWS = 2048;
constexpr int COLSUM_THREADS = 256;

__global__ void k_colsum_partial(const double* __restrict__ x, int N, long long rows, double* __restrict__ partial)
{
    __shared__ double red[COLSUM_THREADS];
    const long long r0 = (long long)blockIdx.x * COLSUM_ROWS;
    const long long r1 = min(rows, r0 + COLSUM_ROWS);
    const int tid = threadIdx.x;
    if (N <= COLSUM_THREADS) {
        const int nsub = COLSUM_THREADS / N;         // row sub-groups that run in parallel
        const int sub = tid / N, i = tid - sub * N;
        double s = 0.0;
        if (sub < nsub)
            for (long long r = r0 + sub; r < r1; r += nsub) s += x[r * N + i];
        red[tid] = s;
        __syncthreads();
        if (tid < N) {
            double tot = 0.0;
            for (int k = 0; k < nsub; ++k) tot += red[k * N + tid];
            partial[(long long)blockIdx.x * N + tid] = tot;
        }
    } else {
        for (int i = tid; i < N; i += COLSUM_THREADS) {
            double s = 0.0;
            for (long long r = r0; r < r1; ++r) s += x[r * N + i];
            partial[(long long)blockIdx.x * N + i] = s;
        }
    }
}

// stage 2: fixed-order sum over the first dimension of partial[nb][n]
__global__ void k_sum_partials(const double* __restrict__ partial, int nb, int n, double* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[(long long)b * n + k];
    out[k] = s;
}

// xi normaliser of frame t: S_t = sum_{i,j} ((alpha_t[i]*A[i][j])*p_{t+1}[j])*beta_{t+1}[j], added in the
// reference's row-major order (_hidden.c:168-176) without FMA contraction.
__global__ void k_xi_norm(const double* __restrict__ alpha, const double* __restrict__ beta,
                          const double* __restrict__ A, const double* __restrict__ pobs, int N, int T,
                          double* __restrict__ S)
{
    extern __shared__ double A_s[];
    for (int k = threadIdx.x; k < N * N; k += blockDim.x) A_s[k] = A[k];
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T - 1) return;
    const double* a = alpha + (long long)t * N;
    const double* b = beta + (long long)(t + 1) * N;
    const double* p = pobs + (long long)(t + 1) * N;
    double s = 0.0;
    for (int i = 0; i < N; ++i) {
        const double ai = a[i];
        for (int j = 0; j < N; ++j)
            s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(ai, A_s[i * N + j]), p[j]), b[j]));
    }
    S[t] = s;
}

constexpr int XI_FRAMES = 512;   // frames per block in k_xi_accumulate

// C_partial[b][i][j] = sum_{t in tile b} xi_t[i,j] / S_t      (thread per (i,j) pair, t ascending)
__global__ void k_xi_accumulate(const double* __restrict__ alpha, const double* __restrict__ beta,
                                const double* __restrict__ A, const double* __restrict__ pobs,
                                const double* __restrict__ S, int N, int T, double* __restrict__ partial)
{
    const int t0 = blockIdx.x * XI_FRAMES;
    const int t1 = min(T - 1, t0 + XI_FRAMES);
    for (int k = threadIdx.x; k < N * N; k += blockDim.x) {
        const int i = k / N, j = k - i * N;
        const double aij = A[k];
        double acc = 0.0;
        for (int t = t0; t < t1; ++t) {
            const double xi = __dmul_rn(__dmul_rn(__dmul_rn(alpha[(long long)t * N + i], aij),
                                                  pobs[(long long)(t + 1) * N + j]),
                                        beta[(long long)(t + 1) * N + j]);
            acc = __dadd_rn(acc, xi / S[t]);
        }
        partial[(long long)blockIdx.x * N * N + k] = acc;
    }
}

__global__ void k_update_pout(const int* __restrict__ sym, const double* __restrict__ w, long long rows, int N,
                              int M, double* __restrict__ pout)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= rows * N) return;
    const long long r = k / N;
    const int i = (int)(k - r * N);
    atomicAdd(pout + (long long)i * M + sym[r], w[k]);
}

__global__ void k_transpose(const double* __restrict__ in, int R, int Cc, double* __restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (long long)R * Cc) return;
    const int r = (int)(k / Cc), c = (int)(k - (long long)r * Cc);
    out[(long long)c * R + r] = in[k];
}

// M-step from the packed (all-reduced) E-step statistics [loglik | gamma0 (N) | C (N*N) | sum gamma | sum gamma d | sum gamma d^2]
// (d = o - mu_old).  Thread i owns row i: A[i,:] = C[i,:] / sum_j C[i,j] with the row sum taken left to right like numpy's
// C.sum(axis=1) on a short row; pi = gamma0 / sum gamma0 (maximum_likelihood.py:318-320); with `gauss`: mu = mu_old + wd / w,
// sigma = sqrt(wdd / w - (wd / w)^2), which equals the reference's two-pass estimate about the NEW mean (gaussian.py:244-270).
// out = [A (N*N) | pi (N) | mu (N) | sigma (N) | flags (1) | loglik (1)]; flags (a double holding small integers) counts what the host
// must look at: +1 per count C[i,j] <= mincount (the general-connectivity estimator applies, _tmatrix_disconnected.py:68-123),
// +1024 per sigma below machine epsilon or NaN (gaussian.py:271-272 raises).
__global__ void k_mstep_hmm(const double* __restrict__ stats, const double* __restrict__ mu_old, int N, int gauss,
                            double mincount, double* __restrict__ out)
{
    __shared__ double g0sum;
    __shared__ int flags;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int j = 0; j < N; ++j) s += stats[1 + j];
        g0sum = s;
        flags = 0;
    }
    __syncthreads();
    const double* Cm = stats + 1 + N;
    double* A = out;
    double* pi = out + (long long)N * N;
    double* mu = pi + N;
    double* sg = mu + N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double rs = 0.0;
        int low = 0;
        for (int j = 0; j < N; ++j) {
            const double c = Cm[(long long)i * N + j];
            rs += c;
            low += !(c > mincount);
        }
        for (int j = 0; j < N; ++j) A[(long long)i * N + j] = Cm[(long long)i * N + j] / rs;
        pi[i] = stats[1 + i] / g0sum;
        int f = low;
        if (gauss) {
            const double* w = stats + 1 + N + (long long)N * N;
            const double shift = w[N + i] / w[i];
            const double var = w[2 * N + i] / w[i] - shift * shift;
            const double s = sqrt(var > 0.0 ? var : 0.0);
            mu[i] = mu_old[i] + shift;
            sg[i] = s;
            if (!(s >= 2.220446049250313e-16)) f += 1024;
        } else {
            mu[i] = 0.0;
            sg[i] = 0.0;
        }
        if (f) atomicAdd(&flags, f);
    }
    __syncthreads();
    if (threadIdx.x == 0) { sg[N] = (double)flags; sg[N + 1] = stats[0]; }
}

// Row normalisation of the B numerators (discrete.py:214-215): one warp per state, fixed-order partial sums per lane and a
// fixed shuffle tree (deterministic).  Also writes the transposed table Bt (M,N) the discrete emission kernels gather from.
__global__ void k_mstep_rows(const double* __restrict__ Bnum, int N, int M, double* __restrict__ B, double* __restrict__ Bt)
{
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= N) return;
    const double* row = Bnum + (long long)i * M;
    double s = 0.0;
    for (int k = lane; k < M; k += 32) s += row[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    for (int k = lane; k < M; k += 32) {
        const double b = row[k] / s;
        B[(long long)i * M + k] = b;
        if (Bt) Bt[(long long)k * N + i] = b;
    }
}

inline unsigned nblk(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

int launch_gaussian_pobs(const double* obs, const double* mu, const double* sigma, int N, long long rows,
                         int ignore_outliers, double* pobs, cudaStream_t st)
{
    if (rows <= 0) return BHMM_OK;
    k_gaussian_pobs<<<nblk(rows, 128), 128, 0, st>>>(obs, mu, sigma, N, rows, ignore_outliers, pobs);
    return BHMM_OK;
}

int launch_discrete_pobs(const int* sym, const double* Bt, int N, int M, long long rows, int ignore_outliers,
                         double* pobs, cudaStream_t st)
{
    if (rows <= 0) return BHMM_OK;
    k_discrete_pobs<<<nblk(rows, 128), 128, 0, st>>>(sym, Bt, N, M, rows, ignore_outliers, pobs);
    return BHMM_OK;
}

int launch_state_probabilities(const double* alpha, const double* beta, int N, long long rows, double* gamma,
                               cudaStream_t st)
{
    if (rows <= 0) return BHMM_OK;
    k_state_probabilities<<<nblk(rows, 128), 128, 0, st>>>(alpha, beta, N, rows, gamma);
    return BHMM_OK;
}

int state_counts_blocks(long long rows) { return (int)((rows + COLSUM_ROWS - 1) / COLSUM_ROWS); }

int launch_state_counts(const double* gamma, int N, long long rows, double* counts, double* scratch, int* blocks,
                        cudaStream_t st)
{
    const int nb = state_counts_blocks(rows);
    if (blocks) *blocks = nb;
    if (nb <= 0) { cudaMemsetAsync(counts, 0, sizeof(double) * N, st); return BHMM_OK; }
    k_colsum_partial<<<nb, COLSUM_THREADS, 0, st>>>(gamma, N, rows, scratch);
    k_sum_partials<<<nblk(N, 128), 128, 0, st>>>(scratch, nb, N, counts);
    return BHMM_OK;
}

int transition_counts_blocks(int T) { return (T - 1 + XI_FRAMES - 1) / XI_FRAMES; }

int launch_transition_counts(const double* alpha, const double* beta, const double* A, const double* pobs, int N,
                             int T, double* C, double* scratch, cudaStream_t st)
{
    // scratch: [S (T) | partial (blocks * N*N)]
    if (T <= 1) { cudaMemsetAsync(C, 0, sizeof(double) * N * N, st); return BHMM_OK; }
    const size_t smem = sizeof(double) * (size_t)N * N;
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) return BHMM_ERR_UNSUPPORTED;
        if (cudaFuncSetAttribute(k_xi_norm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return BHMM_ERR_CUDA;
    }
    double* S = scratch;
    double* partial = scratch + T;
    const int nb = transition_counts_blocks(T);
    k_xi_norm<<<nblk(T - 1, 64), 64, smem, st>>>(alpha, beta, A, pobs, N, T, S);
    k_xi_accumulate<<<nb, 256, 0, st>>>(alpha, beta, A, pobs, S, N, T, partial);
    k_sum_partials<<<nblk((long long)N * N, 128), 128, 0, st>>>(partial, nb, N * N, C);
    return BHMM_OK;
}

int launch_update_pout(const int* sym, const double* w, long long rows, int N, int M, double* pout, cudaStream_t st)
{
    if (rows <= 0) return BHMM_OK;
    k_update_pout<<<nblk(rows * N, 256), 256, 0, st>>>(sym, w, rows, N, M, pout);
    return BHMM_OK;
}

int launch_mstep_hmm(const double* stats, const double* mu_old, int N, int gauss, double mincount, double* out, cudaStream_t st)
{
    k_mstep_hmm<<<1, 256, 0, st>>>(stats, mu_old, N, gauss, mincount, out);
    return BHMM_OK;
}

int launch_mstep_rows(const double* Bnum, int N, int M, double* B, double* Bt, cudaStream_t st)
{
    if (N <= 0 || M <= 0) return BHMM_OK;
    k_mstep_rows<<<(N + 3) / 4, 128, 0, st>>>(Bnum, N, M, B, Bt);
    return BHMM_OK;
}

int launch_transpose(const double* in, int R, int Cc, double* out, cudaStream_t st)
{
    if ((long long)R * Cc <= 0) return BHMM_OK;
    k_transpose<<<nblk((long long)R * Cc, 256), 256, 0, st>>>(in, R, Cc, out);
    return BHMM_OK;
}
