// bhmm_b200/csrc/lane_dispatch.cu -- run-time dispatch of the lane family over the number of states.
#include "common.cuh"
#include "kernels.h"

#define DECL(NN) int launch_lane_##NN(const LaneArgs&, const LaneHostParams&, int, int, cudaStream_t);
DECL(1) DECL(2) DECL(3) DECL(4) DECL(5) DECL(6) DECL(7) DECL(8) DECL(9) DECL(10) DECL(11) DECL(12) DECL(13) DECL(14) DECL(15) DECL(16)
#undef DECL

namespace {
constexpr int LANE_THREADS = 64;   // must match lane_kernels.cuh

// gamma0 = sum over the chains that start a trajectory of their gamma at frame 0.  One block per state; fixed partition
// of the chains over 256 threads and a fixed tree: deterministic.
__global__ void k_add_gamma0(Chains ch, int n_total, int N, const double* __restrict__ g0buf, double* __restrict__ stats)
{
    __shared__ double red[256];
    const int i = blockIdx.x;
    const int per = (n_total + 255) / 256;
    const int lo = threadIdx.x * per, hi = min(n_total, lo + per);
    double s = 0.0;
    for (int c = lo; c < hi; ++c)
        if (ch.t0[c] == 0) s += g0buf[(long long)c * N + i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) stats[1 + i] += red[0];
}

// sums[k] = sum over the partial rows of column k (fixed order): per-state sum o and sum o^2 of the sampled paths
__global__ void k_sum_moments(const double* __restrict__ partials, int rows, int n, double* __restrict__ sums)
{
    const int k = threadIdx.x;
    if (k >= n) return;
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += partials[(long long)r * n + k];
    sums[k] = s;
}
}  // namespace

int lane_blocks(int n_chains) { return (n_chains + LANE_THREADS - 1) / LANE_THREADS; }

bool lane_supported(int N, int em) { return N >= 1 && N <= LANE_MAX_N && (em == EM_GAUSS || em == EM_DISC); }

int launch_lane(const LaneArgs& a, const LaneHostParams& hp, int N, int em, int what, cudaStream_t st)
{
    if (em == EM_GAUSS && hp.sigma) {
        // the lane kernels reproduce the reference's zero densities (and its outlier rule) exactly only while the logarithm
        // of the normalisation constant stays below 37 (lane_kernels.cuh:emission_gauss), and the lazily rescaled recursion
        // has head-room for one step's growth of N / (sigma sqrt(2 pi)) up to 1e100: smaller sigmas belong to the team
        // kernels (bhmm_b200_batch_set_family(b, 0); engine.TrajectoryBatch switches by itself)
        for (int j = 0; j < N; ++j) {
            if (!(hp.sigma[j] >= 1e-16)) {
                bhmm_set_error(BHMM_ERR_UNSUPPORTED, "lane kernels: every sigma must be >= 1e-16 (and not NaN); use the team family (bhmm_b200_batch_set_family)");
                return BHMM_ERR_UNSUPPORTED;
            }
        }
    }
    switch (N) {
#define CASE(NN) case NN: return launch_lane_##NN(a, hp, em, what, st);
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
        CASE(14) CASE(15) CASE(16)
#undef CASE
    }
    return BHMM_ERR_UNSUPPORTED;
}

int launch_add_gamma0(const Chains& ch, int n_total, int N, const double* g0buf, double* stats, cudaStream_t st)
{
    k_add_gamma0<<<N, 256, 0, st>>>(ch, n_total, N, g0buf, stats);
    return BHMM_OK;
}

int launch_lane_sum_moments(const double* partials, int rows, int N, double* sums, cudaStream_t st)
{
    k_sum_moments<<<1, 64, 0, st>>>(partials, rows, 2 * N, sums);
    return BHMM_OK;
}

int lane_threads() { return LANE_THREADS; }

int lane_blocks_per_sm(int N, int em)
{
    static int cache[LANE_MAX_N + 1][3] = {};
    if (N < 1 || N > LANE_MAX_N || em < 0 || em > 2) return 4;
    if (cache[N][em] == 0) {
        LaneArgs a{};
        LaneHostParams hp{nullptr, nullptr, nullptr, nullptr};
        const int r = launch_lane(a, hp, N, em, LANE_QUERY_BLOCKS, nullptr);
        cache[N][em] = (r < 0) ? -r : 4;
    }
    return cache[N][em];
}
