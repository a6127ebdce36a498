// bhmm_b200/csrc/certify.cu -- certification of chain hand-overs and the deterministic E-step reduction.
//
// Time-chunking is only allowed to change results below the stated tolerance.  A chain that does not
// start its trajectory is started from a warmed-up filter; afterwards its hand-over vector (the alpha at
// the frame before the chain, or the beta at the frame after it) is compared, component by component and
// RELATIVELY, with the value the neighbouring chain actually computed.  Chains that disagree by more than
// `tol` are listed and re-run from the exact neighbour value (Chains.exact = 1) until the list is empty;
// in the worst case (a non-mixing model) this degenerates to the sequential recursion, never to a wrong
// answer.  The reference has no counterpart (it is serial in t, _hidden.c:41-63).
#include "common.cuh"
#include "kernels.h"

namespace {

// Besides pass / fail the kernel estimates how long a warm-up each hand-over NEEDS: the mismatch decays geometrically
// with the warm-up length (m ~ rho^w), so w' = w log(tol) / log(m) frames would have sufficed.  The maximum over all
// chains (out[2]) lets the host keep the warm-up a safety factor above the need instead of discovering it by failing
// (a failed backward pass costs a whole extra pass).  Two tolerances: `tol` is what the warm-up length is steered to (the
// need estimate), `tol_fail` >= tol is the mismatch above which a chain is listed for repair.
__global__ void k_certify(Chains ch, int n_total, int N, int dir, const double* __restrict__ hand_used,
                          const double* __restrict__ hand_end, double tol, double tol_fail, int* __restrict__ fail_list,
                          unsigned long long* __restrict__ out)
{
    // One verdict per chain, reduced over the warp before it touches the three global words: with one atomic per chain the
    // 37 888 chains of a C3 wave serialised on the same addresses and the kernel took 55 us, twice per E-step.
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = c < n_total;
    int nb = 0;
    if (active) {
        if (dir > 0) {
            nb = c - 1;
            // a chain that starts its trajectory is exact by construction; the first chain of an owned range (time-sharded
            // trajectory) has its neighbour on another shard, the hand-over is certified there (bhmm_b200_batch_border_handovers)
            if (ch.t0[c] == 0 || c == 0 || ch.row0[nb] + ch.len[nb] != ch.row0[c]) active = false;
        } else {
            nb = c + 1;
            if (ch.t0[c] + ch.len[c] >= ch.T[c]) active = false;          // ends its trajectory: exact by construction
            else if (nb >= n_total || ch.row0[c] + ch.len[c] != ch.row0[nb]) active = false;
        }
    }
    unsigned long long worst_bits = 0ULL, need_bits = 0ULL;
    if (active) {
        const double* u = hand_used + (long long)c * N;
        const double* v = hand_end + (long long)nb * N;
        double worst = 0.0;
        for (int i = 0; i < N; ++i) worst = fmax(worst, rel_mismatch(u[i], v[i]));
        worst_bits = (unsigned long long)__double_as_longlong(worst);       // non-negative doubles order like their bits
        if (worst > tol_fail) {
            const unsigned long long slot = atomicAdd(out, 1ULL);           // (rare)
            fail_list[slot] = c;
        }
        // warm-up actually available to this chain (it is cut short at the trajectory's ends, where the start is exact)
        const int avail = dir > 0 ? ch.t0[c] : ch.T[c] - (ch.t0[c] + ch.len[c]);
        const int w = min(ch.warmv ? ch.warmv[c] : ch.warm, avail);
        double need = 0.0;
        if (worst >= 0.5) need = 2.0 * w + 32.0;
        else if (worst > 0.0) need = w * log(tol) / log(worst);
        need_bits = (unsigned long long)need;
    }
    for (int off = 16; off > 0; off >>= 1) {
        worst_bits = max(worst_bits, __shfl_xor_sync(0xffffffffu, worst_bits, off));
        need_bits = max(need_bits, __shfl_xor_sync(0xffffffffu, need_bits, off));
    }
    if ((threadIdx.x & 31) == 0) {
        if (worst_bits) atomicMax(out + 1, worst_bits);
        if (need_bits) atomicMax(out + 2, need_bits);
    }
}

// stats = [loglik | gamma0 (N) | C (N*N) | sum gamma (N) | sum gamma d (N) | sum gamma d^2 (N)]
// partial rows are [C' (N*N) | gamma0 | sum gamma | sum gamma d | sum gamma d^2]; C = A o C'.
// One warp per statistic: lanes stride over the partial rows, fixed shuffle tree => deterministic.  The last block
// reduces the per-chain log-likelihoods (fixed partition + fixed tree).
__global__ void k_finalize_stats(const double* __restrict__ partials, int grid, const double* __restrict__ chain_ll,
                                 int n_chains, const double* __restrict__ A, int N, double* __restrict__ stats)
{
    __shared__ double red[256];
    const int nstat = N * N + 4 * N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x < gridDim.x - 1) {
        const int k = blockIdx.x * 8 + warp;
        if (k >= nstat) return;
        double s = 0.0;
        for (int b = lane; b < grid; b += 32) s += partials[(long long)b * nstat + k];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) {
            if (k < N * N) stats[1 + N + k] = A[k] * s;
            else if (k < N * N + N) stats[1 + (k - N * N)] = s;
            else stats[1 + k] = s;
        }
        return;
    }
    const int per = (n_chains + blockDim.x - 1) / blockDim.x;
    double s = 0.0;
    const int lo = threadIdx.x * per, hi = min(n_chains, lo + per);
    for (int c = lo; c < hi; ++c) s += chain_ll[c];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) stats[0] = red[0];
}

}  // namespace

int launch_certify(const Chains& ch, int n_total, int N, int dir, const double* hand_used, const double* hand_end,
                   double tol, double tol_fail, int* fail_list, unsigned long long* out, cudaStream_t st)
{
    cudaMemsetAsync(out, 0, 4 * sizeof(unsigned long long), st);
    if (n_total <= 0) return BHMM_OK;
    k_certify<<<(n_total + 127) / 128, 128, 0, st>>>(ch, n_total, N, dir, hand_used, hand_end, tol, fmax(tol, tol_fail), fail_list, out);
    return BHMM_OK;
}

int launch_finalize_stats(const double* partials, int grid, const double* chain_ll, int n_chains, const double* A,
                          int N, double* stats, cudaStream_t st)
{
    const int nstat = N * N + 4 * N;
    k_finalize_stats<<<(nstat + 7) / 8 + 1, 256, 0, st>>>(partials, grid, chain_ll, n_chains, A, N, stats);
    return BHMM_OK;
}
