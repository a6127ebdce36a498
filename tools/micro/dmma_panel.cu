// Prototype of the round-2 panel kernel for N = 32: one warp carries 8 chains, the product
// alpha' = alpha A is 32 mma.sync.m8n8k4.f64 per frame with the B fragments of A resident in registers, and the k-steps are
// ordered so that the accumulator fragment a lane holds after the product IS its A operand for the next frame
// (k-step ks = 2 nt + r  <->  state 8 nt + 2 (lane % 4) + r): no shuffles, no shared memory in the recursion.
// The program (1) checks one step against a plain loop and (2) times the dependent recursion (product, multiply by an
// emission-like factor, power-of-two free normalisation by the chain's sum) to see how close to the DMMA peak it gets.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_panel dmma_panel.cu && ./dmma_panel
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int N = 32;

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// alpha: [warp][chain 0..7][state 0..31] in and out; A: 32 x 32 row-major; steps of the recursion
__global__ void k_panel(const double* __restrict__ A, double* __restrict__ alpha, const double* __restrict__ pscale, int steps)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double* al = alpha + (size_t)warp * 8 * N + g * N;
    // B fragments: Bf[ks][nt] = A[sigma(ks, q)][8 nt + g], sigma(ks, q) = 8 (ks / 2) + 2 q + (ks % 2)
    double Bf[8][4];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) Bf[ks][nt] = A[(8 * (ks >> 1) + 2 * q + (ks & 1)) * N + 8 * nt + g];
    // the lane's 8 states of its chain: a[nt][r] = alpha[g][8 nt + 2 q + r]
    double a[4][2], p[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            a[nt][r] = al[8 * nt + 2 * q + r];
            p[nt][r] = pscale[8 * nt + 2 * q + r];
        }
    for (int s = 0; s < steps; ++s) {
        double d[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { d[nt][0] = 0.0; d[nt][1] = 0.0; }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma(d[nt][0], d[nt][1], a[ks >> 1][ks & 1], Bf[ks][nt]);
        double sum = 0.0;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int r = 0; r < 2; ++r) { d[nt][r] *= p[nt][r]; sum += d[nt][r]; }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const double rs = 1.0 / sum;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int r = 0; r < 2; ++r) a[nt][r] = d[nt][r] * rs;
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int r = 0; r < 2; ++r) al[8 * nt + 2 * q + r] = a[nt][r];
}

int main()
{
    int sms = 148, khz = 1965000;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    std::vector<double> hA(N * N), hp(N);
    srand(1);
    for (int i = 0; i < N; ++i) {
        double s = 0;
        for (int j = 0; j < N; ++j) { hA[i * N + j] = 0.05 + rand() / (double)RAND_MAX + (i == j ? 8.0 : 0.0); s += hA[i * N + j]; }
        for (int j = 0; j < N; ++j) hA[i * N + j] /= s;
        hp[i] = 0.2 + rand() / (double)RAND_MAX;
    }
    for (int wps : {4, 8, 16}) {
        const int warps = sms * wps, chains = warps * 8;
        std::vector<double> h0((size_t)chains * N), h1((size_t)chains * N);
        for (size_t k = 0; k < h0.size(); ++k) h0[k] = 0.1 + rand() / (double)RAND_MAX;
        double *dA, *dal, *dp;
        cudaMalloc(&dA, sizeof(double) * N * N); cudaMalloc(&dp, sizeof(double) * N);
        cudaMalloc(&dal, sizeof(double) * h0.size());
        cudaMemcpy(dA, hA.data(), sizeof(double) * N * N, cudaMemcpyHostToDevice);
        cudaMemcpy(dp, hp.data(), sizeof(double) * N, cudaMemcpyHostToDevice);
        // (1) one step against the plain loop
        cudaMemcpy(dal, h0.data(), sizeof(double) * h0.size(), cudaMemcpyHostToDevice);
        k_panel<<<sms, wps * 32>>>(dA, dal, dp, 1);
        cudaMemcpy(h1.data(), dal, sizeof(double) * h1.size(), cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int c = 0; c < chains; c += 97) {
            double v[N], s = 0;
            for (int j = 0; j < N; ++j) {
                double m = 0;
                for (int i = 0; i < N; ++i) m += h0[(size_t)c * N + i] * hA[i * N + j];
                v[j] = m * hp[j]; s += v[j];
            }
            for (int j = 0; j < N; ++j) worst = fmax(worst, fabs(v[j] / s - h1[(size_t)c * N + j]) / (v[j] / s));
        }
        // (2) timing of the dependent recursion
        const int steps = 20000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_panel<<<sms, wps * 32>>>(dA, dal, dp, 100);
        cudaEventRecord(e0);
        k_panel<<<sms, wps * 32>>>(dA, dal, dp, steps);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double chain_steps = (double)chains * steps;
        printf("%2d warps/SM: one-step max rel. error %.2e; %.3f ms for %d steps: %.2f G chain-frames/s, %.1f TFLOP/s in the "
               "products, %.0f cycles per step and warp (cudaError %d)\n", wps, worst, ms, steps, chain_steps / ms / 1e6,
               chain_steps * 2.0 * N * N / ms / 1e9, ms * 1e-3 * khz * 1e3 / steps, (int)cudaGetLastError());
        cudaFree(dA); cudaFree(dal); cudaFree(dp);
    }
    return 0;
}
