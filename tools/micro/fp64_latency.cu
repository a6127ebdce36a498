// FP64 DFMA dependent-issue latency and the ILP needed to saturate the pipe, per warp scheduler (one warp per SMSP).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double x, double y)
{
    double f[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) f[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) f[i] = fma(f[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int warps_per_sm)
{
    double* out; cudaMalloc(&out, 8 * 148 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<ILP><<<148, warps_per_sm * 32>>>(out, iters / 10, 0.999, 1e-3);
    cudaEventRecord(e0);
    k<ILP><<<148, warps_per_sm * 32>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n_per_warp = (double)iters * 8 * ILP;
    // cycles per DFMA issued by one scheduler (assumes ~1.9 GHz): warps per SMSP = warps_per_sm/4
    const double per_smsp = n_per_warp * (warps_per_sm / 4.0);
    printf("warps/SM %2d ILP %2d: %.3f ms, %.2f ns per dependent step, %.2f cycles@1.9GHz per DFMA per scheduler, %.1f TFLOP/s\n",
           warps_per_sm, ILP, ms, ms * 1e6 / (iters * 8.0), ms * 1e-3 * 1.9e9 / per_smsp,
           148.0 * warps_per_sm * 32 * n_per_warp * 2 / ms / 1e9);
    cudaFree(out);
}
int main()
{
    run<1>(4); run<2>(4); run<4>(4); run<8>(4); run<10>(4); run<16>(4); run<24>(4);
    run<1>(8); run<4>(8); run<10>(8); run<16>(8);
    run<10>(12); run<10>(16);
    return 0;
}
