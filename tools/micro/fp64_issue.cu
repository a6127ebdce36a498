// Does an FP64 warp instruction block the scheduler's issue port for its whole 2-cycle slot, or can integer
// instructions issue in its shadow?  8 independent DFMA chains per thread, R independent integer ops (LOP3/IADD mix,
// 4 chains) per DFMA, for several warps per scheduler.  Prints cycles per DFMA per scheduler.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu && ./fp64_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int R>
__global__ void k(double* out, int iters, double x, double y, unsigned m)
{
    double f[8];
    unsigned g[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            f[i] = fma(f[i], x, y);
#pragma unroll
            for (int r = 0; r < R; ++r) g[(i * R + r) & 3] = (g[(i * R + r) & 3] ^ m) + (unsigned)it;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + g[0] + g[1] + g[2] + g[3];
}

template <int R>
void run(int sms, int wps, int iters, double ghz)
{
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * wps * 32);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<R><<<sms, wps * 32>>>(out, iters / 10, 0.999, 1e-3, 0x5bd1e995u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<R><<<sms, wps * 32>>>(out, iters, 0.999, 1e-3, 0x5bd1e995u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma_per_sched = (double)wps / 4 * iters * 8.0;
    printf("warps/SM %2d  int ops per DFMA %d (x2 SASS: LOP3+IADD): %.3f ms, %.2f cycles per DFMA per scheduler (at %.3f GHz)\n",
           wps, R, ms, ms * 1e-3 * ghz * 1e9 / dfma_per_sched, ghz);
    cudaFree(out);
}

int main()
{
    int sms = 148, khz = 1965000;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    for (int wps : {4, 8, 12, 16}) {
        run<0>(sms, wps, 40000, ghz);
        run<1>(sms, wps, 40000, ghz);
        run<2>(sms, wps, 40000, ghz);
    }
    return 0;
}
