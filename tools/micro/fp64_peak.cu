// Measures the FP64 rates the roofline of the N >= 10 kernels needs ("FP64 peak is not in
// MEASURED_PEAKS.json - builder must measure it"):
//   1. vector DFMA throughput (independent chains per thread),
//   2. tensor DMMA m8n8k4 throughput,
//   3. both interleaved in the same warps: do the two pipes overlap?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>   // 0: DFMA only, 1: DMMA only, 2: both
__global__ void k(double* out, int iters, double x, double y)
{
    double f[8], c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x * 1e-3 + i; c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE != 1) { f[i] = fma(f[i], x, y); f[i] = fma(f[i], x, y); f[i] = fma(f[i], x, y); f[i] = fma(f[i], x, y); }
            if (MODE != 0) dmma(c[i][0], c[i][1], x, y);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i] + c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int blocks, int threads, int iters)
{
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, iters / 10, 0.999, 1e-3);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * threads / 32;
    const double dfma = (MODE != 1) ? warps * iters * 8.0 * 4 * 32 * 2 : 0;       // flops
    const double dmm = (MODE != 0) ? warps * iters * 8.0 * (8 * 8 * 4 * 2) : 0;   // flops
    printf("%-22s blocks=%d threads=%d: %.3f ms  DFMA %.2f TFLOP/s  DMMA %.2f TFLOP/s  (cudaError %d)\n", name, blocks,
           threads, ms, dfma / ms / 1e9, dmm / ms / 1e9, (int)cudaGetLastError());
    cudaFree(out);
}

int main()
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int wps : {4, 8, 16, 32}) {
        printf("--- %d warps per SM\n", wps);
        run<0>("DFMA only", sms, wps * 32, 20000);
        run<1>("DMMA only", sms, wps * 32, 20000);
        run<2>("DFMA + DMMA interleaved", sms, wps * 32, 20000);
    }
    return 0;
}
