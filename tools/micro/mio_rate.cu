// Throughput of the warp-uniform operand delivery paths the lane kernels use, per SM, on B200:
//   broadcast LDS.64 / LDS.128 (all lanes read the same shared-memory address), SHFL.IDX, and each of them mixed with
//   independent DFMA chains (does the register-file write-back of the loads steal FP64 issue cycles?).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mio_rate mio_rate.cu && ./mio_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int NF>   // MODE 0: none, 1: LDS.64 bcast, 2: LDS.128 bcast, 3: SHFL.IDX; NF DFMAs per 4 MIO ops
__global__ void k(double* out, int iters, double x, double y)
{
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 1e-3;
    __syncthreads();
    double f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-3 + i;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    int lane = threadIdx.x & 31;
    double acc = 0.0;
    int iacc = 0;
    int vals[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) vals[i] = lane * (i + 3);
    for (int it = 0; it < iters; ++it) {
        const unsigned addr = base + ((it & 7) << 8);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                if (MODE == 1) {
                    double a;
                    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(a) : "r"(addr + (u * 4 + m) * 16));
                    asm volatile("" ::"d"(a));
                } else if (MODE == 2) {
                    double a, b;
                    asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr + (u * 4 + m) * 16));
                    asm volatile("" ::"d"(a), "d"(b));
                } else if (MODE == 3) {
                    vals[u * 4 + m] = __shfl_sync(0xffffffffu, vals[u * 4 + m], it + m);
                }
            }
#pragma unroll
            for (int q = 0; q < NF; ++q) f[(u * NF + q) & 7] = fma(f[(u * NF + q) & 7], x, y);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += f[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) iacc ^= vals[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + iacc;
}

template <int MODE, int NF>
void run(const char* name, int sms, int wps, int iters, double ghz)
{
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * wps * 32);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, NF><<<sms, wps * 32>>>(out, iters / 10, 0.999, 1e-3);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE, NF><<<sms, wps * 32>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * ghz * 1e9;
    const double mio_per_sm = (MODE ? 16.0 : 0.0) * iters * wps;
    const double dfma_per_sched = 4.0 * NF * iters * wps / 4.0;
    printf("%-12s +%2d DFMA per 4 ops, %2d warps/SM: %8.3f ms", name, NF, wps, ms);
    if (MODE) printf("  %.2f cycles per MIO op per SM", cyc / mio_per_sm);
    if (NF) printf("  %.2f cycles per DFMA per scheduler", cyc / dfma_per_sched);
    printf("\n");
    cudaFree(out);
}

int main()
{
    int sms = 148, khz = 1965000;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    for (int wps : {8, 16}) {
        run<1, 0>("LDS.64 bc", sms, wps, 20000, ghz);
        run<2, 0>("LDS.128 bc", sms, wps, 20000, ghz);
        run<3, 0>("SHFL.IDX", sms, wps, 20000, ghz);
        run<0, 8>("none", sms, wps, 20000, ghz);
        run<1, 8>("LDS.64 bc", sms, wps, 20000, ghz);
        run<2, 8>("LDS.128 bc", sms, wps, 20000, ghz);
        run<3, 8>("SHFL.IDX", sms, wps, 20000, ghz);
        run<2, 4>("LDS.128 bc", sms, wps, 20000, ghz);
        run<2, 16>("LDS.128 bc", sms, wps, 20000, ghz);
    }
    return 0;
}
