// DFMA issue rate per scheduler when the three source operands are distinct registers (the outer-product and
// matrix-vector patterns of the lane kernels) instead of one running value and two loop constants.
//   MODE 0: f[i] = fma(f[i], x, y)                (constants: reuse cache / uniform registers)
//   MODE 1: c[i][j] = fma(u[i], w[j], c[i][j])    (3 x 10 outer product, accumulators in registers)
//   MODE 2: b[i] = fma(A[i][j], w[j], b[i])       (A from shared memory: LDS.128 broadcast, 10 x 10)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_operands fp64_operands.cu && ./fp64_operands
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* out, int iters, double x, double y)
{
    __shared__ double As[100];
    for (int i = threadIdx.x; i < 100; i += blockDim.x) As[i] = 0.01 + 1e-4 * i;
    __syncthreads();
    double f[10], u[3], w[10], c[3][10];
#pragma unroll
    for (int i = 0; i < 10; ++i) { f[i] = threadIdx.x * 1e-3 + i; w[i] = 1.0 + 1e-3 * i + threadIdx.x * 1e-6; }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        u[i] = 1e-3 * (i + 1);
#pragma unroll
        for (int j = 0; j < 10; ++j) c[i][j] = 0.0;
    }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int i = 0; i < 10; ++i) f[i] = fma(f[i], x, y);
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 10; ++j)
#pragma unroll
                for (int i = 0; i < 3; ++i) c[i][j] = fma(u[i], w[j], c[i][j]);
            u[0] += x * 1e-9;
        } else {
#pragma unroll
            for (int i = 0; i < 10; ++i) f[i] = As[i * 10] * w[0];
#pragma unroll
            for (int j = 1; j < 10; ++j)
#pragma unroll
                for (int i = 0; i < 10; ++i) f[i] = fma(As[i * 10 + j], w[j], f[i]);
#pragma unroll
            for (int i = 0; i < 10; ++i) w[i] = f[i] * x;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) s += f[i] + w[i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 10; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int sms, int wps, int iters, double ghz, double dfma_per_iter)
{
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * wps * 32);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, wps * 32>>>(out, iters / 10, 0.999, 1e-3);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<sms, wps * 32>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * ghz * 1e9;
    printf("%-34s %2d warps/SM: %7.3f ms  %.2f cycles per FP64 instruction per scheduler, %.1f cycles per iteration per warp\n",
           name, wps, ms, cyc / (dfma_per_iter * iters * wps / 4.0), cyc / iters);
    cudaFree(out);
}

int main()
{
    int sms = 148, khz = 1965000;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    for (int wps : {4, 8, 12}) {
        run<0>("fma(f, const, const) x30", sms, wps, 40000, ghz, 30);
        run<1>("outer product 3x10 (registers)", sms, wps, 40000, ghz, 30);
        run<2>("matvec 10x10, A from shared memory", sms, wps, 20000, ghz, 110);
    }
    return 0;
}
