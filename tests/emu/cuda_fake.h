// tests/emu/cuda_fake.h -- the few device intrinsics warp_emu.h does not have yet, launch configuration and dynamic shared
// memory for hostified sources (hostify.py).  The fake CUDA runtime itself (cudaMalloc = malloc ...) is cuda_fake.cpp.
#pragma once
#include "warp_emu.h"

#undef __maxnreg__
#define __maxnreg__(...)

template <class T>
inline cudaError_t cudaFuncSetAttribute(T*, cudaFuncAttribute, int) { return cudaSuccess; }

namespace emu {
struct Cfg { int grid, block; size_t smem; };
inline Cfg cfg(int grid, int block, size_t smem = 0, cudaStream_t = 0) { return Cfg{grid, block, smem}; }
inline std::vector<double>& dyn_buf() { static std::vector<double> b; return b; }
inline double* dyn_smem() { return dyn_buf().data(); }
inline void launch(const Cfg& c, const std::function<void()>& body)
{
    dyn_buf().assign(c.smem / sizeof(double) + 2, 0.0);
    static const bool trace = getenv("EMU_TRACE") != nullptr;
    if (trace) fprintf(stderr, "emu launch: grid %d block %d smem %zu\n", c.grid, c.block, c.smem);
    launch(c.grid, c.block, body);
}
}  // namespace emu

inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline long long atomicAdd(long long* p, long long v) { const long long o = *p; *p = o + v; return o; }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; if (v > o) *p = v; return o; }
inline int atomicMax(int* p, int v) { const int o = *p; if (v > o) *p = v; return o; }
inline int atomicExch(int* p, int v) { const int o = *p; *p = v; return o; }
inline unsigned int __umulhi(unsigned int a, unsigned int b) { return (unsigned int)(((unsigned long long)a * b) >> 32); }
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __shfl_sync(unsigned, int v, int src)
{
    emu::Warp& w = emu::my_warp();
    w.si[emu::lane_id()] = v;
    emu::rendezvous(w.g);
    const int r = w.si[src & 31];
    emu::rendezvous(w.g);
    return r;
}
inline double __shfl_sync(unsigned, double v, int src)
{
    emu::Warp& w = emu::my_warp();
    w.sd[emu::lane_id()] = v;
    emu::rendezvous(w.g);
    const double r = w.sd[src & 31];
    emu::rendezvous(w.g);
    return r;
}
inline void __threadfence_block() {}
inline void __threadfence() {}
