// tests/emu/warp_emu.h -- runs a CUDA kernel's SOURCE on the CPU, one block at a time, every CUDA thread a fiber.
//
// Test infrastructure (like oracle/): it exists because the panel kernels were written with no GPU time left; executing
// their actual source -- control flow, indexing, collectives -- is worth more than reading it again.  Every warp-level
// intrinsic the kernels use (__shfl_xor_sync, __ballot_sync, __any_sync, __reduce_max_sync, __syncwarp, mma.sync
// m8n8k4.f64 through the kernels' own dmma() hook) and __syncthreads is a rendez-vous of the warp's / block's fibers;
// fibers are switched cooperatively in lane order (ucontext), so a run is deterministic and atomics are trivially atomic.
// A rendez-vous that can never complete (a divergent collective -- a hang on hardware) aborts with a message.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#include <cuda_runtime.h>   // vector types; __device__, __global__ ... expand to nothing under g++

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

using std::max;
using std::min;

namespace emu {

struct Idx { unsigned x = 0, y = 0, z = 0; };
struct Group { int n = 0, count = 0; unsigned long gen = 0; };
struct Warp {
    Group g;
    double sd[32];
    double sa[32], sb[32];
    unsigned su[32];
    int si[32];
};
struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    Idx tid;
    bool done = false;
};

inline std::vector<Fiber>& fibers() { static std::vector<Fiber> f; return f; }
inline std::vector<Warp>& warps() { static std::vector<Warp> w; return w; }
inline Group& block_group() { static Group g; return g; }
inline ucontext_t& sched_ctx() { static ucontext_t c; return c; }
inline int& current() { static int c = 0; return c; }
inline unsigned long& progress() { static unsigned long p = 0; return p; }
inline Idx& block_idx() { static Idx i; return i; }
inline Idx& block_dim() { static Idx i; return i; }
inline Idx& grid_dim() { static Idx i; return i; }
inline std::function<void()>& body() { static std::function<void()> b; return b; }

inline Fiber& cur() { return fibers()[current()]; }
inline Warp& my_warp() { return warps()[cur().tid.x >> 5]; }
inline int lane_id() { return (int)(cur().tid.x & 31); }
inline void yield() { swapcontext(&cur().ctx, &sched_ctx()); }

inline void rendezvous(Group& g)
{
    const unsigned long my = g.gen;
    if (++g.count == g.n) { g.count = 0; ++g.gen; ++progress(); }
    else while (g.gen == my) yield();
}

inline void trampoline()
{
    body()();
    cur().done = true;
    ++progress();
    swapcontext(&cur().ctx, &sched_ctx());
}

// run `kernel_body` as grid x threads CUDA threads, block after block
inline void launch(int grid, int threads, const std::function<void()>& kernel_body)
{
    body() = kernel_body;
    grid_dim().x = (unsigned)grid;
    block_dim().x = (unsigned)threads;
    for (int b = 0; b < grid; ++b) {
        block_idx().x = (unsigned)b;
        fibers().assign(threads, Fiber());
        warps().assign((threads + 31) / 32, Warp());
        for (int w = 0; w < (int)warps().size(); ++w) warps()[w].g.n = std::min(32, threads - 32 * w);
        block_group() = Group();
        block_group().n = threads;
        for (int t = 0; t < threads; ++t) {
            Fiber& f = fibers()[t];
            f.tid.x = (unsigned)t;
            f.stack.resize(512 * 1024);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack.data();
            f.ctx.uc_stack.ss_size = f.stack.size();
            f.ctx.uc_link = &sched_ctx();
            makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        for (;;) {
            const unsigned long before = progress();
            int alive = 0;
            for (int t = 0; t < threads; ++t) {
                if (fibers()[t].done) continue;
                ++alive;
                current() = t;
                swapcontext(&sched_ctx(), &fibers()[t].ctx);
            }
            if (alive == 0) break;
            if (progress() == before) {
                fprintf(stderr, "warp_emu: deadlock in block %d (a collective some threads never reach)\n", b);
                abort();
            }
        }
    }
}

}  // namespace emu

#define threadIdx (emu::cur().tid)
#define blockIdx (emu::block_idx())
#define blockDim (emu::block_dim())
#define gridDim (emu::grid_dim())

inline void __syncthreads() { emu::rendezvous(emu::block_group()); }
inline int __syncthreads_or(int pred)
{
    static int acc[2];
    static unsigned long phase = 0;
    // two rendez-vous: everybody contributes, everybody reads; the accumulator alternates so that a fast fiber's next call
    // cannot clear a value a slow fiber has not read yet
    const unsigned long my = phase;
    if (pred) acc[my & 1] = 1;
    emu::rendezvous(emu::block_group());
    const int r = acc[my & 1];
    if (phase == my) { phase = my + 1; acc[(my + 1) & 1] = 0; }
    emu::rendezvous(emu::block_group());
    return r;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::rendezvous(emu::my_warp().g); }

inline double __shfl_xor_sync(unsigned, double v, int off)
{
    emu::Warp& w = emu::my_warp();
    const int l = emu::lane_id();
    w.sd[l] = v;
    emu::rendezvous(w.g);
    const double r = w.sd[l ^ off];
    emu::rendezvous(w.g);
    return r;
}
inline unsigned long long __shfl_xor_sync(unsigned, unsigned long long v, int off)
{
    // 64-bit integer payload through the same slots, bit for bit (memcpy: no conversion, no NaN canonicalisation)
    emu::Warp& w = emu::my_warp();
    const int l = emu::lane_id();
    memcpy(&w.sd[l], &v, sizeof(v));
    emu::rendezvous(w.g);
    unsigned long long r;
    memcpy(&r, &w.sd[l ^ off], sizeof(r));
    emu::rendezvous(w.g);
    return r;
}
inline unsigned __ballot_sync(unsigned, bool pred)
{
    emu::Warp& w = emu::my_warp();
    w.su[emu::lane_id()] = pred ? 1u : 0u;
    emu::rendezvous(w.g);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= w.su[i] << i;
    emu::rendezvous(w.g);
    return r;
}
inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0u; }
inline int __reduce_max_sync(unsigned, int v)
{
    emu::Warp& w = emu::my_warp();
    w.si[emu::lane_id()] = v;
    emu::rendezvous(w.g);
    int r = w.si[0];
    for (int i = 1; i < 32; ++i) r = std::max(r, w.si[i]);
    emu::rendezvous(w.g);
    return r;
}
// mma.sync.aligned.m8n8k4.row.col.f64: A[row = lane / 4][k = lane % 4], B[k = lane % 4][n = lane / 4],
// C/D[row = lane / 4][col = 2 (lane % 4) + {0, 1}]
inline void emu_dmma(double& d0, double& d1, double a, double b)
{
    emu::Warp& w = emu::my_warp();
    const int l = emu::lane_id(), g = l >> 2, q = l & 3;
    w.sa[l] = a;
    w.sb[l] = b;
    emu::rendezvous(w.g);
    for (int k = 0; k < 4; ++k) {
        d0 = std::fma(w.sa[4 * g + k], w.sb[4 * (2 * q) + k], d0);
        d1 = std::fma(w.sa[4 * g + k], w.sb[4 * (2 * q + 1) + k], d1);
    }
    emu::rendezvous(w.g);
}
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
inline double atomicAdd(double* p, double v) { const double o = *p; *p = o + v; return o; }
inline int __double2hiint(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b >> 32); }
inline int __double2loint(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b & 0xffffffff); }
inline long long __double_as_longlong(double x) { long long b; memcpy(&b, &x, 8); return b; }
inline double __longlong_as_double(long long b) { double x; memcpy(&x, &b, 8); return x; }
