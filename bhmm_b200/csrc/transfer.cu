// bhmm_b200/csrc/transfer.cu -- host <-> device movers for the estimators' one-off transfers.
//
// MaximumLikelihoodEstimator / BayesianHMMSampler receive a LIST of pageable host arrays (maximum_likelihood.py:60-144) and
// hand back one hidden-state path per trajectory (maximum_likelihood.py:332-352, :439).  A cudaMemcpy from / to pageable
// memory is staged by the driver through its own bounce buffer on ONE thread: measured on the pool's B200 boxes 5.8 GB/s up
// (C3: 819 MB in 0.14 s) and 2 GB/s down into a fresh allocation (410 MB of paths in 0.2 s, most of it page faults of the
// destination) -- together more than the 20 EM iterations in between (0.17 s).  Here a few worker threads each own two pinned
// staging slots and a stream: a worker copies its piece of the concatenated byte stream between the caller's arrays and a
// slot (this is also what touches the destination's fresh pages, in parallel) while the DMA engine moves its other slot.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "host_common.h"

#if defined(__x86_64__) || defined(__SSE2__)
#include <emmintrin.h>
#define BHMM_HAVE_SSE2 1
#endif

namespace {

// memcpy with streaming (non-temporal) stores.  Neither destination is read again by this CPU -- a staging slot is consumed
// by the DMA engine, a caller's fresh array by whoever uses the paths later -- so ordinary stores would first fetch every
// destination line into the cache (read-for-ownership) and evict it again: one extra pass over the data through the host's
// memory system, which is what the transfers of 8 ranks on one host compete for.  BHMM_B200_NT_COPY=0: plain memcpy.
int g_nt_copy = -1;
void stream_copy(char* dst, const char* src, size_t n)
{
#ifdef BHMM_HAVE_SSE2
    if (g_nt_copy && n >= 4096) {
        const size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
        if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
        size_t blocks = n / 64;
        while (blocks--) {
            const __m128i a = _mm_loadu_si128((const __m128i*)src), b = _mm_loadu_si128((const __m128i*)(src + 16));
            const __m128i c = _mm_loadu_si128((const __m128i*)(src + 32)), d = _mm_loadu_si128((const __m128i*)(src + 48));
            _mm_stream_si128((__m128i*)dst, a);
            _mm_stream_si128((__m128i*)(dst + 16), b);
            _mm_stream_si128((__m128i*)(dst + 32), c);
            _mm_stream_si128((__m128i*)(dst + 48), d);
            src += 64; dst += 64;
        }
        n &= 63;
        if (n) memcpy(dst, src, n);
        _mm_sfence();
        return;
    }
#endif
    memcpy(dst, src, n);
}

size_t SLOT_BYTES = (size_t)2 << 20;               // one staging slot, 2 per worker (BHMM_B200_STAGE_KB overrides, first use)
constexpr int MAX_WORKERS = 16;

struct Staging {
    char* base = nullptr;       // MAX... workers x 2 slots, pinned, allocated on first use and kept
    int workers = 0;
    int device = -1;
    cudaStream_t stream[MAX_WORKERS] = {};
    cudaEvent_t done[MAX_WORKERS][2] = {};
};
Staging g_stage;
std::mutex g_stage_mutex;       // one transfer at a time per process (the slots are shared)

int stage_ensure(int workers)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (g_stage.base && (g_stage.workers < workers || g_stage.device != dev)) {
        for (int w = 0; w < g_stage.workers; ++w) {
            cudaStreamDestroy(g_stage.stream[w]);
            cudaEventDestroy(g_stage.done[w][0]);
            cudaEventDestroy(g_stage.done[w][1]);
        }
        cudaFreeHost(g_stage.base);
        g_stage = Staging();
    }
    if (g_nt_copy < 0) { const char* e = getenv("BHMM_B200_NT_COPY"); g_nt_copy = (e && e[0] == '0') ? 0 : 1; }
    static bool env_read = false;
    if (!env_read) {
        env_read = true;
        if (const char* e = getenv("BHMM_B200_STAGE_KB")) {
            const long kb = atol(e);
            if (kb >= 64 && kb <= (1 << 20)) SLOT_BYTES = (size_t)kb << 10;
        }
    }
    if (!g_stage.base) {
        CUDA_TRY(cudaHostAlloc((void**)&g_stage.base, SLOT_BYTES * 2 * workers, cudaHostAllocDefault));
        for (int w = 0; w < workers; ++w) {
            CUDA_TRY(cudaStreamCreateWithFlags(&g_stage.stream[w], cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&g_stage.done[w][0], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&g_stage.done[w][1], cudaEventDisableTiming));
        }
        g_stage.workers = workers;
        g_stage.device = dev;
    }
    return BHMM_OK;
}

// workers the caller asked for (slots are allocated for all of them at the first transfer, however small that one is: a tiny
// first transfer must not leave the pinned allocation of the other workers' slots to the first big one)
int asked_workers(int threads)
{
    const int w = threads > 0 ? threads : (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
    return std::max(1, std::min(w, MAX_WORKERS));
}
// ... and those that have a piece to move
int pick_workers(int threads, long long total_bytes)
{
    const long long pieces = (total_bytes + (long long)SLOT_BYTES - 1) / (long long)SLOT_BYTES;
    return (int)std::max<long long>(1, std::min<long long>(asked_workers(threads), pieces));
}

// The concatenated byte stream of K host arrays, addressed by global offset.
struct Ragged {
    const void* const* ptr;
    std::vector<long long> start;       // K + 1 prefix sums of the byte counts
    int K;
    // copy bytes [off, off + n) of the stream to / from `buf`
    template <bool TO_STAGE>
    void copy(char* buf, long long off, long long n) const
    {
        int k = (int)(std::upper_bound(start.begin(), start.end(), off) - start.begin()) - 1;
        while (n > 0) {
            const long long in = off - start[k];
            const long long take = std::min(n, start[k + 1] - off);
            char* host = (char*)ptr[k] + in;
            if (TO_STAGE) stream_copy(buf, host, (size_t)take);
            else stream_copy(host, buf, (size_t)take);
            buf += take; off += take; n -= take;
            ++k;
            while (n > 0 && k < K && start[k + 1] == start[k]) ++k;
        }
    }
};

// upload == true: host arrays -> d_base; false: d_base -> host arrays
int run_transfer(bool upload, char* d_base, const Ragged& rg, int threads)
{
    const long long total = rg.start[rg.K];
    if (total == 0) return BHMM_OK;
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    RC_TRY(stage_ensure(asked_workers(threads)));          // (may change SLOT_BYTES at the first call: before pick_workers)
    const int W = pick_workers(threads, total);
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    const long long pieces = (total + (long long)SLOT_BYTES - 1) / (long long)SLOT_BYTES;
    std::atomic<long long> next(0);
    std::atomic<int> failed(0);
    cudaError_t first_error[MAX_WORKERS];
    auto work = [&](int w) {
        cudaError_t err = cudaSetDevice(dev);
        char* slot[2] = {g_stage.base + SLOT_BYTES * (2 * w), g_stage.base + SLOT_BYTES * (2 * w + 1)};
        cudaStream_t st = g_stage.stream[w];
        // download: the piece whose DMA is in flight in the other slot is copied out one turn later
        long long pend_off[2] = {-1, -1}, pend_n[2] = {0, 0};
        int turn = 0;
        while (err == cudaSuccess && !failed.load(std::memory_order_relaxed)) {
            const long long p = next.fetch_add(1);
            if (p >= pieces) break;
            const long long off = p * (long long)SLOT_BYTES, n = std::min<long long>(SLOT_BYTES, total - off);
            const int s = turn & 1;
            ++turn;
            if (upload) {
                err = cudaEventSynchronize(g_stage.done[w][s]);          // the slot's previous DMA has drained
                if (err != cudaSuccess) break;
                rg.copy<true>(slot[s], off, n);
                err = cudaMemcpyAsync(d_base + off, slot[s], (size_t)n, cudaMemcpyHostToDevice, st);
                if (err == cudaSuccess) err = cudaEventRecord(g_stage.done[w][s], st);
            } else {
                if (pend_off[s] >= 0) {                                     // empty the slot before it is refilled
                    err = cudaEventSynchronize(g_stage.done[w][s]);
                    if (err != cudaSuccess) break;
                    rg.copy<false>(slot[s], pend_off[s], pend_n[s]);
                }
                err = cudaMemcpyAsync(slot[s], d_base + off, (size_t)n, cudaMemcpyDeviceToHost, st);
                if (err == cudaSuccess) err = cudaEventRecord(g_stage.done[w][s], st);
                pend_off[s] = off; pend_n[s] = n;
            }
        }
        if (err == cudaSuccess) err = cudaStreamSynchronize(st);
        if (err == cudaSuccess && !upload) {
            for (int k = 0; k < 2; ++k) {
                const int s = (turn + k) & 1;                               // older slot first
                if (pend_off[s] >= 0) rg.copy<false>(slot[s], pend_off[s], pend_n[s]);
            }
        }
        first_error[w] = err;
        if (err != cudaSuccess) failed.store(1);
    };
    std::vector<std::thread> pool;
    for (int w = 1; w < W; ++w) pool.emplace_back(work, w);
    work(0);
    for (auto& t : pool) t.join();
    if (failed.load()) {
        for (int w = 0; w < W; ++w)
            if (first_error[w] != cudaSuccess) { bhmm_set_error(BHMM_ERR_CUDA, cudaGetErrorString(first_error[w])); break; }
        return BHMM_ERR_CUDA;
    }
    return BHMM_OK;
}

int make_ragged(Ragged& rg, const void* const* ptrs, const long long* nbytes, int K)
{
    if (K < 0 || (K > 0 && (!ptrs || !nbytes))) { bhmm_set_error(BHMM_ERR_INVALID, "transfer: null array table"); return BHMM_ERR_INVALID; }
    rg.ptr = ptrs;
    rg.K = K;
    rg.start.assign((size_t)K + 1, 0);
    for (int k = 0; k < K; ++k) {
        if (nbytes[k] < 0 || (nbytes[k] > 0 && !ptrs[k])) { bhmm_set_error(BHMM_ERR_INVALID, "transfer: bad array"); return BHMM_ERR_INVALID; }
        rg.start[k + 1] = rg.start[k] + nbytes[k];
    }
    return BHMM_OK;
}

}  // namespace

// Write-touch every page of a fresh host allocation with `threads` threads (0 = 2), so that its page faults -- most of what
// copying 410 MB of paths into a new array costs, and on a busy host by far the most variable part -- happen while the GPU is
// still iterating: MaximumLikelihoodEstimator.fit() calls this from a helper thread on the array it will return the paths in.
// Pure host code (no CUDA call).
extern "C" int bhmm_b200_prefault(void* p, long long nbytes, int threads)
{
    if (nbytes <= 0) return BHMM_OK;
    if (!p) { bhmm_set_error(BHMM_ERR_INVALID, "prefault: null pointer"); return BHMM_ERR_INVALID; }
    const int W = std::max(1, std::min(threads > 0 ? threads : 2, MAX_WORKERS));
    const long long page = 4096, pages = (nbytes + page - 1) / page;
    auto work = [&](int w) {
        volatile char* c = (volatile char*)p;
        const long long lo = pages * w / W, hi = pages * (w + 1) / W;
        for (long long k = lo; k < hi; ++k) c[std::min(k * page, nbytes - 1)] = 0;
    };
    std::vector<std::thread> pool;
    for (int w = 1; w < W; ++w) pool.emplace_back(work, w);
    work(0);
    for (auto& t : pool) t.join();
    return BHMM_OK;
}

// Tuning knobs of the mover (measurement scripts; the defaults are what the estimators use): staging slot size in KiB
// (0 = keep) and streaming stores on / off (negative = keep).  Frees the staging slots; the next transfer re-creates them.
extern "C" int bhmm_b200_transfer_config(int stage_kb, int nt_copy)
{
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    if (g_stage.base) {
        for (int w = 0; w < g_stage.workers; ++w) {
            cudaStreamDestroy(g_stage.stream[w]);
            cudaEventDestroy(g_stage.done[w][0]);
            cudaEventDestroy(g_stage.done[w][1]);
        }
        cudaFreeHost(g_stage.base);
        g_stage = Staging();
    }
    if (stage_kb >= 64 && stage_kb <= (1 << 20)) SLOT_BYTES = (size_t)stage_kb << 10;
    if (nt_copy >= 0) g_nt_copy = nt_copy ? 1 : 0;
    return BHMM_OK;
}

// Concatenate K host arrays into device memory: d_dst[sum_{k'<k} nbytes[k'] ...] = srcs[k].  `stream` (may be NULL) is
// synchronised first: work queued on it may still be using the destination.  Returns after the data has arrived.
extern "C" int bhmm_b200_upload_ragged(void* d_dst, const void* const* srcs, const long long* nbytes, int K, int threads,
                                       void* stream)
{
    Ragged rg;
    RC_TRY(make_ragged(rg, srcs, nbytes, K));
    if (rg.start[K] > 0 && !d_dst) { bhmm_set_error(BHMM_ERR_INVALID, "upload: null destination"); return BHMM_ERR_INVALID; }
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return run_transfer(true, (char*)d_dst, rg, threads);
}

// The inverse: cut device memory back into K host arrays (dsts[k] receives nbytes[k] bytes).  `stream` (the one the
// producer of d_src ran on) is synchronised first.
extern "C" int bhmm_b200_download_ragged(void* const* dsts, const void* d_src, const long long* nbytes, int K, int threads,
                                         void* stream)
{
    Ragged rg;
    RC_TRY(make_ragged(rg, (const void* const*)dsts, nbytes, K));
    if (rg.start[K] > 0 && !d_src) { bhmm_set_error(BHMM_ERR_INVALID, "download: null source"); return BHMM_ERR_INVALID; }
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return run_transfer(false, (char*)const_cast<void*>(d_src), rg, threads);
}
