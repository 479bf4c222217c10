// cgx-b200: shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#define CGX_MAX_RULE_SPAN 15       // ComTypes.h:42-43
#define CGX_MAX_RULE_SYMBOLS 5     // ComTypes.h:44
#define CGX_MAXSCORE 99.0f         // ComTypes.h:51
#define CGX_PRECOMP 100            // ComTypes.h:55
#define CGX_SAMPLER 300            // ComTypes.h:63
#define CGX_SAMPLER_ONEGAP 65      // ComTypes.h:64
#define CGX_SAMPLER_TWOGAP 70      // ComTypes.h:65
#define CGX_LONGEST_SRC 5          // ExtractPair.cu:16

#define CGX_NUM_SMS 148

struct CgxError {
    std::string msg;
    int code = 1;        // return value of the C ABI entry point (CGX_E_* in include/cgx_b200.h)
};

#define CUDA_CHECK(expr)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char b_[512];                                                                             \
            snprintf(b_, sizeof b_, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            throw CgxError{b_};                                                                       \
        }                                                                                             \
    } while (0)

#define CGX_REQUIRE(cond, ...)                                                    \
    do {                                                                          \
        if (!(cond)) {                                                            \
            char b_[512];                                                         \
            snprintf(b_, sizeof b_, __VA_ARGS__);                                 \
            throw CgxError{std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + b_}; \
        }                                                                         \
    } while (0)

// the batch does not fit the 31-bit hit / 32-bit cell indices: the caller splits it (CGX_E_BATCH_TOO_LARGE)
#define CGX_REQUIRE_BATCH(cond, ...)                                              \
    do {                                                                          \
        if (!(cond)) {                                                            \
            char b_[512];                                                         \
            snprintf(b_, sizeof b_, __VA_ARGS__);                                 \
            throw CgxError{std::string("batch too large: ") + b_, 3};             \
        }                                                                         \
    } while (0)

// Grow-only device buffer: allocations are reused across query batches so that the steady-state hot
// path performs no cudaMalloc.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool owned = true;
    template <typename T>
    T *get(size_t count) {
        size_t bytes = count * sizeof(T);
        if (bytes > cap) {
            if (p && owned) CUDA_CHECK(cudaFree(p));
            p = nullptr;
            cap = 0;
            size_t want = bytes + bytes / 4 + 256;
            cudaError_t e = cudaMalloc(&p, want);
            if (e == cudaErrorMemoryAllocation) {                  // without the slack, then give up: the caller's batch is too large for
                (void)cudaGetLastError();                          // what the index leaves of the HBM -- CGX_E_BATCH_TOO_LARGE, callers split it
                want = bytes + 256;
                e = cudaMalloc(&p, want);
                if (e == cudaErrorMemoryAllocation) {
                    (void)cudaGetLastError();
                    p = nullptr;
                    char b_[160];
                    snprintf(b_, sizeof b_, "batch too large: out of device memory (%zu MB buffer)", want >> 20);
                    throw CgxError{b_, 3};
                }
            }
            CUDA_CHECK(e);
            cap = want;
            owned = true;
        }
        return reinterpret_cast<T *>(p);
    }
    template <typename T>
    T *ptr() const { return reinterpret_cast<T *>(p); }
    void release() {
        if (p && owned) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // adopt caller-owned device memory (multi-GPU import path)
    void adopt(void *q, size_t bytes) {
        release();
        p = q;
        cap = bytes;
        owned = false;
    }
};

// Grow-only pinned host buffer (result staging: D2H at full PCIe rate, reused across batches).
struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    template <typename T>
    T *get(size_t count) {
        size_t bytes = count * sizeof(T);
        if (bytes > cap) {
            if (p) CUDA_CHECK(cudaFreeHost(p));
            size_t want = bytes + bytes / 4 + 4096;
            CUDA_CHECK(cudaMallocHost(&p, want));
            cap = want;
        }
        return reinterpret_cast<T *>(p);
    }
    template <typename T>
    T *ptr() const { return reinterpret_cast<T *>(p); }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

static inline unsigned cgx_div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

static inline int cgx_bits_for(uint64_t v) {   // number of bits needed to represent values in [0, v]
    int b = 0;
    while (v) { b++; v >>= 1; }
    return b ? b : 1;
}

#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// Small device -> host read-back (counts the host needs to size the next stage) that stays off the copy engines: one warp
// stores the words into mapped pinned memory, the stream is synchronised, the host reads them.  A cudaMemcpyAsync D2H
// would queue in the D2H engine behind whatever result array of the previous batch is still travelling (FIFO per
// engine, 1.3 GB = 25-50 ms) and stall a pipelined batch at every count (cgx_extract_begin).
static __global__ void cgx_mailbox_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, int words) {
    const int i = threadIdx.x;
    if (i < words) dst[i] = src[i];
}
static inline void cgx_read_back(void *host_dst, const void *dev_src, size_t bytes, cudaStream_t stream) {
    thread_local uint32_t *box = nullptr;                 // one mailbox per host thread (one thread drives one GPU)
    constexpr int WORDS = 64;
    if (!box) CUDA_CHECK(cudaHostAlloc((void **)&box, WORDS * sizeof(uint32_t), cudaHostAllocPortable | cudaHostAllocMapped));
    if (bytes % 4 != 0 || bytes > WORDS * sizeof(uint32_t)) throw CgxError{"cgx_read_back: unsupported size"};
    cgx_mailbox_kernel<<<1, WORDS, 0, stream>>>(box, (const uint32_t *)dev_src, (int)(bytes / 4));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    memcpy(host_dst, box, bytes);
}

// streaming (read-once) 128-bit load that does not pollute L1
__device__ __forceinline__ int4 ld_nc_int4(const int4 *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
#endif
