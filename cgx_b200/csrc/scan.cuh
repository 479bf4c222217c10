// cgx-b200: device-wide exclusive prefix sum (hand-written; no Thrust / CUB).
// Reduce-then-scan with 2048-element tiles, applied recursively to the tile sums.  Traffic: the input
// is read twice and the output written once (12 B/element for uint32) -- HBM-bound.
#pragma once
#include "common.cuh"
#include "prof.h"
#include <algorithm>

namespace cgx {

constexpr int SC_BLOCK = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_BLOCK * SC_ITEMS;

struct ScanTemp {
    DevBuf level[4];
};

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= (unsigned)o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across a CTA of SC_BLOCK threads; returns exclusive prefix,
// *total receives the CTA sum.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t s_w[SC_BLOCK / 32 + 1];
    uint32_t inc = warp_incl_scan(v);
    unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t x = lane < SC_BLOCK / 32 ? s_w[lane] : 0;
        uint32_t xi = warp_incl_scan(x);
        if (lane < SC_BLOCK / 32) s_w[lane] = xi - x;
        if (lane == SC_BLOCK / 32 - 1) s_w[SC_BLOCK / 32] = xi;
    }
    __syncthreads();
    *total = s_w[SC_BLOCK / 32];
    return inc - v + s_w[w];
}

static __global__ void __launch_bounds__(SC_BLOCK) sc_reduce_kernel(const uint32_t *__restrict__ in, size_t n, uint32_t *__restrict__ sums) {
    size_t base = (size_t)blockIdx.x * SC_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        size_t idx = base + (size_t)i * SC_BLOCK + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    uint32_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// out[i] = offset(tile) + exclusive scan inside the tile; offsets == nullptr means a single tile.
static __global__ void __launch_bounds__(SC_BLOCK) sc_scan_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t n,
                                                           const uint32_t *__restrict__ offsets, uint32_t *__restrict__ total_out) {
    size_t base = (size_t)blockIdx.x * SC_TILE + (size_t)threadIdx.x * SC_ITEMS;
    uint32_t v[SC_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        s += v[i];
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan(s, &tot) + (offsets ? offsets[blockIdx.x] : 0u);
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == SC_BLOCK - 1) *total_out = ex;
}

// Exclusive scan; in and out may alias.  If total_out != nullptr the grand total is written there (device).
static void exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n, uint32_t *total_out, cudaStream_t stream, ScanTemp &tmp,
                               int level = 0, int *launches = nullptr) {
    if (n == 0) {
        if (total_out) CUDA_CHECK(cudaMemsetAsync(total_out, 0, sizeof(uint32_t), stream));
        return;
    }
    size_t tiles = (n + SC_TILE - 1) / SC_TILE;
    if (tiles == 1) {
        PROF("scan", (double)n * 8, sc_scan_kernel<<<1, SC_BLOCK, 0, stream>>>(in, out, n, nullptr, total_out));
        if (launches) *launches += 1;
        return;
    }
    CGX_REQUIRE(level < 4, "scan: too many levels");
    uint32_t *sums = tmp.level[level].get<uint32_t>(tiles);
    PROF("scan", (double)n * 4, sc_reduce_kernel<<<(unsigned)tiles, SC_BLOCK, 0, stream>>>(in, n, sums));
    if (launches) *launches += 1;
    exclusive_scan_u32(sums, sums, tiles, nullptr, stream, tmp, level + 1, launches);
    PROF("scan", (double)n * 8, sc_scan_kernel<<<(unsigned)tiles, SC_BLOCK, 0, stream>>>(in, out, n, sums, total_out));
    if (launches) *launches += 1;
}

#endif

}  // namespace cgx
