// cgx-b200: GPU suffix-array construction by prefix doubling over packed (rank[i], rank[i+h]) keys,
// sorted with the hand-written onesweep radix sort (radix_sort.cuh).
//
// Replaces the reference's single-threaded CPU DC3/skew (SuffixArray.c:51-129 suffixArrayInt, called
// from :196 suffixArrayConstruct).  The answer is unique -- plain lexicographic order on token ids,
// text padded with 0, the final symbol V+2 unique (Start.cu:321-327) -- so parity is a memcmp.
//
// Round structure (h = 2, 4, 8, ...):  key[i] = rank[i] << B | rank[i+h]  (rank 0 = "past the end"),
// radix sort (key, i), head flags where adjacent keys differ, prefix sum -> new dense ranks,
// scatter rank[sa[k]].  Stop when every key is distinct.  Algorithmic traffic per round:
// 16 B/suffix (8 B key read, 4 B SA write, 4 B rank write) -- the figure SURVEY.md 8(d) charges.
#include "radix_sort.cuh"
#include "scan.cuh"
#include "index.h"

namespace cgx {

__global__ void sa_init_keys_kernel(const int32_t *__restrict__ str, size_t n, int tokbits, uint64_t *__restrict__ keys,
                                    uint32_t *__restrict__ vals) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // tokens >= 0; +1 so that the 0 padding past the end still sorts first and equals "nothing"
    uint64_t a = (uint64_t)(uint32_t)str[i], b = (uint64_t)(uint32_t)str[i + 1];
    keys[i] = (a << tokbits) | b;
    vals[i] = (uint32_t)i;
}

__global__ void sa_round_keys_kernel(const uint32_t *__restrict__ rank, size_t n, size_t h, int rbits, uint64_t *__restrict__ keys,
                                     uint32_t *__restrict__ vals) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t a = rank[i];
    uint64_t b = (i + h < n) ? rank[i + h] : 0u;
    keys[i] = (a << rbits) | b;
    vals[i] = (uint32_t)i;
}

__global__ void sa_head_flags_kernel(const uint64_t *__restrict__ keys, size_t n, uint32_t *__restrict__ flags) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    flags[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1u : 0u;
}

// rank[sa[k]] = (exclusive scan of flags)[k] + flags[k]  (1-based dense rank)
__global__ void sa_scatter_rank_kernel(const uint32_t *__restrict__ sa, const uint32_t *__restrict__ excl, const uint64_t *__restrict__ keys,
                                       size_t n, uint32_t *__restrict__ rank) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t f = (k == 0 || keys[k] != keys[k - 1]) ? 1u : 0u;
    rank[sa[k]] = excl[k] + f;
}

void build_suffix_array(const int32_t *d_str, size_t n, int32_t maxtok, int32_t *d_sa_out, SaWorkspace &ws, cudaStream_t stream,
                        SaStats *stats) {
    CGX_REQUIRE(n >= 2 && n < (1ull << 30), "suffix array: n=%zu out of range", n);
    uint64_t *keys = ws.keys.get<uint64_t>(n), *keys_tmp = ws.keys_tmp.get<uint64_t>(n);
    uint32_t *vals = ws.vals.get<uint32_t>(n), *vals_tmp = ws.vals_tmp.get<uint32_t>(n);
    uint32_t *rank = ws.rank.get<uint32_t>(n), *flags = ws.flags.get<uint32_t>(n);
    uint32_t *d_total = ws.total.get<uint32_t>(4);
    const int tokbits = cgx_bits_for((uint64_t)maxtok);
    const int rbits = cgx_bits_for((uint64_t)n);
    const unsigned grid = cgx_div_up(n, 256);
    int launches = 0, rounds = 0;
    uint64_t *ks;
    uint32_t *vs;
    sa_init_keys_kernel<<<grid, 256, 0, stream>>>(d_str, n, tokbits, keys, vals);
    launches++;
    radix_sort<uint64_t>(keys, keys_tmp, vals, vals_tmp, n, 0, 2 * tokbits, stream, ws.radix, &ks, &vs, &launches);
    size_t h = 2;
    while (true) {
        rounds++;
        sa_head_flags_kernel<<<grid, 256, 0, stream>>>(ks, n, flags);
        exclusive_scan_u32(flags, flags, n, d_total, stream, ws.scan, 0, &launches);
        uint32_t distinct = 0;
        cgx_read_back(&distinct, d_total, sizeof(uint32_t), stream);
        launches += 1;
        if ((size_t)distinct == n) break;
        CGX_REQUIRE(h < 2 * n, "suffix array: did not converge (text without a unique final symbol?)");
        sa_scatter_rank_kernel<<<grid, 256, 0, stream>>>(vs, flags, ks, n, rank);
        sa_round_keys_kernel<<<grid, 256, 0, stream>>>(rank, n, h, rbits, keys, vals);
        launches += 2;
        radix_sort<uint64_t>(keys, keys_tmp, vals, vals_tmp, n, 0, 2 * rbits, stream, ws.radix, &ks, &vs, &launches);
        h *= 2;
    }
    CUDA_CHECK(cudaMemcpyAsync(d_sa_out, vs, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, stream));
    if (stats) {
        stats->rounds = rounds;
        stats->launches = launches;
        stats->key_bits = 2 * rbits;
    }
}

}  // namespace cgx
