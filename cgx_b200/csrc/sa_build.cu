// cgx-b200: GPU suffix-array construction by prefix doubling over packed (rank[i], rank[i+h]) keys,
// sorted with the hand-written onesweep radix sort (radix_sort.cuh).
//
// Replaces the reference's single-threaded CPU DC3/skew (SuffixArray.c:51-129 suffixArrayInt, called
// from :196 suffixArrayConstruct).  The answer is unique -- plain lexicographic order on token ids,
// text padded with 0, the final symbol V+2 unique (Start.cu:321-327) -- so parity is a memcmp.
//
// Round structure (h = 2, 4, 8, ...):  key[i] = rank[i] << B | rank[i+h]  (rank 0 = "past the end"),
// radix sort (key, i), head flags where adjacent keys differ, rank = position of the group head.  A suffix that is alone in
// its group is final: once fewer than half of the suffixes are still tied, a round gathers, sorts and writes back only
// those (Larsson-Sadakane discarding) -- on the synthetic corpora the last three of five rounds touch a few per cent of n.
// Stop when every group is a singleton.  Algorithmic traffic per round: 16 B/suffix (8 B key read, 4 B SA write, 4 B rank
// write) -- the figure SURVEY.md 8(d) charges.
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

// head[k] = 1 where the sorted key at k opens a new group (differs from its predecessor)
__global__ void sa_head_flags_kernel(const uint64_t *__restrict__ keys, size_t n, uint32_t *__restrict__ flags) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    flags[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1u : 0u;
}

// Ranks are GROUP HEAD POSITIONS (+1): rank[i] = 1 + SA position of the first suffix of i's group.  A suffix that is alone
// in its group keeps that rank for good, which is what lets later rounds leave it out (below).
// headpos[g] = SA position of the head of dense group g   (excl = exclusive scan of the head flags)
__global__ void sa_headpos_kernel(const uint32_t *__restrict__ head, const uint32_t *__restrict__ excl, const uint32_t *__restrict__ pos, size_t n,
                                  uint32_t *__restrict__ headpos) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (head[k]) headpos[excl[k]] = pos ? pos[k] : (uint32_t)k;
}
// rank[sa[k]] = headpos[group of k] + 1
__global__ void sa_scatter_rank_kernel(const uint32_t *__restrict__ sa, const uint32_t *__restrict__ head, const uint32_t *__restrict__ excl,
                                       const uint32_t *__restrict__ headpos, size_t n, uint32_t *__restrict__ rank) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    rank[sa[k]] = headpos[excl[k] + head[k] - 1] + 1u;
}

// ---- rounds over the suffixes that are not yet alone in their group -----------------------------------------------------
// active[k] = 1 unless the suffix at SA position k is a singleton group (head here and head right after)
__global__ void sa_active_kernel(const uint32_t *__restrict__ head, size_t n, uint32_t *__restrict__ active) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const bool single = head[k] && (k + 1 == n || head[k + 1]);
    active[k] = single ? 0u : 1u;
}
// compacted list of the active SA positions, their suffixes and their (rank[i], rank[i+h]) keys.  The list is in SA order,
// so it is already sorted by the first key component; sorting it by the whole key permutes suffixes inside their groups only.
__global__ void sa_active_keys_kernel(const uint32_t *__restrict__ head, const uint32_t *__restrict__ aexcl, const uint32_t *__restrict__ sa,
                                      const uint32_t *__restrict__ rank, size_t n, size_t h, int rbits, uint32_t *__restrict__ apos,
                                      uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const bool single = head[k] && (k + 1 == n || head[k + 1]);
    if (single) return;
    const uint32_t j = aexcl[k];
    const uint32_t i = sa[k];
    const uint64_t a = rank[i];
    const uint64_t b = ((size_t)i + h < n) ? rank[(size_t)i + h] : 0u;
    apos[j] = (uint32_t)k;
    keys[j] = (a << rbits) | b;
    vals[j] = i;
}
// the re-sorted suffixes go back to the active SA positions (ascending), and the head flags there are refreshed
__global__ void sa_active_writeback_kernel(const uint32_t *__restrict__ apos, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ ahead,
                                           size_t m, uint32_t *__restrict__ sa, uint32_t *__restrict__ head) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t k = apos[j];
    sa[k] = vals[j];
    head[k] = ahead[j];
}

static void rank_from_heads(const uint32_t *sa_sorted, const uint32_t *head, const uint32_t *pos, size_t cnt, uint32_t *excl, uint32_t *headpos,
                            uint32_t *rank, uint32_t *d_total, cudaStream_t stream, ScanTemp &scan, int *launches) {
    const unsigned grid = cgx_div_up(cnt, 256);
    exclusive_scan_u32(head, excl, cnt, d_total, stream, scan, 0, launches);
    sa_headpos_kernel<<<grid, 256, 0, stream>>>(head, excl, pos, cnt, headpos);
    sa_scatter_rank_kernel<<<grid, 256, 0, stream>>>(sa_sorted, head, excl, headpos, cnt, rank);
    *launches += 2;
}

void build_suffix_array(const int32_t *d_str, size_t n, int32_t maxtok, int32_t *d_sa_out, SaWorkspace &ws, cudaStream_t stream,
                        SaStats *stats) {
    CGX_REQUIRE(n >= 2 && n < (1ull << 30), "suffix array: n=%zu out of range", n);
    uint64_t *keys = ws.keys.get<uint64_t>(n), *keys_tmp = ws.keys_tmp.get<uint64_t>(n);
    uint32_t *vals = ws.vals.get<uint32_t>(n), *vals_tmp = ws.vals_tmp.get<uint32_t>(n);
    uint32_t *rank = ws.rank.get<uint32_t>(n), *flags = ws.flags.get<uint32_t>(n);
    uint32_t *head = ws.head.get<uint32_t>(n + 1), *excl = ws.excl.get<uint32_t>(n + 1), *headpos = ws.headpos.get<uint32_t>(n + 1), *apos = ws.apos.get<uint32_t>(n + 1);
    uint32_t *d_total = ws.total.get<uint32_t>(4);
    uint32_t *sa = reinterpret_cast<uint32_t *>(d_sa_out);
    const int tokbits = cgx_bits_for((uint64_t)maxtok);
    const int rbits = cgx_bits_for((uint64_t)n);
    const unsigned grid = cgx_div_up(n, 256);
    int launches = 0, rounds = 0;
    uint64_t *ks;
    uint32_t *vs;
    // ---- first round: all suffixes by their first two tokens
    sa_init_keys_kernel<<<grid, 256, 0, stream>>>(d_str, n, tokbits, keys, vals);
    launches++;
    radix_sort<uint64_t>(keys, keys_tmp, vals, vals_tmp, n, 0, 2 * tokbits, stream, ws.radix, &ks, &vs, &launches);
    CUDA_CHECK(cudaMemcpyAsync(sa, vs, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, stream));
    sa_head_flags_kernel<<<grid, 256, 0, stream>>>(ks, n, head);
    launches++;
    rank_from_heads(sa, head, nullptr, n, excl, headpos, rank, d_total, stream, ws.scan, &launches);
    size_t h = 2;
    while (true) {
        rounds++;
        // the suffixes that still share a group
        sa_active_kernel<<<grid, 256, 0, stream>>>(head, n, flags);
        exclusive_scan_u32(flags, flags, n, d_total + 1, stream, ws.scan, 0, &launches);
        launches += 1;
        uint32_t m32 = 0;
        cgx_read_back(&m32, d_total + 1, sizeof(m32), stream);
        const size_t m = m32;
        if (m == 0) break;
        CGX_REQUIRE(h < 2 * n, "suffix array: did not converge (text without a unique final symbol?)");
        if (m * 2 >= n) {
            // most suffixes are still tied: a coalesced pass over all of them beats gathering the active ones
            sa_round_keys_kernel<<<grid, 256, 0, stream>>>(rank, n, h, rbits, keys, vals);
            launches += 1;
            radix_sort<uint64_t>(keys, keys_tmp, vals, vals_tmp, n, 0, 2 * rbits, stream, ws.radix, &ks, &vs, &launches);
            CUDA_CHECK(cudaMemcpyAsync(sa, vs, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, stream));
            sa_head_flags_kernel<<<grid, 256, 0, stream>>>(ks, n, head);
            launches += 1;
            rank_from_heads(sa, head, nullptr, n, excl, headpos, rank, d_total, stream, ws.scan, &launches);
        } else {
            // only the m tied suffixes are gathered, re-sorted and re-ranked; singletons keep their SA position and rank
            sa_active_keys_kernel<<<grid, 256, 0, stream>>>(head, flags, sa, rank, n, h, rbits, apos, keys, vals);
            radix_sort<uint64_t>(keys, keys_tmp, vals, vals_tmp, m, 0, 2 * rbits, stream, ws.radix, &ks, &vs, &launches);
            uint32_t *ahead = flags;                       // head flags of the sorted active list (the active offsets are spent)
            sa_head_flags_kernel<<<cgx_div_up(m, 256), 256, 0, stream>>>(ks, m, ahead);
            sa_active_writeback_kernel<<<cgx_div_up(m, 256), 256, 0, stream>>>(apos, vs, ahead, m, sa, head);
            launches += 3;
            rank_from_heads(vs, ahead, apos, m, excl, headpos, rank, d_total, stream, ws.scan, &launches);
        }
        h *= 2;
    }
    if (stats) {
        stats->rounds = rounds;
        stats->launches = launches;
        stats->key_bits = 2 * rbits;
    }
}

}  // namespace cgx
