// cgx-b200: gappy-phrase matching as sorted-occurrence band joins.
//
// Replaces oneGapLookUpSA (GappyLook.cu:128-474), twoGapLookUpSA (:476-737), the frequent-pair
// precomputation they lean on (precomp, :740-870; preComputation, SuffixArray.cu:1132-1340), the
// thrust sorts of the hit lists (SuffixArray.cu:1836,2205) and the host scans that turn them into
// per-pattern ranges (:1854-1875, :2214-2233).
//
// Hit set of aXb (identical for the reference's three strategies -- forward scan from a, backward scan
// from b, walk of the precomputed pair list -- see oracle/cgx_oracle.c onegap_lookup):
//   { (p, L) : a at p, b at p+ls+g, g >= 1, ls+g+le <= 15, every token of the gap >= 2,
//              checkBoundaryGap(p+ls, p+ls+g-1) },   L = ls+g+le-1
// Both occurrence lists are POSITION-SORTED slices of the index (inv[len-1][up..down]), so the join is
// a band merge: the shorter list drives, cut into tiles of JN_TILE elements (a load-balanced tile list
// over all patterns, found by binary search on the scanned tile counts -- merge-path style
// partitioning); per tile two binary searches bracket the slice of the other list that can fall into
// the band of the tile, and every lane then only searches that bracket.  Hits are appended with
// warp-aggregated atomics as packed 64-bit keys (pattern | position | length) and put in
// (pattern, position, length) order by one onesweep radix sort; a boundary kernel derives the
// per-pattern ranges.  The reference's 100x100 frequent-pair cache is not needed for speed; the one
// place where it leaks into results (featureMissingCount is added to SampleCountF for patterns made of
// two frequent tokens, ExtractPair.c:900-908) is reproduced by counting, in the same join, the
// candidates that fail only the alignment check.
#include "batch.h"
#include "prof.h"

namespace cgx {

constexpr int JN_TILE = 128;

__device__ __forceinline__ int lower_bound_i32(const int32_t *__restrict__ a, int lo, int hi, int x) {   // first idx in [lo,hi) with a[idx] >= x
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(&a[mid]) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// GappyLook.cu:43-126 checkBoundaryGap, preceded by the "every gap token >= 2" scan of the callers
// (:341-351, :403-418), answered from the precomputed gap-consistency word of the span's first position
// (index.cu ix_gap_words_kernel): one 4-byte load per candidate instead of a walk over the gap's RLP words and
// its target window.  Returns 0 = a token < 2 in the gap, 1 = alignment check failed, 2 = ok.
__device__ __forceinline__ int gap_check(const uint32_t *__restrict__ gapw, int start, int ender) {
    const uint32_t w = __ldg(&gapw[start]);
    const int g = ender - start + 1;
    if ((int)((w >> 16) & 15u) < g) return 0;
    return ((w >> (g - 1)) & 1u) ? 2 : 1;
}

__device__ __forceinline__ void append_hit(uint64_t key, unsigned long long *counter, uint64_t *out, size_t cap) {
    unsigned m = __activemask();
    int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    size_t slot = (size_t)base + __popc(m & lanemask_lt());
    if (slot < cap) out[slot] = key;
}

// ------------------------------------------------------------------------------------------------
__global__ void j1_tiles_kernel(const Pat1Dev *__restrict__ patd, int D1, uint32_t *__restrict__ tiles) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D1) return;
    Pat1Dev p = patd[d];
    int nA = p.down_a - p.up_a + 1, nB = p.down_b - p.up_b + 1;
    tiles[d] = (uint32_t)((min(nA, nB) + JN_TILE - 1) / JN_TILE);
}

__device__ __forceinline__ int find_owner(const uint32_t *__restrict__ off, int n, uint32_t tile) {   // largest d with off[d] <= tile
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(&off[mid]) <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(JN_TILE) j1_join_kernel(const Pat1 *__restrict__ pat, const Pat1Dev *__restrict__ patd, int D1,
                                                          const uint32_t *__restrict__ tile_off, const int32_t *__restrict__ inv1,
                                                          const int32_t *__restrict__ inv2, const int32_t *__restrict__ inv3,
                                                          const uint32_t *__restrict__ gapw,
                                                          unsigned long long *__restrict__ counter, uint64_t *__restrict__ hits, size_t cap,
                                                          int32_t *__restrict__ missing) {
    __shared__ int s_d, s_olo, s_ohi, s_missing;
    if (threadIdx.x == 0) {
        s_d = find_owner(tile_off, D1, blockIdx.x);
        s_missing = 0;
    }
    __syncthreads();
    const int d = s_d;
    const Pat1 p = pat[d];
    const Pat1Dev pd = patd[d];
    const int ls = p.ls, le = p.le;
    const int nA = pd.down_a - pd.up_a + 1, nB = pd.down_b - pd.up_b + 1;
    const bool driveA = nA <= nB;                                // GappyLook.cu:241 (dis <= dis2 -> forward)
    const int32_t *invA = (ls == 1 ? inv1 : ls == 2 ? inv2 : inv3) + pd.up_a;
    const int32_t *invB = (le == 1 ? inv1 : le == 2 ? inv2 : inv3) + pd.up_b;
    const int32_t *drv = driveA ? invA : invB;
    const int32_t *oth = driveA ? invB : invA;
    const int ndrv = driveA ? nA : nB, noth = driveA ? nB : nA;
    const int e0 = (int)(blockIdx.x - tile_off[d]) * JN_TILE;
    const int e1 = min(e0 + JN_TILE, ndrv);
    // band of the other list relative to a driver position x:  [x+lo_off, x+hi_off]
    const int lo_off = driveA ? ls + 1 : -(CGX_MAX_RULE_SPAN - le);
    const int hi_off = driveA ? CGX_MAX_RULE_SPAN - le : -(ls + 1);
    if (threadIdx.x == 0) {
        int xf = drv[e0], xl = drv[e1 - 1];
        s_olo = lower_bound_i32(oth, 0, noth, xf + lo_off);
        s_ohi = lower_bound_i32(oth, s_olo, noth, xl + hi_off + 1);
    }
    __syncthreads();
    const int e = e0 + threadIdx.x;
    if (e < e1 && s_ohi > s_olo) {
        const int x = drv[e];
        int j = lower_bound_i32(oth, s_olo, s_ohi, x + lo_off);
        for (; j < s_ohi; j++) {
            int y = __ldg(&oth[j]);
            if (y > x + hi_off) break;
            int a_p = driveA ? x : y, b_p = driveA ? y : x;
            int r = gap_check(gapw, a_p + ls, b_p - 1);
            if (r == 2) {
                uint64_t key = ((uint64_t)(uint32_t)d << 34) | ((uint64_t)(uint32_t)a_p << 4) | (uint64_t)(b_p + le - 1 - a_p);
                append_hit(key, counter, hits, cap);
            } else if (r == 1 && p.marker_pair >= 0) {
                atomicAdd(&s_missing, 1);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_missing) atomicAdd(&missing[d], s_missing);
}

// per-pattern [start,count] in the sorted hit list
template <int SHIFT>
__global__ void hit_ranges_kernel(const uint64_t *__restrict__ hits, size_t n, int32_t *__restrict__ start_count, int stride_ints) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t d = (uint32_t)(hits[k] >> SHIFT);
    if (k == 0 || (uint32_t)(hits[k - 1] >> SHIFT) != d) start_count[(size_t)d * stride_ints + 0] = (int32_t)k;
    if (k == n - 1 || (uint32_t)(hits[k + 1] >> SHIFT) != d) {
        // count = k - start + 1 ; start may be written by another thread -> derive it by searching backwards is costly;
        // store end in the count slot and fix up in a second kernel
        start_count[(size_t)d * stride_ints + 1] = (int32_t)k;
    }
}
__global__ void hit_ranges_fix_kernel(int32_t *__restrict__ start_count, int n_pat, int stride_ints) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_pat) return;
    int s = start_count[(size_t)d * stride_ints + 0];
    if (s >= 0) start_count[(size_t)d * stride_ints + 1] = start_count[(size_t)d * stride_ints + 1] - s + 1;
}

__global__ void j1_missing_kernel(Pat1 *__restrict__ pat, const int32_t *__restrict__ missing, int D1) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < D1) pat[d].fs_extra = missing[d];
}

static unsigned long long read_u64(const unsigned long long *d, cudaStream_t stream) {
    unsigned long long v = 0;
    CUDA_CHECK(cudaMemcpyAsync(&v, d, sizeof(v), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    return v;
}

void stage_onegap_join(const Index &ix, Batch &b, cudaStream_t stream) {
    b.hits1 = 0;
    const int D1 = b.D1;
    if (D1 == 0) return;
    CGX_REQUIRE(D1 < (1 << 30) && ix.n < (1ull << 30), "one-gap join: pattern or corpus size exceeds the 30-bit key fields");
    uint32_t *tiles = b.j_tiles.get<uint32_t>((size_t)D1 + 2);
    uint32_t *tot = b.counters.get<uint32_t>(16);
    unsigned long long *ctr = (unsigned long long *)(tot + 4);
    j1_tiles_kernel<<<cgx_div_up(D1, 256), 256, 0, stream>>>(b.pat1_dev.ptr<Pat1Dev>(), D1, tiles);
    exclusive_scan_u32(tiles, tiles, (size_t)D1, tot, stream, b.scan, 0, &b.launches);
    uint32_t n_tiles = 0;
    CUDA_CHECK(cudaMemcpyAsync(&n_tiles, tot, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    int32_t *missing = b.missing.get<int32_t>((size_t)D1);
    if (b.hit_cap == 0) b.hit_cap = 1u << 22;
    while (true) {
        uint64_t *hits = b.hit_keys.get<uint64_t>(b.hit_cap);
        CUDA_CHECK(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), stream));
        CUDA_CHECK(cudaMemsetAsync(missing, 0, sizeof(int32_t) * (size_t)D1, stream));
        if (n_tiles)
            PROF("join_onegap", 0.0, (j1_join_kernel<<<n_tiles, JN_TILE, 0, stream>>>(b.pat1.ptr<Pat1>(), b.pat1_dev.ptr<Pat1Dev>(), D1, tiles, ix.inv[0].ptr<int32_t>(),
                                                           ix.inv[1].ptr<int32_t>(), ix.inv[2].ptr<int32_t>(), ix.gapw.ptr<uint32_t>(), ctr, hits, b.hit_cap, missing)));
        b.launches += 2;
        unsigned long long H = read_u64(ctr, stream);
        if (H <= b.hit_cap) { b.hits1 = (int64_t)H; break; }
        b.hit_cap = (size_t)H + (size_t)H / 8 + 1024;        // grow once to the exact need and redo the join
    }
    j1_missing_kernel<<<cgx_div_up(D1, 256), 256, 0, stream>>>(b.pat1.ptr<Pat1>(), missing, D1);
    b.launches++;
    if (b.hits1 == 0) return;
    const size_t H = (size_t)b.hits1;
    uint64_t *hits = b.hit_keys.ptr<uint64_t>();
    uint64_t *tmp = b.hit_keys_tmp.get<uint64_t>(H);
    uint64_t *hs;
    radix_sort<uint64_t>(hits, tmp, nullptr, nullptr, H, 0, 34 + cgx_bits_for((uint64_t)D1), stream, b.radix, &hs, nullptr, &b.launches);
    uint64_t *dst = b.hits1_sorted.get<uint64_t>(H);
    CUDA_CHECK(cudaMemcpyAsync(dst, hs, sizeof(uint64_t) * H, cudaMemcpyDeviceToDevice, stream));
    static_assert(sizeof(Pat1) == 32, "Pat1 layout");
    hit_ranges_kernel<34><<<cgx_div_up(H, 256), 256, 0, stream>>>(dst, H, &b.pat1.ptr<int32_t>()[4], 8);
    hit_ranges_fix_kernel<<<cgx_div_up(D1, 256), 256, 0, stream>>>(&b.pat1.ptr<int32_t>()[4], D1, 8);
    b.launches += 2;
}

// ------------------------------------------------------------------------------------------------
// two-gap: parent hits (p, L) of aXb joined with the occurrences of the single token c
//   c at p+L+1+g2, g2 >= 1, (L+1)+g2+1 <= 15  ->  c_pos in [p+L+2, p+14]        (GappyLook.cu:595-655)
// ------------------------------------------------------------------------------------------------
__global__ void j2_tiles_kernel(const Pat2 *__restrict__ pat2, const Pat1 *__restrict__ pat1, const int32_t *__restrict__ tok_start, int D2,
                                uint32_t *__restrict__ tiles) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D2) return;
    Pat2 p = pat2[d];
    int nH = pat1[p.pat1].hit_count;
    int nC = tok_start[p.ctok + 1] - tok_start[p.ctok];
    tiles[d] = nH > 0 ? (uint32_t)((min(nH, nC) + JN_TILE - 1) / JN_TILE) : 0u;
}

__device__ __forceinline__ int lower_bound_hitpos(const uint64_t *__restrict__ h, int lo, int hi, int x) {   // first idx with pos >= x
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        int pos = (int)((h[mid] >> 4) & 0x3fffffffu);
        if (pos < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(JN_TILE) j2_join_kernel(const Pat2 *__restrict__ pat2, const Pat1 *__restrict__ pat1, int D2,
                                                          const uint32_t *__restrict__ tile_off, const uint64_t *__restrict__ hits1,
                                                          const int32_t *__restrict__ inv1, const int32_t *__restrict__ tok_start,
                                                          const uint32_t *__restrict__ gapw,
                                                          unsigned long long *__restrict__ counter, uint64_t *__restrict__ hits, size_t cap) {
    __shared__ int s_d, s_olo, s_ohi;
    if (threadIdx.x == 0) s_d = find_owner(tile_off, D2, blockIdx.x);
    __syncthreads();
    const int d = s_d;
    const Pat2 p2 = pat2[d];
    const Pat1 p1 = pat1[p2.pat1];
    const uint64_t *H = hits1 + p1.hit_start;
    const int nH = p1.hit_count;
    const int32_t *C = inv1 + tok_start[p2.ctok];
    const int nC = tok_start[p2.ctok + 1] - tok_start[p2.ctok];
    const bool driveH = nH <= nC;
    const int e0 = (int)(blockIdx.x - tile_off[d]) * JN_TILE;
    const int e1 = min(e0 + JN_TILE, driveH ? nH : nC);
    if (threadIdx.x == 0) {
        if (driveH) {
            int pf = (int)((H[e0] >> 4) & 0x3fffffffu), pl = (int)((H[e1 - 1] >> 4) & 0x3fffffffu);
            s_olo = lower_bound_i32(C, 0, nC, pf + 4);                          // L >= 2 -> c_pos >= p+4
            s_ohi = lower_bound_i32(C, s_olo, nC, pl + CGX_MAX_RULE_SPAN);      // c_pos <= p+14
        } else {
            int cf = C[e0], cl = C[e1 - 1];
            s_olo = lower_bound_hitpos(H, 0, nH, cf - (CGX_MAX_RULE_SPAN - 1));
            s_ohi = lower_bound_hitpos(H, s_olo, nH, cl - 4 + 1);
        }
    }
    __syncthreads();
    const int e = e0 + threadIdx.x;
    if (e >= e1 || s_ohi <= s_olo) return;
    if (driveH) {
        uint64_t hk = H[e];
        int p = (int)((hk >> 4) & 0x3fffffffu), L = (int)(hk & 15);
        int j = lower_bound_i32(C, s_olo, s_ohi, p + L + 2);
        for (; j < s_ohi; j++) {
            int c = __ldg(&C[j]);
            if (c > p + CGX_MAX_RULE_SPAN - 1) break;
            if (gap_check(gapw, p + L + 1, c - 1) == 2)
                append_hit(((uint64_t)(uint32_t)d << 38) | ((uint64_t)(uint32_t)p << 8) | ((uint64_t)L << 4) | (uint64_t)(c - p), counter, hits, cap);
        }
    } else {
        int c = C[e];
        int j = lower_bound_hitpos(H, s_olo, s_ohi, c - (CGX_MAX_RULE_SPAN - 1));
        for (; j < s_ohi; j++) {
            uint64_t hk = H[j];
            int p = (int)((hk >> 4) & 0x3fffffffu), L = (int)(hk & 15);
            if (p > c - 4) break;
            if (c < p + L + 2) continue;
            if (gap_check(gapw, p + L + 1, c - 1) == 2)
                append_hit(((uint64_t)(uint32_t)d << 38) | ((uint64_t)(uint32_t)p << 8) | ((uint64_t)L << 4) | (uint64_t)(c - p), counter, hits, cap);
        }
    }
}

void stage_twogap_join(const Index &ix, Batch &b, cudaStream_t stream) {
    b.hits2 = 0;
    const int D2 = b.D2;
    if (D2 == 0 || b.hits1 == 0) return;
    CGX_REQUIRE(D2 < (1 << 26), "two-gap join: %d distinct patterns exceed the 26-bit key field", D2);
    uint32_t *tiles = b.j_tiles.get<uint32_t>((size_t)D2 + 2);
    uint32_t *tot = b.counters.get<uint32_t>(16);
    unsigned long long *ctr = (unsigned long long *)(tot + 4);
    j2_tiles_kernel<<<cgx_div_up(D2, 256), 256, 0, stream>>>(b.pat2.ptr<Pat2>(), b.pat1.ptr<Pat1>(), ix.tok_start.ptr<int32_t>(), D2, tiles);
    exclusive_scan_u32(tiles, tiles, (size_t)D2, tot, stream, b.scan, 0, &b.launches);
    uint32_t n_tiles = 0;
    CUDA_CHECK(cudaMemcpyAsync(&n_tiles, tot, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    while (true) {
        uint64_t *hits = b.hit_keys.get<uint64_t>(b.hit_cap);
        CUDA_CHECK(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), stream));
        if (n_tiles)
            PROF("join_twogap", 0.0, (j2_join_kernel<<<n_tiles, JN_TILE, 0, stream>>>(b.pat2.ptr<Pat2>(), b.pat1.ptr<Pat1>(), D2, tiles, b.hits1_sorted.ptr<uint64_t>(),
                                                           ix.inv[0].ptr<int32_t>(), ix.tok_start.ptr<int32_t>(), ix.gapw.ptr<uint32_t>(), ctr, hits, b.hit_cap)));
        b.launches += 2;
        unsigned long long H = read_u64(ctr, stream);
        if (H <= b.hit_cap) { b.hits2 = (int64_t)H; break; }
        b.hit_cap = (size_t)H + (size_t)H / 8 + 1024;
    }
    if (b.hits2 == 0) return;
    const size_t H = (size_t)b.hits2;
    uint64_t *hits = b.hit_keys.ptr<uint64_t>();
    uint64_t *tmp = b.hit_keys_tmp.get<uint64_t>(H);
    uint64_t *hs;
    radix_sort<uint64_t>(hits, tmp, nullptr, nullptr, H, 0, 38 + cgx_bits_for((uint64_t)D2), stream, b.radix, &hs, nullptr, &b.launches);
    uint64_t *dst = b.hits2_sorted.get<uint64_t>(H);
    CUDA_CHECK(cudaMemcpyAsync(dst, hs, sizeof(uint64_t) * H, cudaMemcpyDeviceToDevice, stream));
    static_assert(sizeof(Pat2) == 16, "Pat2 layout");
    hit_ranges_kernel<38><<<cgx_div_up(H, 256), 256, 0, stream>>>(dst, H, &b.pat2.ptr<int32_t>()[2], 4);
    hit_ranges_fix_kernel<<<cgx_div_up(D2, 256), 256, 0, stream>>>(&b.pat2.ptr<int32_t>()[2], D2, 4);
    b.launches += 2;
}

}  // namespace cgx
