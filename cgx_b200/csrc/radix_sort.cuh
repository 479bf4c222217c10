// cgx-b200: hand-written onesweep LSD radix sort for sm_100a (no Thrust / CUB).
//
// One upfront kernel builds the digit histograms of every pass in a single read of the keys; each
// pass is then ONE kernel ("onesweep", Adinets & Merrill 2022): a tile of keys is ranked with
// warp-level match/ballot multisplit, the tile's per-digit counts are published, the global prefix of
// every digit is obtained by decoupled look-back over the preceding tiles' published counts, and the
// keys (and optional 32-bit payloads) are staged through shared memory so that the scatter to HBM is
// written in digit-contiguous, coalesced runs.  Per pass the traffic is therefore one read and one
// write of the data: (sizeof(K)+payload) * 2 bytes per element (ncu: DRAM bytes = exactly that).  On B200 the pass is
// bound by instruction issue, not by HBM: ~93 SASS instructions per key, 2.9 TB/s = 0.45 of the measured copy peak
// (profiles/README.md r01e); the ranking (one ballot per digit bit) is half of them.
//
// The sort is stable.  Keys are uint32_t or uint64_t; only bits [begin_bit, end_bit) are sorted on
// (callers know their key widths: packed (rank,rank) pairs, (pattern,position) pairs ...), split
// into 8-bit digits (compile-time width) plus one narrower remainder pass (runtime width).
#pragma once
#include "common.cuh"
#include "prof.h"
#include <algorithm>

namespace cgx {

// tile shape, measured on B200 with tools/sort_bench.py (2^27 keys; 52-bit keys only / 64-bit keys + payload, ms per sort):
//   512x8 (2 CTAs/SM) 8.19 / 10.9    256x8 (4) 9.84 / 12.3    256x16 (3) 7.31 / 9.42    256x16 (4) 6.68 / 10.15    256x24 (2) 7.44 / 9.77
#ifndef CGX_RS_BLOCK
#define CGX_RS_BLOCK 256
#endif
#ifndef CGX_RS_ITEMS
#define CGX_RS_ITEMS 16
#endif
constexpr int RS_BLOCK = CGX_RS_BLOCK;        // threads per CTA (>= 256: one thread per digit in the look-back)
constexpr int RS_WARPS = RS_BLOCK / 32;
constexpr int RS_ITEMS = CGX_RS_ITEMS;        // keys per thread
constexpr int RS_TILE = RS_BLOCK * RS_ITEMS;  // 4096 keys per tile
constexpr int RS_BINS = 256;
constexpr int RS_MAX_PASSES = 8;

// look-back status word: two flag bits above the count.  32-bit words (30-bit counts) up to 2^30 keys, 64-bit words beyond
template <typename S> struct RsStatus;
template <> struct RsStatus<uint32_t> {
    static constexpr uint32_t PARTIAL = 1u << 30, INCLUSIVE = 2u << 30, VALUE_MASK = (1u << 30) - 1;
};
template <> struct RsStatus<uint64_t> {
    static constexpr uint64_t PARTIAL = 1ull << 62, INCLUSIVE = 2ull << 62, VALUE_MASK = (1ull << 62) - 1;
};

struct RadixPlan {
    int num_passes;
    int shift[RS_MAX_PASSES];
    int bits[RS_MAX_PASSES];
};

static inline RadixPlan make_radix_plan(int begin_bit, int end_bit) {
    RadixPlan p;
    int total = end_bit - begin_bit;
    if (total < 1) total = 1;
    p.num_passes = (total + 7) / 8;
    // full 8-bit digits (the kernel specialised at compile time: half the ranking instructions of the runtime-width form),
    // the remainder in the most significant pass
    int s = begin_bit;
    for (int i = 0; i < p.num_passes; i++) {
        p.bits[i] = total - 8 * i >= 8 ? 8 : total - 8 * i;
        p.shift[i] = s;
        s += p.bits[i];
    }
    return p;
}

struct RadixTemp {
    DevBuf hist;      // [passes][256] uint32: digit histograms, then exclusive offsets
    DevBuf status;    // [tiles][256] look-back words (uint32, uint64 from 2^30 keys on)
    DevBuf counters;  // [passes] dynamic tile counters
};

#ifdef __CUDACC__

// peers of this lane's digit: one ballot per digit bit (cf. the MatchAny of CUB's radix rank).  MATCH.ANY occupies the XU
// pipe for ~200 cycles per warp instruction on sm_100 (ncu, round 1b: pipe_xu 73 % busy, the limiter of the whole pass).
__device__ __forceinline__ unsigned rs_match_digit(uint32_t d, int bits) {
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        if (b < bits) {
            const bool p = (d >> b) & 1u;
            const unsigned m = __ballot_sync(0xffffffffu, p);
            peers &= p ? m : ~m;
        }
    }
    return peers;
}
// the same with the digit width known at compile time: per bit one bit-field extract, one ballot, one three-input logic op
// (the runtime-width form spends ~7 instructions per bit on predicate bookkeeping -- cuobjdump, round 1c)
// peers = AND of the ballots of my set bits, minus OR of the ballots of my clear bits: per bit one ballot and two predicated
// logic ops (ptxas moves the digit's bits into predicates with one R2P)
template <int B>
__device__ __forceinline__ void rs_match_bit(unsigned &pos, unsigned &neg, uint32_t d) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 t, m;\n\t"
        "and.b32 t, %2, %3;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
        "@p and.b32 %0, %0, m;\n\t"
        "@!p or.b32 %1, %1, m;\n\t"
        "}"
        : "+r"(pos), "+r"(neg)
        : "r"(d), "n"(1u << B));
}
template <int BITS>
__device__ __forceinline__ unsigned rs_match_digit_ct(uint32_t d) {
    unsigned pos = 0xffffffffu, neg = 0u;
    if (BITS > 0) rs_match_bit<0>(pos, neg, d);
    if (BITS > 1) rs_match_bit<1>(pos, neg, d);
    if (BITS > 2) rs_match_bit<2>(pos, neg, d);
    if (BITS > 3) rs_match_bit<3>(pos, neg, d);
    if (BITS > 4) rs_match_bit<4>(pos, neg, d);
    if (BITS > 5) rs_match_bit<5>(pos, neg, d);
    if (BITS > 6) rs_match_bit<6>(pos, neg, d);
    if (BITS > 7) rs_match_bit<7>(pos, neg, d);
    return pos & ~neg;
}
__device__ __forceinline__ uint32_t rs_ld_status(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void rs_st_status(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t rs_ld_status(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void rs_st_status(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}


// Decoupled look-back of one digit (one thread per digit): sum of the counts the preceding tiles published for this digit.
// ncu of round 1's kernel put 38 % of all stall samples on this loop: it walked back one tile per DEPENDENT global load
// (~9 tiles deep on average, every step a full L2 round trip).  Now the status words of the next RS_LB predecessors are
// loaded together and consumed in order -- the chain is a quarter as long; a word that is not published yet is re-polled alone.
constexpr int RS_LB = 4;
template <typename S>
__device__ __forceinline__ uint32_t rs_look_back(const S *__restrict__ status, uint32_t tile, unsigned digit) {
    using ST = RsStatus<S>;
    uint32_t excl = 0;
    long long t = (long long)tile - 1;
    while (t >= 0) {
        S v[RS_LB];
#pragma unroll
        for (int k = 0; k < RS_LB; k++) v[k] = t - k >= 0 ? rs_ld_status(status + (size_t)(t - k) * RS_BINS + digit) : (S)ST::INCLUSIVE;
#pragma unroll
        for (int k = 0; k < RS_LB; k++) {
            while ((v[k] & (ST::PARTIAL | ST::INCLUSIVE)) == 0) { __nanosleep(20); v[k] = rs_ld_status(status + (size_t)(t - k) * RS_BINS + digit); }
            excl += (uint32_t)(v[k] & ST::VALUE_MASK);               // counts stay below 2^32 (n < 2^32)
            if (v[k] & ST::INCLUSIVE) return excl;
        }
        t -= RS_LB;
    }
    return excl;
}

template <typename K>
__global__ void __launch_bounds__(RS_BLOCK) rs_histogram_kernel(const K *__restrict__ keys, size_t n, RadixPlan plan,
                                                                uint32_t *__restrict__ hist) {
    __shared__ uint32_t sh[RS_MAX_PASSES * RS_BINS];
    for (int i = threadIdx.x; i < plan.num_passes * RS_BINS; i += RS_BLOCK) sh[i] = 0;
    __syncthreads();
    size_t stride = (size_t)gridDim.x * RS_BLOCK;
    for (size_t i = (size_t)blockIdx.x * RS_BLOCK + threadIdx.x; i < n; i += stride) {
        K k = keys[i];
#pragma unroll
        for (int p = 0; p < RS_MAX_PASSES; p++) {
            if (p < plan.num_passes) {
                uint32_t d = (uint32_t)(k >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u);
                atomicAdd(&sh[p * RS_BINS + d], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.num_passes * RS_BINS; i += RS_BLOCK)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// exclusive scan of each pass's 256-bin histogram (one CTA of 256 threads per pass)
static __global__ void __launch_bounds__(RS_BINS) rs_scan_hist_kernel(uint32_t *hist) {
    __shared__ uint32_t sh[RS_BINS];
    uint32_t *h = hist + blockIdx.x * RS_BINS;
    uint32_t v = h[threadIdx.x];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < RS_BINS; off <<= 1) {
        uint32_t t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    h[threadIdx.x] = sh[threadIdx.x] - v;
}

// One onesweep pass.  256 threads x 16 keys = one 4096-key tile; the tile's keys (and payloads) are exchanged through a
// shared-memory buffer that aliases the per-warp digit counters of the ranking phase, so a CTA needs 32 KB (+16 KB
// with payloads) and 64-80 registers per thread: four (three with payloads) CTAs per SM.  (Round 1a kept the keys, their
// ranks, staging positions and 64-bit output offsets in registers: 157 registers with payloads = one CTA per SM,
// 12 % occupancy, 1.3 TB/s.)
// CT_BITS = 8: digit width fixed at compile time (bits == 8); CT_BITS = 0: runtime width (the remainder pass, tiny keys)
template <typename K, bool HAS_VALUES, typename S = uint32_t, int CT_BITS = 0>
__global__ void __launch_bounds__(RS_BLOCK, HAS_VALUES ? 3 : 4) rs_onesweep_kernel(const K *__restrict__ keys_in, K *__restrict__ keys_out,
                                                                  const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ vals_out,
                                                                  size_t n, int shift, int bits, const uint32_t *__restrict__ digit_offset,
                                                                  S *status, uint32_t *tile_counter) {
    using ST = RsStatus<S>;
    extern __shared__ __align__(16) unsigned char s_raw[];      // RS_TILE keys (+ RS_TILE payloads): rs_smem_bytes<K>(HAS_VALUES)
    __shared__ uint32_t s_bin_start[RS_BINS];
    __shared__ uint32_t s_global_base[RS_BINS];
    __shared__ uint32_t s_scan_tot[RS_BINS / 32];
    __shared__ uint32_t s_tile;
    uint16_t(*s_warp_hist)[RS_BINS] = reinterpret_cast<uint16_t(*)[RS_BINS]>(s_raw);      // [RS_WARPS][RS_BINS], phase 1-3 only
    K *s_keys = reinterpret_cast<K *>(s_raw);
    uint32_t *s_vals = reinterpret_cast<uint32_t *>(s_raw + RS_TILE * sizeof(K));
    static_assert(RS_WARPS * RS_BINS * sizeof(uint16_t) <= RS_TILE * sizeof(uint32_t), "digit counters must fit the exchange buffer");

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mask = CT_BITS ? (1u << CT_BITS) - 1u : (1u << bits) - 1u;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < RS_WARPS * RS_BINS / 2; i += RS_BLOCK) reinterpret_cast<uint32_t *>(s_raw)[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const size_t warp_base = (size_t)tile * RS_TILE + (size_t)warp * (32 * RS_ITEMS);

    K key[RS_ITEMS];
    uint32_t val[HAS_VALUES ? RS_ITEMS : 1];
    uint32_t rank[RS_ITEMS];
    const bool full = ((size_t)tile + 1) * RS_TILE <= n;         // all tiles but the last: no bounds checks
    if (full) {
        const K *kp = keys_in + warp_base + lane;
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) key[i] = kp[i * 32];
        if (HAS_VALUES) {
            const uint32_t *vp = vals_in + warp_base + lane;
#pragma unroll
            for (int i = 0; i < RS_ITEMS; i++) val[i] = vp[i * 32];
        }
    } else {
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) {
            const size_t idx = warp_base + (size_t)i * 32 + lane;
            key[i] = idx < n ? keys_in[idx] : (K)~(K)0;
            if (HAS_VALUES) val[i] = idx < n ? vals_in[idx] : 0u;
        }
    }
    // ---- warp-level multisplit: stable rank of every key among the warp's keys with the same digit
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        const unsigned peers = CT_BITS ? rs_match_digit_ct<CT_BITS ? CT_BITS : 1>(d) : rs_match_digit(d, bits);
        const uint32_t base = s_warp_hist[warp][d];            // every peer reads the same counter (broadcast)
        __syncwarp();
        if ((peers & lt) == 0) s_warp_hist[warp][d] = (uint16_t)(base + __popc(peers));   // lowest peer lane updates it
        rank[i] = base + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    // ---- per digit (threads 0..255): exclusive offsets across warps, tile count published for the look-back
    uint32_t tile_count = 0, incl = 0;
    if (tid < RS_BINS) {
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            const uint32_t t = s_warp_hist[w][tid];
            s_warp_hist[w][tid] = (uint16_t)sum;
            sum += t;
        }
        tile_count = sum;
        rs_st_status(status + (size_t)tile * RS_BINS + tid, (S)sum | (tile == 0 ? ST::INCLUSIVE : ST::PARTIAL));
        // block-wide exclusive scan of the 256 tile counts: shuffle scan per warp, then the 8 warp totals
        incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        if (lane == 31) s_scan_tot[warp] = incl;
    }
    __syncthreads();
    if (tid < RS_BINS) {
        uint32_t pre = 0;
#pragma unroll
        for (int w = 0; w < RS_BINS / 32; w++) pre += (w < (int)warp) ? s_scan_tot[w] : 0u;
        const uint32_t bin_start = pre + incl - tile_count;
        s_bin_start[tid] = bin_start;
        // decoupled look-back over the preceding tiles' published counts
        uint32_t excl = 0;
        if (tile > 0) {
            excl = rs_look_back<S>(status, tile, tid);
            rs_st_status(status + (size_t)tile * RS_BINS + tid, (S)(excl + tile_count) | ST::INCLUSIVE);
        }
        s_global_base[tid] = digit_offset[tid] + excl - bin_start;
    }
    __syncthreads();
    // ---- staging position of every key (last use of the digit counters), then exchange through shared memory
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        rank[i] += s_bin_start[d] + s_warp_hist[warp][d];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        s_keys[rank[i]] = key[i];
        if (HAS_VALUES) s_vals[rank[i]] = val[i];
    }
    __syncthreads();
    // ---- digit-contiguous, coalesced stores.  (A separate check-free loop for full tiles was measured: 1488 -> 1376 SASS
    // instructions but 4.88 -> 5.12 ms per 2^27-key sort; the single predicated loop stays.)
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const unsigned j = tid + i * RS_BLOCK;
        const K k = s_keys[j];
        const uint32_t d = (uint32_t)(k >> shift) & mask;
        const size_t o = (size_t)s_global_base[d] + j;
        if (full || o < n) {
            keys_out[o] = k;
            if (HAS_VALUES) vals_out[o] = s_vals[j];
        }
    }
}


// ------------------------------------------------------------------------------------------------
// sm_100a asynchronous-copy plumbing: 1-D bulk copies global -> shared (cp.async.bulk, SASS UBLKCP) that complete on an mbarrier
// (SYNCS).  One elected thread arms the barrier with the byte count and issues the copy; every thread of the CTA waits on
// the barrier's phase parity.  fence.proxy.async orders the CTA's earlier generic-proxy accesses to the buffer (the exchange of
// the previous tile) before the async-proxy write that refills it.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rs_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rs_mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rs_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rs_mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rs_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rs_mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(rs_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void rs_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rs_smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(rs_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void rs_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void rs_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// One onesweep pass, Blackwell form: PERSISTENT CTAs (grid = resident CTAs) take tiles from the dynamic tile counter; the tile's
// keys (and payloads) arrive in shared memory by cp.async.bulk into one of two stage buffers while the previous tile is still
// being ranked and scattered, so no warp ever waits on a global load of keys (ncu of the one-tile-per-CTA form above: 6.1 of
// 15.3 stall cycles per issued instruction were long-scoreboard waits on the 16 LDGs at the head of every CTA).  The stage
// buffer a tile arrived in is also its exchange buffer (keys are in registers by then), the other one is being refilled.
// Ranking, look-back and the digit-contiguous store are those of rs_onesweep_kernel.  Forward progress of the look-back: a
// CTA holds at most its current and its prefetched next tile, both taken from the counter in ascending order, so the
// smallest unfinished tile is always some running CTA's CURRENT tile and waits on nothing.
template <typename K, bool HAS_VALUES, typename S = uint32_t, int CT_BITS = 0>
__global__ void __launch_bounds__(RS_BLOCK, HAS_VALUES ? 2 : 3) rs_onesweep_bulk_kernel(const K *__restrict__ keys_in, K *__restrict__ keys_out,
                                                                       const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ vals_out,
                                                                       size_t n, uint32_t num_tiles, int shift, int bits,
                                                                       const uint32_t *__restrict__ digit_offset, S *status, uint32_t *tile_counter) {
    using ST = RsStatus<S>;
    constexpr unsigned KEY_BYTES = RS_TILE * sizeof(K), VAL_BYTES = HAS_VALUES ? RS_TILE * sizeof(uint32_t) : 0, STAGE_BYTES = KEY_BYTES + VAL_BYTES;
    extern __shared__ __align__(128) unsigned char s_stages[];      // two stages: [keys | payloads] each
    __shared__ uint16_t s_warp_hist[RS_WARPS][RS_BINS];
    __shared__ uint32_t s_bin_start[RS_BINS];
    __shared__ uint32_t s_global_base[RS_BINS];
    __shared__ uint32_t s_scan_tot[RS_BINS / 32];
    __shared__ uint32_t s_next;
    __shared__ __align__(8) uint64_t s_bar[2];

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mask = CT_BITS ? (1u << CT_BITS) - 1u : (1u << bits) - 1u;
    const unsigned lt = lanemask_lt();
    if (tid == 0) {
        rs_mbar_init(&s_bar[0], 1);
        rs_mbar_init(&s_bar[1], 1);
        rs_fence_mbar_init();
        s_next = atomicAdd(tile_counter, 1u);
    }
    __syncthreads();
    uint32_t tile = s_next;
    unsigned stage = 0, parity0 = 0, parity1 = 0;
    auto issue = [&](uint32_t t, unsigned st) {                  // one thread: refill stage st with tile t (full tiles only)
        unsigned char *dst = s_stages + (size_t)st * STAGE_BYTES;
        rs_fence_proxy_async();
        rs_mbar_expect_tx(&s_bar[st], STAGE_BYTES);
        rs_bulk_g2s(dst, keys_in + (size_t)t * RS_TILE, KEY_BYTES, &s_bar[st]);
        if (HAS_VALUES) rs_bulk_g2s(dst + KEY_BYTES, vals_in + (size_t)t * RS_TILE, VAL_BYTES, &s_bar[st]);
    };
    if (tid == 0 && tile < num_tiles && ((size_t)tile + 1) * RS_TILE <= n) issue(tile, 0);
    __syncthreads();                                             // s_next read by everyone before thread 0 overwrites it

    while (tile < num_tiles) {
        for (int i = tid; i < RS_WARPS * RS_BINS / 2; i += RS_BLOCK) reinterpret_cast<uint32_t *>(&s_warp_hist[0][0])[i] = 0;
        const bool full = ((size_t)tile + 1) * RS_TILE <= n;     // all tiles but the last
        K *s_keys = reinterpret_cast<K *>(s_stages + (size_t)stage * STAGE_BYTES);
        uint32_t *s_vals = reinterpret_cast<uint32_t *>(s_stages + (size_t)stage * STAGE_BYTES + KEY_BYTES);
        const unsigned wbase = warp * (32 * RS_ITEMS) + lane;

        K key[RS_ITEMS];
        uint32_t val[HAS_VALUES ? RS_ITEMS : 1];
        uint32_t rank[RS_ITEMS];
        if (full) {
            rs_mbar_wait(&s_bar[stage], stage ? parity1 : parity0);
            if (stage) parity1 ^= 1u; else parity0 ^= 1u;
#pragma unroll
            for (int i = 0; i < RS_ITEMS; i++) key[i] = s_keys[wbase + i * 32];
            if (HAS_VALUES) {
#pragma unroll
                for (int i = 0; i < RS_ITEMS; i++) val[i] = s_vals[wbase + i * 32];
            }
        } else {
            const size_t gbase = (size_t)tile * RS_TILE + wbase;
#pragma unroll
            for (int i = 0; i < RS_ITEMS; i++) {
                const size_t idx = gbase + (size_t)i * 32;
                key[i] = idx < n ? keys_in[idx] : (K)~(K)0;
                if (HAS_VALUES) val[i] = idx < n ? vals_in[idx] : 0u;
            }
        }
        __syncthreads();                                         // keys are in registers: this stage is free for the exchange
        // ---- warp-level multisplit: stable rank of every key among the warp's keys with the same digit
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
            const unsigned peers = CT_BITS ? rs_match_digit_ct<CT_BITS ? CT_BITS : 1>(d) : rs_match_digit(d, bits);
            const uint32_t base = s_warp_hist[warp][d];
            __syncwarp();
            if ((peers & lt) == 0) s_warp_hist[warp][d] = (uint16_t)(base + __popc(peers));
            rank[i] = base + __popc(peers & lt);
            __syncwarp();
        }
        __syncthreads();
        // ---- per digit: exclusive offsets across warps, tile count published, decoupled look-back
        uint32_t tile_count = 0, incl = 0;
        if (tid < RS_BINS) {
            uint32_t sum = 0;
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) {
                const uint32_t t = s_warp_hist[w][tid];
                s_warp_hist[w][tid] = (uint16_t)sum;
                sum += t;
            }
            tile_count = sum;
            rs_st_status(status + (size_t)tile * RS_BINS + tid, (S)sum | (tile == 0 ? ST::INCLUSIVE : ST::PARTIAL));
            incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += t;
            }
            if (lane == 31) s_scan_tot[warp] = incl;
        }
        __syncthreads();
        if (tid < RS_BINS) {
            uint32_t pre = 0;
#pragma unroll
            for (int w = 0; w < RS_BINS / 32; w++) pre += (w < (int)warp) ? s_scan_tot[w] : 0u;
            const uint32_t bin_start = pre + incl - tile_count;
            s_bin_start[tid] = bin_start;
            // the next tile id is taken only here (its round trip overlaps the look-back): a CTA that held two ids from the start of
            // a tile (round 2a) made every successor poll for a tile that had not begun -- 40 polls per tile in ncu
            uint32_t nxt = 0;
            if (tid == 0) nxt = atomicAdd(tile_counter, 1u);
            uint32_t excl = 0;
            if (tile > 0) {
                excl = rs_look_back<S>(status, tile, tid);
                rs_st_status(status + (size_t)tile * RS_BINS + tid, (S)(excl + tile_count) | ST::INCLUSIVE);
            }
            s_global_base[tid] = __ldg(&digit_offset[tid]) + excl - bin_start;
            if (tid == 0) s_next = nxt;
        }
        __syncthreads();
        const uint32_t next = s_next;
        if (tid == 0 && next < num_tiles && ((size_t)next + 1) * RS_TILE <= n) issue(next, stage ^ 1u);   // lands during the exchange and the stores
        // ---- staging position of every key, exchange through this tile's stage buffer
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
            const uint32_t r = rank[i] + s_bin_start[d] + s_warp_hist[warp][d];
            s_keys[r] = key[i];
            if (HAS_VALUES) s_vals[r] = val[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) {
            const unsigned j = tid + i * RS_BLOCK;
            const K k = s_keys[j];
            const uint32_t d = (uint32_t)(k >> shift) & mask;
            const size_t o = (size_t)s_global_base[d] + j;
            if (full || o < n) {
                keys_out[o] = k;
                if (HAS_VALUES) vals_out[o] = s_vals[j];
            }
        }
        __syncthreads();                                         // the stage, the digit tables and s_next are free again
        tile = next;
        stage ^= 1u;
    }
}

template <typename K>
static inline size_t rs_smem_bytes(bool has_values) { return RS_TILE * sizeof(K) + (has_values ? RS_TILE * sizeof(uint32_t) : 0); }

// grid of the persistent form: every CTA resident at once (SMs x CTAs per SM for this instantiation and stage size), capped by
// the tile count.  The occupancy query and the shared-memory opt-in happen once per instantiation and device.
struct RsGridCache { const void *fn; int dev; size_t smem; size_t resident; };
template <typename F>
static inline unsigned rs_bulk_grid(F kernel, size_t dyn_smem, size_t tiles) {
    // keyed by the kernel's ADDRESS: instantiations that differ only in non-type template arguments share one function type
    static thread_local std::vector<RsGridCache> cache;
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    const void *fn = reinterpret_cast<const void *>(kernel);
    for (const RsGridCache &c : cache)
        if (c.fn == fn && c.dev == dev && c.smem == dyn_smem) return (unsigned)std::min(tiles, c.resident);
    int per_sm = 0, sms = 0;
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, RS_BLOCK, dyn_smem));
    CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CGX_REQUIRE(per_sm >= 1, "onesweep: a %zu-byte stage pair does not fit one SM", dyn_smem);
    cache.push_back({fn, dev, dyn_smem, (size_t)per_sm * (size_t)sms});
    return (unsigned)std::min(tiles, cache.back().resident);
}

// Sorts n keys (and optional payloads) on bits [begin_bit, end_bit).  keys/vals and the *_tmp buffers
// ping-pong; the sorted data ends in *keys_sorted / *vals_sorted (one of the two buffers).
template <typename K>
void radix_sort(K *keys, K *keys_tmp, uint32_t *vals, uint32_t *vals_tmp, size_t n, int begin_bit, int end_bit,
                cudaStream_t stream, RadixTemp &tmp, K **keys_sorted, uint32_t **vals_sorted, int *launches = nullptr) {
    *keys_sorted = keys;
    if (vals_sorted) *vals_sorted = vals;
    if (n <= 1) return;
    CGX_REQUIRE(n < (1ull << 32) - RS_TILE, "radix_sort: n=%zu does not fit the 32-bit digit offsets", n);
    const bool wide = n >= (size_t)RsStatus<uint32_t>::VALUE_MASK;      // 64-bit look-back words from 2^30 keys on
    CGX_REQUIRE(begin_bit >= 0 && end_bit <= (int)(8 * sizeof(K)) && (end_bit - begin_bit + 7) / 8 <= RS_MAX_PASSES,
                "radix_sort: bits [%d, %d) do not fit a %zu-bit key / %d passes", begin_bit, end_bit, 8 * sizeof(K), RS_MAX_PASSES);
    RadixPlan plan = make_radix_plan(begin_bit, end_bit);
    size_t tiles = (n + RS_TILE - 1) / RS_TILE;
    uint32_t *hist = tmp.hist.get<uint32_t>(RS_MAX_PASSES * RS_BINS);
    const size_t status_bytes = tiles * RS_BINS * (wide ? sizeof(uint64_t) : sizeof(uint32_t));
    void *status = tmp.status.get<unsigned char>(status_bytes);
    uint32_t *counters = tmp.counters.get<uint32_t>(RS_MAX_PASSES);
    CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * RS_MAX_PASSES * RS_BINS, stream));
    CUDA_CHECK(cudaMemsetAsync(counters, 0, sizeof(uint32_t) * RS_MAX_PASSES, stream));
    unsigned hgrid = (unsigned)std::min<size_t>((n + RS_BLOCK * 8 - 1) / (RS_BLOCK * 8), (size_t)CGX_NUM_SMS * 8);
    PROF("radix_histogram", (double)n * sizeof(K), rs_histogram_kernel<K><<<hgrid, RS_BLOCK, 0, stream>>>(keys, n, plan, hist));
    rs_scan_hist_kernel<<<plan.num_passes, RS_BINS, 0, stream>>>(hist);
    if (launches) *launches += 2;
    K *kin = keys, *kout = keys_tmp;
    uint32_t *vin = vals, *vout = vals_tmp;
    const size_t smem = rs_smem_bytes<K>(vals != nullptr);
    // one-tile-per-CTA form (default), or the persistent cp.async.bulk form (CGX_RS_BULK=1, 16-byte aligned buffers): measured on
    // B200 the bulk form is the slower one (profiles/README.md r02): the pass is bound by the look-back chain and by issue slots,
    // not by the latency of the key loads, and two 32-KB stages cost a quarter of the occupancy
    static const bool want_bulk = getenv("CGX_RS_BULK") != nullptr;
    const bool aligned = (((uintptr_t)keys | (uintptr_t)keys_tmp | (uintptr_t)vals | (uintptr_t)vals_tmp) & 15u) == 0;
    const bool bulk = want_bulk && aligned;
    if (!bulk) {
        CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<K, true, uint32_t, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes<K>(true)));
        CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<K, false, uint32_t, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes<K>(false)));
        CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<K, true, uint32_t, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes<K>(true)));
        CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<K, false, uint32_t, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes<K>(false)));
        CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<K, true, uint64_t, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes<K>(true)));
        CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<K, false, uint64_t, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes<K>(false)));
    }
    for (int p = 0; p < plan.num_passes; p++) {
        CUDA_CHECK(cudaMemsetAsync(status, 0, status_bytes, stream));
        const double bytes = (double)n * 2.0 * (sizeof(K) + (vals ? 4 : 0));
        const uint32_t *off = hist + p * RS_BINS;
        const bool ct8 = plan.bits[p] == 8 && !wide;
#define CGX_RS_LAUNCH(V, ST, CT, vi, vo)                                                                                                                      \
        do {                                                                                                                                              \
            if (bulk) {                                                                                                                                   \
                const unsigned g_ = rs_bulk_grid(rs_onesweep_bulk_kernel<K, V, ST, CT>, 2 * smem, tiles);                                                \
                PROF("radix_onesweep", bytes, (rs_onesweep_bulk_kernel<K, V, ST, CT><<<g_, RS_BLOCK, 2 * smem, stream>>>(kin, kout, vi, vo, n, (uint32_t)tiles, plan.shift[p], plan.bits[p], off, (ST *)status, counters + p))); \
            } else {                                                                                                                                      \
                PROF("radix_onesweep", bytes, (rs_onesweep_kernel<K, V, ST, CT><<<(unsigned)tiles, RS_BLOCK, smem, stream>>>(kin, kout, vi, vo, n, plan.shift[p], plan.bits[p], off, (ST *)status, counters + p))); \
            }                                                                                                                                             \
        } while (0)
        if (vals && ct8) CGX_RS_LAUNCH(true, uint32_t, 8, vin, vout);
        else if (!vals && ct8) CGX_RS_LAUNCH(false, uint32_t, 8, nullptr, nullptr);
        else if (vals && !wide) CGX_RS_LAUNCH(true, uint32_t, 0, vin, vout);
        else if (!vals && !wide) CGX_RS_LAUNCH(false, uint32_t, 0, nullptr, nullptr);
        else if (vals) CGX_RS_LAUNCH(true, uint64_t, 0, vin, vout);
        else CGX_RS_LAUNCH(false, uint64_t, 0, nullptr, nullptr);
#undef CGX_RS_LAUNCH
        if (launches) *launches += 1;
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    CUDA_CHECK(cudaGetLastError());
    *keys_sorted = kin;
    if (vals_sorted) *vals_sorted = vin;
}

#endif  // __CUDACC__

}  // namespace cgx
