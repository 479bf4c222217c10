// cgx-b200: gappy-phrase matching as a window scan over the occurrences of the patterns' first phrases.
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
//
// The reference (and round 1a of this file) joins the two occurrence lists of EVERY pattern: work
// sum_d min(|occ a|, |occ b|) = 6.8e9 list elements at C2 for 2.4e8 hits.  Here the loop is turned inside out:
//   * every distinct first phrase `a` of the batch is walked ONCE over its position-sorted occurrence list
//     inv[ls][up_a..down_a] (a flat, load-balanced element list: sum_a |occ a| <= 3n, 5e7 at C2);
//   * one half-warp takes one occurrence p: lane g-1 owns gap width g.  One load of the gap-consistency word
//     gapw[p+ls] (index.cu) answers "gap tokens >= 2" and checkBoundaryGap for all 13 widths at once; the lanes
//     whose width passes read the canonical m-gram ids bkt[le][p+ls+g] (coalesced: 13 consecutive words per
//     half-warp and le) -- "which phrase starts at q" without searching anything;
//   * (a, le, id of the le-gram at q) is looked up in a per-batch open-addressing hash table of the batch's
//     patterns (16-byte slots: key + pattern id in one sector), after a 1-bit-per-bucket filter ("is this m-gram
//     the second phrase of any pattern", n/8 bytes per le: L2-resident) has rejected most candidates.
// Hits are appended with warp-aggregated atomics as packed 64-bit keys (pattern | position | length) and put
// in (pattern, position, length) order by one onesweep radix sort; a boundary kernel derives the per-pattern
// ranges.  The reference's 100x100 frequent-pair cache is not needed; the one place where it leaks into
// results (featureMissingCount is added to SampleCountF for patterns made of two frequent tokens,
// ExtractPair.c:900-908) is reproduced by counting, in the same scan, the candidates that fail only the
// alignment check.
//
// aXbXc (a, b, c single tokens; GappyLook.cu:595-655): one thread per parent hit (p, L) of aXb whose pattern has
// children; the second gap's word gapw[p+L+1] yields the admissible widths g2 <= 13-L, and for each the token
// c = str[p+L+1+g2] is looked up as (parent pattern, c) in a second hash table.
#include "batch.h"
#include "hash.cuh"
#include "prof.h"
#include <utility>

namespace cgx {

// key of a one-gap pattern in the per-batch table (hash.cuh QTab): A = canonical m-gram id of the second phrase (a corpus
// position: pbits bits), B = first phrase id << 2 | length of the second phrase (bits_for(G) + 2 bits)
__device__ __forceinline__ uint32_t key1_b(uint32_t phrase_a, int le) { return (phrase_a << 2) | (uint32_t)le; }

// Hits are staged per warp in shared memory and flushed with ONE global atomic per ~200 hits (a single counter
// hit by one atomic per few hits serialises the whole grid: 56 % of the stall samples of round 1b's kernel).
constexpr int ST_CAP = 256;                    // staged keys per warp

__device__ __forceinline__ void stage_push(bool found, uint64_t key, uint64_t *__restrict__ buf, int &count) {   // all 32 lanes call
    const unsigned m = __ballot_sync(0xffffffffu, found);
    if (found) buf[count + __popc(m & lanemask_lt())] = key;
    count += __popc(m);
}

__device__ __forceinline__ void stage_flush(uint64_t *__restrict__ buf, int &count, unsigned long long *__restrict__ counter, uint64_t *__restrict__ out,
                                            size_t cap) {                                                    // all 32 lanes call
    if (count == 0) return;
    __syncwarp();
    const unsigned lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)count);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int i = (int)lane; i < count; i += 32)
        if ((size_t)base + i < cap) out[(size_t)base + i] = buf[i];
    __syncwarp();
    count = 0;
}

__device__ __forceinline__ bool bit_test(const uint32_t *__restrict__ bm, uint32_t i) { return (__ldg(&bm[i >> 5]) >> (i & 31)) & 1u; }

// ------------------------------------------------------------------------------------------------
// one-gap: per-batch tables
// ------------------------------------------------------------------------------------------------
// aflag[g]: bit 0 = phrase g is the first phrase of some pattern; bit 1 = it is a single top-100 token (only then
// can a pattern be a "marker pair" of the reference's frequent-pair table)
__global__ void j1_setup_kernel(const Pat1 *__restrict__ pat, const Pat1Dev *__restrict__ patd, const int32_t *__restrict__ pat_ga, int D1,
                                const int32_t *__restrict__ str, const uint8_t *__restrict__ freq_rank, const QTab tab, int pbits, uint32_t *__restrict__ overflow,
                                uint32_t *__restrict__ bm1, uint32_t *__restrict__ bm2, uint32_t *__restrict__ bm3, uint32_t *__restrict__ bm_marker,
                                uint32_t *__restrict__ aflag, uint32_t *__restrict__ bma, size_t bm_words, uint32_t *__restrict__ aid, size_t n) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D1) return;
    const Pat1 p = pat[d];
    const uint32_t ub = (uint32_t)patd[d].up_b;
    const uint32_t ga = (uint32_t)pat_ga[d];
    if (!qt_insert(tab, ub, key1_b(ga, p.le), ((uint32_t)d << 1) | (p.marker_pair >= 0 ? 1u : 0u))) *overflow = 1u;   // value: pattern id, marker-pair flag
    uint32_t *bm = p.le == 1 ? bm1 : p.le == 2 ? bm2 : bm3;
    atomicOr(&bm[ub >> 5], 1u << (ub & 31));
    if (p.marker_pair >= 0) atomicOr(&bm_marker[ub >> 5], 1u << (ub & 31));
    const uint32_t f = 1u | ((p.ls == 1 && freq_rank[str[p.a_pos]]) ? 2u : 0u);
    if (aflag[ga] != f) aflag[ga] = f;          // every writer of one slot writes the same value
    // position-major scan: "an ls-gram with bucket up_a is a first phrase" (bitmap) and its phrase id (valid where the bit is set;
    // entries of earlier batches are never read because the bitmap is cleared per batch)
    const uint32_t ua = (uint32_t)patd[d].up_a;
    uint32_t *ba = bma + (size_t)(p.ls - 1) * bm_words;
    if (!((ba[ua >> 5] >> (ua & 31)) & 1u)) atomicOr(&ba[ua >> 5], 1u << (ua & 31));
    const uint32_t av = ga | ((f & 2u) << 30);
    uint32_t *ai = aid + (size_t)(p.ls - 1) * n;
    if (ai[ua] != av) ai[ua] = av;
}

// elements per phrase: its occurrence count when it is the first phrase of some pattern, else 0
__global__ void j1_counts_kernel(const int32_t *__restrict__ phrases, const uint32_t *__restrict__ aflag, int G, uint32_t *__restrict__ cnt) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G) cnt[g] = (aflag[g] & 1u) ? (uint32_t)(phrases[g * 4 + 1] - phrases[g * 4] + 1) : 0u;
}

__device__ __forceinline__ int find_owner(const uint32_t *__restrict__ off, int n, uint32_t e) {   // largest g with off[g] <= e
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(&off[mid]) <= e) lo = mid; else hi = mid;
    }
    return lo;
}

constexpr int J1_BLOCK = 256;

struct J1Args {
    const int32_t *phrases;
    const uint32_t *aflag, *elem_off;
    int G;
    uint32_t n_elems;
    const int32_t *inv[3];
    const int32_t *bkt[3];
    const uint32_t *gapw;
    const uint32_t *bm[3];
    const uint32_t *bm_marker;
    QTab tab;
    int pshift;                      // hit key = pattern << pshift | position << 4 | (length-1); pshift - 4 = bits of a position
    uint32_t n;
    unsigned long long *counter;     // [0] hits, [1] bucket words read (algorithmic-byte account)
    uint64_t *hits;
    size_t cap;
    int32_t *missing;
};

__global__ void __launch_bounds__(J1_BLOCK) j1_scan_kernel(const J1Args a) {
    __shared__ uint64_t s_stage[J1_BLOCK / 32][ST_CAP];
    uint64_t *stage = s_stage[threadIdx.x >> 5];
    int staged = 0;
    const unsigned lane = threadIdx.x & 31;
    const uint32_t e = (blockIdx.x * (uint32_t)J1_BLOCK + threadIdx.x);      // one element per lane, then 2 elements per warp step
    // ---- lane-parallel: owner phrase, corpus position and gap word of 32 consecutive elements
    int my_g = -1, my_len = 0, my_p = 0;
    uint32_t my_w = 0;
    if (e < a.n_elems) {
        my_g = find_owner(a.elem_off, a.G, e);
        const int up = a.phrases[my_g * 4];
        my_len = a.phrases[my_g * 4 + 2];
        my_p = __ldg(&a.inv[my_len - 1][up + (int)(e - a.elem_off[my_g])]);
        my_w = __ldg(&a.gapw[my_p + my_len]) | (((a.aflag[my_g] >> 1) & 1u) << 31);      // bit 31: a is a frequent single token
    }
    const unsigned half = lane >> 4, h = lane & 15;
    const int g = (int)h + 1;                                              // gap width of this lane
    unsigned probes = 0;
#pragma unroll 2
    for (int it = 0; it < 16; it++) {
        const int src = it * 2 + (int)half;
        const int ga = __shfl_sync(0xffffffffu, my_g, src);
        const int ls = __shfl_sync(0xffffffffu, my_len, src);
        const int p = __shfl_sync(0xffffffffu, my_p, src);
        const uint32_t w = __shfl_sync(0xffffffffu, my_w, src);
        const int run = (int)((w >> 16) & 15u);
        // lane is live when its width leaves every gap token >= 2 and room for b
        const bool live = ga >= 0 && g <= run && g <= CGX_MAX_RULE_SPAN - 1 - ls;
        const bool ok = live && ((w >> (g - 1)) & 1u);
        const bool miss = live && !ok && (w >> 31);                       // frequent single token a: count what the pair table would miss
        const uint32_t q = (uint32_t)(p + ls + g);
        const int le_max = min(min(3, CGX_MAX_RULE_SYMBOLS - 1 - ls), CGX_MAX_RULE_SPAN - ls - g);
        // independent loads first: the canonical ids of the 1-, 2-, 3-grams at q, then their filter bits
        uint32_t ub[3];
        bool cand[3];
#pragma unroll
        for (int le = 1; le <= 3; le++) {
            cand[le - 1] = (ok && le <= le_max && q + (uint32_t)le <= a.n) || (miss && le == 1 && q < a.n);
            ub[le - 1] = cand[le - 1] ? (uint32_t)__ldg(&a.bkt[le - 1][q]) : 0u;
            probes += cand[le - 1] ? 1u : 0u;
        }
#pragma unroll
        for (int le = 1; le <= 3; le++)
            if (cand[le - 1]) cand[le - 1] = bit_test((le == 1 && miss) ? a.bm_marker : a.bm[le - 1], ub[le - 1]);
        uint32_t ss[3];
        uint32_t sw[3];
#pragma unroll
        for (int le = 1; le <= 3; le++)
            if (cand[le - 1]) qt_touch(a.tab, ub[le - 1], key1_b((uint32_t)ga, le), &ss[le - 1], &sw[le - 1]);
#pragma unroll
        for (int le = 1; le <= 3; le++) {
            if (!__any_sync(0xffffffffu, cand[le - 1])) continue;          // warp-uniform
            uint32_t v = 0;
            bool found = cand[le - 1] && qt_resolve(a.tab, ss[le - 1], sw[le - 1], &v);
            if (found && miss) {                                           // only le == 1 reaches here with miss set
                if (v & 1u) atomicAdd(&a.missing[v >> 1], 1);
                found = false;
            }
            const uint64_t key = ((uint64_t)(v >> 1) << a.pshift) | ((uint64_t)(uint32_t)p << 4) | (uint64_t)(ls + g + le - 1);
            stage_push(found, key, stage, staged);
        }
        if (staged > ST_CAP - 96) stage_flush(stage, staged, &a.counter[0], a.hits, a.cap);
    }
    stage_flush(stage, staged, &a.counter[0], a.hits, a.cap);
    for (int o = 16; o; o >>= 1) probes += __shfl_xor_sync(0xffffffffu, probes, o);
    if (lane == 0 && probes) atomicAdd(&a.counter[1], (unsigned long long)probes);
}

// ------------------------------------------------------------------------------------------------
// Position-major variant (large batches).  The phrase-major scan above fetches the 15-token window of every occurrence
// of every first phrase separately -- random 32-byte sectors, and B200 sustains only ~47 G random sectors/s
// (tools/gather_bench.cu: 1.5 TB/s against 6.5 TB/s streaming).  When most corpus positions start a first phrase anyway,
// it is cheaper to STREAM the interleaved window array jwin = {bkt1, bkt2, bkt3, gapw} once through shared memory: a CTA
// stages 256+17 consecutive positions, each half-warp takes 16 of them; lane h tests "is the m-gram at position P+h a first
// phrase" for m = 1..3 (three L2-resident bitmaps), and every (position, m) that is one runs the same 13-lane window
// logic as above on shared memory.  HBM traffic: 16 B per corpus position + the pattern-table probes.
// ------------------------------------------------------------------------------------------------
// (Measured and dropped, round 1d: candidates that pass the bitmap filters pushed to a per-warp FIFO in shared memory and probed 32
// at a time -- the probe / resolve code runs with one or two active lanes in place -- 13.1 -> 17.2 ms: the in-place form issues
// the three first-slot loads of a step together and overlaps them with the next step's window work, the drained form waits for
// every probe.)
constexpr int JP_TILE = 256;         // positions per CTA
constexpr int JP_HALO = 17;          // q <= p + 3 + 13

struct JPArgs {
    const int4 *jwin;
    uint32_t n;
    const uint32_t *bma[3];          // first-phrase bitmaps by length
    const uint32_t *aid[3];          // bucket -> phrase id | frequent-single-token flag << 31
    const uint32_t *bm[3];
    const uint32_t *bm_marker;
    QTab tab;
    int pshift;
    int mask_shift;                  // bit of the key where the 10-bit second-gap word goes (H1_MASK_SHIFT), 64 = not carried
    unsigned long long *counter;     // [0] hits, [1] table lookups, [2] elements
    uint64_t *hits;
    size_t cap;
    int32_t *missing;
};

__global__ void __launch_bounds__(JP_TILE) j1_pos_kernel(const JPArgs a) {
    __shared__ int4 s_win[JP_TILE + JP_HALO];
    __shared__ uint64_t s_stage[JP_TILE / 32][ST_CAP];
    uint64_t *stage = s_stage[threadIdx.x >> 5];
    int staged = 0;
    const unsigned lane = threadIdx.x & 31, half = lane >> 4, h = lane & 15;
    const uint32_t P0 = blockIdx.x * (uint32_t)JP_TILE;
    for (int i = threadIdx.x; i < JP_TILE + JP_HALO; i += JP_TILE) {
        const uint32_t p = P0 + i;
        s_win[i] = p < a.n ? __ldg(&a.jwin[p]) : make_int4(0, 0, 0, 0);
    }
    __syncthreads();
    // this half-warp's 16 positions: lane h owns position hp + h while the first-phrase tests run
    const int hw = (int)(threadIdx.x >> 4);                     // 0..15
    const int my = hw * 16 + (int)h;                            // tile-relative position of this lane
    const int4 mine = s_win[my];
    uint32_t my_aid[3];
    unsigned my_mask = 0;
    const bool inside = P0 + my < a.n;
#pragma unroll
    for (int m = 0; m < 3; m++) {
        const uint32_t ub = (uint32_t)(m == 0 ? mine.x : m == 1 ? mine.y : mine.z);
        my_aid[m] = 0;
        if (inside && bit_test(a.bma[m], ub)) { my_mask |= 1u << m; my_aid[m] = __ldg(&a.aid[m][ub]); }
    }
    const int g = (int)h + 1;
    unsigned lookups = 0, elems = 0;
    // walk the 16 positions of both half-warps in lock step (warp-uniform trip count)
    for (int k = 0; k < 16; k++) {
        const int src = (int)(half * 16) + k;
        const unsigned pm = __shfl_sync(0xffffffffu, my_mask, src);
        const uint32_t a0 = __shfl_sync(0xffffffffu, my_aid[0], src), a1 = __shfl_sync(0xffffffffu, my_aid[1], src), a2 = __shfl_sync(0xffffffffu, my_aid[2], src);
        if (!__any_sync(0xffffffffu, pm != 0)) continue;
        const int rel = hw * 16 + k;                              // tile-relative position p
        const int p = (int)P0 + rel;
#pragma unroll
        for (int ls = 1; ls <= 3; ls++) {
            const bool on = (pm >> (ls - 1)) & 1u;
            if (!__any_sync(0xffffffffu, on)) continue;
            const uint32_t av = ls == 1 ? a0 : ls == 2 ? a1 : a2;
            const uint32_t ga = av & 0x7fffffffu;
            const uint32_t w = on ? (uint32_t)s_win[rel + ls].w : 0u;
            const int run = (int)((w >> 16) & 15u);
            const bool live = on && g <= run && g <= CGX_MAX_RULE_SPAN - 1 - ls;
            const bool ok = live && ((w >> (g - 1)) & 1u);
            const bool miss = live && !ok && (av >> 31);
            const int qrel = rel + ls + g;
            const uint32_t q = (uint32_t)(p + ls + g);
            const int le_max = min(min(3, CGX_MAX_RULE_SYMBOLS - 1 - ls), CGX_MAX_RULE_SPAN - ls - g);
            if (h == 0 && on) elems++;
            uint32_t ub[3];
            bool cand[3];
            const int4 wq = (ok || miss) ? s_win[qrel] : make_int4(0, 0, 0, 0);
#pragma unroll
            for (int le = 1; le <= 3; le++) {
                cand[le - 1] = (ok && le <= le_max && q + (uint32_t)le <= a.n) || (miss && le == 1 && q < a.n);
                ub[le - 1] = (uint32_t)(le == 1 ? wq.x : le == 2 ? wq.y : wq.z);
            }
#pragma unroll
            for (int le = 1; le <= 3; le++)
                if (cand[le - 1]) cand[le - 1] = bit_test((le == 1 && miss) ? a.bm_marker : a.bm[le - 1], ub[le - 1]);
                uint32_t ss[3];
            uint32_t sw[3];
#pragma unroll
            for (int le = 1; le <= 3; le++)
                if (cand[le - 1]) { qt_touch(a.tab, ub[le - 1], key1_b(ga, le), &ss[le - 1], &sw[le - 1]); lookups++; }
#pragma unroll
            for (int le = 1; le <= 3; le++) {
                if (!__any_sync(0xffffffffu, cand[le - 1])) continue;
                uint32_t v = 0;
                bool found = cand[le - 1] && qt_resolve(a.tab, ss[le - 1], sw[le - 1], &v);
                if (found && miss) {
                    if (v & 1u) atomicAdd(&a.missing[v >> 1], 1);
                    found = false;
                }
                const uint64_t key = ((uint64_t)(v >> 1) << a.pshift) | ((uint64_t)(uint32_t)p << 4) | (uint64_t)(ls + g + le - 1);
                stage_push(found, key, stage, staged);
            }
            if (staged > ST_CAP - 96) stage_flush(stage, staged, &a.counter[0], a.hits, a.cap);
        }
    }
    stage_flush(stage, staged, &a.counter[0], a.hits, a.cap);
    for (int o = 16; o; o >>= 1) { lookups += __shfl_xor_sync(0xffffffffu, lookups, o); elems += __shfl_xor_sync(0xffffffffu, elems, o); }
    if (lane == 0 && (lookups | elems)) { atomicAdd(&a.counter[1], (unsigned long long)lookups); atomicAdd(&a.counter[2], (unsigned long long)elems); }
}


// ------------------------------------------------------------------------------------------------
// Ordered emission (round 2).  Round 1 appended the hits with atomics in arbitrary order and put them in (pattern, position,
// length) order with a stable radix sort over pattern AND position bits: six 8-bit passes over 2.4e8 keys at C2, seven at C3.
// The position-major scan visits the corpus in position order anyway: a tile (256 positions) now stages ALL its hits in shared
// memory in (position, length) order per pattern, takes one contiguous segment of the output with a single atomicAdd, and
// records (segment start, count) under its tile number.  A prefix sum over the counts in tile order and one streaming copy
// (16 B per hit, ~1 ms at C2) then lay the segments out in position order, so the stable sort only has to cover the pattern
// bits: three passes instead of six (seven).  (First attempt, measured and dropped: a decoupled look-back over the per-tile
// counts, as in the onesweep pass, so that tiles write straight to their final place.  A tile knows its count only when its
// walk is finished and tile durations are heavy-tailed -- hits per tile vary by orders of magnitude -- so every tile behind
// a slow one spun in the look-back holding its SM slot: join_onegap 13.0 -> 20.4 ms, join_twogap 11.0 -> 19.0 ms at C2.)
// ------------------------------------------------------------------------------------------------
constexpr int JO_TILE = 256;          // positions per CTA (8 warps x 32 positions)
constexpr int JO_CAP = 1024;          // most staged hits per warp the kernel may be given (6 bytes of dynamic shared memory each)
constexpr int JO_CAP_DEFAULT = 512;   // default: 24 KB per CTA, five CTAs per SM; a fuller stage is flushed as one more chunk of the warp

// A one-gap hit key: [second-gap word : 10][...][pattern : dbits][position : pbits][length - 1 : 4].  The top 10 bits carry the
// admissible widths 1..10 of a gap that would follow the hit (what gapw[p + L + 1] says, cut to the rule span) when the pattern,
// position and length fields leave room for them (Batch::h1_mask); the two-gap join reads them instead of gathering the word.
constexpr int H1_MASK_SHIFT = 54;

struct JOArgs {
    JPArgs p;
    uint32_t *seg_base;               // per (warp, chunk): start of the chunk in the (unordered) output
    uint32_t *seg_count;              // per (warp, chunk): hits (zeroed before the launch)
    unsigned chunks_per_warp;         // chunk records per warp: enough for the most hits 32 positions can have
    const uint8_t *flags;             // per position: first- / second-phrase bits (j1_flags_kernel)
    unsigned stage_cap;               // staged hits per chunk, 32..JO_CAP (tests lower it to drive warps through many chunks)
};

// Per-batch position flags for the ordered scan: bit m-1 (m = 1..3) = the m-gram at this position is the first phrase of some
// pattern, bit 2+m = it is the second phrase of some pattern, bit 6 = it is the second token of a frequent-pair pattern.  One
// streamed pass (16 B read, 1 B written per position, seven L2-resident bitmap probes with every lane busy) takes the bitmap
// probes out of the walk, where each was a dependent load in front of the table probe (11 % of its stall samples, r02b).
__global__ void __launch_bounds__(256) j1_flags_kernel(const int4 *__restrict__ jwin, uint32_t n, const uint32_t *__restrict__ bma0, const uint32_t *__restrict__ bma1,
                                                       const uint32_t *__restrict__ bma2, const uint32_t *__restrict__ bm0, const uint32_t *__restrict__ bm1,
                                                       const uint32_t *__restrict__ bm2, const uint32_t *__restrict__ bm_marker, uint8_t *__restrict__ flags) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int4 w = __ldg(&jwin[p]);
    unsigned f = 0;
    f |= bit_test(bma0, (uint32_t)w.x) ? 1u : 0u;
    f |= bit_test(bma1, (uint32_t)w.y) ? 2u : 0u;
    f |= bit_test(bma2, (uint32_t)w.z) ? 4u : 0u;
    f |= bit_test(bm0, (uint32_t)w.x) ? 8u : 0u;
    f |= bit_test(bm1, (uint32_t)w.y) ? 16u : 0u;
    f |= bit_test(bm2, (uint32_t)w.z) ? 32u : 0u;
    f |= bit_test(bm_marker, (uint32_t)w.x) ? 64u : 0u;
    flags[p] = (uint8_t)f;
}

// One walk over the 32 positions of this warp.  DIRECT = false: hits go to the warp's shared-memory stage (pattern id +
// position-in-tile/length word) while they fit, and are counted either way; the featureMissingCount side effect and the
// statistics happen here.  DIRECT = true (second walk of a warp whose stage overflowed): hits are written to their final place.
//
// Candidates -- (position, first phrase, gap width, second phrase) combinations that passed the gap word and the bitmaps -- are not
// probed where they are found (8 of 32 lanes active there, and a third of the kernel's stall samples sat on those loads:
// profiles/r02a) but queued per warp in shared memory, in the order the walk finds them, and probed 32 at a time in two steps:
// a batch's bucket sectors are requested when it leaves the queue and read when the NEXT batch leaves it, so their latency
// hides behind the walk in between.  Hits are staged in queue order = (position, first-phrase length, second-phrase length,
// gap width) order; two hits of ONE pattern differ in position or width only, so every pattern's hits stay in (position, length) order.
struct JQueue {
    uint32_t *a, *b;      // table key halves (QTab): m-gram id of the second phrase / first phrase id << 2 | le
    uint16_t *m;          // tile-relative position << 4 | length - 1 ... bit 15: "miss" candidate (featureMissingCount)
};
constexpr int JQ_CAP = 64;

// where a warp's hits go: a stage in shared memory that is flushed to the output as one CHUNK whenever the next batch of probes
// might not fit -- a chunk takes its place in the output with one atomicAdd and is recorded as (start, count) under
// (warp number, chunk number); chunks are later copied into (warp, chunk) order, i.e. position order
struct JStage {
    uint32_t *pat;                   // staged pattern ids (< 2^26); bits 26..31: high 6 bits of the second-gap word (H1_MASK_SHIFT)
    uint16_t *pl;                    // staged tile-relative position << 4 | length - 1; bits 12..15: low 4 bits of the second-gap word
    unsigned cap;                    // staged hits per chunk at most (>= 32)
    uint32_t *seg_base, *seg_count;  // this warp's chunk records (chunks_per_warp of each)
};

__device__ __forceinline__ unsigned jo_walk(const JPArgs &a, const int4 *__restrict__ s_win, const uint8_t *__restrict__ s_flags, int wrel, uint32_t P0, unsigned my_mask, const uint32_t my_aid[3],
                                            const JStage st, const JQueue q, unsigned &lookups, unsigned &elems) {
    const unsigned lane = threadIdx.x & 31, half = lane >> 4, h = lane & 15;
    const int g = (int)h + 1;
    unsigned count = 0, staged = 0, chunk = 0, queued = 0;         // warp-uniform
    uint32_t pend_home = 0, pend_tag = 0, pend_meta = 0;           // this lane's probe in flight
    bool pend = false;
    auto flush = [&]() {                                           // all lanes call: the stage becomes the warp's next chunk
        if (staged == 0) return;
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(&a.counter[0], (unsigned long long)staged);
            st.seg_base[chunk] = (uint32_t)base;                   // hits per batch < 2^31 (hit_limit); a batch beyond is refused before its list is read
            st.seg_count[chunk] = staged;
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (unsigned i = lane; i < staged; i += 32) {
            const unsigned long long dst = base + i;
            const unsigned pl = st.pl[i];
            const uint32_t pt = st.pat[i];
            const uint64_t mk = (uint64_t)(((pt >> 26) << 4) | (pl >> 12));          // second-gap word (0 when not carried)
            if (dst < a.cap) a.hits[dst] = (a.mask_shift < 64 ? mk << a.mask_shift : 0ull) | ((uint64_t)(a.mask_shift < 64 ? pt & 0x3FFFFFFu : pt) << a.pshift) | ((uint64_t)(P0 + ((pl >> 4) & 0xFFu)) << 4) | (uint64_t)(pl & 15u);
        }
        __syncwarp();
        count += staged;
        staged = 0;
        chunk++;
    };
    auto resolve_pending = [&]() {                                 // all lanes call
        uint32_t v = 0;
        bool found = pend && qt_resolve(a.tab, pend_home, pend_tag, &v);
        if (found && (pend_meta & 0x8000u)) {                      // frequent single-token pair failing only the alignment check
            if (v & 1u) atomicAdd(&a.missing[v >> 1], 1);
            found = false;
        }
        const unsigned m = __ballot_sync(0xffffffffu, found);
        if (staged + __popc(m) > st.cap) flush();
        if (found) {
            const unsigned idx = staged + __popc(m & lanemask_lt());
            // the gap word behind the hit is in the window: the two-gap join then needs no gather for it (at C3 that gather, one
            // random DRAM sector per parent hit, was half of what bound j2_ordered_kernel).  It rides in the spare bits of the two
            // stage words (mask_shift = 64: no room in the key, nothing carried).
            const unsigned len_ = pend_meta & 15u;
            unsigned mk = 0;
            if (a.mask_shift < 64) {
                const uint32_t w2 = (uint32_t)s_win[((pend_meta & 0x7fffu) >> 4) + len_ + 1].w;
                const int gmax = min((int)((w2 >> 16) & 15u), CGX_MAX_RULE_SPAN - 2 - (int)len_);
                mk = gmax > 0 ? (w2 & ((1u << gmax) - 1u) & 0x3FFu) : 0u;
            }
            st.pat[idx] = (v >> 1) | ((mk >> 4) << 26);
            st.pl[idx] = (uint16_t)((pend_meta & 0x0fffu) | ((mk & 15u) << 12));
        }
        staged += __popc(m);
        pend = false;
    };
    auto drain = [&](unsigned n_take) {                            // all lanes call; n_take <= 32 entries leave the queue
        resolve_pending();
        if (lane < n_take) {
            qt_touch(a.tab, q.a[lane], q.b[lane], &pend_home, &pend_tag);
            pend_meta = q.m[lane];
            pend = true;
        }
        __syncwarp();
        const unsigned rest = queued - n_take;
        uint32_t ta = 0, tb = 0; uint16_t tm = 0;
        if (lane < rest) { ta = q.a[n_take + lane]; tb = q.b[n_take + lane]; tm = q.m[n_take + lane]; }
        __syncwarp();
        if (lane < rest) { q.a[lane] = ta; q.b[lane] = tb; q.m[lane] = tm; }
        __syncwarp();
        queued = rest;
    };
    for (int k = 0; k < 16; k++) {
        const int src = 2 * k + (int)half;                         // half 0: even positions, half 1: odd ones -> pushes are in position order
        const unsigned pm = __shfl_sync(0xffffffffu, my_mask, src);
        const uint32_t a0 = __shfl_sync(0xffffffffu, my_aid[0], src), a1 = __shfl_sync(0xffffffffu, my_aid[1], src), a2 = __shfl_sync(0xffffffffu, my_aid[2], src);
        if (!__any_sync(0xffffffffu, pm != 0)) continue;
        const int rel = wrel + src;                                // tile-relative position p
        const int p = (int)P0 + rel;
#pragma unroll 1
        for (int ls = 1; ls <= 3; ls++) {
            const bool on = (pm >> (ls - 1)) & 1u;
            if (!__any_sync(0xffffffffu, on)) continue;
            const uint32_t av = ls == 1 ? a0 : ls == 2 ? a1 : a2;
            const uint32_t ga = av & 0x7fffffffu;
            const uint32_t w = on ? (uint32_t)s_win[rel + ls].w : 0u;
            const int run = (int)((w >> 16) & 15u);
            const bool live = on && g <= run && g <= CGX_MAX_RULE_SPAN - 1 - ls;
            const bool ok = live && ((w >> (g - 1)) & 1u);
            const bool miss = live && !ok && (av >> 31);
            const int qrel = rel + ls + g;
            const uint32_t qpos = (uint32_t)(p + ls + g);
            const int le_max = min(min(3, CGX_MAX_RULE_SYMBOLS - 1 - ls), CGX_MAX_RULE_SPAN - ls - g);
            if (h == 0 && on) elems++;
            uint32_t ub[3];
            bool cand[3];
            const int4 wq = (ok || miss) ? s_win[qrel] : make_int4(0, 0, 0, 0);
#pragma unroll
            for (int le = 1; le <= 3; le++) {
                cand[le - 1] = (ok && le <= le_max && qpos + (uint32_t)le <= a.n) || (miss && le == 1 && qpos < a.n);
                ub[le - 1] = (uint32_t)(le == 1 ? wq.x : le == 2 ? wq.y : wq.z);
            }
            const unsigned qf = (ok || miss) ? (unsigned)s_flags[qrel] : 0u;      // "second phrase of some pattern" bits of position q
#pragma unroll
            for (int le = 1; le <= 3; le++)
                cand[le - 1] = cand[le - 1] && ((qf >> ((le == 1 && miss) ? 6 : 2 + le)) & 1u);
#pragma unroll
            for (int le = 1; le <= 3; le++) {
                const unsigned m = __ballot_sync(0xffffffffu, cand[le - 1]);
                if (!m) continue;                                  // warp-uniform
                if (cand[le - 1]) {
                    const unsigned at = queued + __popc(m & lanemask_lt());
                    q.a[at] = ub[le - 1];
                    q.b[at] = key1_b(ga, le);
                    q.m[at] = (uint16_t)((rel << 4) | (ls + g + le - 1) | (miss ? 0x8000 : 0));
                    lookups++;
                }
                queued += __popc(m);
                __syncwarp();
                if (queued >= 32) drain(32);
            }
        }
    }
    if (queued) drain(queued);
    resolve_pending();
    flush();
    return count;
}

__global__ void __launch_bounds__(JO_TILE) j1_pos_ordered_kernel(const JOArgs o) {
    const JPArgs &a = o.p;
    __shared__ __align__(16) int4 s_win[JO_TILE + JP_HALO];
    extern __shared__ __align__(16) unsigned char s_dyn[];          // per warp: stage_cap pattern ids (u32); then per warp stage_cap position/length words (u16)
    __shared__ uint8_t s_flags[JO_TILE + JP_HALO + 7];
    __shared__ uint32_t s_qa[JO_TILE / 32][JQ_CAP], s_qb[JO_TILE / 32][JQ_CAP];
    __shared__ uint16_t s_qm[JO_TILE / 32][JQ_CAP];
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const JQueue jq{s_qa[warp], s_qb[warp], s_qm[warp]};
    uint32_t *st_pat = reinterpret_cast<uint32_t *>(s_dyn) + (size_t)warp * o.stage_cap;
    uint16_t *st_pl = reinterpret_cast<uint16_t *>(s_dyn + sizeof(uint32_t) * o.stage_cap * (JO_TILE / 32)) + (size_t)warp * o.stage_cap;
    const uint32_t tile = blockIdx.x;
    const uint32_t P0 = tile * (uint32_t)JO_TILE;
    // the tile's window -- 273 consecutive 16-byte position records, 4.4 KB -- arrives by ONE bulk copy (cp.async.bulk, the TMA
    // unit's 1-D form; SASS UBLKCP) issued by one thread and signalled on an mbarrier, instead of 273 LDG.128 + STS.128 with
    // their address arithmetic; positions beyond the corpus are zero-filled by the threads meanwhile
    __shared__ uint64_t s_bar;
    const unsigned n_valid = min((unsigned)(JO_TILE + JP_HALO), a.n - P0);
    if (tid == 0) { rs_mbar_init(&s_bar, 1); rs_fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
        rs_mbar_expect_tx(&s_bar, n_valid * (unsigned)sizeof(int4));
        rs_bulk_g2s(s_win, a.jwin + P0, n_valid * (unsigned)sizeof(int4), &s_bar);
    }
    for (unsigned i = n_valid + tid; i < (unsigned)(JO_TILE + JP_HALO); i += JO_TILE) s_win[i] = make_int4(0, 0, 0, 0);
    for (unsigned i = tid; i < (unsigned)(JO_TILE + JP_HALO); i += JO_TILE) s_flags[i] = i < n_valid ? __ldg(&o.flags[P0 + i]) : (uint8_t)0;
    rs_mbar_wait(&s_bar, 0);
    __syncthreads();                                                // the last CTA-wide barrier: the warps are independent from here on
    // lane l of warp w tests "does a first phrase start at position 32 w + l" for the three lengths
    const int wrel = (int)warp * 32;
    const int4 mine = s_win[wrel + (int)lane];
    const unsigned my_flags = s_flags[wrel + (int)lane];
    uint32_t my_aid[3];
    unsigned my_mask = 0;
    const bool inside = P0 + (uint32_t)wrel + lane < a.n;
#pragma unroll
    for (int m = 0; m < 3; m++) {
        const uint32_t ub = (uint32_t)(m == 0 ? mine.x : m == 1 ? mine.y : mine.z);
        my_aid[m] = 0;
        if (inside && ((my_flags >> m) & 1u)) { my_mask |= 1u << m; my_aid[m] = __ldg(&a.aid[m][ub]); }
    }
    unsigned lookups = 0, elems = 0;
    // output chunks are per WARP (32 positions).  (Round 2a took one segment per tile: the warps of a tile then met at a barrier to add
    // their counts up, and since hits per warp are heavy-tailed, 15 % of the kernel's stall samples were warps waiting there for
    // the slowest of the eight.  Round 2b staged a whole warp and walked a second time, writing directly, when the stage
    // overflowed: 1024-hit stages cost occupancy, 512-hit stages cost second walks -- 10.1 vs 13.6 ms.)
    const size_t seg = ((size_t)tile * (JO_TILE / 32) + warp) * o.chunks_per_warp;
    const JStage st{st_pat, st_pl, o.stage_cap, o.seg_base + seg, o.seg_count + seg};
    jo_walk(a, s_win, s_flags, wrel, P0, my_mask, my_aid, st, jq, lookups, elems);
    for (int off = 16; off; off >>= 1) { lookups += __shfl_xor_sync(0xffffffffu, lookups, off); elems += __shfl_xor_sync(0xffffffffu, elems, off); }
    if (lane == 0 && (lookups | elems)) { atomicAdd(&a.counter[1], (unsigned long long)lookups); atomicAdd(&a.counter[2], (unsigned long long)elems); }
}

// segments (in arrival order) -> their own order: dst[seg_dst[t] + i] = src[seg_base[t] + i].  One warp per group of `per`
// consecutive segment records (the chunks of one scanning warp, filled in order: the first empty one ends the group)
template <class BaseT>
__global__ void __launch_bounds__(256) seg_copy_kernel(const uint64_t *__restrict__ src, const BaseT *__restrict__ seg_base, const uint32_t *__restrict__ seg_dst,
                                                       const uint32_t *__restrict__ seg_count, uint32_t n_groups, uint32_t per, uint64_t *__restrict__ dst) {
    const uint32_t grp = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (grp >= n_groups) return;
    for (uint32_t k = 0; k < per; k++) {
        const size_t t = (size_t)grp * per + k;
        const uint32_t cnt = seg_count[t];
        if (cnt == 0) { if (per > 1) break; else return; }
        const uint64_t *s = src + seg_base[t];
        uint64_t *d = dst + seg_dst[t];
        for (uint32_t i = threadIdx.x & 31; i < cnt; i += 32) d[i] = s[i];
    }
}

// per-pattern [start,count] in the sorted hit list
__global__ void hit_ranges_kernel(const uint64_t *__restrict__ hits, size_t n, int shift, uint32_t dmask, int32_t *__restrict__ start_count, int stride_ints) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t d = (uint32_t)(hits[k] >> shift) & dmask;
    if (k == 0 || ((uint32_t)(hits[k - 1] >> shift) & dmask) != d) start_count[(size_t)d * stride_ints + 0] = (int32_t)k;
    if (k == n - 1 || ((uint32_t)(hits[k + 1] >> shift) & dmask) != d) start_count[(size_t)d * stride_ints + 1] = (int32_t)k;   // end; fixed up below
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

static void read_u64s(unsigned long long *dst, const unsigned long long *d, int count, cudaStream_t stream) {
    cgx_read_back(dst, d, sizeof(unsigned long long) * count, stream);
}

static void prof_add_bytes(const char *name, double bytes) {
    if (g_prof && g_prof->enabled) g_prof->table[name].bytes += bytes;
}

void stage_onegap_join(const Index &ix, Batch &b, cudaStream_t stream) {
    b.hits1 = 0;
    const int D1 = b.D1, G = b.G;
    b.pbits = cgx_bits_for((uint64_t)ix.n);
    if (D1 == 0) return;
    CGX_REQUIRE(cgx_bits_for((uint64_t)D1) + b.pbits + 4 <= 64 && ix.n < (1ull << 30) && D1 < (1 << 30), "one-gap join: pattern/corpus size exceeds the hit-key fields");
    const size_t bm_words = (ix.n + 31) / 32 + 1;
    uint32_t *bm = b.j_bitmaps.get<uint32_t>(7 * bm_words);            // [0..2] second phrases by length, [3] marker, [4..6] first phrases
    uint32_t *aid = b.j_aid.get<uint32_t>(3 * ix.n);                   // bucket -> first-phrase id, valid where the first-phrase bit is set
    uint32_t *aflag = b.j_aflag.get<uint32_t>((size_t)G + 1);
    uint32_t *eoff = b.j_tiles.get<uint32_t>((size_t)G + 2);
    // per-batch pattern table (hash.cuh QTab): key = (first phrase id, le, m-gram id of the second phrase), value = pattern id, marker flag
    QTab tab;
    tab.abits = std::max(b.pbits, 10); tab.bbits = cgx_bits_for((uint64_t)G) + 2;
    CGX_REQUIRE_BATCH(tab.bbits <= 24 && D1 < (1 << 30), "%d phrases / %d one-gap patterns exceed the pattern-table fields", G, D1);
    uint32_t buckets = qt_buckets_for((size_t)D1, tab.abits + tab.bbits);
    if (b.j1_buckets > buckets && b.j1_buckets <= 4 * buckets) buckets = b.j1_buckets;       // an earlier batch of this size needed more room
    uint32_t *tot = b.counters.get<uint32_t>(32);
    unsigned long long *ctr = (unsigned long long *)(tot + 4);         // [0] hits [1] bucket words read
    uint32_t n_elems = 0;
    while (true) {
        tab.rb = tab.abits + tab.bbits - qt_log2(buckets); tab.bmask = buckets - 1;
        CGX_REQUIRE(tab.rb >= 1 && tab.rb <= 27, "one-gap pattern table: %d remainder bits", tab.rb);
        tab.slots = b.j_hash.get<unsigned long long>((size_t)buckets * 4);
        CUDA_CHECK(cudaMemsetAsync(bm, 0, sizeof(uint32_t) * 7 * bm_words, stream));
        CUDA_CHECK(cudaMemsetAsync(aflag, 0, sizeof(uint32_t) * ((size_t)G + 1), stream));
        CUDA_CHECK(cudaMemsetAsync(tab.slots, 0xff, sizeof(unsigned long long) * (size_t)buckets * 4, stream));
        CUDA_CHECK(cudaMemsetAsync(tot + 1, 0, sizeof(uint32_t), stream));
        PROF("join_setup", (double)D1 * (32 + 16 + 4 + 8), (j1_setup_kernel<<<cgx_div_up(D1, 256), 256, 0, stream>>>(b.pat1.ptr<Pat1>(), b.pat1_dev.ptr<Pat1Dev>(), b.pat1_ga.ptr<int32_t>(), D1, ix.str.ptr<int32_t>(),
                                                                ix.freq_flag.ptr<uint8_t>(), tab, b.pbits, tot + 1, bm, bm + bm_words, bm + 2 * bm_words, bm + 3 * bm_words, aflag,
                                                                bm + 4 * bm_words, bm_words, aid, ix.n)));
        j1_counts_kernel<<<cgx_div_up(G, 256), 256, 0, stream>>>(b.phrases.ptr<int32_t>(), aflag, G, eoff);
        exclusive_scan_u32(eoff, eoff, (size_t)G, tot, stream, b.scan, 0, &b.launches);
        b.launches += 2;
        uint32_t back[2] = {0, 0};                                                             // [0] elements, [1] an insert found no room
        cgx_read_back(back, tot, sizeof(back), stream);
        n_elems = back[0];
        if (!back[1]) break;
        buckets *= 2;                                                                          // (a cluster of full buckets: rare at <= 3 entries per bucket)
        CGX_REQUIRE(buckets <= (1u << 28), "one-gap pattern table does not settle");
    }
    b.j1_buckets = buckets;
    int32_t *missing = b.missing.get<int32_t>((size_t)D1);
    if (b.hit_cap == 0) b.hit_cap = 1u << 22;
    J1Args a;
    a.phrases = b.phrases.ptr<int32_t>(); a.aflag = aflag; a.elem_off = eoff; a.G = G; a.n_elems = n_elems;
    for (int k = 0; k < 3; k++) { a.inv[k] = ix.inv[k].ptr<int32_t>(); a.bkt[k] = ix.bkt[k].ptr<int32_t>(); a.bm[k] = bm + (size_t)k * bm_words; }
    a.gapw = ix.gapw.ptr<uint32_t>(); a.bm_marker = bm + 3 * bm_words; a.tab = tab;
    a.pshift = b.pbits + 4; a.n = (uint32_t)ix.n; a.counter = ctr; a.missing = missing;
    // variant: stream the whole corpus once (position-major) when the first phrases cover a good part of it, else walk
    // the occurrence lists (phrase-major).  CGX_JOIN_MODE=phrase|position forces one (tests exercise both).
    bool position_major = (size_t)n_elems * 4 >= ix.n;
    if (const char *e = getenv("CGX_JOIN_MODE")) { if (!strcmp(e, "phrase")) position_major = false; else if (!strcmp(e, "position")) position_major = true; }
    JPArgs ap;
    ap.jwin = ix.jwin.ptr<int4>(); ap.n = (uint32_t)ix.n;
    for (int k = 0; k < 3; k++) { ap.bma[k] = bm + (size_t)(4 + k) * bm_words; ap.aid[k] = aid + (size_t)k * ix.n; ap.bm[k] = bm + (size_t)k * bm_words; }
    ap.bm_marker = bm + 3 * bm_words; ap.tab = tab; ap.pshift = b.pbits + 4; ap.counter = ctr; ap.missing = missing;
    ap.mask_shift = 64;
    unsigned long long host_ctr[3] = {0, 0, 0};
    // position-major scans emit in position order (tile segments + ordered copy): the sort below then covers the pattern bits only.
    // CGX_JOIN_ORDERED=0 keeps round 1's unordered append + full (pattern, position) sort for that variant too.
    bool ordered = position_major;
    if (const char *e = getenv("CGX_JOIN_ORDERED")) { if (!strcmp(e, "0")) ordered = false; }
    const uint32_t n_tiles = cgx_div_up(ix.n, JO_TILE);
    JOArgs ao;
    uint32_t *seg_dst = nullptr;
    size_t n_segs = 0;                                                                         // chunk records: warps x chunks per warp
    if (ordered) {
        ao.stage_cap = JO_CAP_DEFAULT;
        if (const char *e = getenv("CGX_JOIN_STAGE_CAP")) { const unsigned v = (unsigned)strtoul(e, nullptr, 10); if (v >= 32 && v <= (unsigned)JO_CAP) ao.stage_cap = v; }
        // a chunk leaves the stage with at least cap - 31 hits; 32 positions have at most 32 x 3 x 13 x 3 candidates
        ao.chunks_per_warp = (32u * 3u * 13u * 3u) / (ao.stage_cap - 31u) + 2u;
        n_segs = (size_t)n_tiles * (JO_TILE / 32) * ao.chunks_per_warp;
        CGX_REQUIRE_BATCH(n_segs < (1ull << 32), "%zu output chunks", n_segs);
        ao.seg_base = b.j_status.get<uint32_t>(n_segs + 1);
        ao.seg_count = b.j_segcnt.get<uint32_t>(2 * n_segs + 2);
        seg_dst = ao.seg_count + n_segs + 1;
        if (!b.j1_smem_opt_in) {      // per context = per device: the opt-in is a per-device attribute of the kernel
            CUDA_CHECK(cudaFuncSetAttribute(j1_pos_ordered_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, JO_TILE / 32 * JO_CAP * 6));
            b.j1_smem_opt_in = true;
        }
    }
    while (true) {
        a.hits = ap.hits = b.hit_keys.get<uint64_t>(b.hit_cap);
        a.cap = ap.cap = b.hit_cap;
        CUDA_CHECK(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long) * 3, stream));
        CUDA_CHECK(cudaMemsetAsync(missing, 0, sizeof(int32_t) * (size_t)D1, stream));
        b.h1_mask = false;
        if (n_elems && ordered) {
            b.h1_mask = cgx_bits_for((uint64_t)D1) + b.pbits + 4 <= H1_MASK_SHIFT && cgx_bits_for((uint64_t)D1) <= 26;
            if (const char *e = getenv("CGX_JOIN_CARRY_GAPW")) { if (!strcmp(e, "0")) b.h1_mask = false; }
            ap.mask_shift = b.h1_mask ? H1_MASK_SHIFT : 64;
            ao.p = ap;
            uint8_t *fl = b.j_flags.get<uint8_t>((size_t)ix.n + 64);
            PROF("join_setup", 17.0 * (double)ix.n, (j1_flags_kernel<<<cgx_div_up(ix.n, 256), 256, 0, stream>>>(ap.jwin, ap.n, ap.bma[0], ap.bma[1], ap.bma[2], ap.bm[0], ap.bm[1], ap.bm[2],
                                                                                                   ap.bm_marker, fl)));
            b.launches++;
            ao.flags = fl;
            CUDA_CHECK(cudaMemsetAsync(ao.seg_count, 0, sizeof(uint32_t) * n_segs, stream));        // unused chunk records count zero hits
            PROF("join_onegap", 0.0, (j1_pos_ordered_kernel<<<n_tiles, JO_TILE, (size_t)(JO_TILE / 32) * ao.stage_cap * 6, stream>>>(ao)));
        } else if (n_elems && position_major) PROF("join_onegap", 0.0, (j1_pos_kernel<<<cgx_div_up(ix.n, JP_TILE), JP_TILE, 0, stream>>>(ap)));
        else if (n_elems) PROF("join_onegap", 0.0, (j1_scan_kernel<<<cgx_div_up(n_elems, J1_BLOCK), J1_BLOCK, 0, stream>>>(a)));
        b.launches += 1;
        read_u64s(host_ctr, ctr, 3, stream);
        CGX_REQUIRE_BATCH(host_ctr[0] < hit_limit(), "%llu one-gap hits", host_ctr[0]);
        if (host_ctr[0] <= b.hit_cap) { b.hits1 = (int64_t)host_ctr[0]; break; }
        b.hit_cap = (size_t)host_ctr[0] + (size_t)host_ctr[0] / 8 + 1024;        // grow once to the exact need and redo the scan
    }
    // algorithmic bytes.  phrase-major: per element its position + gap word, every bucket word read, every hit
    // written; position-major: the window array streamed once, the phrase id per element, one 8-byte slot per table lookup, hits
    if (position_major) prof_add_bytes("join_onegap", 16.0 * (double)ix.n + 4.0 * (double)host_ctr[2] + 8.0 * (double)host_ctr[1] + 8.0 * (double)host_ctr[0]);
    else prof_add_bytes("join_onegap", 8.0 * (double)n_elems + 4.0 * (double)host_ctr[1] + 8.0 * (double)host_ctr[0]);
    b.j1_elems = (int64_t)n_elems;
    j1_missing_kernel<<<cgx_div_up(D1, 256), 256, 0, stream>>>(b.pat1.ptr<Pat1>(), missing, D1);
    b.launches++;
    if (b.hits1 == 0) return;
    const size_t H = (size_t)b.hits1;
    uint64_t *hits = b.hit_keys.ptr<uint64_t>();
    uint64_t *tmp = b.hit_keys_tmp.get<uint64_t>(H);
    if (ordered) {      // the tiles' segments, in tile (= position) order
        exclusive_scan_u32(ao.seg_count, seg_dst, n_segs, nullptr, stream, b.scan, 0, &b.launches);
        PROF("join_seg_copy", 16.0 * (double)H, (seg_copy_kernel<uint32_t><<<cgx_div_up(n_segs / ao.chunks_per_warp, 8), 256, 0, stream>>>(hits, ao.seg_base, seg_dst, ao.seg_count,
                                                                                                                                  (uint32_t)(n_segs / ao.chunks_per_warp), ao.chunks_per_warp, tmp)));
        b.launches++;
        std::swap(hits, tmp);
    }
    uint64_t *hs;
    // The sort is stable and hits of one (pattern, position) leave the scan in ascending length (lane = gap width), so the 4
    // length bits are never sorted on; after an ordered emission every pattern's hits are in (position, length) order already
    // and only the pattern bits are.
    radix_sort<uint64_t>(hits, tmp, nullptr, nullptr, H, ordered ? b.pbits + 4 : 4, b.pbits + 4 + cgx_bits_for((uint64_t)D1), stream, b.radix, &hs, nullptr, &b.launches);
    // the buffer the sort ended in becomes hits1_sorted (buffers rotate instead of a 16 B/hit copy)
    std::swap(b.hits1_sorted, hs == b.hit_keys.ptr<uint64_t>() ? b.hit_keys : b.hit_keys_tmp);
    uint64_t *dst = b.hits1_sorted.ptr<uint64_t>();
    static_assert(sizeof(Pat1) == 32, "Pat1 layout");
    hit_ranges_kernel<<<cgx_div_up(H, 256), 256, 0, stream>>>(dst, H, b.pbits + 4, (1u << cgx_bits_for((uint64_t)D1)) - 1u, &b.pat1.ptr<int32_t>()[4], 8);
    hit_ranges_fix_kernel<<<cgx_div_up(D1, 256), 256, 0, stream>>>(&b.pat1.ptr<int32_t>()[4], D1, 8);
    b.launches += 2;
}

// ------------------------------------------------------------------------------------------------
// two-gap: parent hits (p, L) of aXb extended by the single token c
//   c at p+L+1+g2, g2 >= 1, (L+1)+g2+1 <= 15  ->  g2 <= 13-L                      (GappyLook.cu:595-655)
// ------------------------------------------------------------------------------------------------
// child_sig[d1]: 64-bit signature of the tokens c for which aXbXc is a pattern of the batch (bit = 6-bit hash of c); 0 = the
// parent has no child.  A parent has ~5 children on average, so the signature rejects ~90 % of the candidate tokens before
// they cost a probe of the pattern table (round 1c: every admissible width of every parent hit probed the table).
__device__ __forceinline__ unsigned sig_bit(uint32_t c) { return (c * 0x9E3779B1u) >> 26; }

__global__ void j2_setup_kernel(const Pat2 *__restrict__ pat2, int D2, QTab tab, uint32_t *__restrict__ overflow, unsigned long long *__restrict__ child_sig) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D2) return;
    const Pat2 p = pat2[d];
    if (!qt_insert(tab, (uint32_t)p.pat1, (uint32_t)p.ctok, (uint32_t)d)) *overflow = 1u;      // key: (parent pattern, token c) -> two-gap pattern id
    atomicOr(&child_sig[p.pat1], 1ull << sig_bit((uint32_t)p.ctok));
}

__global__ void __launch_bounds__(256) j2_scan_kernel(const uint64_t *__restrict__ hits1, size_t H1, int pbits, const unsigned long long *__restrict__ child_sig,
                                                      const int32_t *__restrict__ str, const uint32_t *__restrict__ gapw,
                                                      const QTab tab, uint32_t dmask,
                                                      unsigned long long *__restrict__ counter, uint64_t *__restrict__ hits, size_t cap) {
    __shared__ uint64_t s_stage[256 / 32][ST_CAP];
    uint64_t *stage = s_stage[threadIdx.x >> 5];
    int staged = 0;
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned active = 0, probes = 0;
    uint32_t bits = 0, d1 = 0;
    unsigned long long sig = 0;
    int p = 0, L = 0;
    if (k < H1) {
        const uint64_t hk = hits1[k];
        d1 = (uint32_t)(hk >> (pbits + 4)) & dmask;
        sig = __ldg(&child_sig[d1]);
        if (sig) {
            p = (int)((hk >> 4) & ((1ull << pbits) - 1)); L = (int)(hk & 15);
            const uint32_t w = __ldg(&gapw[p + L + 1]);
            const int run = (int)((w >> 16) & 15u);
            const int gmax = min(run, CGX_MAX_RULE_SPAN - 2 - L);
            bits = gmax > 0 ? (w & ((1u << gmax) - 1u)) : 0u;
            active = 1;
        }
    }
    while (__any_sync(0xffffffffu, bits != 0)) {       // every round each lane tries its next (up to) four admissible widths:
        int rr[4];                                     // tokens, then first table probes, issued together
        uint32_t cc[4], ss[4];
        uint32_t st[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            rr[u] = 0;
            if (bits) { rr[u] = p + L + 1 + __ffs(bits); bits &= bits - 1; probes++; }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) if (rr[u]) cc[u] = (uint32_t)__ldg(&str[rr[u]]);
#pragma unroll
        for (int u = 0; u < 4; u++) if (rr[u] && !((sig >> sig_bit(cc[u])) & 1ull)) rr[u] = 0;      // not a child token of this parent
#pragma unroll
        for (int u = 0; u < 4; u++) if (rr[u]) qt_touch(tab, d1, cc[u], &ss[u], &st[u]);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (!__any_sync(0xffffffffu, rr[u] != 0)) continue;
            uint32_t d2 = 0;
            const bool found = rr[u] && qt_resolve(tab, ss[u], st[u], &d2);
            const uint64_t key = ((uint64_t)d2 << (pbits + 8)) | ((uint64_t)(uint32_t)p << 8) | ((uint64_t)(rr[u] - p - L - 1) << 4) | (uint64_t)L;   // (pattern, position, width, length)
            stage_push(found, key, stage, staged);
        }
        if (staged > ST_CAP - 128) stage_flush(stage, staged, &counter[0], hits, cap);
    }
    stage_flush(stage, staged, &counter[0], hits, cap);
    for (int o = 16; o; o >>= 1) { active += __shfl_xor_sync(0xffffffffu, active, o); probes += __shfl_xor_sync(0xffffffffu, probes, o); }
    if ((threadIdx.x & 31) == 0 && active) { atomicAdd(&counter[1], (unsigned long long)probes); atomicAdd(&counter[2], (unsigned long long)active); }
}


// Ordered form of j2_scan_kernel: a thread owns one parent hit of the (pattern, position, length)-sorted one-gap list; its (up
// to 13) two-gap hits are parked in shared memory by ascending width, the tile takes one segment of the output with a single
// atomicAdd (as above), and the hits are written in (parent pattern, position, WIDTH, LENGTH) order: the parents of one
// (pattern, position) -- a "run" of at most 13 adjacent threads, ascending length -- interleave their hits width-major.  That
// is the order twoGapLookUpSA's lock-step warp hands out places in (lanes = adjacent parents, loop variable = width), which
// thrust's stable sort on (pattern, position) keeps (oracle/cgx_oracle.c cmp_hit4).  After the segments are copied into tile
// order a stable sort on the two-gap pattern bits alone (a pattern aXbXc has ONE parent) yields (pattern, position, width,
// length): three passes instead of seven (eight at C3).
// Tiles are cut on run boundaries: a tile is nominally J2O_STRIDE parents; its first parents are skipped when they continue
// the previous tile's last run, and threads J2O_STRIDE..255 take the parents that continue its own last run.
constexpr int J2O_SLOTS = CGX_MAX_RULE_SPAN - 2;       // widths 1..13
constexpr int J2O_BLOCK = 256;                         // threads per tile (128, to shorten the waits at the tile's two barriers, measured slower: 7.8 -> 9.0 ms at C2)
constexpr int J2O_STRIDE = J2O_BLOCK - 16;             // >= J2O_BLOCK - (longest run - 1)
__global__ void __launch_bounds__(J2O_BLOCK) j2_ordered_kernel(const uint64_t *__restrict__ hits1, size_t H1, int pbits, const unsigned long long *__restrict__ child_sig,
                                                         const int32_t *__restrict__ str, const uint32_t *__restrict__ gapw, const QTab tab, uint32_t dmask, int carried,
                                                         unsigned long long *__restrict__ counter, unsigned long long *__restrict__ seg_base,
                                                         uint32_t *__restrict__ seg_count, uint64_t *__restrict__ hits, size_t cap) {
    __shared__ uint32_t s_d2[J2O_SLOTS][J2O_BLOCK];           // hits parked by [width - 1][owner thread]
    __shared__ unsigned s_mask[J2O_BLOCK];
    __shared__ uint16_t s_excl[J2O_BLOCK];
    __shared__ uint32_t s_qtok[J2O_BLOCK / 32][64];                  // per-warp candidate queue: token / owner lane << 4 | width - 1
    __shared__ uint16_t s_qown[J2O_BLOCK / 32][64];
    __shared__ uint8_t s_head[J2O_BLOCK + 1];
    __shared__ unsigned s_wsum[J2O_BLOCK / 32];
    __shared__ unsigned long long s_base;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const size_t base = (size_t)tile * J2O_STRIDE;
    const size_t k = base + tid;
    unsigned active = 0, probes = 0;
    uint32_t bits = 0, d1 = 0;
    unsigned long long sig = 0;
    int p = 0, L = 0;
    bool mine = false, head = true;
    if (k < H1) {
        const uint64_t kmask = carried ? (1ull << H1_MASK_SHIFT) - 1ull : ~0ull;          // the key without the carried second-gap word
        const uint64_t hk = hits1[k];
        const uint64_t run_key = (hk & kmask) >> 4;                            // (pattern, position)
        const uint64_t before = base ? (hits1[base - 1] & kmask) >> 4 : ~0ull;  // last parent of the previous tile's nominal range
        mine = run_key != before;                                              // not a continuation of the previous tile's last run
        if (tid >= J2O_STRIDE) mine = mine && run_key == ((hits1[base + J2O_STRIDE - 1] & kmask) >> 4);
        head = !mine || tid == 0 || run_key != ((hits1[k - 1] & kmask) >> 4);
        if (mine) {
            d1 = (uint32_t)(hk >> (pbits + 4)) & dmask;
            sig = __ldg(&child_sig[d1]);
        }
        if (sig) {
            p = (int)((hk >> 4) & ((1ull << pbits) - 1)); L = (int)(hk & 15);
            if (carried && L > 2) bits = (uint32_t)(hk >> H1_MASK_SHIFT);      // admissible widths 1..10: all there are for L >= 3
            else {                                                             // (L = 2 may admit an 11th width: read the word itself)
                const uint32_t w = __ldg(&gapw[p + L + 1]);
                const int run = (int)((w >> 16) & 15u);
                const int gmax = min(run, CGX_MAX_RULE_SPAN - 2 - L);
                bits = gmax > 0 ? (w & ((1u << gmax) - 1u)) : 0u;
            }
            active = 1;
        }
    }
    // Candidate tokens: every lane walks its admissible widths (two per round), reads the token there and tests it against the
    // parent's child signature.  The ~0.7 candidates per parent that pass are queued per warp in shared memory (owner lane, width,
    // token) and the pattern table is probed for 32 of them at a time -- in place the probes ran with 3 of 32 lanes active and a
    // third of the kernel's stall samples sat on their loads (profiles/r02a).  A hit is parked under its owner and WIDTH
    // (s_d2[width][owner], s_mask[owner] bit), so the order in which probes complete does not matter.
    uint32_t *q_tok = s_qtok[warp];
    uint16_t *q_own = s_qown[warp];
    s_mask[tid] = 0;
    __syncwarp();
    unsigned queued = 0;                                // warp-uniform
    uint32_t pend_home = 0, pend_tag = 0, pend_own = 0;
    bool pend = false;
    auto resolve_pending = [&]() {                      // the probes requested by the previous drain: their sectors have had the walk in between to arrive
        uint32_t d2 = 0;
        if (pend && qt_resolve(tab, pend_home, pend_tag, &d2)) {
            const unsigned ot = (tid & ~31u) + (pend_own >> 4);
            s_d2[pend_own & 15u][ot] = d2;
            atomicOr(&s_mask[ot], 1u << (pend_own & 15u));
        }
        pend = false;
    };
    auto drain = [&](unsigned n_take) {                 // the first n_take (<= 32) queue entries leave the queue: request their buckets, shift the rest down
        resolve_pending();
        const bool on = lane < n_take;
        const unsigned own = on ? q_own[lane] : 0u;     // owner lane << 4 | width - 1
        const uint32_t tokc = on ? q_tok[lane] : 0u;
        const uint32_t d1o = __shfl_sync(0xffffffffu, d1, own >> 4);
        if (on) { qt_touch(tab, d1o, tokc, &pend_home, &pend_tag); pend_own = own; pend = true; }
        __syncwarp();
        const unsigned rest = queued - n_take;
        uint32_t t2 = 0; uint16_t o2 = 0;
        if (lane < rest) { t2 = q_tok[n_take + lane]; o2 = q_own[n_take + lane]; }
        __syncwarp();
        if (lane < rest) { q_tok[lane] = t2; q_own[lane] = o2; }
        __syncwarp();
        queued = rest;
    };
    while (__any_sync(0xffffffffu, bits != 0)) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            int g = -1;
            if (bits) { g = __ffs(bits) - 1; bits &= bits - 1; probes++; }
            uint32_t tokc = 0;
            if (g >= 0) tokc = (uint32_t)__ldg(&str[p + L + 2 + g]);
            const bool cand = g >= 0 && ((sig >> sig_bit(tokc)) & 1ull);
            const unsigned m = __ballot_sync(0xffffffffu, cand);
            if (cand) {
                const unsigned at = queued + __popc(m & lanemask_lt());
                q_tok[at] = tokc;
                q_own[at] = (uint16_t)((lane << 4) | (unsigned)g);
            }
            queued += __popc(m);
            __syncwarp();
            if (queued >= 32) drain(32);
        }
    }
    if (queued) drain(queued);
    resolve_pending();
    __syncwarp();
    unsigned found_mask = s_mask[tid];
    const unsigned c = __popc(found_mask);              // widths that hit (bit g2-1), their count
    s_head[tid] = head ? 1 : 0;
    if (tid == 0) s_head[J2O_BLOCK] = 1;
    // exclusive prefix of c over the CTA
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    unsigned excl = incl - c;
#pragma unroll
    for (int w = 0; w < J2O_BLOCK / 32; w++) if (w < (int)warp) excl += s_wsum[w];
    s_excl[tid] = (uint16_t)excl;
    if (tid == 0) {
        unsigned tot = 0;
#pragma unroll
        for (int w = 0; w < J2O_BLOCK / 32; w++) tot += s_wsum[w];
        const unsigned long long seg = tot ? atomicAdd(&counter[0], (unsigned long long)tot) : 0ull;
        seg_base[tile] = seg;
        seg_count[tile] = tot;
        s_base = seg;
    }
    __syncthreads();
    if (c) {
        const bool alone = head && s_head[tid + 1];                   // the run is this thread only: its hits are consecutive
        int rs = (int)tid, re = (int)tid;                             // run = threads rs..re
        if (!alone) {
            while (!s_head[rs]) rs--;
            while (!s_head[re + 1]) re++;
        }
        const unsigned long long o = s_base + s_excl[rs];
        for (unsigned j = 0; j < c; j++) {
            const int g = __ffs(found_mask) - 1;                      // j-th smallest width of this thread, 0-based
            found_mask &= found_mask - 1;
            unsigned at = j;
            if (!alone) {
                at = 0;
                const unsigned below = (1u << g) - 1u;
                for (int m = rs; m <= re; m++) {
                    const unsigned mm = s_mask[m];
                    at += __popc(mm & below) + ((m < (int)tid) ? ((mm >> g) & 1u) : 0u);
                }
            }
            if (o + at < cap)
                hits[o + at] = ((uint64_t)s_d2[g][tid] << (pbits + 8)) | ((uint64_t)(uint32_t)p << 8) | ((uint64_t)(g + 1) << 4) | (uint64_t)L;
        }
    }
    for (int off = 16; off; off >>= 1) { active += __shfl_xor_sync(0xffffffffu, active, off); probes += __shfl_xor_sync(0xffffffffu, probes, off); }
    if (lane == 0 && active) { atomicAdd(&counter[1], (unsigned long long)probes); atomicAdd(&counter[2], (unsigned long long)active); }
}

void stage_twogap_join(const Index &ix, Batch &b, cudaStream_t stream) {
    b.hits2 = 0;
    const int D2 = b.D2;
    if (D2 == 0 || b.hits1 == 0) return;
    CGX_REQUIRE(cgx_bits_for((uint64_t)D2) + b.pbits + 8 <= 64, "two-gap join: %d distinct patterns exceed the hit-key field", D2);
    // per-batch table of the two-gap patterns (hash.cuh QTab): key = (parent one-gap pattern, token c), value = pattern id.  Buckets
    // of four 8-byte slots: a probe is one sector, and a miss -- one in ten tokens passes the child signature without being a
    // child -- ends at the first bucket that is not full (the packed linear-probing table of round 1 walked ~6 slots for a miss
    // at its load factor: a fifth of the kernel's stall samples, r02b)
    QTab tab;
    tab.abits = std::max(cgx_bits_for((uint64_t)b.D1), 10); tab.bbits = cgx_bits_for((uint64_t)ix.maxtok);      // the wider half first (hash.cuh)
    CGX_REQUIRE_BATCH(tab.bbits <= 24 && tab.bbits >= 1 && D2 < (1 << 30), "%d x %d patterns exceed the two-gap pattern-table fields", b.D1, D2);
    uint32_t buckets = qt_buckets_for((size_t)D2, tab.abits + tab.bbits);
    if (b.j2_buckets > buckets && b.j2_buckets <= 4 * buckets) buckets = b.j2_buckets;
    unsigned long long *child_sig = b.j_aflag.get<unsigned long long>((size_t)b.D1 + 1);
    uint32_t *tot = b.counters.get<uint32_t>(32);
    unsigned long long *ctr = (unsigned long long *)(tot + 4);
    while (true) {
        tab.rb = tab.abits + tab.bbits - qt_log2(buckets); tab.bmask = buckets - 1;
        CGX_REQUIRE(tab.rb >= 1 && tab.rb <= 27, "two-gap pattern table: %d remainder bits", tab.rb);
        tab.slots = b.j_hash.get<unsigned long long>((size_t)buckets * 4);
        CUDA_CHECK(cudaMemsetAsync(tab.slots, 0xff, sizeof(unsigned long long) * (size_t)buckets * 4, stream));
        CUDA_CHECK(cudaMemsetAsync(child_sig, 0, sizeof(unsigned long long) * (size_t)b.D1, stream));
        CUDA_CHECK(cudaMemsetAsync(tot + 1, 0, sizeof(uint32_t), stream));
        PROF("join_setup", (double)D2 * (16 + 8), (j2_setup_kernel<<<cgx_div_up(D2, 256), 256, 0, stream>>>(b.pat2.ptr<Pat2>(), D2, tab, tot + 1, child_sig)));
        b.launches++;
        uint32_t over = 0;
        cgx_read_back(&over, tot + 1, sizeof(over), stream);
        if (!over) break;
        buckets *= 2;
        CGX_REQUIRE(buckets <= (1u << 28), "two-gap pattern table does not settle");
    }
    b.j2_buckets = buckets;
    const size_t H1 = (size_t)b.hits1;
    unsigned long long host_ctr[3] = {0, 0, 0};
    bool ordered = true;                                   // CGX_JOIN_ORDERED=0: round 1's unordered append + full sort
    if (const char *e = getenv("CGX_JOIN_ORDERED")) { if (!strcmp(e, "0")) ordered = false; }
    const uint32_t n_tiles = cgx_div_up(H1, J2O_STRIDE);
    unsigned long long *seg_base = ordered ? b.j_status.get<unsigned long long>((size_t)n_tiles + 1) : nullptr;
    uint32_t *seg_count = ordered ? b.j_segcnt.get<uint32_t>((size_t)2 * n_tiles + 2) : nullptr;
    uint32_t *seg_dst = ordered ? seg_count + n_tiles + 1 : nullptr;
    while (true) {
        uint64_t *hits = b.hit_keys.get<uint64_t>(b.hit_cap);
        CUDA_CHECK(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long) * 3, stream));
        if (ordered) {
            PROF("join_twogap", 0.0, (j2_ordered_kernel<<<n_tiles, J2O_BLOCK, 0, stream>>>(b.hits1_sorted.ptr<uint64_t>(), H1, b.pbits, child_sig, ix.str.ptr<int32_t>(),
                                                               ix.gapw.ptr<uint32_t>(), tab, (1u << cgx_bits_for((uint64_t)b.D1)) - 1u, b.h1_mask ? 1 : 0, ctr, seg_base, seg_count, hits, b.hit_cap)));
        } else
        PROF("join_twogap", 0.0, (j2_scan_kernel<<<cgx_div_up(H1, 256), 256, 0, stream>>>(b.hits1_sorted.ptr<uint64_t>(), H1, b.pbits, child_sig, ix.str.ptr<int32_t>(),
                                                           ix.gapw.ptr<uint32_t>(), tab, (1u << cgx_bits_for((uint64_t)b.D1)) - 1u, ctr, hits, b.hit_cap)));
        b.launches += 1;
        read_u64s(host_ctr, ctr, 3, stream);
        CGX_REQUIRE_BATCH(host_ctr[0] < hit_limit(), "%llu two-gap hits", host_ctr[0]);
        if (host_ctr[0] <= b.hit_cap) { b.hits2 = (int64_t)host_ctr[0]; break; }
        b.hit_cap = (size_t)host_ctr[0] + (size_t)host_ctr[0] / 8 + 1024;
    }
    // algorithmic bytes: every parent hit read, gap word of the active ones (unless the hit keys carry it), each candidate token, each hit written
    prof_add_bytes("join_twogap", 8.0 * (double)H1 + (b.h1_mask ? 0.0 : 4.0 * (double)host_ctr[2]) + 4.0 * (double)host_ctr[1] + 8.0 * (double)host_ctr[0]);
    if (b.hits2 == 0) return;
    const size_t H = (size_t)b.hits2;
    uint64_t *hits = b.hit_keys.ptr<uint64_t>();
    uint64_t *tmp = b.hit_keys_tmp.get<uint64_t>(H);
    if (ordered) {
        exclusive_scan_u32(seg_count, seg_dst, (size_t)n_tiles, nullptr, stream, b.scan, 0, &b.launches);
        PROF("join_seg_copy", 16.0 * (double)H, (seg_copy_kernel<unsigned long long><<<cgx_div_up(n_tiles, 8), 256, 0, stream>>>(hits, seg_base, seg_dst, seg_count, n_tiles, 1u, tmp)));
        b.launches++;
        std::swap(hits, tmp);
    }
    uint64_t *hs;
    radix_sort<uint64_t>(hits, tmp, nullptr, nullptr, H, ordered ? b.pbits + 8 : 0, b.pbits + 8 + cgx_bits_for((uint64_t)D2), stream, b.radix, &hs, nullptr, &b.launches);
    std::swap(b.hits2_sorted, hs == b.hit_keys.ptr<uint64_t>() ? b.hit_keys : b.hit_keys_tmp);
    uint64_t *dst = b.hits2_sorted.ptr<uint64_t>();
    static_assert(sizeof(Pat2) == 16, "Pat2 layout");
    hit_ranges_kernel<<<cgx_div_up(H, 256), 256, 0, stream>>>(dst, H, b.pbits + 8, 0xffffffffu, &b.pat2.ptr<int32_t>()[2], 4);
    hit_ranges_fix_kernel<<<cgx_div_up(D2, 256), 256, 0, stream>>>(&b.pat2.ptr<int32_t>()[2], D2, 4);
    b.launches += 2;
}

}  // namespace cgx
