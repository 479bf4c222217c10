// cgx-b200: on-device rule aggregation and lexical scoring.
//
// Replaces the single-threaded host loops createLexiconFast / createLexiconGappyFast /
// createLexiconTwoGapFast (ExtractPair.c:515-1276: sprintf/strcat a "src ||| tgt" string per extracted
// record and de-duplicate it in uthash / std::map), extractGlobalPairsUpDown (ExtractPair.cu:2082-2106)
// and lexicalTaskMaxEF + searchLexFile (:2108-2432).
//
// Rule identity = (converted source id, target symbol sequence with each gap collapsed to one marker)
// (ExtractPair.c:813-848, :1141-1173).  The extraction kernels leave their records in slot-indexed cells that are
// already grouped by source id (extract.cu): all records of one id sit in the <= 300 (65 / 70 for gappy seeds)
// consecutive cells of its pattern.  So no global sort is needed (round 1a-1c: a 64-bit hash sort + an id sort of
// 1.2e8 records = 30 GB of radix passes per batch): every record gets a 64-bit hash of its target symbol sequence
// (agg_hash_kernel), and one thread per cell scans the 4-byte hash tags of the cells BEFORE it in its segment (broadcast
// loads, the segment is shared by the neighbouring threads) for the first equal one = the head of its rule.  Every record
// then reports to its head with two commutative atomics -- paircount += 1, representative = min(target start, cell) -- and
// heads add f = records of the id (agg_group_kernel).  Heads are flagged, prefix-summed and compacted into rules in
// (id, first cell) order -- deterministic, as are count and representative.  Exactness does not rest on the hash: every
// non-head record is compared symbol-by-symbol with its head, and a mismatch raises a flag on which the host re-runs the
// aggregation with another hash seed.  Rules leave the device packed (16 bytes + one word per id, include/cgx_b200.h).
// Lexical weights: one thread per distinct rule; every (f,e) pair is ONE probe of the lexical hash table (16-byte
// slots holding both directions' values; 2^21 slots = 32 MB at C2, L2-resident) instead of a 20-step binary search;
// -log10 through the same lg2.approx path the reference's -use_fast_math build takes.
#include "batch.h"
#include "hash.cuh"
#include "prof.h"

namespace cgx {

struct AggIdx {
    const int32_t *str, *tgt;
    const int32_t *phrases;
    const Pat1 *pat1;
    const Pat2 *pat2;
    int G, D1, D2;
};

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 33; z *= 0xff51afd7ed558ccdULL; z ^= z >> 33; z *= 0xc4ceb9fe1a85ec53ULL; z ^= z >> 33;
    return z;
}

// target symbol sequence of a record: its target tokens outside the gaps, gap1 -> 0xFFFFFFFF, gap2 -> 0xFFFFFFFE (one symbol
// per gap); walked by agg_hash_kernel (hash) and next_symbol() (verification)
// cells of one kind: up to 4 slot-indexed regions in ascending id order (extract.cu)
struct AggRegion {
    uint32_t base;                 // first cell
    int32_t id_base;               // converted id of pattern 0 of the region
    const uint32_t *slot_off;      // pattern -> first slot (n_patterns + 1 entries)
};
struct AggLayout {
    AggRegion r[4];
    int n_regions;
    uint32_t cells;
};

// hash[i] = 0 for an empty cell, else the (odd) hash of the record's target symbol sequence; tag[i] = its low word (the
// 4-byte filter the grouping loop scans); live[i] = 1 for a record (its exclusive scan gives the records per segment = f)
// Also clears the per-cell accumulators of agg_group (two separate memsets of 12 B per cell before).
__global__ void agg_hash_kernel(const RuleRec *__restrict__ rec, uint32_t cells, const int32_t *__restrict__ tgt, uint64_t seed, uint64_t *__restrict__ hash,
                                uint32_t *__restrict__ tag, uint32_t *__restrict__ live, unsigned long long *__restrict__ acc_best,
                                uint32_t *__restrict__ acc_cnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cells) {
        const RuleRec r = rec[i];
        uint64_t h = 0;
        const bool is_live = r.id >= 0;
        if (is_live) {
            uint32_t tok[15];                                     // the span is <= 15 tokens: issue every load before using any
#pragma unroll
            for (int j = 0; j < 15; j++) tok[j] = j <= (int)r.end ? (uint32_t)__ldg(&tgt[r.tgt_start + j]) : 0u;
            int ns = 0, skip_to = -1;                             // the target symbol sequence (see above)
            // two independent 32-bit multiplicative lanes per symbol (4 instructions; a 64-bit murmur round per symbol was
            // ~20 and made this kernel issue-bound), one 64-bit finaliser per record.  Equal hashes are verified symbol by
            // symbol in agg_group, so the hash only has to make collisions inside a <= 300-cell segment rare.
            uint32_t h1 = (uint32_t)seed, h2 = (uint32_t)(seed >> 32);
#pragma unroll
            for (int j = 0; j < 15; j++) {
                if (j > (int)r.end) break;
                if (j <= skip_to) continue;
                uint32_t sym;
                if (r.gap1 != 255 && j >= (int)r.gap1 && j <= (int)r.gap1_1) { sym = 0xFFFFFFFFu; skip_to = r.gap1_1; }
                else if (r.gap2 != 255 && j >= (int)r.gap2 && j <= (int)r.gap2_1) { sym = 0xFFFFFFFEu; skip_to = r.gap2_1; }
                else sym = tok[j];
                h1 = (h1 ^ sym) * 0x9E3779B1u;
                h2 = (h2 + sym) * 0x85EBCA77u + 0x165667B1u;
                ns++;
            }
            h = mix64((((uint64_t)h1 << 32) | (uint64_t)h2) ^ (uint64_t)ns) | 1ull;
        }
        hash[i] = h;
        tag[i] = (uint32_t)h;
        live[i] = is_live ? 1u : 0u;
        acc_best[i] = ~0ull;
        acc_cnt[i] = 0u;
    }
}

// target side of a record, unpacked to scalars (byte fields of a by-reference RuleRec end up in local memory)
struct TgtSpan {
    int ts, end, g1, g1e, g2, g2e;       // g1 / g2 = -1 when the gap is absent
    __device__ __forceinline__ explicit TgtSpan(const RuleRec &r)
        : ts(r.tgt_start), end(r.end), g1(r.gap1 == 255 ? -1 : (int)r.gap1), g1e(r.gap1_1), g2(r.gap2 == 255 ? -1 : (int)r.gap2), g2e(r.gap2_1) {}
};
// next target symbol from span offset j on (tokens outside the gaps, gap1 -> 0xFFFFFFFF, gap2 -> 0xFFFFFFFE); false past the end.
// The same symbol sequence agg_hash_kernel hashes.
__device__ __forceinline__ bool next_symbol(const int32_t *__restrict__ tgt, const TgtSpan &r, int &j, uint32_t &sym) {
    if (j > r.end) return false;
    if (r.g1 >= 0 && j >= r.g1 && j <= r.g1e) { sym = 0xFFFFFFFFu; j = r.g1e + 1; }
    else if (r.g2 >= 0 && j >= r.g2 && j <= r.g2e) { sym = 0xFFFFFFFEu; j = r.g2e + 1; }
    else { sym = (uint32_t)__ldg(&tgt[r.ts + j]); j++; }
    return true;
}
__device__ __forceinline__ bool same_target(const int32_t *__restrict__ tgt, const RuleRec &rx, const RuleRec &ry) {
    const TgtSpan x(rx), y(ry);
    int jx = 0, jy = 0;
    while (true) {
        uint32_t sx = 0, sy = 0;
        const bool hx = next_symbol(tgt, x, jx, sx), hy = next_symbol(tgt, y, jy, sy);
        if (hx != hy) return false;
        if (!hx) return true;
        if (sx != sy) return false;
    }
}

// One thread per cell: group the records of its segment (= the cells of its source id) by target sequence.
//   flags[i] = 1 when cell i is the first cell of its rule (no equal cell before it in the segment).
//   Every record then reports to its rule head h (itself for a head) with two commutative atomics:
//     acc_cnt[h]  += 1 per non-head member  (+ f << 16 by the head: f = records of the id = non-empty cells of the segment)
//     acc_best[h]  = min(tgt_start << 32 | cell)   -> the representative = member with the smallest target start
//   so heads need no forward scan of the segment for their paircount (round 1b: heads, 60 % of the records, scanned the
//   whole segment twice; 8 of 32 lanes active on average).  The backward scan reads the 4-byte tags (shared with the
//   neighbouring threads: L1 broadcast) four at a time and stops at the first equal cell.
// (Three quarters of the cells are empty -- a slot emits a record of a shape or not -- and round 1 ran this kernel over all of them
// with 8 of 32 lanes alive.  Since round 2 the live cells are listed first (agg_live_list_kernel: the exclusive scan of the live
// flags is needed for the f counts anyway) and thread j takes the j-th live cell: full warps, neighbouring lanes in one segment.)
__global__ void __launch_bounds__(256) agg_live_list_kernel(const uint64_t *__restrict__ hash, const uint32_t *__restrict__ live_before, uint32_t cells,
                                                            uint32_t *__restrict__ live_list, uint32_t *__restrict__ flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    if (hash[i] != 0) live_list[live_before[i]] = i;
    else flags[i] = 0;                                       // an empty cell heads no rule
}

__global__ void __launch_bounds__(256) agg_group_kernel(AggLayout lay, const RuleRec *__restrict__ rec, const uint64_t *__restrict__ hash,
                                                        const uint32_t *__restrict__ tag, const uint32_t *__restrict__ live_before,
                                                        const uint32_t *__restrict__ live_list,
                                                        const int32_t *__restrict__ tgt, uint32_t *__restrict__ flags,
                                                        unsigned long long *__restrict__ acc_best, uint32_t *__restrict__ acc_cnt,
                                                        int *__restrict__ collision) {
    const uint32_t j0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (j0 >= live_before[lay.cells]) return;                // live_before[cells] = number of live cells
    const uint32_t i = live_list[j0];
    const uint64_t h = hash[i];
    const uint32_t t = (uint32_t)h;
    const RuleRec r = rec[i];
    int reg = 0;
#pragma unroll
    for (int k = 1; k < 4; k++) if (k < lay.n_regions && i >= lay.r[k].base) reg = k;
    const uint32_t pat = (uint32_t)(r.id - lay.r[reg].id_base);
    const uint32_t s0 = lay.r[reg].base + __ldg(&lay.r[reg].slot_off[pat]);
    // the tags are scanned four at a time (aligned 16-byte loads).  Only the first and the last group of four can hold cells
    // outside [s0, i): they are checked with position masks, the groups in between with four bare compares.
    uint32_t first = i;
    auto check4 = [&](uint32_t base, bool masked) {
        const uint4 t4 = __ldg(reinterpret_cast<const uint4 *>(tag + base));
        if (t4.x != t && t4.y != t && t4.z != t && t4.w != t) return;
        const uint32_t tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t j = base + k;
            if (tt[k] == t && (!masked || (j >= s0 && j < i)) && first == i && __ldg(&hash[j]) == h) first = j;
        }
    };
    {
        uint32_t base = s0 & ~3u;
        const uint32_t last = (i - 1) & ~3u;            // group of cell i - 1 (i > s0 below)
        if (i > s0) {
            check4(base, true);
            for (base += 4; base < last && first == i; base += 4) check4(base, false);
            if (first == i && last > (s0 & ~3u)) check4(last, true);
        }
    }
    const unsigned long long key = ((unsigned long long)(uint32_t)r.tgt_start << 32) | (unsigned long long)i;
    atomicMin(&acc_best[first], key);
    if (first == i) {
        const uint32_t s1 = lay.r[reg].base + __ldg(&lay.r[reg].slot_off[pat + 1]);
        const uint32_t f = __ldg(&live_before[s1]) - __ldg(&live_before[s0]);      // records of the id = non-empty cells of the segment
        flags[i] = 1;
        atomicAdd(&acc_cnt[i], f << 16);
    } else {
        flags[i] = 0;
        atomicAdd(&acc_cnt[first], 1u);
        const RuleRec q = rec[first];
        if (!same_target(tgt, r, q)) atomicExch(collision, 1);
    }
}

// all_suffix_fsample before the cap (ExtractPair.c:637, :891-908, :1211-1247)
__device__ __forceinline__ int fsample_of(const AggIdx &a, int kind, int id) {
    int blk = -1, p1 = -1, p2 = -1;
    if (kind == 0) blk = id;
    else if (kind == 1) { if (id < a.G) blk = id; else if (id < 2 * a.G) blk = id - a.G; else p1 = id - 2 * a.G; }
    else { if (id < a.G) blk = id; else if (id < a.G + a.D2) p2 = id - a.G; else if (id < a.G + a.D2 + a.D1) p1 = id - a.G - a.D2; else p1 = id - a.G - a.D2 - a.D1; }
    if (blk >= 0) return a.phrases[blk * 4 + 1] - a.phrases[blk * 4] + 1;
    if (p2 >= 0) return a.pat2[p2].hit_count;
    Pat1 p = a.pat1[p1];
    return p.hit_count + (p.marker_pair >= 0 ? p.fs_extra : 0);
}

// source terminals of a converted id (the F set of lexicalTaskMaxEF)
__device__ __forceinline__ int source_terminals(const AggIdx &a, int kind, int id, int32_t out[8]) {
    int blk = -1, p1 = -1, p2 = -1, n = 0;
    if (kind == 0) blk = id;
    else if (kind == 1) { if (id < a.G) blk = id; else if (id < 2 * a.G) blk = id - a.G; else p1 = id - 2 * a.G; }
    else { if (id < a.G) blk = id; else if (id < a.G + a.D2) p2 = id - a.G; else if (id < a.G + a.D2 + a.D1) p1 = id - a.G - a.D2; else p1 = id - a.G - a.D2 - a.D1; }
    if (blk >= 0) { int s = a.phrases[blk * 4 + 3], l = a.phrases[blk * 4 + 2]; for (int i = 0; i < l; i++) out[n++] = __ldg(&a.str[s + i]); }
    if (p2 >= 0) p1 = a.pat2[p2].pat1;
    if (p1 >= 0) {
        Pat1 p = a.pat1[p1];
        for (int i = 0; i < p.ls; i++) out[n++] = __ldg(&a.str[p.a_pos + i]);
        for (int i = 0; i < p.le; i++) out[n++] = __ldg(&a.str[p.b_pos + i]);
    }
    if (p2 >= 0) out[n++] = a.pat2[p2].ctok;
    return n;
}

// ExtractPair.cu:2108-2142 searchLexFile: both values of (f,e) in one probe of the lexical hash table; absent -> 0
__device__ __forceinline__ void lex_get(const ulonglong2 *__restrict__ slots, uint32_t mask, int f, int e, float *v1, float *v2) {
    const uint64_t k = ((uint64_t)(uint32_t)(f + 1) << 32) | (uint64_t)(uint32_t)(e + 1);
    uint64_t pay;
    if (ht_find(slots, mask, k, &pay)) { *v1 = __uint_as_float((uint32_t)pay); *v2 = __uint_as_float((uint32_t)(pay >> 32)); }
    else { *v1 = 0.f; *v2 = 0.f; }
}

// asks for the sector lex_get(f, e) will read first (prefetch: no destination register), so that the probes of one target
// terminal are in flight together; issued back to back with their loads they waited for one another (50 % of the kernel's
// stall samples, profiles/r02a), and holding their first slots in registers cost more than it hid
__device__ __forceinline__ void lex_touch(const ulonglong2 *__restrict__ slots, uint32_t mask, int f, int e) {
    const uint64_t k = ((uint64_t)(uint32_t)(f + 1) << 32) | (uint64_t)(uint32_t)(e + 1);
    asm volatile("prefetch.global.L1 [%0];" ::"l"(slots + (ht_mix(k) & mask)));
}

// head_cell[r] = cell of the first record of rule r (excl = exclusive scan of the head flags)
__global__ void agg_head_cell_kernel(const uint32_t *__restrict__ excl, uint32_t cells, uint32_t n_rules, uint32_t *__restrict__ head_cell) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    const uint32_t e = excl[i];
    const uint32_t nxt = (i + 1 < cells) ? excl[i + 1] : n_rules;
    if (nxt != e) head_cell[e] = i;
}

// One thread per distinct rule: paircount, f, fs, representative record, lexical weights.  (One thread per CELL with the
// heads writing their rule keeps the reads streaming but leaves 73 % of the lanes idle in the probe-heavy part: 2x slower.)
__global__ void __launch_bounds__(128) agg_rules_kernel(AggIdx a, int kind, const RuleRec *__restrict__ rec, const uint32_t *__restrict__ head_cell,
                                                        const unsigned long long *__restrict__ acc_best, const uint32_t *__restrict__ acc_cnt, uint32_t n_rules,
                                                        const ulonglong2 *__restrict__ lex, uint32_t lex_mask, cgx_rule_t *__restrict__ rules,
                                                        int32_t *__restrict__ rule_id, uint32_t *__restrict__ idinfo) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rules) return;
    const uint32_t hc = head_cell[r];
    const uint32_t cw = acc_cnt[hc];                                     // members - 1 | f << 16
    const RuleRec best = rec[(uint32_t)acc_best[hc]];                    // low word of min(tgt_start << 32 | cell)
    const int pc = (int)(cw & 0xffffu) + 1;
    cgx_rule_t out;
    out.tgt_start = best.tgt_start;
    const uint32_t g1 = best.gap1 == 255 ? 15u : (uint32_t)best.gap1, g1e = best.gap1 == 255 ? 15u : (uint32_t)best.gap1_1;
    const uint32_t g2 = best.gap2 == 255 ? 15u : (uint32_t)best.gap2, g2e = best.gap2 == 255 ? 15u : (uint32_t)best.gap2_1;
    out.span = (uint32_t)best.end | (g1 << 4) | (g1e << 8) | (g2 << 12) | (g2e << 16) | ((uint32_t)pc << 20);   // include/cgx_b200.h
    int fs = fsample_of(a, kind, best.id);
    fs = fs > CGX_SAMPLER ? CGX_SAMPLER : fs;                                           // ExtractPair.c:638,910,1249
    rule_id[r] = best.id;
    idinfo[best.id] = (cw >> 16) | ((uint32_t)fs << 9);        // f | fs << 9 (both <= 300): every rule of the id writes the same word; the rule count is added later
    // ---- lexicalTaskMaxEF (ExtractPair.cu:2144-2432): for every source terminal the best MaxLexFgivenE over the target
    // terminals (and NULL), for every target terminal the best MaxLexEgivenF over the source terminals (and NULL).  One
    // table probe serves both directions of a (f, e) pair.
    int32_t F[8];
    const int nf = source_terminals(a, kind, best.id, F);
    float mxf[8];
#pragma unroll
    for (int j = 0; j < 8; j++) mxf[j] = 0.f;
    float egivenf = 0.f, v1, v2;
    const int ts = best.tgt_start;
    // target terminals as a bit mask of span offsets (gaps cleared), walked with ffs in ascending order: the loop runs
    // "number of terminals" times instead of "span length" times with the gap positions idle (6 of 32 lanes were active)
    uint32_t tmask = (2u << best.end) - 1u;
    if (best.gap1 != 255) tmask &= ~(((2u << best.gap1_1) - 1u) ^ ((1u << best.gap1) - 1u));
    if (best.gap2 != 255) tmask &= ~(((2u << best.gap2_1) - 1u) ^ ((1u << best.gap2) - 1u));
    const bool any_e = tmask != 0;
    while (tmask) {
        const int jj = __ffs(tmask) - 1;
        tmask &= tmask - 1;
        const int e = __ldg(&a.tgt[ts + jj]);
        float mx = 0.f;
        if (nf > 0) lex_touch(lex, lex_mask, -1, e);
#pragma unroll
        for (int j = 0; j < 5; j++) if (j < nf) lex_touch(lex, lex_mask, F[j], e);
        if (nf > 0) { lex_get(lex, lex_mask, -1, e, &v1, &v2); mx = fmaxf(mx, v1); }
#pragma unroll
        for (int j = 0; j < 5; j++) {                                       // nf <= 5 (CGX_LONGEST_SRC / MAX_rule_symbols)
            if (j >= nf) break;
            lex_get(lex, lex_mask, F[j], e, &v1, &v2);
            mx = fmaxf(mx, v1);
            mxf[j] = fmaxf(mxf[j], v2);
        }
        egivenf += mx > 0.f ? -__log10f(mx) : CGX_MAXSCORE;
    }
    float fgivene = 0.f;
    if (any_e) {
#pragma unroll
        for (int j = 0; j < 5; j++) if (j < nf) lex_touch(lex, lex_mask, F[j], -1);
    }
#pragma unroll
    for (int j = 0; j < 5; j++) {
        if (j >= nf) break;
        float mx = mxf[j];
        if (any_e) { lex_get(lex, lex_mask, F[j], -1, &v1, &v2); mx = fmaxf(mx, v2); }
        fgivene += mx > 0.f ? -__log10f(mx) : CGX_MAXSCORE;
    }
    out.max_lex_f_given_e = fgivene;
    out.max_lex_e_given_f = egivenf;
    rules[r] = out;
}

// per converted id: index of its first rule (the rules come out in ascending id order, so an id's rules are consecutive:
// globalOnPairsUpDown*, ExtractPair.cu:3745-3756, :3805-3816, :2082) and, in bits 18..26 of its idinfo word, how many there are.
// The count is last - first + 1 < 512, added modulo 512 in two halves by the threads at the two ends of the id's run.
static_assert(CGX_SAMPLER < 512, "idinfo packs f, fs and the rule count of an id in nine bits each: all three are bounded by the sampler");
__global__ void agg_updown_kernel(const int32_t *__restrict__ rule_id, uint32_t n_rules, int32_t *__restrict__ first, uint32_t *__restrict__ idinfo) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rules) return;
    const int32_t id = rule_id[r];
    if (r == 0 || rule_id[r - 1] != id) { first[id] = (int32_t)r; atomicAdd(&idinfo[id], ((513u - (r & 511u)) & 511u) << 18); }
    if (r == n_rules - 1 || rule_id[r + 1] != id) atomicAdd(&idinfo[id], (r & 511u) << 18);
}

void stage_aggregate(const Index &ix, Batch &b, cudaStream_t stream) {
    AggIdx a{ix.str.ptr<int32_t>(), ix.tgt.ptr<int32_t>(), b.phrases.ptr<int32_t>(), b.pat1.ptr<Pat1>(), b.pat2.ptr<Pat2>(), b.G, b.D1, b.D2};
    const int G = b.G, D1 = b.D1, D2 = b.D2;
    const int nids[3] = {G, 2 * G + D1, G + D2 + 2 * D1};
    const uint32_t ns0 = (uint32_t)b.n_slots[0], ns1 = (uint32_t)b.n_slots[1], ns2 = (uint32_t)b.n_slots[2];
    const uint32_t *so0 = b.slot_off[0].ptr<uint32_t>(), *so1 = b.slot_off[1].ptr<uint32_t>(), *so2 = b.slot_off[2].ptr<uint32_t>();
    AggLayout lay[3];
    lay[0].n_regions = 1; lay[0].r[0] = {0u, 0, so0};
    lay[1].n_regions = 3; lay[1].r[0] = {0u, 0, so0}; lay[1].r[1] = {ns0, G, so0}; lay[1].r[2] = {2 * ns0, 2 * G, so1};
    lay[2].n_regions = 4; lay[2].r[0] = {0u, 0, so0}; lay[2].r[1] = {ns0, G, so2}; lay[2].r[2] = {ns0 + ns2, G + D2, so1}; lay[2].r[3] = {ns0 + ns2 + ns1, G + D2 + D1, so1};
    uint32_t *tot = b.counters.get<uint32_t>(32);
    int *collision = (int *)(tot + 14);
    for (int kk = 0; kk < 3; kk++) {
        const int kind = 2 - kk;                           // largest result first: its D2H overlaps the other kinds' kernels
        const uint32_t N = (uint32_t)b.rec_cells[kind];
        lay[kind].cells = N;
        for (int k = lay[kind].n_regions; k < 4; k++) lay[kind].r[k] = lay[kind].r[0];
        b.n_ids[kind] = nids[kind];
        b.n_rules[kind] = 0;
        int32_t *h_ud = b.h_updown[kind].get<int32_t>((size_t)nids[kind] + 2);
        uint32_t *h_ii = b.h_idinfo[kind].get<uint32_t>((size_t)nids[kind] + 1);
        b.h_rules[kind].get<cgx_rule_t>(1);
        if (N == 0 || nids[kind] == 0) {
            memset(h_ud, 0xff, sizeof(int32_t) * (size_t)nids[kind]);
            memset(h_ii, 0, sizeof(uint32_t) * (size_t)nids[kind]);
            continue;
        }
        const RuleRec *rec = b.rec[kind].ptr<RuleRec>();
        uint64_t *hash = b.rec_hash.get<uint64_t>((size_t)N);
        uint32_t *flags = b.rec_flags.get<uint32_t>((size_t)N + 2);
        unsigned long long *acc_best = b.rec_meta.get<unsigned long long>((size_t)N);
        uint32_t *acc_cnt = b.rec_cnt.get<uint32_t>((size_t)N);
        uint32_t *tag = b.rec_tag.get<uint32_t>((size_t)N + 8);
        uint32_t *live = b.rec_live.get<uint32_t>((size_t)N + 2);
        int32_t *updown = b.updown[kind].get<int32_t>((size_t)nids[kind] + 1);      // first rule of every id (-1: none)
        uint32_t *idinfo = b.idinfo[kind].get<uint32_t>((size_t)nids[kind] + 1);
        uint64_t seed = 0x243f6a8885a308d3ULL;
        uint32_t R = 0;
        for (int attempt = 0; attempt < 8; attempt++, seed = seed * 6364136223846793005ULL + 1442695040888963407ULL) {
            CUDA_CHECK(cudaMemsetAsync(collision, 0, sizeof(int), stream));
            PROF("agg_hash", (double)N * 44, (agg_hash_kernel<<<cgx_div_up(N, 256), 256, 0, stream>>>(rec, N, ix.tgt.ptr<int32_t>(), seed, hash, tag, live, acc_best, acc_cnt)));
            exclusive_scan_u32(live, live, N, tot + 16, stream, b.scan, 0, &b.launches);          // live[N] = total below
            CUDA_CHECK(cudaMemcpyAsync(live + N, tot + 16, sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
            uint32_t *live_list = b.rec_list.get<uint32_t>((size_t)N + 2);
            PROF("agg_group", (double)N * (8 + 4 + 4), (agg_live_list_kernel<<<cgx_div_up(N, 256), 256, 0, stream>>>(hash, live, N, live_list, flags)));
            PROF("agg_group", 0.0, (agg_group_kernel<<<cgx_div_up(N, 256), 256, 0, stream>>>(lay[kind], rec, hash, tag, live, live_list, ix.tgt.ptr<int32_t>(), flags, acc_best, acc_cnt, collision)));
            b.launches++;
            exclusive_scan_u32(flags, flags, N, tot, stream, b.scan, 0, &b.launches);
            b.launches += 2;
            uint32_t hostv[20];
            cgx_read_back(hostv, tot, sizeof(hostv), stream);
            if (hostv[14] == 0) {
                R = hostv[0];
                const unsigned long long nr = hostv[16];
                b.n_rec[kind] = (int64_t)nr;
                if (g_prof && g_prof->enabled) {          // records actually grouped: 16 B record + 8 B meta per head, 20 B per record compared
                    g_prof->table["agg_group"].bytes += (double)nr * 16 + (double)R * 8;
                    g_prof->table["agg_hash"].bytes += (double)nr * 12;
                }
                if (R == 0) break;
                cgx_rule_t *rules = b.rules[kind].get<cgx_rule_t>(R);
                int32_t *rule_id = b.rule_id.get<int32_t>((size_t)R + 1);
                CUDA_CHECK(cudaMemsetAsync(idinfo, 0, sizeof(uint32_t) * (size_t)nids[kind], stream));
                uint32_t *head_cell = b.rule_head.get<uint32_t>((size_t)R + 2);
                agg_head_cell_kernel<<<cgx_div_up(N, 256), 256, 0, stream>>>(flags, N, R, head_cell);
                // (issuing the nf + 1 probes of a target terminal together was measured: 80-96 registers, 12.0 -> 14.9 ms; the table is
                // L2-resident and the kernel is bound by instruction issue, not by the probe latency)
                PROF("agg_rules", (double)R * (4 + 8 + 16 + 16 + 4) + (double)R * 13 * 16, (agg_rules_kernel<<<cgx_div_up(R, 128), 128, 0, stream>>>(a, kind, rec, head_cell, acc_best, acc_cnt, R,
                                                                       ix.lex_hash.ptr<ulonglong2>(), ix.lex_hash_mask, rules, rule_id, idinfo)));
                CUDA_CHECK(cudaMemsetAsync(updown, 0xff, sizeof(int32_t) * (size_t)nids[kind], stream));
                agg_updown_kernel<<<cgx_div_up(R, 256), 256, 0, stream>>>(rule_id, R, updown, idinfo);
                b.launches += 3;
                break;
            }
            CGX_REQUIRE(attempt < 7, "aggregation: hash collisions persisted over 8 seeds");
        }
        b.n_rules[kind] = (int32_t)R;
        if (b.fetch_results) {
            cgx_rule_t *h_r = b.h_rules[kind].get<cgx_rule_t>((size_t)R + 1);
            if (R) fetch_async(b, h_r, b.rules[kind].ptr<cgx_rule_t>(), sizeof(cgx_rule_t) * R, stream);
            if (R) fetch_async(b, h_ud, updown, sizeof(int32_t) * (size_t)nids[kind], stream);
            else memset(h_ud, 0xff, sizeof(int32_t) * (size_t)nids[kind]);          // no rule at all: every id empty
            if (R) fetch_async(b, h_ii, idinfo, sizeof(uint32_t) * (size_t)nids[kind], stream);
            else memset(h_ii, 0, sizeof(uint32_t) * (size_t)nids[kind]);
        }
    }
}

}  // namespace cgx
