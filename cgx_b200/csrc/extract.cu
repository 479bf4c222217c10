// cgx-b200: per-occurrence alignment-consistency checking and rule extraction.
//
// Replaces extractConsistentPairs_Gappy (ExtractPair.cu:1055-1795: ab, Xab, abX, XabX from the sampled
// occurrences of every distinct contiguous phrase), extractConsistentPairs_OneGap (:351-889: aXb, XaXb,
// aXbX from the sampled hits of every aXb) and extractConsistentPairs_TwoGap (:891-1053: aXbXc), their
// device helpers consistent / checkBoundary / checkBoundaryFast / checkBoundaryFast2 (:103-342), and the
// thrust sorts of their outputs (:3417-3441, :3528, :3652-3662).
//
// Work decomposition: the reference launches one 512-thread CTA per pattern and lets every thread run a
// <= 300-iteration loop per candidate just to decide whether the candidate is sampled (:1151-1160).  Here
// the sampled index set is enumerated directly -- slot j of a pattern with n > S occurrences is
// occurrence (int)((double)((float)j * ((float)n * (1.0f/S))) + 0.5), the exact arithmetic of the
// reference under -use_fast_math (single FMUL by the rounded reciprocal, verified in its sm_100 SASS) --
// and ONE thread is launched per sampled occurrence over a flat, load-balanced slot list (prefix sum of
// min(n, S) per pattern).  Every occurrence emits at most one record of each rule shape, so each (shape, slot) has its
// OWN 16-byte cell: record arrays are slot-indexed regions laid out in ascending converted-id order (no atomics, no
// counters, deterministic, and already grouped by source id for the aggregation -- aggregate.cu); cells that stay
// empty keep id = -1 from the memset that precedes the kernels.
//   kind 0 (ab)   : [ab: ns0]
//   kind 1        : [Xab: ns0][abX: ns0][aXb: ns1]
//   kind 2        : [XabX: ns0][aXbXc: ns2][XaXb: ns1][aXbX: ns1]          (ns0/1/2 = contiguous / one-gap / two-gap slots)
//
// Measured and dropped (round 1c, C2): the extension loops run with 6-8 of 32 lanes active (ncu: 5.9 / 8.5 active threads per
// warp, SM pipes 75-80 % busy), so the survivors of the seed phase were compacted (a) inside the CTA through shared memory --
// the retired threads keep their warp slots until the CTA ends, occupancy collapses: extract_onegap 9.8 -> 16.6 ms -- and
// (b) through a global queue into a second kernel with full warps -- the extension phase then re-fetches the window sectors the
// seed phase had just pulled into L1, and latency, not issue slots, is what it waits for: 9.8 -> 14.9 ms.  A slot -> pattern
// array instead of the binary search costs 1.0 ms to fill and saves 0.2 ms.  (c) Persistent warps that queue the survivors of 8 chunks
// of 32 slots in shared memory and run the extension phase on 32 of them at a time: extract_onegap 9.8 -> 23.8 ms, extract_contig
// 6.1 -> 6.8 ms (every full warp then waits for its longest extension chain).  The one-thread-per-occurrence form stays.
#include "batch.h"
#include "align_fields.cuh"
#include "prof.h"

namespace cgx {

// xw[k] = RLP[k] | 1 for a word (token >= 2), 0 at EOS / padding: the extension loops test "is a word" and read its aligned
// span with ONE load (the reference reads str[k], then RLP[k]: two dependent random sectors); lrq[j] = range-minimum table over {L_tar, R_tar} (index.cu).
// RLP itself is only read for the target offset stored at the previous EOS.
template <class A>
struct ExtractIdx {
    const int32_t *sa;
    const typename A::word_t *xw, *RLP;
    const typename A::lrq_t *lrq;
    int n;
};

// ExtractPair.cu:103-133 consistent: min L_tar / max R_tar over the target window [start, end] (unaligned tokens skipped) must
// be exactly the source span [start_chk, end_chk].  The window is the union of two entries of the range-minimum table lrq
// (index.cu): two independent 8-byte loads.  Callers never pass more than CGX_MAX_RULE_SPAN tokens; longer windows take the loop.
template <class A>
__device__ __forceinline__ bool consistent(const ExtractIdx<A> &x, int start, int end, int start_chk, int end_chk, int startpos_source) {
    unsigned mn = A::UNAL, mx = 0;
    const int len = end - start + 1;
    if (len > 0) {
        const int k = min(3, 31 - __clz(len)), step = 1 << k;
        const typename A::lrq_t a = __ldg(&x.lrq[start]), b = __ldg(&x.lrq[end - step + 1]);
        unsigned mna, mxa, mnb, mxb;
        A::lrq_get(a, k, mna, mxa);
        A::lrq_get(b, k, mnb, mxb);
        mn = min(mna, mnb);
        mx = max(mxa, mxb);
        for (int j = start + step; j <= end - step; j += step) {                     // runs only when len > 15
            A::lrq_get(__ldg(&x.lrq[j]), k, mna, mxa);
            mn = min(mn, mna); mx = max(mx, mxa);
        }
    }
    return !(startpos_source + (int)mn != start_chk || startpos_source + (int)mx != end_chk);
}

__device__ __forceinline__ void emit(RuleRec *__restrict__ out, size_t cell, int id, unsigned ts, unsigned te, int g1s, int g1e, int g2s, int g2e) {
    RuleRec r;
    r.id = id; r.tgt_start = (int32_t)ts; r.end = (uint8_t)(te - ts);
    r.gap1 = g1s < 0 ? 255 : (uint8_t)(g1s - (int)ts); r.gap1_1 = g1s < 0 ? 255 : (uint8_t)(g1e - (int)ts);
    r.gap2 = g2s < 0 ? 255 : (uint8_t)(g2s - (int)ts); r.gap2_1 = g2s < 0 ? 255 : (uint8_t)(g2e - (int)ts);
    r.pad[0] = r.pad[1] = r.pad[2] = 0;
    *reinterpret_cast<uint4 *>(&out[cell]) = *reinterpret_cast<const uint4 *>(&r);
}

// ExtractPair.cu:1133-1160 / :445-471 / :946-972 sampling.  Returns the occurrence index of slot j, or -1.
__device__ __forceinline__ int sample_index(int j, int n, int S, float rcp) {
    if (n <= S) return j;
    float step = __fmul_rn((float)n, rcp);
    int idx = (int)((double)__fmul_rn((float)j, step) + 0.5);
    if (j > 0 && (int)((double)__fmul_rn((float)(j - 1), step) + 0.5) == idx) return -1;
    return idx < n ? idx : -1;
}

// owner of a slot = largest i with off[i] <= slot (off: exclusive prefix of the slots per pattern, n + 1 entries).  hint[b] is the
// owner of slot 128 b (owner_hints_kernel), so the search runs over [hint[b], hint[b + 1]] -- a handful of patterns -- instead
// of all n: 3-7 dependent loads instead of 23 at C2 (the full search was 15 % of extract_onegap's stall samples, r02a)
constexpr int OWNER_BLOCK_LOG = 7;
__device__ __forceinline__ int find_owner_u32(const uint32_t *__restrict__ off, const uint32_t *__restrict__ hint, uint32_t slot) {
    int lo = (int)__ldg(&hint[slot >> OWNER_BLOCK_LOG]), hi = (int)__ldg(&hint[(slot >> OWNER_BLOCK_LOG) + 1]) + 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(&off[mid]) <= slot) lo = mid; else hi = mid;
    }
    return lo;
}
// hint[b] = pattern that owns slot 128 b, for every block start below the total; hint[blocks] = the last pattern
__global__ void owner_hints_kernel(const uint32_t *__restrict__ off, int n, uint32_t total, uint32_t *__restrict__ hint) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    const uint32_t s = off[d], e = off[d + 1];
    for (uint32_t b = (s + (1u << OWNER_BLOCK_LOG) - 1) >> OWNER_BLOCK_LOG; e > s && (b << OWNER_BLOCK_LOG) < e; b++) hint[b] = (uint32_t)d;
    if (d == n - 1) hint[((total + (1u << OWNER_BLOCK_LOG) - 1) >> OWNER_BLOCK_LOG)] = (uint32_t)(n - 1);
}

// ------------------------------------------------------------------------------------------------
// contiguous phrases: ab, Xab, abX, XabX      (ExtractPair.cu:1163-1792)
// ------------------------------------------------------------------------------------------------
__global__ void slots_contig_kernel(const int32_t *__restrict__ phrases, int G, uint32_t *__restrict__ cnt) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G) cnt[g] = (uint32_t)min(phrases[g * 4 + 1] - phrases[g * 4] + 1, CGX_SAMPLER);
}

template <class A>
__global__ void __launch_bounds__(128) extract_contig_kernel(ExtractIdx<A> x, const int32_t *__restrict__ phrases, int G, const uint32_t *__restrict__ slot_off, const uint32_t *__restrict__ hint,
                                                             uint32_t n_slots, RuleRec *__restrict__ rec_ab, RuleRec *__restrict__ rec_Xab,
                                                             RuleRec *__restrict__ rec_abX, RuleRec *__restrict__ rec_XabX) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const int bnum = find_owner_u32(slot_off, hint, slot);
    const int start = phrases[bnum * 4], end = phrases[bnum * 4 + 1], longestmatch = phrases[bnum * 4 + 2];
    const int occ = sample_index((int)(slot - slot_off[bnum]), end - start + 1, CGX_SAMPLER, 1.0f / (float)CGX_SAMPLER);
    if (occ < 0) return;
    const int current_str = __ldg(&x.sa[start + occ]);
    const int globalc = G;
    const int SPAN = CGX_MAX_RULE_SPAN;

    unsigned L, R;
    typename A::word_t temp;
    int sen_target_begin = -1, tempind = 0;
    unsigned min_L = A::UNAL, max_R = 0;
    unsigned gap1_start = 0, gap1_end = 0, gap2_start = 0, gap2_end = 0, target_start = 0, target_end = 0;
    bool next = true, abX = true, Xab = true, XabX = true, ab = true, XabNoSuccess = true, abXNoSuccess = true;
    int XabCount = 0, abXCount = 0;
    unsigned min_L_Xab = A::UNAL, max_R_Xab = 0, min_L_abX = A::UNAL, max_R_abX = 0, min_L_XabX = A::UNAL, max_R_XabX = 0;

    for (int k = current_str; k < current_str + longestmatch; k++) {
        temp = __ldg(&x.xw[k]);
        L = A::L(temp); R = A::R(temp);
        if (k == current_str) {
            tempind = k - (int)A::P(temp) - 1;
            sen_target_begin = tempind == -1 ? 0 : (int)(uint32_t)__ldg(&x.RLP[tempind]);
        }
        if ((L == A::UNAL || R == A::UNAL) && (k == current_str || k == current_str + longestmatch - 1)) {
            ab = false;
            if (k == current_str) abXNoSuccess = false; else XabNoSuccess = false;
        } else if (L == A::UNAL || R == A::UNAL) {
        } else { min_L = min(min_L, L); max_R = max(max_R, R); }
    }
    if (min_L > max_R || max_R - min_L >= (unsigned)SPAN) { abX = false; Xab = false; XabX = false; ab = false; }
    tempind++;
    const int ender = current_str + longestmatch - 1;
    if (ab && consistent(x, (int)min_L + sen_target_begin, (int)max_R + sen_target_begin, current_str, ender, tempind))
        emit(rec_ab, slot, bnum, min_L + sen_target_begin, max_R + sen_target_begin, -1, -1, -1, -1);
    if (longestmatch + 1 > CGX_MAX_RULE_SYMBOLS) { abX = false; Xab = false; }
    if (longestmatch + 2 > CGX_MAX_RULE_SYMBOLS) XabX = false;

    for (int i = 1; longestmatch + i <= SPAN && (abXNoSuccess || XabNoSuccess || XabX); i++) {
        // ---- X on the left: tokens current_str-i .. current_str-1 ----
        temp = (Xab && current_str - i >= 0) ? __ldg(&x.xw[current_str - i]) : 0u;
        if (temp & 1u) {
            next = true;
            L = A::L(temp); R = A::R(temp);
            if (L == A::UNAL || R == A::UNAL) { next = false; if (i == 1) { Xab = false; XabX = false; } }
            else { min_L_Xab = min(min_L_Xab, L); max_R_Xab = max(max_R_Xab, R); }
            if (next && min_L_Xab > max_R_Xab) return;
            if ((int)max_R_Xab - (int)min_L_Xab >= SPAN) { next = false; Xab = false; }
            if (next) {
                gap1_start = sen_target_begin + min_L_Xab; gap1_end = sen_target_begin + max_R_Xab;
                next = consistent(x, (int)gap1_start, (int)gap1_end, current_str - i, current_str - 1, tempind);
                if (next) XabCount = i;
            }
            if (XabNoSuccess && next) {
                target_start = sen_target_begin + min(min_L_Xab, min_L);
                target_end = sen_target_begin + max(max_R_Xab, max_R);
                if (target_end - target_start >= (unsigned)SPAN) { next = false; Xab = false; }
                if (next) next = consistent(x, (int)target_start, (int)target_end, current_str - i, ender, tempind);
            }
            if (XabNoSuccess && next) {
                emit(rec_Xab, slot, bnum, target_start, target_end, (int)gap1_start, (int)gap1_end, -1, -1);
                XabNoSuccess = false;
            }
        } else Xab = false;
        // ---- X on the right: tokens ender+1 .. ender+i ----
        temp = abX ? __ldg(&x.xw[ender + i]) : 0u;
        if (temp & 1u) {
            next = true;
            L = A::L(temp); R = A::R(temp);
            if (L == A::UNAL || R == A::UNAL) { next = false; if (i == 1) { abX = false; XabX = false; } }
            else { min_L_abX = min(min_L_abX, L); max_R_abX = max(max_R_abX, R); }
            if (next && min_L_abX > max_R_abX) return;
            if ((int)max_R_abX - (int)min_L_abX >= SPAN) { next = false; abX = false; }
            if (next) {
                gap1_start = sen_target_begin + min_L_abX; gap1_end = sen_target_begin + max_R_abX;
                next = consistent(x, (int)gap1_start, (int)gap1_end, ender + 1, ender + i, tempind);
                if (next) abXCount = i;
            }
            if (abXNoSuccess && next) {
                target_start = sen_target_begin + min(min_L_abX, min_L);
                target_end = sen_target_begin + max(max_R_abX, max_R);
                if (target_end - target_start >= (unsigned)SPAN) { next = false; abX = false; }
                if (next) next = consistent(x, (int)target_start, (int)target_end, current_str, ender + i, tempind);
            }
            if (abXNoSuccess && next) {
                emit(rec_abX, slot, globalc + bnum, target_start, target_end, (int)gap1_start, (int)gap1_end, -1, -1);
                abXNoSuccess = false;
            }
        } else abX = false;
        // ---- XabX ----
        if (XabX && (abX || Xab)) {
            if (XabCount == i) {          // left gap just validated; look for the smallest valid right gap
                min_L_XabX = A::UNAL; max_R_XabX = 0;
                for (int icount = 1; XabX && icount <= abXCount; icount++) {
                    next = true;
                    if (icount + XabCount + longestmatch <= SPAN) {
                        temp = __ldg(&x.xw[ender + icount]);
                        L = A::L(temp); R = A::R(temp);
                        if (L == A::UNAL || R == A::UNAL) { next = false; if (i == 1) return; }
                        else { min_L_XabX = min(min_L_XabX, L); max_R_XabX = max(max_R_XabX, R); }
                    } else { next = false; icount = abXCount + 1; }
                    if (next && (int)max_R_XabX - (int)min_L_XabX >= SPAN) { next = false; icount = abXCount + 1; }
                    if (next) {
                        gap2_start = sen_target_begin + min_L_XabX; gap2_end = sen_target_begin + max_R_XabX;
                        if (min_L_XabX > max_R_XabX) return;
                        next = consistent(x, (int)gap2_start, (int)gap2_end, ender + 1, ender + icount, tempind);
                    }
                    if (next) {
                        target_start = sen_target_begin + min(min(min_L_XabX, min_L_Xab), min_L);
                        target_end = sen_target_begin + max(max(max_R_XabX, max_R_Xab), max_R);
                        if (target_end - target_start >= (unsigned)SPAN) { next = false; icount = abXCount + 1; }
                        if (next) next = consistent(x, (int)target_start, (int)target_end, current_str - XabCount, ender + icount, tempind);
                        if (next) {
                            gap1_start = sen_target_begin + min_L_Xab; gap1_end = sen_target_begin + max_R_Xab;
                            emit(rec_XabX, slot, bnum, target_start, target_end, (int)gap1_start, (int)gap1_end, (int)gap2_start, (int)gap2_end);
                            XabX = false;
                        }
                    }
                }
            }
            if (XabX && abXCount == i) {  // right gap just validated; look for the smallest valid left gap
                min_L_XabX = A::UNAL; max_R_XabX = 0;
                for (int icount = 1; XabX && icount <= XabCount; icount++) {
                    next = true;
                    if (icount + abXCount + longestmatch <= SPAN) {
                        temp = __ldg(&x.xw[current_str - icount]);
                        L = A::L(temp); R = A::R(temp);
                        if (L == A::UNAL || R == A::UNAL) { next = false; if (i == 1) return; }
                        else { min_L_XabX = min(min_L_XabX, L); max_R_XabX = max(max_R_XabX, R); }
                    } else { icount = XabCount + 1; next = false; }
                    if (next && (int)max_R_XabX - (int)min_L_XabX >= SPAN) { icount = XabCount + 1; next = false; }
                    if (next) {
                        gap1_start = sen_target_begin + min_L_XabX; gap1_end = sen_target_begin + max_R_XabX;
                        if (min_L_XabX > max_R_XabX) return;
                        next = consistent(x, (int)gap1_start, (int)gap1_end, current_str - icount, current_str - 1, tempind);
                    }
                    if (next) {
                        target_start = sen_target_begin + min(min(min_L_XabX, min_L_abX), min_L);
                        target_end = sen_target_begin + max(max(max_R_XabX, max_R_abX), max_R);
                        if (target_end - target_start >= (unsigned)SPAN) { next = false; icount = XabCount + 1; }
                        if (next) next = consistent(x, (int)target_start, (int)target_end, current_str - icount, ender + abXCount, tempind);
                        if (next) {
                            gap2_start = sen_target_begin + min_L_abX; gap2_end = sen_target_begin + max_R_abX;
                            emit(rec_XabX, slot, bnum, target_start, target_end, (int)gap1_start, (int)gap1_end, (int)gap2_start, (int)gap2_end);
                            XabX = false;
                        }
                    }
                }
            }
        } else XabX = false;
        if (!XabX) {                                          // ExtractPair.cu:1782-1789
            if (!Xab && XabNoSuccess) XabNoSuccess = false;
            if (!abX && abXNoSuccess) abXNoSuccess = false;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// boundary helpers for the gappy seeds
// ------------------------------------------------------------------------------------------------
// ExtractPair.cu:135-194 checkBoundaryFast / :196-250 checkBoundaryFast2 (same scan; Fast2 reports absolute target span)
template <class A>
__device__ __forceinline__ bool boundary_fast(const ExtractIdx<A> &x, int start, int ender, unsigned *min_LL, unsigned *max_RR, int *sen_target_begin,
                                              int *tempind) {
    unsigned min_L = A::UNAL, max_R = 0;
    *sen_target_begin = -1; *tempind = 0;
    for (int k = start; k <= ender; k++) {
        typename A::word_t w = __ldg(&x.xw[k]);
        unsigned L = A::L(w), R = A::R(w);
        if ((L == A::UNAL || R == A::UNAL) && (k == start || k == ender)) return false;
        if (L == A::UNAL || R == A::UNAL) continue;
        if (k == start) {
            *tempind = k - (int)A::P(w) - 1;
            *sen_target_begin = (*tempind == -1) ? 0 : (int)(uint32_t)__ldg(&x.RLP[*tempind]);
        }
        min_L = min(min_L, L); max_R = max(max_R, R);
    }
    *min_LL = min_L; *max_RR = max_R;
    if (min_L <= max_R && max_R - min_L < CGX_MAX_RULE_SPAN) { (*tempind)++; return true; }
    return false;
}

// ExtractPair.cu:252-342 checkBoundary: 0 normal false, 1 ok, 2 first token unaligned, 3 last, 4 both
template <class A>
__device__ __forceinline__ int check_boundary(const ExtractIdx<A> &x, int start, int ender, unsigned *target_start, unsigned *target_end) {
    unsigned min_L = A::UNAL, max_R = 0;
    int sen_target_begin = -1, tempind = 0, wrong = 0;
    for (int k = start; k <= ender; k++) {
        typename A::word_t w = __ldg(&x.xw[k]);
        unsigned L = A::L(w), R = A::R(w);
        bool un = (L == A::UNAL || R == A::UNAL);
        if (k == start) {
            tempind = k - (int)A::P(w) - 1;
            sen_target_begin = tempind == -1 ? 0 : (int)(uint32_t)__ldg(&x.RLP[tempind]);
        }
        if (un && (k == start || k == ender)) {
            if (start == ender && wrong == 0) wrong = 4;
            else if (wrong == 0 && k == start) wrong = 2;
            else if (wrong == 0 && k == ender) wrong = 3;
            else if (wrong != 0) wrong = 4;
        } else if (!un) { min_L = min(min_L, L); max_R = max(max_R, R); }
    }
    *target_start = min_L + sen_target_begin; *target_end = max_R + sen_target_begin;
    if (wrong) return wrong;
    if (min_L <= max_R && max_R - min_L < CGX_MAX_RULE_SPAN) {
        tempind++;
        if (consistent(x, (int)*target_start, (int)*target_end, start, ender, tempind)) return 1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// one-gap seeds: aXb, XaXb, aXbX          (ExtractPair.cu:458-887)
// ------------------------------------------------------------------------------------------------
__global__ void slots_pat1_kernel(const Pat1 *__restrict__ pat, int D1, uint32_t *__restrict__ cnt) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < D1) cnt[d] = (uint32_t)min(pat[d].hit_count, CGX_SAMPLER_ONEGAP);
}

template <class A>
__global__ void __launch_bounds__(128) extract_onegap_kernel(ExtractIdx<A> x, const Pat1 *__restrict__ pat, int D1, const uint64_t *__restrict__ hits1,
                                                             const uint32_t *__restrict__ slot_off, const uint32_t *__restrict__ hint, uint32_t n_slots, int G, int D2, int pbits,
                                                             RuleRec *__restrict__ rec_aXb, RuleRec *__restrict__ rec_XaXb,
                                                             RuleRec *__restrict__ rec_aXbX) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const int d = find_owner_u32(slot_off, hint, slot);
    const Pat1 p = pat[d];
    const int occ = sample_index((int)(slot - slot_off[d]), p.hit_count, CGX_SAMPLER_ONEGAP, 1.0f / (float)CGX_SAMPLER_ONEGAP);
    if (occ < 0) return;
    const uint64_t hk = hits1[(size_t)p.hit_start + occ];
    const int current_str = (int)((hk >> 4) & ((1ull << pbits) - 1)), firstEnd = (int)(hk & 15);
    const int startLen = p.ls, endLen = p.le;
    const int SPAN = CGX_MAX_RULE_SPAN;
    const int ender = current_str + firstEnd;
    unsigned min_L, max_R;
    int sen_target_begin, tempind;
    if (!boundary_fast(x, current_str + startLen, ender - endLen, &min_L, &max_R, &sen_target_begin, &tempind)) return;
    unsigned gap1_start = min_L + sen_target_begin, gap1_end = max_R + sen_target_begin;
    unsigned target_start = 0, target_end = 0;
    bool next = true, left = true, right = true;
    int re = check_boundary(x, current_str, ender, &target_start, &target_end);
    min_L = (target_start - (unsigned)sen_target_begin) & A::UNAL;
    max_R = (target_end - (unsigned)sen_target_begin) & A::UNAL;
    if (re == 0) next = false;
    else if (re == 2) { next = false; right = false; }
    else if (re == 3) { next = false; left = false; }
    else if (re == 4) { next = false; left = false; right = false; }
    if ((target_start == 0 && target_end == 0) || min_L > max_R || gap1_start < target_start || gap1_end > target_end) return;   // :591-595
    if (next) emit(rec_aXb, slot, 2 * G + d, target_start, target_end, (int)gap1_start, (int)gap1_end, -1, -1);
    if (startLen + endLen + 2 > CGX_MAX_RULE_SYMBOLS) return;
    const unsigned originalGapStart = gap1_start, originalGapEnd = gap1_end;
    unsigned min_XaXb = A::UNAL, max_XaXb = 0, min_aXbX = A::UNAL, max_aXbX = 0, L, R;
    typename A::word_t temp;
    for (int i = 1; firstEnd + 1 + i <= SPAN && (left || right); i++) {
        temp = (left && current_str - i >= 0) ? __ldg(&x.xw[current_str - i]) : 0u;
        if (temp & 1u) {
            next = true;
            L = A::L(temp); R = A::R(temp);
            if (L == A::UNAL || R == A::UNAL) { next = false; if (i == 1) left = false; }
            else { min_XaXb = min(min_XaXb, L); max_XaXb = max(max_XaXb, R); }
            if (next && min_XaXb > max_XaXb) return;
            if ((int)max_XaXb - (int)min_XaXb >= SPAN) { next = false; left = false; }
            unsigned g_s = 0, g_e = 0;
            if (next) {
                g_s = sen_target_begin + min_XaXb; g_e = sen_target_begin + max_XaXb;
                next = consistent(x, (int)g_s, (int)g_e, current_str - i, current_str - 1, tempind);
            }
            if (next) {
                target_start = sen_target_begin + min(min_XaXb, min_L);
                target_end = sen_target_begin + max(max_XaXb, max_R);
                if (target_end - target_start >= (unsigned)SPAN) { next = false; left = false; }
                if (next) next = consistent(x, (int)target_start, (int)target_end, current_str - i, ender, tempind);
            }
            if (next) {
                emit(rec_XaXb, slot, G + D2 + d, target_start, target_end, (int)g_s, (int)g_e, (int)originalGapStart, (int)originalGapEnd);
                left = false;
            }
        } else left = false;
        temp = right ? __ldg(&x.xw[ender + i]) : 0u;
        if (temp & 1u) {
            next = true;
            L = A::L(temp); R = A::R(temp);
            if (L == A::UNAL || R == A::UNAL) { next = false; if (i == 1) right = false; }
            else { min_aXbX = min(min_aXbX, L); max_aXbX = max(max_aXbX, R); }
            if (next && min_aXbX > max_aXbX) return;
            if ((int)max_aXbX - (int)min_aXbX >= SPAN) { next = false; right = false; }
            unsigned g_s = 0, g_e = 0;
            if (next) {
                g_s = sen_target_begin + min_aXbX; g_e = sen_target_begin + max_aXbX;
                next = consistent(x, (int)g_s, (int)g_e, ender + 1, ender + i, tempind);
            }
            if (next) {
                target_start = sen_target_begin + min(min_aXbX, min_L);
                target_end = sen_target_begin + max(max_aXbX, max_R);
                if (target_end - target_start >= (unsigned)SPAN) { next = false; right = false; }
                if (next) next = consistent(x, (int)target_start, (int)target_end, current_str, ender + i, tempind);
            }
            if (next) {
                emit(rec_aXbX, slot, G + D2 + D1 + d, target_start, target_end, (int)originalGapStart, (int)originalGapEnd, (int)g_s, (int)g_e);
                right = false;
            }
        } else right = false;
    }
}

// ------------------------------------------------------------------------------------------------
// two-gap seeds: aXbXc                      (ExtractPair.cu:959-1052)
// ------------------------------------------------------------------------------------------------
__global__ void slots_pat2_kernel(const Pat2 *__restrict__ pat, int D2, uint32_t *__restrict__ cnt) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < D2) cnt[d] = (uint32_t)min(pat[d].hit_count, CGX_SAMPLER_TWOGAP);
}

template <class A>
__global__ void __launch_bounds__(128) extract_twogap_kernel(ExtractIdx<A> x, const Pat2 *__restrict__ pat2, const Pat1 *__restrict__ pat1, int D2,
                                                             const uint64_t *__restrict__ hits2, const uint32_t *__restrict__ slot_off, const uint32_t *__restrict__ hint, uint32_t n_slots,
                                                             int G, int pbits, RuleRec *__restrict__ rec_aXbXc) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const int d = find_owner_u32(slot_off, hint, slot);
    const Pat2 p2 = pat2[d];
    const int occ = sample_index((int)(slot - slot_off[d]), p2.hit_count, CGX_SAMPLER_TWOGAP, 1.0f / (float)CGX_SAMPLER_TWOGAP);
    if (occ < 0) return;
    const uint64_t hk = hits2[(size_t)p2.hit_start + occ];
    const int current_str = (int)((hk >> 8) & ((1ull << pbits) - 1)), firstEnd = (int)(hk & 15), secondEnd = firstEnd + 1 + (int)((hk >> 4) & 15);   // key: width of the second gap, then length
    const Pat1 p1 = pat1[p2.pat1];
    unsigned mnL, mxR;
    int stb, ti;
    // first and second gap (checkBoundaryFast2 without the width test failing = "not possible" in the reference)
    if (!boundary_fast(x, current_str + p1.ls, current_str + firstEnd - p1.le, &mnL, &mxR, &stb, &ti)) return;
    unsigned g1s = mnL + stb, g1e = mxR + stb;
    if (!boundary_fast(x, current_str + firstEnd + 1, current_str + secondEnd - 1, &mnL, &mxR, &stb, &ti)) return;
    unsigned g2s = mnL + stb, g2e = mxR + stb;
    unsigned ts, te;
    if (check_boundary(x, current_str, current_str + secondEnd, &ts, &te) == 1)
        emit(rec_aXbXc, slot, G + d, ts, te, (int)g1s, (int)g1e, (int)g2s, (int)g2e);
}

// the three kernels of one batch on one field layout (align_fields.cuh)
template <class A>
static void launch_extract(const Index &ix, Batch &b, cudaStream_t stream, const uint32_t *ns, const uint32_t *so0, const uint32_t *so1, const uint32_t *so2,
                           RuleRec *r0, RuleRec *r1, RuleRec *r2) {
    const int G = b.G, D1 = b.D1, D2 = b.D2;
    ExtractIdx<A> x{ix.sa.ptr<int32_t>(), ix.xw.ptr<typename A::word_t>(), ix.RLP.ptr<typename A::word_t>(), ix.lr.ptr<typename A::lrq_t>(), (int)ix.n};
    // slot -> pattern hints, one per 128 slots (find_owner_u32)
    const size_t nb[3] = {((size_t)ns[0] >> OWNER_BLOCK_LOG) + 3, ((size_t)ns[1] >> OWNER_BLOCK_LOG) + 3, ((size_t)ns[2] >> OWNER_BLOCK_LOG) + 3};
    uint32_t *h0 = b.slot_hint.get<uint32_t>(nb[0] + nb[1] + nb[2]), *h1 = h0 + nb[0], *h2 = h1 + nb[1];
    if (ns[0]) owner_hints_kernel<<<cgx_div_up(G, 256), 256, 0, stream>>>(so0, G, ns[0], h0);
    if (ns[1]) owner_hints_kernel<<<cgx_div_up(D1, 256), 256, 0, stream>>>(so1, D1, ns[1], h1);
    if (ns[2]) owner_hints_kernel<<<cgx_div_up(D2, 256), 256, 0, stream>>>(so2, D2, ns[2], h2);
    if (ns[0]) PROF("extract_contig", (double)ns[0] * 56, (extract_contig_kernel<A><<<cgx_div_up(ns[0], 128), 128, 0, stream>>>(x, b.phrases.ptr<int32_t>(), G, so0, h0, ns[0], r0, r1, r1 + ns[0], r2)));
    if (ns[2]) PROF("extract_twogap", (double)ns[2] * 56, (extract_twogap_kernel<A><<<cgx_div_up(ns[2], 128), 128, 0, stream>>>(x, b.pat2.ptr<Pat2>(), b.pat1.ptr<Pat1>(), D2, b.hits2_sorted.ptr<uint64_t>(), so2, h2, ns[2], G, b.pbits, r2 + ns[0])));
    if (ns[1]) PROF("extract_onegap", (double)ns[1] * 56, (extract_onegap_kernel<A><<<cgx_div_up(ns[1], 128), 128, 0, stream>>>(x, b.pat1.ptr<Pat1>(), D1, b.hits1_sorted.ptr<uint64_t>(), so1, h1, ns[1], G, D2, b.pbits, r1 + (size_t)2 * ns[0], r2 + (size_t)ns[0] + ns[2], r2 + (size_t)ns[0] + ns[2] + ns[1])));
}

void stage_extract(const Index &ix, Batch &b, cudaStream_t stream) {
    const int G = b.G, D1 = b.D1, D2 = b.D2;
    b.n_rec[0] = b.n_rec[1] = b.n_rec[2] = 0;
    b.rec_cells[0] = b.rec_cells[1] = b.rec_cells[2] = 0;
    b.n_slots[0] = b.n_slots[1] = b.n_slots[2] = 0;
    b.samples = 0;
    if (G == 0) return;
    uint32_t *tot = b.counters.get<uint32_t>(32);
    // slot offsets: G+1 / D1+1 / D2+1 entries (the last one = total), also read by the aggregation
    uint32_t *so0 = b.slot_off[0].get<uint32_t>((size_t)G + 2);
    slots_contig_kernel<<<cgx_div_up(G, 256), 256, 0, stream>>>(b.phrases.ptr<int32_t>(), G, so0);
    exclusive_scan_u32(so0, so0, (size_t)G, tot, stream, b.scan, 0, &b.launches);
    CUDA_CHECK(cudaMemcpyAsync(so0 + G, tot, sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
    uint32_t *so1 = b.slot_off[1].get<uint32_t>((size_t)D1 + 2), *so2 = b.slot_off[2].get<uint32_t>((size_t)D2 + 2);
    if (D1) {
        slots_pat1_kernel<<<cgx_div_up(D1, 256), 256, 0, stream>>>(b.pat1.ptr<Pat1>(), D1, so1);
        exclusive_scan_u32(so1, so1, (size_t)D1, tot + 1, stream, b.scan, 0, &b.launches);
        CUDA_CHECK(cudaMemcpyAsync(so1 + D1, tot + 1, sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
    }
    if (D2) {
        slots_pat2_kernel<<<cgx_div_up(D2, 256), 256, 0, stream>>>(b.pat2.ptr<Pat2>(), D2, so2);
        exclusive_scan_u32(so2, so2, (size_t)D2, tot + 2, stream, b.scan, 0, &b.launches);
        CUDA_CHECK(cudaMemcpyAsync(so2 + D2, tot + 2, sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
    }
    uint32_t ns[3] = {0, 0, 0};
    cgx_read_back(ns, tot, sizeof(uint32_t) * 3, stream);
    if (!D1) ns[1] = 0;
    if (!D2) ns[2] = 0;
    b.launches += 3;
    b.samples = (int64_t)ns[0] + ns[1] + ns[2];
    for (int k = 0; k < 3; k++) b.n_slots[k] = ns[k];
    // one cell per (shape, slot)
    const size_t c0 = ns[0], c1 = (size_t)2 * ns[0] + ns[1], c2 = (size_t)ns[0] + ns[2] + (size_t)2 * ns[1];
    // slot totals come from 32-bit scans: the three kinds are bounded separately (a sum that wrapped would be smaller than a part)
    CGX_REQUIRE_BATCH((uint64_t)ns[0] + ns[1] + ns[2] < (1ull << 31) && c1 < (1ull << 32) && c2 < (1ull << 32),
                      "%zu / %zu record cells exceed the 32-bit cell index", c1, c2);
    b.rec_cells[0] = c0; b.rec_cells[1] = c1; b.rec_cells[2] = c2;
    RuleRec *r0 = b.rec[0].get<RuleRec>(c0 + 1), *r1 = b.rec[1].get<RuleRec>(c1 + 1), *r2 = b.rec[2].get<RuleRec>(c2 + 1);
    CUDA_CHECK(cudaMemsetAsync(r0, 0xff, sizeof(RuleRec) * c0, stream));
    CUDA_CHECK(cudaMemsetAsync(r1, 0xff, sizeof(RuleRec) * c1, stream));
    CUDA_CHECK(cudaMemsetAsync(r2, 0xff, sizeof(RuleRec) * c2, stream));
    // algorithmic bytes (SURVEY 8d B_ext, lower bound): per sampled occurrence its SA / hit entry and slot owner (8 B), the
    // RLP + text words of the smallest source window it must inspect (phrase + one extension token per side: 8 B x 5)
    // and the L/R bytes of a 4-token target window (2 x 4) = 56 B; emitted cells are not counted
    if (ix.wide) launch_extract<AlignWide>(ix, b, stream, ns, so0, so1, so2, r0, r1, r2);
    else launch_extract<AlignNarrow>(ix, b, stream, ns, so0, so1, so2, r0, r1, r2);
    b.launches += 6;
}

}  // namespace cgx
