// cgx-b200: per-batch device workspace and host-side result mirrors.
#pragma once
#include "../../include/cgx_b200.h"
#include "index.h"
#include <algorithm>
#include <utility>

namespace cgx {

// 16-byte extracted rule record (replaces res_phrase_t 9 B / rule_onegap 11 B / rule_twogap 13 B,
// ComTypes.h:224-242,349-353; padded so that records move as one 128-bit word).
struct __align__(16) RuleRec {
    int32_t id;          // converted id within its array (ExtractPair.c:723-729 / :999-1006)
    int32_t tgt_start;   // absolute start in the target text
    uint8_t end;         // target span length - 1
    uint8_t gap1, gap1_1, gap2, gap2_1;   // offsets from tgt_start, 255 = none
    uint8_t pad[3];
};
static_assert(sizeof(RuleRec) == 16, "RuleRec must be 16 bytes");

// one-gap pattern (gappy_search, ComTypes.h:167-177, re-laid out as 8 ints)
struct Pat1 {
    int32_t a_pos, ls, b_pos, le;        // corpus position of one occurrence of a / b and their lengths
    int32_t hit_start, hit_count;        // range in the sorted hit list (-1,0 when no hit)
    int32_t marker_pair;                 // >= 0: both tokens frequent and ls = le = 1 (the reference's precomputed-pair path)
    int32_t fs_extra;                    // featureMissingCount for marker patterns (SuffixArray.cu:1290)
};
struct Pat1Dev {                         // device-only companion
    int32_t up_a, down_a, up_b, down_b;
};
struct Pat2 {
    int32_t pat1, ctok, hit_start, hit_count;
};

// Everything cgx_result() exposes for one batch -- device arrays, their pinned host mirrors and the counts -- so that
// a finished batch can be parked while the next one runs: its D2H then overlaps the next batch's kernels and the host
// may still be writing the batch before it (cgx_extract_begin / cgx_result_at; three sets = one computing, one
// travelling, one being consumed).
constexpr int CGX_RESULT_SETS = 3;
struct ResultSet {
    DevBuf phrase_id, phrases, pat1, pat2, q1_off, q1_ids, q2_off, q2_ids, rules[3], updown[3], idinfo[3];
    PinnedBuf h_phrase_id, h_phrases, h_pat1, h_pat2, h_q1_off, h_q1_ids, h_q2_off, h_q2_ids, h_rules[3], h_updown[3], h_idinfo[3];
    int32_t Q = 0, T = 0, G = 0, D1 = 0, D2 = 0;
    int32_t n_rules[3] = {0, 0, 0}, n_ids[3] = {0, 0, 0};
    cgx_batch_info_t info;
    cudaEvent_t done = nullptr;      // recorded on the copy stream behind the batch's last D2H
    bool fetched = false;            // the batch copied its results to the host mirrors
    bool valid = false;
    void release() {
        DevBuf *d[] = {&phrase_id, &phrases, &pat1, &pat2, &q1_off, &q1_ids, &q2_off, &q2_ids, &rules[0], &rules[1], &rules[2], &updown[0], &updown[1], &updown[2],
                       &idinfo[0], &idinfo[1], &idinfo[2]};
        for (auto *x : d) x->release();
        PinnedBuf *h[] = {&h_phrase_id, &h_phrases, &h_pat1, &h_pat2, &h_q1_off, &h_q1_ids, &h_q2_off, &h_q2_ids, &h_rules[0], &h_rules[1], &h_rules[2], &h_updown[0], &h_updown[1], &h_updown[2],
                          &h_idinfo[0], &h_idinfo[1], &h_idinfo[2]};
        for (auto *x : h) x->release();
        if (done) cudaEventDestroy(done);
        done = nullptr;
    }
};

struct Batch {
    int32_t Q = 0, T = 0;
    // inputs
    DevBuf q_tok, q_off, tok2q;
    // lookup
    DevBuf longest, iv;
    // phrases
    DevBuf ph_keys, ph_keys_tmp, ph_vals, ph_vals_tmp, ph_flags, phrase_id, phrases;
    int32_t G = 0;
    // one-gap enumeration
    DevBuf e1_count, e1_inst, e1_keys, e1_keys_tmp, e1_vals, e1_vals_tmp, e1_flags, e1_pid, pat1, pat1_dev, pat1_pos;
    DevBuf ql_keys, ql_keys_tmp, q1_off, q1_ids, q2_off, q2_ids;
    int32_t enu1 = 0, D1 = 0;
    // joins
    DevBuf j_tiles, j_bitmaps, j_aflag, j_aid, j_hash, j_status, j_segcnt, j_flags, pat1_ga, hit_keys, hit_keys_tmp, counters, missing;
    int64_t hits1 = 0, hits2 = 0, j1_elems = 0;
    int pbits = 30;                        // position field width of the packed hit keys (bits needed for n)
    size_t hit_cap = 0;
    bool h1_mask = false;                  // the one-gap hit keys of this batch carry the second-gap word in their top 10 bits (join.cu H1_MASK_SHIFT)
    int32_t adv_refused_q = 0, adv_ok_q = 0;   // cgx_batch_advice: smallest batch refused so far, size and hits of the last finished one
    double adv_ok_hits = 0.0;
    uint32_t j1_buckets = 0, j2_buckets = 0;              // buckets of the last one-gap pattern table (kept when a batch had to grow it)
    bool j1_smem_opt_in = false;          // j1_pos_ordered_kernel's dynamic shared memory opted in on this device
    DevBuf hits1_sorted, hits2_sorted;     // uint64 keys: pattern << (pbits+4) | pos << 4 | len-1  /  pattern << (pbits+8) | pos << 8 | g2 << 4 | L  (g2 = width of the second gap: c at pos+L+1+g2)
    // two-gap enumeration
    DevBuf e2_count, e2_keys, e2_keys_tmp, e2_vals, e2_vals_tmp, e2_flags, pat2;
    int32_t enu2 = 0, D2 = 0;
    // extraction
    DevBuf slot_off[3], slot_hint, rec[3], rec_hash, rec_tag, rec_live, rec_list, rec_flags, rec_meta, rec_cnt;   // rec[k]: slot-indexed cells (extract.cu), empty cells have id = -1
    size_t rec_cells[3] = {0, 0, 0};       // cells per kind
    int64_t n_slots[3] = {0, 0, 0};        // sampled-occurrence slots: contiguous / one-gap / two-gap
    int64_t n_rec[3] = {0, 0, 0};          // non-empty cells per kind
    int64_t samples = 0;
    // rules
    DevBuf rules[3], updown[3], idinfo[3], id_count[3], rule_head, rule_id;
    int32_t n_rules[3] = {0, 0, 0};
    int32_t n_ids[3] = {0, 0, 0};
    // temp
    RadixTemp radix;
    ScanTemp scan;
    DevBuf scratch, scratch2;
    // pinned host staging of the inputs
    int32_t *h_pinned = nullptr;
    size_t h_pinned_cap = 0;
    // host mirrors
    PinnedBuf h_phrase_id, h_phrases, h_pat1, h_pat2, h_q1_off, h_q1_ids, h_q2_off, h_q2_ids;
    PinnedBuf h_rules[3], h_updown[3], h_idinfo[3];
    bool fetch_results = true;       // false: results stay on the device (device-resident throughput measurement)
    cgx_batch_info_t info;
    // the two batches before this one (parked[0] = previous); the current batch's arrays are the named members above
    ResultSet parked[CGX_RESULT_SETS - 1];
    cudaEvent_t done_ev = nullptr;   // current batch: behind its last D2H on the copy stream
    bool valid = false;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // results travel to the (pinned) host mirrors on their own stream, overlapping the kernels that follow their producer
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_ev = nullptr;
    int launches = 0;
};

// D2H of a finished result array: ordered after everything enqueued so far on `stream`, runs on the copy stream
static inline void fetch_async(Batch &b, void *dst, const void *src, size_t bytes, cudaStream_t stream) {
    if (!bytes) return;
    CUDA_CHECK(cudaEventRecord(b.copy_ev, stream));
    CUDA_CHECK(cudaStreamWaitEvent(b.copy_stream, b.copy_ev, 0));
    // In pieces: the D2H engine serves one copy at a time, and the 4-byte count read-backs of the NEXT batch (pipelined
    // callers) would otherwise queue behind a 1.3 GB rule array and stall its kernels for the whole copy.
    static const size_t piece = [] { const char *e = getenv("CGX_D2H_PIECE"); size_t v = e ? strtoull(e, nullptr, 10) : 0; return v ? v : (size_t)4 << 20; }();
    for (size_t o = 0; o < bytes; o += piece)
        CUDA_CHECK(cudaMemcpyAsync((char *)dst + o, (const char *)src + o, bytes - o < piece ? bytes - o : piece, cudaMemcpyDeviceToHost, b.copy_stream));
}

// the same for the first `width` bytes of every `spitch`-byte record (the host needs less of a record than the device keeps)
static inline void fetch_async_2d(Batch &b, void *dst, size_t width, const void *src, size_t spitch, size_t rows, cudaStream_t stream) {
    if (!rows) return;
    CUDA_CHECK(cudaEventRecord(b.copy_ev, stream));
    CUDA_CHECK(cudaStreamWaitEvent(b.copy_stream, b.copy_ev, 0));
    const size_t piece = std::max<size_t>(1, ((size_t)4 << 20) / width);           // rows per piece (~4 MB, see fetch_async)
    for (size_t r = 0; r < rows; r += piece)
        CUDA_CHECK(cudaMemcpy2DAsync((char *)dst + r * width, width, (const char *)src + r * spitch, spitch, width, rows - r < piece ? rows - r : piece,
                                     cudaMemcpyDeviceToHost, b.copy_stream));
}

// current result arrays <-> a parked set (pointer swaps only)
static inline void swap_results(Batch &b, ResultSet &r) {
    std::swap(b.phrase_id, r.phrase_id); std::swap(b.phrases, r.phrases); std::swap(b.pat1, r.pat1); std::swap(b.pat2, r.pat2);
    std::swap(b.q1_off, r.q1_off); std::swap(b.q1_ids, r.q1_ids); std::swap(b.q2_off, r.q2_off); std::swap(b.q2_ids, r.q2_ids);
    std::swap(b.h_phrase_id, r.h_phrase_id); std::swap(b.h_phrases, r.h_phrases); std::swap(b.h_pat1, r.h_pat1); std::swap(b.h_pat2, r.h_pat2);
    std::swap(b.h_q1_off, r.h_q1_off); std::swap(b.h_q1_ids, r.h_q1_ids); std::swap(b.h_q2_off, r.h_q2_off); std::swap(b.h_q2_ids, r.h_q2_ids);
    for (int k = 0; k < 3; k++) {
        std::swap(b.rules[k], r.rules[k]); std::swap(b.updown[k], r.updown[k]); std::swap(b.h_rules[k], r.h_rules[k]); std::swap(b.h_updown[k], r.h_updown[k]);
        std::swap(b.idinfo[k], r.idinfo[k]); std::swap(b.h_idinfo[k], r.h_idinfo[k]);
        std::swap(b.n_rules[k], r.n_rules[k]); std::swap(b.n_ids[k], r.n_ids[k]);
    }
    std::swap(b.Q, r.Q); std::swap(b.T, r.T); std::swap(b.G, r.G); std::swap(b.D1, r.D1); std::swap(b.D2, r.D2);
    std::swap(b.info, r.info); std::swap(b.done_ev, r.done); std::swap(b.fetch_results, r.fetched); std::swap(b.valid, r.valid);
}
// A new batch begins: current -> parked[0] -> parked[1] -> current.  The set that becomes current held the batch three
// begins ago; its copies have long finished (waited for all the same before its arrays are overwritten).
static inline void rotate_results(Batch &b) {
    swap_results(b, b.parked[0]);          // current = old parked[0], parked[0] = the batch that just finished
    swap_results(b, b.parked[1]);          // current = old parked[1], parked[1] = old parked[0]
    if (b.done_ev) CUDA_CHECK(cudaEventSynchronize(b.done_ev));
}

// hits per batch and kind: hit_start / hit_count of the pattern tables are int32.  CGX_HIT_LIMIT lowers it (tests of the
// callers' batch splitting).
static inline unsigned long long hit_limit() {
    unsigned long long lim = (1ull << 31) - 1;
    if (const char *e = getenv("CGX_HIT_LIMIT")) { unsigned long long v = strtoull(e, nullptr, 10); if (v && v < lim) lim = v; }
    return lim;
}

// stages (each enqueues on `stream`; those that need a count on the host synchronise once)
void stage_lookup(const Index &ix, Batch &b, cudaStream_t stream);
void stage_phrases(const Index &ix, Batch &b, cudaStream_t stream);
void stage_onegap_enumerate(const Index &ix, Batch &b, cudaStream_t stream);
void stage_onegap_join(const Index &ix, Batch &b, cudaStream_t stream);
void stage_twogap_enumerate(const Index &ix, Batch &b, cudaStream_t stream);
void stage_twogap_join(const Index &ix, Batch &b, cudaStream_t stream);
void stage_extract(const Index &ix, Batch &b, cudaStream_t stream);
void stage_aggregate(const Index &ix, Batch &b, cudaStream_t stream);

}  // namespace cgx
