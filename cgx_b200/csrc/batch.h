// cgx-b200: per-batch device workspace and host-side result mirrors.
#pragma once
#include "../../include/cgx_b200.h"
#include "index.h"

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
    DevBuf j_tiles, j_bitmaps, j_aflag, j_aid, j_hash, pat1_ga, hit_keys, hit_keys_tmp, counters, missing;
    int64_t hits1 = 0, hits2 = 0, j1_elems = 0;
    int pbits = 30;                        // position field width of the packed hit keys (bits needed for n)
    size_t hit_cap = 0;
    DevBuf hits1_sorted, hits2_sorted;     // uint64 keys: pattern << (pbits+4) | pos << 4 | len-1  /  pattern << (pbits+8) | pos << 8 | L << 4 | c-pos
    // two-gap enumeration
    DevBuf e2_count, e2_keys, e2_keys_tmp, e2_vals, e2_vals_tmp, e2_flags, pat2;
    int32_t enu2 = 0, D2 = 0;
    // extraction
    DevBuf slot_off[3], rec[3], rec_hash, rec_tag, rec_live, rec_flags, rec_meta;   // rec[k]: slot-indexed cells (extract.cu), empty cells have id = -1
    size_t rec_cells[3] = {0, 0, 0};       // cells per kind
    int64_t n_slots[3] = {0, 0, 0};        // sampled-occurrence slots: contiguous / one-gap / two-gap
    int64_t n_rec[3] = {0, 0, 0};          // non-empty cells per kind
    int64_t samples = 0;
    // rules
    DevBuf rules[3], updown[3], id_count[3], rule_head;
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
    PinnedBuf h_rules[3], h_updown[3];
    bool fetch_results = true;       // false: results stay on the device (device-resident throughput measurement)
    cgx_batch_info_t info;
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
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, b.copy_stream));
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
