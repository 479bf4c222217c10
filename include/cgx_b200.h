/* cgx_b200.h -- C ABI of the B200-native hierarchical grammar extractor (libcgx_b200.so).
 *
 * The reference (hohoCode/cgx) has no plugin / FFI surface: its only stable interface is the
 * `strmatchcuda` executable (Main.c:28-61) and, internally, the C-to-CUDA call sequence of
 * start() (Start.cu:488-629).  This header is that internal boundary re-drawn as a thin C ABI:
 * plain pointers and sizes, opaque handle, no C++ / torch types.  Each entry point names the
 * reference function(s) it replaces.
 *
 * Conventions: every function returns 0 on success, non-zero on failure (cgx_last_error() has the
 * text).  Host pointers unless a name ends in _dev.  One context per GPU; a context is not
 * thread-safe, different contexts are independent.  There is NO CPU fallback: every entry point
 * fails if the CUDA device is unavailable.
 */
#ifndef CGX_B200_H
#define CGX_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cgx_ctx cgx_ctx_t;

/* return codes */
#define CGX_OK 0
#define CGX_E_FAILED 1            /* cgx_last_error() has the text */
#define CGX_E_EXCEPTION 2
#define CGX_E_BATCH_TOO_LARGE 3   /* cgx_extract*: the batch's hit lists / record cells exceed the 31-bit indices of the result
                                     tables (hits grow with corpus size x batch size): split the batch and call again.  The
                                     reference has the same limit without the check (ONEGAP_PREALLOCATION 60 M hits,
                                     ComTypes.h:56; SURVEY.md 8c "operating envelope"). */

/* ---- context --------------------------------------------------------------------------------- */
/* replaces suffixArraySearchInit (SuffixArray.cu:769) / cudaSetDevice */
int cgx_create(int device, cgx_ctx_t **out);
void cgx_destroy(cgx_ctx_t *ctx);                     /* suffixArraySearchFinalize (:807), extractPairFinalize (ExtractPair.cu:2434) */
const char *cgx_last_error(const cgx_ctx_t *ctx);
int cgx_version(void);

/* ---- index (one-time per corpus) ------------------------------------------------------------- */
/* Replaces suffixArrayConstruct (SuffixArray.c:196, CPU DC3 + LCP tables), the H2D copies of
 * suffixArraySearch (SuffixArray.cu:1396-1412) and preComputation (:1132-1340).
 *   src   : n+3 ints, layout of initRefSet (Start.cu:240-380): ids >= 2, EOS = 1, trailer "1, V+2", 0 0 0
 *   tgt   : m+3 ints, layout of initRefTargetSet (Start.cu:142-238)
 *   RLP, L_tar, R_tar : layout of initAlignment (ExtractPair.cu:2639-2739)
 * Builds the suffix array on the GPU (prefix doubling + onesweep radix sort) and the auxiliary
 * position-sorted n-gram occurrence lists. */
int cgx_index_build(cgx_ctx_t *ctx, const int32_t *src, int64_t n, const int32_t *tgt, int64_t m,
                    const uint32_t *RLP, const uint8_t *L_tar, const uint8_t *R_tar);

/* The same for corpora with sentences of 255 tokens and more, which the reference refuses ("Not possible, too long sentence",
 * ExtractPair.cu:2683; uint8_t position counters, Start.cu:269).  The alignment fields are 16 bits wide:
 *   RLP64[i] = L << 48 | R << 32 | P << 16 for a source token (65535 = unaligned), the target sentence offset at an EOS;
 *   L_tar16 / R_tar16 = min / max aligned source index per target token (65535 = unaligned).
 * Same algorithm, same results as cgx_index_build on a corpus that fits 8 bits (lifted limit).
 * CGX_FORCE_WIDE=1 makes cgx_index_build widen its input and take this path (tests). */
int cgx_index_build_wide(cgx_ctx_t *ctx, const int32_t *src, int64_t n, const int32_t *tgt, int64_t m, const uint64_t *RLP64,
                         const uint16_t *L_tar16, const uint16_t *R_tar16);

/* Replaces initWordPossibilityIntKey (ExtractPair.cu:2442-2554): f/e ids (-1 = NULL word), v1 feeds
 * MaxLexEgivenF, v2 feeds MaxLexFgivenE.  Sorted on the GPU by (f, e). */
int cgx_lex_load(cgx_ctx_t *ctx, const int32_t *f, const int32_t *e, const float *v1, const float *v2, int64_t count);

typedef struct {
    int64_t n, m;
    int32_t sa_rounds;        /* prefix-doubling rounds executed */
    int32_t sa_key_bits;      /* sorted key width */
    int32_t sa_launches;
    float sa_build_ms;        /* tokens on device -> sa[] on device (CUDA events) */
    float aux_build_ms;       /* occurrence lists, token buckets, frequent tokens */
    int64_t index_bytes;      /* resident HBM bytes of the index */
} cgx_index_info_t;
int cgx_index_info(const cgx_ctx_t *ctx, cgx_index_info_t *out);

/* Suffix array only, device pointers in/out (bench: SA-build metric with inputs resident in HBM).
 * str_dev: n+3 ints; sa_dev: n ints. */
int cgx_sa_build_dev(cgx_ctx_t *ctx, const int32_t *str_dev, int64_t n, int32_t max_token, int32_t *sa_dev,
                     int32_t *rounds_out, float *ms_out);

/* Device-resident index arrays, for broadcasting a built index to peer GPUs (NCCL) and adopting it there. */
typedef struct {
    int64_t n, m, lex_count;
    int32_t max_token;
    int32_t freq_list[100];
    int32_t wide;             /* 1: 16-bit alignment fields (cgx_index_build_wide): RLP is 8 bytes per token, L_tar / R_tar 2 bytes */
    void *str, *sa, *inv1, *inv2, *inv3, *bkt1, *bkt2, *bkt3, *tok_start, *RLP, *L_tar, *R_tar, *tgt, *freq_flag, *gapw, *lex_key, *lex_v1, *lex_v2;
} cgx_index_arrays_t;
int cgx_index_export(cgx_ctx_t *ctx, cgx_index_arrays_t *out);            /* pointers stay owned by ctx */
int cgx_index_alloc(cgx_ctx_t *ctx, const cgx_index_arrays_t *shape, cgx_index_arrays_t *out);  /* allocate empty arrays of that shape on this ctx */
int cgx_index_commit(cgx_ctx_t *ctx);                                      /* mark the (filled) arrays as a built index */

/* Single-process multi-GPU: broadcast the index built on ctxs[0] to ctxs[1..n-1] with ncclBroadcast over
 * NVLink / NVSwitch (one communicator per device, ncclCommInitAll).  New relative to the reference, which
 * is single-GPU; queries are then sharded across the contexts by the caller (no steady-state collective).
 * NCCL is dlopen()ed at call time.  Multi-process callers (torchrun) broadcast the exported arrays with
 * their own communicator instead (cgx_index_export / cgx_index_alloc / cgx_index_commit). */
int cgx_index_broadcast(cgx_ctx_t **ctxs, int n);

/* Persisted index: the index inputs as the loaders laid them out (the two token arrays, the alignment arrays, the sorted lexical
 * table: 18 bytes per token pair) written to / read from one file, so that later runs on the same corpus skip parsing the
 * alignment and lexical files.  A load rebuilds the suffix array and the auxiliary arrays on the device (30 ms at 26 M tokens:
 * cheaper than reading them back, which is what the first form of this file did).  The reference only has a dead stub of this
 * (SuffixArray.c:208-230 "sa_precomp.txt"; README.md:85 promises separating the one-time costs).  cgx_index_load replaces
 * cgx_index_build + cgx_lex_load. */
int cgx_index_save(cgx_ctx_t *ctx, const char *path);
int cgx_index_load(cgx_ctx_t *ctx, const char *path);
/* 1 when the resident index was built from exactly these token arrays (same lengths, same FNV-1a checksums, which the index
 * file records); 0 otherwise.  strmatchcuda -i refuses an index file that belongs to another corpus of the same length. */
int cgx_index_matches(const cgx_ctx_t *ctx, const int32_t *src, int64_t n, const int32_t *tgt, int64_t m);

/* parity helpers: copy index arrays to the host */
int cgx_index_copy_sa(cgx_ctx_t *ctx, int32_t *sa_out);                    /* n ints */
int cgx_index_copy_inv(cgx_ctx_t *ctx, int which, int32_t *out);           /* which = 1..3, n ints */
int cgx_index_copy_frequent(cgx_ctx_t *ctx, int32_t *out100);

/* ---- one query batch: match + extract + score ------------------------------------------------- */
/* Replaces suffixArraySearch (SuffixArray.cu:1342-2269) + ExtractPairs_Large_Data_Gappy
 * (ExtractPair.cu:3215-4001) incl. the host aggregation createLexicon*Fast (ExtractPair.c:515-1276)
 * and lexicalTaskMaxEF (ExtractPair.cu:2144).
 *   qry_tok : T ids in the source vocabulary, -1 = OOV (constructQryIndex, Start.cu:50-132)
 *   qry_off : Q+1 offsets into qry_tok
 * Results stay in the context until the next cgx_extract / cgx_destroy. */
int cgx_extract(cgx_ctx_t *ctx, const int32_t *qry_tok, const int32_t *qry_off, int32_t Q);

/* Pipelined form of cgx_extract for a stream of batches (new relative to the reference, which runs one query file per
 * process, Start.cu:558): returns as soon as the batch's kernels have finished, while the tail of its result copy is
 * still travelling to the host; that copy then overlaps the kernels of the next cgx_extract_begin.  The context keeps
 * the results of the last three batches (device arrays and pinned host mirrors rotate), so the caller can still be
 * writing batch i-2 while batch i-1 travels and batch i computes:
 *     cgx_extract_begin(ctx, batch[i]);  cgx_result_at(ctx, 1, &res) -> batch[i-1] complete on the host
 * cgx_result_at(ctx, age, ..) blocks until that batch's copies are done; age 0 = the most recent batch (= cgx_result),
 * 1 = the one before, 2 = two before.  Views of a batch stay valid until the third cgx_extract_begin after its own. */
int cgx_extract_begin(cgx_ctx_t *ctx, const int32_t *qry_tok, const int32_t *qry_off, int32_t Q);

/* Same pipeline with the queries already resident in HBM and the results left there (device pointers:
 * qry_tok_dev T ints, qry_off_dev Q+1 ints, tok2q_dev T ints = query index of every token).  Used to
 * measure the device-resident throughput; cgx_result() is not available after this call. */
int cgx_extract_dev(cgx_ctx_t *ctx, const int32_t *qry_tok_dev, const int32_t *qry_off_dev, const int32_t *tok2q_dev, int32_t Q, int32_t T);

/* Per-kernel CUDA-event timing on the launching stream (off by default).  cgx_profile_report returns a JSON
 * object {kernel: {launches, ms, bytes}} accumulated since cgx_profile_enable(ctx, 1); `bytes` are the
 * algorithmic bytes of DESIGN.md. */
int cgx_profile_enable(cgx_ctx_t *ctx, int on);
const char *cgx_profile_report(cgx_ctx_t *ctx);

typedef struct {
    int32_t Q, T;
    int32_t G;          /* distinct contiguous phrases (GenerateBlocks, ExtractPair.cu:2742) */
    int32_t enu1, D1;   /* one-gap enumeration count / distinct patterns */
    int64_t hits1;      /* one-gap corpus occurrences (oneGapLookUpSA) */
    int32_t enu2, D2;
    int64_t hits2;
    int64_t samples;    /* sampled occurrences extracted */
    int64_t n_ab, n_1gap, n_2gap;      /* rule records per array */
    int32_t rules[3];   /* distinct rules: [0] ab, [1] Xab|abX|aXb, [2] XabX|aXbXc|XaXb|aXbX */
    int32_t launches;   /* kernels launched for this batch */
    float ms_total, ms_lookup, ms_enum, ms_join, ms_extract, ms_aggregate;   /* CUDA-event stage times */
} cgx_batch_info_t;
int cgx_batch_info(const cgx_ctx_t *ctx, cgx_batch_info_t *out);

/* Queries to put into the next batch of a stream, at most `wanted`.  The hit lists of a batch grow with corpus size x batch
 * size and are indexed with 31 bits; a batch beyond that is refused (CGX_E_BATCH_TOO_LARGE) only after its join scan has run
 * and the caller halves it.  This remembers the smallest size refused so far and the hits of the last finished batch, so
 * that a stream pays for a refusal once, not once per batch.  (The reference preallocates 60 M hits, ComTypes.h:56, and fails
 * beyond; its callers run ~5 k queries per process, README.md:76-79.) */
int32_t cgx_batch_advice(const cgx_ctx_t *ctx, int32_t wanted);

/* A distinct scored rule (red_dup_t, ComTypes.h:244-255), packed into 16 bytes: a C2 batch returns 7e7 of them and the
 * device-to-host copy of the rules is most of what a batch sends back (8 ranks on one host share its ingest bandwidth).
 * The converted id of a rule (ExtractPair.c:723-729 / :999-1006) is the range it sits in -- rules come in ascending id order,
 * `first[id]` is the id's first rule and its idinfo word says how many follow -- and the two per-id counts (f,
 * all_suffix_fsample) travel once per id in `idinfo` instead of once per rule. */
typedef struct {
    int32_t tgt_start;     /* representative target span: start in the target text */
    uint32_t span;         /* bits 0-3 end (inclusive length-1); 4-7 gap1, 8-11 gap1_1, 12-15 gap2, 16-19 gap2_1: target gaps as offsets
                              from tgt_start, 15 = no such gap; bits 20-28 paircount (<= 300) -> CountEF, EgivenFCoherent, IsSingletonFE */
    float max_lex_f_given_e, max_lex_e_given_f;
} cgx_rule_t;
#define CGX_RULE_NOGAP 15
#define CGX_RULE_END(r) ((int)((r)->span & 15u))
#define CGX_RULE_GAP1(r) ((int)(((r)->span >> 4) & 15u))
#define CGX_RULE_GAP1_END(r) ((int)(((r)->span >> 8) & 15u))
#define CGX_RULE_GAP2(r) ((int)(((r)->span >> 12) & 15u))
#define CGX_RULE_GAP2_END(r) ((int)(((r)->span >> 16) & 15u))
#define CGX_RULE_PC(r) ((int)(((r)->span >> 20) & 511u))
/* idinfo word of a converted id: extracted pairs with this source id (-> IsSingletonF), all_suffix_fsample capped at 300
 * (-> SampleCountF) and the number of distinct rules of the id (each <= 300: nine bits); 0 for an id without rules */
#define CGX_ID_F(w) ((int)((w) & 0x1ffu))
#define CGX_ID_FS(w) ((int)(((w) >> 9) & 0x1ffu))
#define CGX_ID_RULES(w) ((int)(((w) >> 18) & 0x1ffu))

/* Host views of the batch results (valid until the next cgx_extract; see cgx_extract_begin for the pipelined form):
 *   phrase_id : T*5 ints, id of the contiguous phrase q[t..t+len) for len = 1..5, -1 when absent
 *   phrases   : G x {sa_up, sa_down, len, corpus_pos}      (saind_t, ComTypes.h:342)
 *   pat1      : D1 x {a_pos, ls, b_pos, le}      (what the printer needs of gappy_search; hit ranges, marker and
 *                                                 featureMissingCount stay on the device: cgx_debug_fetch "pat1_full")
 *   pat2      : D2 x {pat1_id, c_token}          (cgx_debug_fetch "pat2_full" adds the hit range)
 *   q1_off/q1_ids, q2_off/q2_ids : per-query lists of one-gap / two-gap pattern ids (ascending id)
 *   rules[k], n_rules[k], and per converted id the index of its first rule (first, -1 when it has none; the rules of id are
 *   first[id] .. first[id] + CGX_ID_RULES(idinfo[id]) - 1 -- globalOnPairsUpDown*, ExtractPair.cu:3745-3756) and its idinfo word */
typedef struct {
    int32_t Q, T, G, D1, D2;
    const int32_t *phrase_id;
    const int32_t *phrases;
    const int32_t *pat1;
    const int32_t *pat2;
    const int32_t *q1_off, *q1_ids, *q2_off, *q2_ids;
    const cgx_rule_t *rules[3];
    int32_t n_rules[3];
    const int32_t *first[3];
    int32_t n_ids[3];
    const uint32_t *idinfo[3];
} cgx_result_t;
int cgx_result(cgx_ctx_t *ctx, cgx_result_t *out);
int cgx_result_at(cgx_ctx_t *ctx, int age, cgx_result_t *out);      /* see cgx_extract_begin */

/* parity helpers (tests): intermediate arrays of the last batch, copied to the host.
 *   what = "longest" (T ints, capped at 5), "intervals" (T*5*2 ints), "hits1" (hits1 x 3: id,pos,len),
 *          "hits2" (hits2 x 4), "rec_ab"/"rec_1"/"rec_2" (7 ints per record, converted ids), "pat1_full" (D1 x 8), "pat2_full" (D2 x 4)
 * Returns the number of int32 written (<= cap) or a negative error. */
int64_t cgx_debug_fetch(cgx_ctx_t *ctx, const char *what, int32_t *out, int64_t cap);

/* The onesweep radix sort on its own (tests / tools/sort_bench.py): sorts n device-resident 64-bit keys on bits
 * [begin_bit, end_bit), optionally carrying 32-bit payloads (vals_dev may be NULL).  keys_dev/vals_dev hold the sorted
 * data on return; ms_out = CUDA-event time of the sort, passes_out = onesweep passes executed. */
int cgx_debug_sort_u64(cgx_ctx_t *ctx, uint64_t *keys_dev, uint32_t *vals_dev, int64_t n, int begin_bit, int end_bit, float *ms_out, int *passes_out);

#ifdef __cplusplus
}
#endif
#endif
