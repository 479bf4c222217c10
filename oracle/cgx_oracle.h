/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the hohoCode/cgx grammar-extraction path.
 *
 * Nothing in the product (cgx_b200/, include/, the strmatchcuda host) may include, link or call
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, as the checker.
 *
 * Parity status: the reference repository owns no tests, golden vectors or fixtures (SURVEY.md
 * section 4).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF: the
 * unmodified reference binary (oracle/_ref/strmatchcuda, built by oracle/build_ref.sh) run on a
 * B200 on synthetic corpora; fixtures generated that way are committed under tests/golden/
 * together with the script that made them (tools/make_golden.py), and the -m gpu tests re-run the
 * reference binary live next to the product.
 *
 * Every function cites the reference file:line it restates.
 */
#ifndef CGX_ORACLE_H
#define CGX_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_RULE_SPAN 15          /* ComTypes.h:42 MAX_rule_span / :43 MAX_rule_span_pattern */
#define ORC_MAX_RULE_SYMBOLS 5        /* ComTypes.h:44 */
#define ORC_MIN_GAP 1                 /* ComTypes.h:47 */
#define ORC_MAXSCORE 99               /* ComTypes.h:51 */
#define ORC_PRECOMP 100               /* ComTypes.h:55 PRECOMPUTECOUNT */
#define ORC_SAMPLER 300               /* ComTypes.h:63 */
#define ORC_SAMPLER_ONEGAP 65         /* ComTypes.h:64 */
#define ORC_SAMPLER_TWOGAP 70         /* ComTypes.h:65 */
#define ORC_LONGEST_SRC 5             /* ExtractPair.cu:16 LONGESTCHSOURCE */
#define ORC_THREADS 512               /* ExtractPair.cu:9 THREADS_PER_BLOCK (matters for the early-return quirks) */

typedef struct orc_s orc_t;
/* alignment fields: 8 bits like the reference, or 16 bits with -DORC_WIDE (libcgx_oracle_wide.so; see cgx_oracle.c) */
#ifdef ORC_WIDE
typedef uint16_t lr_t;
typedef uint64_t rlp_t;
#else
typedef uint8_t lr_t;
typedef uint32_t rlp_t;
#endif
int orc_is_wide(void);

/* ---- construction ------------------------------------------------------------------------- */
/* From the integer layouts of the reference loaders (Start.cu:240-380, ExtractPair.cu:2639-2739).
 * str: n+3 ints (3 trailing zeros), tgt: m+3 ints.  All arrays are copied. */
orc_t *orc_create(const int32_t *str, int32_t n, const int32_t *tgt, int32_t m,
                  const rlp_t *RLP, const lr_t *L_tar, const lr_t *R_tar,
                  const int32_t *lex_f, const int32_t *lex_e, const float *lex_v1, const float *lex_v2,
                  int32_t lex_count);
/* From the six strmatchcuda text inputs (Main.c:55-60).  NULL on failure. */
orc_t *orc_create_from_files(const char *src, const char *tgt, const char *align, const char *lex);
void orc_destroy(orc_t *o);

/* Use an externally computed suffix array (e.g. oracle/_ref/libref_sa.so) instead of the built-in one. */
void orc_set_sa(orc_t *o, const int32_t *sa);
/* Built-in CPU suffix array (prefix doubling; same unique answer as SuffixArray.c:51 suffixArrayInt). */
void orc_build_sa(orc_t *o);

/* ---- queries ------------------------------------------------------------------------------ */
/* qry_tok: T ids (-1 = OOV, Start.cu:97); qry_off: Q+1 offsets.  Runs match + extract + score.
 * Returns 0 on success. */
int orc_run(orc_t *o, const int32_t *qry_tok, const int32_t *qry_off, int32_t Q);
int orc_run_query_file(orc_t *o, const char *qry_path);
/* Write grammar.<qid>.s files (PrintResults.c:407-577). */
int orc_write_grammars(orc_t *o, const char *outdir);

/* ---- introspection (for parity tests) ----------------------------------------------------- */
typedef struct {
    int32_t n, m, Q, T;
    int32_t G;            /* distinct contiguous phrases (GenerateBlocks, ExtractPair.cu:2742) */
    int32_t enu1;         /* one-gap enumeration count (SuffixArray.cu:1586) */
    int32_t D1;           /* distinct one-gap patterns (:1720) */
    int32_t hits1;        /* countOneGapSA (:1817), marker records included */
    int32_t enu2, D2, hits2;
    int32_t precomp_count;/* "Found %u pairs!!" (:1284) */
    int32_t n_ab, n_1gap_contig, n_2gap_contig;   /* ExtractPair.cu:3401-3407 */
    int32_t n_axbxc;      /* :3522 */
    int32_t n_axb, n_2gap_from1;                  /* :3635,3637 */
    int32_t lex_1gap, lex_2gap, lex_ab;           /* :3739,3798,3874 distinct rules */
} orc_counts_t;
void orc_get_counts(const orc_t *o, orc_counts_t *c);

const int32_t *orc_sa(const orc_t *o);
/* per query token: longest match (uncapped), and for m=1..min(longest,cap) the SA interval */
const int32_t *orc_longest(const orc_t *o);
/* returns 1 and fills up/down if m <= longest[t] */
int orc_interval(const orc_t *o, int32_t t, int32_t m, int32_t *up, int32_t *down);

/* distinct contiguous phrases: G x {start,end,matchlen,string_start} (saind_t, ComTypes.h:342) */
const int32_t *orc_blocks(const orc_t *o);
/* one-gap patterns: D1 x {tok[5] (-1 gap, -2 pad), number, ls, le, start_on_salist, end_on_salist} = 10 ints */
const int32_t *orc_onegap_patterns(const orc_t *o);
/* one-gap hits sorted by (id,pos,len): hits1 x {id, pos, len} */
const int32_t *orc_onegap_hits(const orc_t *o);
/* two-gap patterns: D2 x {blockid, ctok, start_on_salist, end_on_salist} */
const int32_t *orc_twogap_patterns(const orc_t *o);
/* two-gap hits sorted: hits2 x {id, pos, len, len2} */
const int32_t *orc_twogap_hits(const orc_t *o);
/* top-100 frequent tokens (ascending id) and featureMissingCount[100*100] */
const int32_t *orc_frequent(const orc_t *o);
const int32_t *orc_feature_missing(const orc_t *o);
/* Per-query id lists in the order the writer walks them (PrintResults.c:451-570): which = 0 distinct contiguous phrases of query qi
 * in first-appearance order (GenerateBlocks), 1 its one-gap patterns, 2 its two-gap patterns.  Returns the count. */
int32_t orc_query_list(const orc_t *o, int which, int32_t qi, const int32_t **out);

/* precomputed pair lists: index[100*100] x {start,end} and list[precomp_count] x {start,len} */
const int32_t *orc_precomp_index(const orc_t *o);
const int32_t *orc_precomp_list(const orc_t *o);

/* Extracted rule records, sorted by (kind, id, then emission order).  7 ints each:
 *   {id, tgt_start, end, gap1, gap1_1, gap2, gap2_1}; unused gap fields are -1.
 * kind 0 = ab (res_phrase_t), 1 = one-gap array (Xab | abX | aXb with the reference's separators),
 * 2 = two-gap array (XabX | aXbXc | XaXb,aXbX).  ids are the *converted* ids of ExtractPair.c:723,999. */
int32_t orc_records(const orc_t *o, int kind, const int32_t **out);

/* Distinct rules (red_dup_t, ComTypes.h:244).  For kind in {0,1,2}: count and arrays. */
typedef struct {
    int32_t id;            /* converted id */
    int32_t rec[6];        /* representative record: tgt_start,end,gap1,gap1_1,gap2,gap2_1 */
    int32_t f;             /* records with this id */
    int32_t fs;            /* all_suffix_fsample (capped at 300) */
    int32_t pc;            /* paircount */
    float aa, score, bb;   /* EgivenFCoherent, SampleCountF, CountEF */
    float mlfe, mlef;      /* MaxLexFgivenE, MaxLexEgivenF */
} orc_rule_t;
int32_t orc_rules(const orc_t *o, int kind, const orc_rule_t **out);

#ifdef __cplusplus
}
#endif
#endif
