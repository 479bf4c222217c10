/* cgx_host.h -- C host side of the B200-native grammar extractor: text loaders, grammar writer and the
 * driver behind the `strmatchcuda` command line.  Plain C; calls the GPU only through include/cgx_b200.h.
 *
 * Mirrors the reference's host interface for this path:
 *   cgxh_corpus_load      <- initRefSet / initRefTargetSet   (Start.cu:240-380, :142-238)
 *   cgxh_alignment_load   <- initAlignment                   (ExtractPair.cu:2639-2739)
 *   cgxh_lex_load         <- initWordPossibilityIntKey (parse) (ExtractPair.cu:2442-2519)
 *   cgxh_queries_load     <- constructQryIndex               (Start.cu:50-132)
 *   cgxh_write_grammars   <- print_query_GPU_Gappy           (PrintResults.c:339-577)
 *   cgxh_run              <- start                           (Start.cu:488-629)
 */
#ifndef CGX_HOST_H
#define CGX_HOST_H
#include <stdint.h>
#include "../../include/cgx_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cgxh_vocab cgxh_vocab_t;

typedef struct {
    int32_t *tok;        /* n+3 ints: ids >= 2, EOS = 1 after each sentence, trailer "1, last", 0 0 0 */
    int64_t n;
    uint8_t *P;          /* n bytes: position in sentence (source side only, else NULL) */
    int32_t *sentenceind;/* n_sent+1 */
    int32_t n_sent;
    cgxh_vocab_t *vocab; /* name <-> id */
    int32_t last;        /* the unique final symbol */
} cgxh_side_t;

typedef struct {
    uint32_t *RLP;       /* n */
    uint8_t *L_tar, *R_tar; /* m */
    int wide;            /* 1: the corpus has a sentence the 8-bit fields cannot hold; the three arrays below are set instead */
    uint64_t *RLP64;     /* n: L << 48 | R << 32 | P << 16 (65535 = unaligned), target sentence offset at an EOS */
    uint16_t *L_tar16, *R_tar16; /* m */
} cgxh_align_t;

typedef struct {
    int32_t *f, *e;
    float *v1, *v2;
    int64_t count;
} cgxh_lex_t;

typedef struct {
    int32_t *tok;        /* T ids, -1 = OOV */
    int32_t *off;        /* Q+1 */
    int32_t Q, T;
    int32_t max_len;
} cgxh_queries_t;

int cgxh_corpus_load(const char *path, int want_P, cgxh_side_t *out);
void cgxh_side_free(cgxh_side_t *s);
const char *cgxh_vocab_name(const cgxh_vocab_t *v, int32_t id);
int32_t cgxh_vocab_id(const cgxh_vocab_t *v, const char *name);   /* -1 when absent */
int32_t cgxh_vocab_size(const cgxh_vocab_t *v);                   /* HASH_COUNT + 2 */

int cgxh_alignment_load(const char *path, const cgxh_side_t *src, const cgxh_side_t *tgt, cgxh_align_t *out);
void cgxh_align_free(cgxh_align_t *a);
int cgxh_lex_load(const char *path, const cgxh_side_t *src, const cgxh_side_t *tgt, cgxh_lex_t *out);
void cgxh_lex_free(cgxh_lex_t *l);
/* the four loaders above over all the files of a corpus, two threads at a time (source || target, then alignment || lexical file);
 * align_path / lex_path NULL = not parsed (*al / *lex zeroed).  Returns the first loader's non-zero code. */
int cgxh_load_files(const char *src_path, const char *tgt_path, const char *align_path, const char *lex_path, cgxh_side_t *src,
                    cgxh_side_t *tgt, cgxh_align_t *al, cgxh_lex_t *lex);
int cgxh_queries_load(const char *path, const cgxh_side_t *src, cgxh_queries_t *out);
void cgxh_queries_free(cgxh_queries_t *q);

/* Writes <outdir>/grammar.<qid_base+q>.s for the queries of one batch (PrintResults.c:434-574).
 * n_threads <= 1: single thread like the reference; > 1: queries are split across POSIX threads. */
int cgxh_write_grammars(const char *outdir, const cgx_result_t *res, const int32_t *qry_off, int32_t qid_base, const cgxh_side_t *src,
                        const cgxh_side_t *tgt, int n_threads);

/* the same with gzip_level 1..9: <outdir>/grammar.<qid>.s.gz, the form cdec reads its per-sentence grammars in (zlib; 0 = plain text) */
int cgxh_write_grammars_ex(const char *outdir, const cgx_result_t *res, const int32_t *qry_off, int32_t qid_base, const cgxh_side_t *src,
                           const cgxh_side_t *tgt, int n_threads, int gzip_level);

/* the writer's printf("%f") replacement for float-valued features (exactly glibc's digits; writer.c); returns the length */
int cgxh_format_f6(float x, char *out);

#define CGXH_DEFAULT_BATCH 10000
typedef struct {
    const char *reffile, *qryfile, *reftargetfile, *align, *wordscdec, *destinationDirectory;   /* options_t, ComTypes.h:67-78 */
    const char *timefile;
    int minmatchlen, fingerlen;
    int n_gpus;          /* extension: queries sharded over this many GPUs of the box (default 1) */
    int batch_queries;   /* extension: queries per GPU batch (0 = even split over the GPUs, at most CGXH_DEFAULT_BATCH) */
    int writer_threads;
    int quiet;
    const char *index_file;   /* extension: persisted GPU index (cgx_index_load when it exists, else build + cgx_index_save) */
    int serve;           /* extension (-S): after the command line's query file, serve "<query file> <output dir>" requests from stdin
                            against the resident index until EOF or "quit" */
    int gzip_level;      /* extension (-z [level]): grammar.<qid>.s.gz instead of plain text */
} cgxh_options_t;
int cgxh_run(const cgxh_options_t *opt);

#ifdef __cplusplus
}
#endif
#endif
