/* cgx-b200 host: cdec-style per-query grammar files, byte-compatible with the reference's line format
 * and rule-group order (PrintResults.c:339-577; strings of ExtractPair.c:743-796, :1021-1123, :1141-1163;
 * host features of ExtractPair.c:641-655). */
#define _GNU_SOURCE
#include "cgx_host.h"
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct {
    const char *outdir;
    const cgx_result_t *res;
    const int32_t *qry_off;
    int32_t qid_base;
    const cgxh_side_t *src, *tgt;
    int32_t q_begin, q_end;
    int gzip_level;
    int rc;
} wjob_t;

typedef struct { char *p; size_t n, cap; } sbuf;
static void sb_reserve(sbuf *b, size_t extra) {
    if (b->n + extra + 1 > b->cap) {
        while (b->n + extra + 1 > b->cap) b->cap = b->cap ? b->cap * 2 : 1 << 20;
        b->p = (char *)realloc(b->p, b->cap);
    }
}
static void sb_puts(sbuf *b, const char *s) {
    size_t l = strlen(s);
    sb_reserve(b, l);
    memcpy(b->p + b->n, s, l);
    b->n += l;
}
static void sb_putc(sbuf *b, char c) { sb_reserve(b, 1); b->p[b->n++] = c; }
static void sb_putn(sbuf *b, const char *s, size_t l) {
    sb_reserve(b, l);
    memcpy(b->p + b->n, s, l);
    b->n += l;
}

/* printf("%f") of a FLOAT-valued feature without printf (one snprintf with five %f conversions per rule was ~1 us = the whole
 * cost of the writer: ~1 M rules/s/thread against 7e7 rules of a C2 batch).  Exactly glibc's digits: a float has a 24-bit
 * significand and 10^6 = 2^6 * 15625 with 15625 < 2^14, so (double)x * 1e6 is EXACT (<= 38 significant bits), and rint() rounds
 * it to nearest-even the way printf rounds the exact decimal expansion.  Sign as printf prints it (also for -0.0 and for
 * negatives that round to zero).  Values >= 2^63 / 1e6, infinities and NaNs go through snprintf. */
static char *put_f6(char *o, float xf) {
    double x = (double)xf;
    if (!(fabs(x) < 9.0e12)) { o += sprintf(o, "%f", x); return o; }
    if (signbit(x)) { *o++ = '-'; x = -x; }
    const uint64_t v = (uint64_t)rint(x * 1e6);
    uint64_t ip = v / 1000000u;
    uint32_t fp = (uint32_t)(v % 1000000u);
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + ip % 10); ip /= 10; } while (ip);
    while (n) *o++ = tmp[--n];
    *o++ = '.';
    for (int d = 5; d >= 0; d--) { o[d] = (char)('0' + fp % 10); fp /= 10; }
    return o + 6;
}
/* test hook: the writer's "%f" (tests compare it with printf over random and edge-case floats) */
int cgxh_format_f6(float x, char *out) { char *e = put_f6(out, x); *e = 0; return (int)(e - out); }
static char *put_lit(char *o, const char *s, size_t l) { memcpy(o, s, l); return o + l; }
#define LIT(o, s) put_lit(o, s, sizeof(s) - 1)

static void put_phrase(sbuf *b, const cgxh_side_t *src, int32_t pos, int32_t len) {
    for (int i = 0; i < len; i++) {
        if (i) sb_putc(b, ' ');
        sb_puts(b, cgxh_vocab_name(src->vocab, src->tok[pos + i]));
    }
}
static void put_pat1(sbuf *b, const cgxh_side_t *src, const int32_t *p, const char *gap) {
    put_phrase(b, src, p[0], p[1]);
    sb_putc(b, ' '); sb_puts(b, gap); sb_putc(b, ' ');
    put_phrase(b, src, p[2], p[3]);
}

/* source side of a converted id */
static void source_string(sbuf *b, const cgx_result_t *r, const cgxh_side_t *src, int kind, int32_t cid) {
    const int32_t G = r->G, D1 = r->D1, D2 = r->D2;
    if (kind == 0) { put_phrase(b, src, r->phrases[4 * cid + 3], r->phrases[4 * cid + 2]); return; }
    if (kind == 1) {
        if (cid < G) { sb_puts(b, "[X,1] "); put_phrase(b, src, r->phrases[4 * cid + 3], r->phrases[4 * cid + 2]); }
        else if (cid < 2 * G) { int32_t g = cid - G; put_phrase(b, src, r->phrases[4 * g + 3], r->phrases[4 * g + 2]); sb_puts(b, " [X,1]"); }
        else put_pat1(b, src, &r->pat1[4 * (cid - 2 * G)], "[X,1]");
        return;
    }
    if (cid < G) { sb_puts(b, "[X,1] "); put_phrase(b, src, r->phrases[4 * cid + 3], r->phrases[4 * cid + 2]); sb_puts(b, " [X,2]"); }
    else if (cid < G + D2) {
        const int32_t *p2 = &r->pat2[2 * (cid - G)];
        put_pat1(b, src, &r->pat1[4 * p2[0]], "[X,1]");
        sb_puts(b, " [X,2] "); sb_puts(b, cgxh_vocab_name(src->vocab, p2[1]));
    } else if (cid < G + D2 + D1) { sb_puts(b, "[X,1] "); put_pat1(b, src, &r->pat1[4 * (cid - G - D2)], "[X,2]"); }
    else { put_pat1(b, src, &r->pat1[4 * (cid - G - D2 - D1)], "[X,1]"); sb_puts(b, " [X,2]"); }
}

static void put_group(sbuf *b, const cgx_result_t *r, const cgxh_side_t *src, const cgxh_side_t *tgt, int kind, int32_t cid, sbuf *srcbuf) {
    if (cid < 0 || cid >= r->n_ids[kind]) return;
    const int32_t lo = r->first[kind][cid];
    if (lo < 0) return;
    const int32_t hi = lo + CGX_ID_RULES(r->idinfo[kind][cid]) - 1;
    srcbuf->n = 0;
    source_string(srcbuf, r, src, kind, cid);
    sb_reserve(srcbuf, 1);
    srcbuf->p[srcbuf->n] = 0;
    const uint32_t idw = r->idinfo[kind][cid];
    const int f = CGX_ID_F(idw), fs = CGX_ID_FS(idw);
    const size_t srclen = srcbuf->n;
    /* features that depend on the id only (ExtractPair.c:641): rendered once per group */
    char idfeat[64], *e = idfeat;
    e = LIT(e, " SampleCountF=");
    e = put_f6(e, (float)log10((double)(1 + fs)));
    const size_t idfeat_len = (size_t)(e - idfeat);
    int last_pc = -1;
    char pcfeat[2][48];                  /* EgivenFCoherent / CountEF of the previous paircount (runs of equal counts are common) */
    size_t pcfeat_len[2] = {0, 0};
    for (int32_t i = lo; i <= hi; i++) {
        const cgx_rule_t *u = &r->rules[kind][i];
        const int end = CGX_RULE_END(u), g1 = CGX_RULE_GAP1(u), g1e = CGX_RULE_GAP1_END(u), g2 = CGX_RULE_GAP2(u), g2e = CGX_RULE_GAP2_END(u);
        const int pc = CGX_RULE_PC(u);
        sb_putn(b, "[X] ||| ", 8); sb_putn(b, srcbuf->p, srclen); sb_putn(b, " ||| ", 5);
        int first = 1;
        for (int j = 0; j <= end; j++) {
            if (!first) sb_putc(b, ' ');
            first = 0;
            if (g1 != CGX_RULE_NOGAP && j >= g1 && j <= g1e) { sb_putn(b, "[X,1]", 5); j = g1e; }
            else if (g2 != CGX_RULE_NOGAP && j >= g2 && j <= g2e) { sb_putn(b, "[X,2]", 5); j = g2e; }
            else sb_puts(b, cgxh_vocab_name(tgt->vocab, tgt->tok[u->tgt_start + j]));
        }
        if (pc != last_pc) {
            /* ExtractPair.c:653-655: float log10 of the ratio, double log10 of the count */
            char *q = pcfeat[0];
            q = LIT(q, " ||| EgivenFCoherent=");
            q = put_f6(q, -log10f((float)pc / (float)fs));
            pcfeat_len[0] = (size_t)(q - pcfeat[0]);
            q = pcfeat[1];
            q = LIT(q, " CountEF=");
            q = put_f6(q, (float)log10((double)(1 + pc)));
            pcfeat_len[1] = (size_t)(q - pcfeat[1]);
            last_pc = pc;
        }
        sb_reserve(b, 256);
        char *o = b->p + b->n;
        o = put_lit(o, pcfeat[0], pcfeat_len[0]);
        o = put_lit(o, idfeat, idfeat_len);
        o = put_lit(o, pcfeat[1], pcfeat_len[1]);
        o = LIT(o, " MaxLexFgivenE=");
        o = put_f6(o, u->max_lex_f_given_e);
        o = LIT(o, " MaxLexEgivenF=");
        o = put_f6(o, u->max_lex_e_given_f);
        o = LIT(o, " IsSingletonF=");
        *o++ = (char)('0' + (f == 1));
        o = LIT(o, " IsSingletonFE=");
        *o++ = (char)('0' + (pc == 1));
        *o++ = '\n';
        b->n = (size_t)(o - b->p);
    }
}

static void *write_range(void *arg) {
    wjob_t *j = (wjob_t *)arg;
    const cgx_result_t *r = j->res;
    sbuf out = {0}, srcbuf = {0};
    int32_t *stamp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(r->G > 0 ? r->G : 1));
    for (int32_t g = 0; g < r->G; g++) stamp[g] = -1;
    char fn[4096];
    for (int32_t q = j->q_begin; q < j->q_end; q++) {
        out.n = 0;
        /* contiguous phrases in first-appearance order inside the query (GenerateBlocks, ExtractPair.cu:2786-2892) */
        for (int32_t t = j->qry_off[q]; t < j->qry_off[q + 1]; t++) {
            for (int m = 0; m < 5; m++) {
                int32_t g = r->phrase_id[(size_t)t * 5 + m];
                if (g < 0 || stamp[g] == q) continue;
                stamp[g] = q;
                put_group(&out, r, j->src, j->tgt, 1, g + r->G, &srcbuf);   /* abX  */
                put_group(&out, r, j->src, j->tgt, 1, g, &srcbuf);          /* Xab  */
                put_group(&out, r, j->src, j->tgt, 2, g, &srcbuf);          /* XabX */
                put_group(&out, r, j->src, j->tgt, 0, g, &srcbuf);          /* ab   */
            }
        }
        for (int32_t k = r->q1_off[q]; k < r->q1_off[q + 1]; k++) {
            int32_t d = r->q1_ids[k];
            put_group(&out, r, j->src, j->tgt, 1, 2 * r->G + d, &srcbuf);               /* aXb  */
            put_group(&out, r, j->src, j->tgt, 2, r->G + r->D2 + d, &srcbuf);           /* XaXb */
            put_group(&out, r, j->src, j->tgt, 2, r->G + r->D2 + r->D1 + d, &srcbuf);   /* aXbX */
        }
        for (int32_t k = r->q2_off[q]; k < r->q2_off[q + 1]; k++)
            put_group(&out, r, j->src, j->tgt, 2, r->G + r->q2_ids[k], &srcbuf);        /* aXbXc */
        if (j->gzip_level > 0) {
            char mode[16];
            snprintf(fn, sizeof fn, "%s/grammar.%d.s.gz", j->outdir, j->qid_base + q);
            snprintf(mode, sizeof mode, "wb%d", j->gzip_level > 9 ? 9 : j->gzip_level);
            gzFile gz = gzopen(fn, mode);
            if (!gz) {
                fprintf(stderr, "Please check your file directory address for grammar rule files output. It is not valid. Program Exits.\n");
                j->rc = 1;
                break;
            }
            size_t done = 0;
            while (done < out.n) {                                                     /* gzwrite takes an unsigned length */
                const size_t piece = out.n - done > (1u << 30) ? (1u << 30) : out.n - done;
                if (gzwrite(gz, out.p + done, (unsigned)piece) <= 0) { j->rc = 1; break; }
                done += piece;
            }
            if (gzclose(gz) != Z_OK) j->rc = 1;
            if (j->rc) break;
            continue;
        }
        snprintf(fn, sizeof fn, "%s/grammar.%d.s", j->outdir, j->qid_base + q);         /* PrintResults.c:437 */
        FILE *fp = fopen(fn, "w");
        if (!fp) {
            fprintf(stderr, "Please check your file directory address for grammar rule files output. It is not valid. Program Exits.\n");
            j->rc = 1;
            break;
        }
        if (out.n) fwrite(out.p, 1, out.n, fp);
        fclose(fp);
    }
    free(stamp); free(out.p); free(srcbuf.p);
    return NULL;
}

int cgxh_write_grammars(const char *outdir, const cgx_result_t *res, const int32_t *qry_off, int32_t qid_base, const cgxh_side_t *src,
                        const cgxh_side_t *tgt, int n_threads) {
    return cgxh_write_grammars_ex(outdir, res, qry_off, qid_base, src, tgt, n_threads, 0);
}

int cgxh_write_grammars_ex(const char *outdir, const cgx_result_t *res, const int32_t *qry_off, int32_t qid_base, const cgxh_side_t *src,
                           const cgxh_side_t *tgt, int n_threads, int gzip_level) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > res->Q) n_threads = res->Q > 0 ? res->Q : 1;
    wjob_t *jobs = (wjob_t *)calloc((size_t)n_threads, sizeof(wjob_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    int rc = 0;
    for (int i = 0; i < n_threads; i++) {
        jobs[i].outdir = outdir; jobs[i].res = res; jobs[i].qry_off = qry_off; jobs[i].qid_base = qid_base; jobs[i].src = src; jobs[i].tgt = tgt; jobs[i].gzip_level = gzip_level;
        jobs[i].q_begin = (int32_t)((int64_t)res->Q * i / n_threads);
        jobs[i].q_end = (int32_t)((int64_t)res->Q * (i + 1) / n_threads);
        if (n_threads == 1) write_range(&jobs[i]);
        else pthread_create(&th[i], NULL, write_range, &jobs[i]);
    }
    for (int i = 0; i < n_threads; i++) {
        if (n_threads > 1) pthread_join(th[i], NULL);
        rc |= jobs[i].rc;
    }
    free(jobs); free(th);
    return rc;
}
