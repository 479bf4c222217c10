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

typedef struct {
    const char *outdir;
    const cgx_result_t *res;
    const int32_t *qry_off;
    int32_t qid_base;
    const cgxh_side_t *src, *tgt;
    int32_t q_begin, q_end;
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
        else put_pat1(b, src, &r->pat1[8 * (cid - 2 * G)], "[X,1]");
        return;
    }
    if (cid < G) { sb_puts(b, "[X,1] "); put_phrase(b, src, r->phrases[4 * cid + 3], r->phrases[4 * cid + 2]); sb_puts(b, " [X,2]"); }
    else if (cid < G + D2) {
        const int32_t *p2 = &r->pat2[4 * (cid - G)];
        put_pat1(b, src, &r->pat1[8 * p2[0]], "[X,1]");
        sb_puts(b, " [X,2] "); sb_puts(b, cgxh_vocab_name(src->vocab, p2[1]));
    } else if (cid < G + D2 + D1) { sb_puts(b, "[X,1] "); put_pat1(b, src, &r->pat1[8 * (cid - G - D2)], "[X,2]"); }
    else { put_pat1(b, src, &r->pat1[8 * (cid - G - D2 - D1)], "[X,1]"); sb_puts(b, " [X,2]"); }
}

static void put_group(sbuf *b, const cgx_result_t *r, const cgxh_side_t *src, const cgxh_side_t *tgt, int kind, int32_t cid, sbuf *srcbuf) {
    if (cid < 0 || cid >= r->n_ids[kind]) return;
    int32_t lo = r->updown[kind][2 * cid], hi = r->updown[kind][2 * cid + 1];
    if (lo < 0 || hi < 0) return;
    srcbuf->n = 0;
    source_string(srcbuf, r, src, kind, cid);
    sb_reserve(srcbuf, 1);
    srcbuf->p[srcbuf->n] = 0;
    char feat[256];
    const uint32_t idw = r->idinfo[kind][cid];
    const int f = CGX_ID_F(idw), fs = CGX_ID_FS(idw);
    for (int32_t i = lo; i <= hi; i++) {
        const cgx_rule_t *u = &r->rules[kind][i];
        const int end = CGX_RULE_END(u), g1 = CGX_RULE_GAP1(u), g1e = CGX_RULE_GAP1_END(u), g2 = CGX_RULE_GAP2(u), g2e = CGX_RULE_GAP2_END(u);
        const int pc = CGX_RULE_PC(u);
        sb_puts(b, "[X] ||| "); sb_puts(b, srcbuf->p); sb_puts(b, " ||| ");
        int first = 1;
        for (int j = 0; j <= end; j++) {
            const char *w;
            if (g1 != CGX_RULE_NOGAP && j >= g1 && j <= g1e) { w = "[X,1]"; j = g1e; }
            else if (g2 != CGX_RULE_NOGAP && j >= g2 && j <= g2e) { w = "[X,2]"; j = g2e; }
            else w = cgxh_vocab_name(tgt->vocab, tgt->tok[u->tgt_start + j]);
            if (!first) sb_putc(b, ' ');
            first = 0;
            sb_puts(b, w);
        }
        /* ExtractPair.c:653-655, :641: float log10 of the ratio, double log10 of the counts */
        float aa = -log10f((float)pc / (float)fs);
        float score = (float)log10((double)(1 + fs));
        float bb = (float)log10((double)(1 + pc));
        snprintf(feat, sizeof feat, " ||| EgivenFCoherent=%f SampleCountF=%f CountEF=%f MaxLexFgivenE=%f MaxLexEgivenF=%f IsSingletonF=%d IsSingletonFE=%d\n",
                 aa, score, bb, u->max_lex_f_given_e, u->max_lex_e_given_f, f == 1, pc == 1);
        sb_puts(b, feat);
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
    if (n_threads < 1) n_threads = 1;
    if (n_threads > res->Q) n_threads = res->Q > 0 ? res->Q : 1;
    wjob_t *jobs = (wjob_t *)calloc((size_t)n_threads, sizeof(wjob_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    int rc = 0;
    for (int i = 0; i < n_threads; i++) {
        jobs[i].outdir = outdir; jobs[i].res = res; jobs[i].qry_off = qry_off; jobs[i].qid_base = qid_base; jobs[i].src = src; jobs[i].tgt = tgt;
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
