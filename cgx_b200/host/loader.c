/* cgx-b200 host: text loaders producing exactly the integer layouts of the reference's loaders, so that
 * the suffix array and every downstream id agree bit-for-bit. */
#define _GNU_SOURCE
#include "cgx_host.h"
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- vocabulary: open-addressing string -> id map (replaces uthash; ids = 2 + first appearance) ---- */
struct cgxh_vocab {
    char **slot_key;
    int32_t *slot_val;
    int64_t cap, count;
    char **names;        /* id -> name */
    int64_t names_cap;
};

static uint64_t hash_str(const char *s) {
    uint64_t h = 1469598103934665603ULL;
    while (*s) { h ^= (unsigned char)*s++; h *= 1099511628211ULL; }
    return h;
}

static cgxh_vocab_t *vocab_new(void) {
    cgxh_vocab_t *v = (cgxh_vocab_t *)calloc(1, sizeof(*v));
    v->cap = 1 << 16;
    v->slot_key = (char **)calloc((size_t)v->cap, sizeof(char *));
    v->slot_val = (int32_t *)malloc(sizeof(int32_t) * (size_t)v->cap);
    v->names_cap = 1 << 16;
    v->names = (char **)calloc((size_t)v->names_cap, sizeof(char *));
    return v;
}

static void vocab_insert_raw(cgxh_vocab_t *v, char *key, int32_t val) {
    uint64_t h = hash_str(key) & (uint64_t)(v->cap - 1);
    while (v->slot_key[h]) h = (h + 1) & (uint64_t)(v->cap - 1);
    v->slot_key[h] = key;
    v->slot_val[h] = val;
}

int32_t cgxh_vocab_id(const cgxh_vocab_t *v, const char *name) {
    uint64_t h = hash_str(name) & (uint64_t)(v->cap - 1);
    while (v->slot_key[h]) {
        if (strcmp(v->slot_key[h], name) == 0) return v->slot_val[h];
        h = (h + 1) & (uint64_t)(v->cap - 1);
    }
    return -1;
}

static int32_t vocab_add(cgxh_vocab_t *v, const char *name) {
    if (v->count * 2 >= v->cap) {
        char **ok = v->slot_key; int32_t *ov = v->slot_val; int64_t oc = v->cap;
        v->cap *= 2;
        v->slot_key = (char **)calloc((size_t)v->cap, sizeof(char *));
        v->slot_val = (int32_t *)malloc(sizeof(int32_t) * (size_t)v->cap);
        for (int64_t i = 0; i < oc; i++) if (ok[i]) vocab_insert_raw(v, ok[i], ov[i]);
        free(ok); free(ov);
    }
    int32_t id = (int32_t)v->count + 2;                         /* Start.cu:288 HASH_COUNT + 2 */
    char *cp = strdup(name);
    vocab_insert_raw(v, cp, id);
    v->count++;
    if (id >= v->names_cap) {
        int64_t nc = v->names_cap * 2;
        v->names = (char **)realloc(v->names, sizeof(char *) * (size_t)nc);
        memset(v->names + v->names_cap, 0, sizeof(char *) * (size_t)(nc - v->names_cap));
        v->names_cap = nc;
    }
    v->names[id] = cp;
    return id;
}

const char *cgxh_vocab_name(const cgxh_vocab_t *v, int32_t id) { return (id >= 0 && id < v->names_cap) ? v->names[id] : NULL; }
int32_t cgxh_vocab_size(const cgxh_vocab_t *v) { return (int32_t)v->count + 2; }

static void vocab_free(cgxh_vocab_t *v) {
    if (!v) return;
    for (int64_t i = 0; i < v->cap; i++) free(v->slot_key[i]);
    free(v->slot_key); free(v->slot_val); free(v->names); free(v);
}

/* ---- growable int buffer ---- */
typedef struct { int32_t *v; int64_t n, cap; } ibuf;
static void ib_push(ibuf *b, int32_t x) {
    if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 1 << 16; b->v = (int32_t *)realloc(b->v, sizeof(int32_t) * (size_t)b->cap); }
    b->v[b->n++] = x;
}

/* Tokenisation of the reference (Start.cu:270-310): getline, strip one trailing '\n', strtok on ' ',
 * stop at the first token that begins with white space, strip a trailing '\n' from a token. */
#define FOR_EACH_TOKEN(line, tokvar) \
    for (char *tokvar = strtok((line), " "); tokvar != NULL && !isspace((unsigned char)*tokvar); tokvar = strtok(NULL, " "))

int cgxh_corpus_load(const char *path, int want_P, cgxh_side_t *out) {
    memset(out, 0, sizeof(*out));
    FILE *fh = fopen(path, "r");
    if (!fh) { fprintf(stderr, "Can not open reference file \"%s\"\n", path); return 1; }
    cgxh_vocab_t *v = vocab_new();
    ibuf tok = {0}, sent = {0}, pos = {0};
    char *line = NULL; size_t cap = 0;
    int32_t last = -1;
    ib_push(&sent, 0);
    while (getline(&line, &cap, fh) != -1) {
        size_t l = strlen(line);
        if (l && line[l - 1] == '\n') line[l - 1] = '\0';
        int local = 0;
        FOR_EACH_TOKEN(line, t) {
            size_t tl = strlen(t);
            if (tl && t[tl - 1] == '\n') t[tl - 1] = '\0';
            int32_t id = cgxh_vocab_id(v, t);
            if (id < 0) { id = vocab_add(v, t); last = id; }
            ib_push(&tok, id);
            if (want_P) ib_push(&pos, local & 0xFF);             /* uint8_t localcount, Start.cu:300 */
            local++;
        }
        ib_push(&tok, 1);                                        /* EOS, Start.cu:306 */
        if (want_P) ib_push(&pos, 0);
        ib_push(&sent, (int32_t)tok.n);
    }
    free(line); fclose(fh);
    ib_push(&tok, 1);                                            /* Start.cu:321-326 */
    if (want_P) ib_push(&pos, 0);
    last++;
    ib_push(&tok, last);
    if (want_P) ib_push(&pos, 0);
    out->n = tok.n;
    ib_push(&tok, 0); ib_push(&tok, 0); ib_push(&tok, 0);        /* Start.cu:354 */
    out->tok = tok.v;
    if (want_P) {
        out->P = (uint8_t *)malloc((size_t)out->n);
        for (int64_t i = 0; i < out->n; i++) out->P[i] = (uint8_t)pos.v[i];
        free(pos.v);
    }
    out->sentenceind = sent.v;
    out->n_sent = (int32_t)sent.n - 1;
    out->vocab = v;
    out->last = last;
    return 0;
}

void cgxh_side_free(cgxh_side_t *s) {
    free(s->tok); free(s->P); free(s->sentenceind); vocab_free(s->vocab);
    memset(s, 0, sizeof(*s));
}

/* ExtractPair.cu:2639-2739.  The reference keeps the aligned spans and the position in the sentence in 8 bits and exits on an
 * alignment point at index 255 or beyond ("Not possible, too long sentence", :2683).  Here the spans are collected in 16 bits;
 * a corpus that fits the reference's layout gets exactly that layout (out->wide = 0: RLP, L_tar, R_tar), any other the 16-bit
 * one (out->wide = 1: RLP64 = L << 48 | R << 32 | P << 16, L_tar16, R_tar16, 65535 = unaligned; cgx_index_build_wide). */
int cgxh_alignment_load(const char *path, const cgxh_side_t *src, const cgxh_side_t *tgt, cgxh_align_t *out) {
    memset(out, 0, sizeof(*out));
    FILE *fh = fopen(path, "r");
    if (!fh) { fprintf(stderr, "Can not open reference file \"%s\"\n", path); return 1; }
    const int64_t n = src->n, m = tgt->n;
    uint16_t *Lt = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)m), *Rt = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)m);
    uint16_t *Ls = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)n), *Rs = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)n);
    memset(Lt, 255, sizeof(uint16_t) * (size_t)m); memset(Rt, 255, sizeof(uint16_t) * (size_t)m);
    memset(Ls, 255, sizeof(uint16_t) * (size_t)n); memset(Rs, 255, sizeof(uint16_t) * (size_t)n);
    char *line = NULL; size_t cap = 0;
    int32_t qcount = -1;
    int rc = 0, wide = 0;
    for (int32_t q = 0; q < src->n_sent; q++)                      /* a sentence the 8-bit position counter cannot count (Start.cu:269,300) */
        if (src->sentenceind[q + 1] - src->sentenceind[q] - 1 > 255) wide = 1;
    while (getline(&line, &cap, fh) != -1) {
        qcount++;
        if (qcount >= src->n_sent || qcount >= tgt->n_sent) { fprintf(stderr, "alignment file has more lines than the corpus; ignoring the rest\n"); break; }
        size_t l = strlen(line);
        if (l && line[l - 1] == '\n') line[l - 1] = '\0';
        char *t = strtok(line, " -");
        while (t != NULL && !isspace((unsigned char)*t)) {
            int s_no = atoi(t);
            t = strtok(NULL, " -");
            if (!t) { fprintf(stderr, "Not possible!\n"); rc = 2; goto done; }
            int t_no = atoi(t);
            if (s_no >= 65535 || t_no >= 65535 || s_no < 0 || t_no < 0) { fprintf(stderr, "Not possible, too long sentence\n"); rc = 3; goto done; }
            if (s_no >= 255 || t_no >= 255) wide = 1;
            int64_t si = (int64_t)src->sentenceind[qcount] + s_no, ti = (int64_t)tgt->sentenceind[qcount] + t_no;
            if (si >= n || ti >= m) { fprintf(stderr, "alignment point outside the corpus at line %d\n", qcount); rc = 4; goto done; }
            if (Ls[si] == 65535 || Rs[si] == 65535) { Ls[si] = (uint16_t)t_no; Rs[si] = (uint16_t)t_no; }
            else if (t_no > Rs[si]) Rs[si] = (uint16_t)t_no;
            else if (t_no < Ls[si]) Ls[si] = (uint16_t)t_no;
            if (Lt[ti] == 65535 || Rt[ti] == 65535) { Lt[ti] = (uint16_t)s_no; Rt[ti] = (uint16_t)s_no; }
            else if (s_no > Rt[ti]) Rt[ti] = (uint16_t)s_no;
            else if (s_no < Lt[ti]) Lt[ti] = (uint16_t)s_no;
            t = strtok(NULL, " -");
        }
    }
    out->wide = wide;
    if (wide) {
        uint64_t *RLP = (uint64_t *)calloc((size_t)n, sizeof(uint64_t));
        int32_t q = 1;
        int64_t sent_start = 0;
        for (int64_t i = 0; i < n - 1; i++) {
            if (q <= src->n_sent && i == (int64_t)src->sentenceind[q] - 1) {
                RLP[i] = (uint64_t)(uint32_t)(q <= tgt->n_sent ? tgt->sentenceind[q] : tgt->sentenceind[tgt->n_sent]);
                sent_start = src->sentenceind[q];
                q++;
            } else {
                const int64_t P = q <= src->n_sent ? i - sent_start : 0;      /* the trailer after the last sentence counts from 0 (Start.cu:324) */
                RLP[i] = ((uint64_t)Ls[i] << 48) | ((uint64_t)Rs[i] << 32) | ((uint64_t)(P & 0xFFFF) << 16);
            }
        }
        out->RLP64 = RLP; out->L_tar16 = Lt; out->R_tar16 = Rt;
        Lt = Rt = NULL;
    } else {
        uint32_t *RLP = (uint32_t *)calloc((size_t)n, sizeof(uint32_t));
        uint8_t *Lt8 = (uint8_t *)malloc((size_t)m), *Rt8 = (uint8_t *)malloc((size_t)m);
        for (int64_t j = 0; j < m; j++) { Lt8[j] = (uint8_t)(Lt[j] == 65535 ? 255 : Lt[j]); Rt8[j] = (uint8_t)(Rt[j] == 65535 ? 255 : Rt[j]); }
        int32_t q = 1;
        for (int64_t i = 0; i < n - 1; i++) {                    /* :2721-2731 */
            if (q <= src->n_sent && i == (int64_t)src->sentenceind[q] - 1) {
                RLP[i] = (uint32_t)(q <= tgt->n_sent ? tgt->sentenceind[q] : tgt->sentenceind[tgt->n_sent]);
                q++;
            } else {
                const uint32_t L8 = Ls[i] == 65535 ? 255u : Ls[i], R8 = Rs[i] == 65535 ? 255u : Rs[i];
                RLP[i] = (L8 << 24) | (R8 << 16) | ((uint32_t)src->P[i] << 8);
            }
        }
        out->RLP = RLP; out->L_tar = Lt8; out->R_tar = Rt8;
    }
done:
    free(line); fclose(fh); free(Ls); free(Rs); free(Lt); free(Rt);
    return rc;
}
void cgxh_align_free(cgxh_align_t *a) { free(a->RLP); free(a->L_tar); free(a->R_tar); free(a->RLP64); free(a->L_tar16); free(a->R_tar16); memset(a, 0, sizeof(*a)); }

/* ExtractPair.cu:2463-2519: four white-space separated columns; words unknown to the corpus are skipped
 * unless they are the literal NULL, which maps to id -1. */
int cgxh_lex_load(const char *path, const cgxh_side_t *src, const cgxh_side_t *tgt, cgxh_lex_t *out) {
    memset(out, 0, sizeof(*out));
    FILE *fh = fopen(path, "r");
    if (!fh) { fprintf(stderr, "The Word Possibility File is not Found!\n"); return 1; }
    int64_t cap = 1 << 16, cnt = 0;
    int32_t *f = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap), *e = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
    float *v1 = (float *)malloc(sizeof(float) * (size_t)cap), *v2 = (float *)malloc(sizeof(float) * (size_t)cap);
    char a[8192], b[8192];
    float x, y;
    while (fscanf(fh, "%8191s %8191s %f %f", a, b, &x, &y) == 4) {
        int32_t fi = cgxh_vocab_id(src->vocab, a), ei = cgxh_vocab_id(tgt->vocab, b);
        if (fi < 0 && strcmp(a, "NULL") != 0) continue;
        if (ei < 0 && strcmp(b, "NULL") != 0) continue;
        if (cnt == cap) {
            cap *= 2;
            f = (int32_t *)realloc(f, sizeof(int32_t) * (size_t)cap); e = (int32_t *)realloc(e, sizeof(int32_t) * (size_t)cap);
            v1 = (float *)realloc(v1, sizeof(float) * (size_t)cap); v2 = (float *)realloc(v2, sizeof(float) * (size_t)cap);
        }
        f[cnt] = fi < 0 ? -1 : fi; e[cnt] = ei < 0 ? -1 : ei; v1[cnt] = x; v2[cnt] = y;
        cnt++;
    }
    fclose(fh);
    out->f = f; out->e = e; out->v1 = v1; out->v2 = v2; out->count = cnt;
    return 0;
}

void cgxh_lex_free(cgxh_lex_t *l) { free(l->f); free(l->e); free(l->v1); free(l->v2); memset(l, 0, sizeof(*l)); }

/* Start.cu:50-132: ids in the SOURCE vocabulary, -1 for unknown words */
int cgxh_queries_load(const char *path, const cgxh_side_t *src, cgxh_queries_t *out) {
    memset(out, 0, sizeof(*out));
    FILE *fh = fopen(path, "r");
    if (!fh) { fprintf(stderr, "Can not open query file \"%s\"\n", path); return 1; }
    ibuf tok = {0}, off = {0};
    char *line = NULL; size_t cap = 0;
    int max_len = -1;
    ib_push(&off, 0);
    while (getline(&line, &cap, fh) != -1) {
        int per = 0;
        FOR_EACH_TOKEN(line, t) {
            size_t tl = strlen(t);
            if (tl && t[tl - 1] == '\n') t[tl - 1] = '\0';
            ib_push(&tok, cgxh_vocab_id(src->vocab, t));
            per++;
        }
        if (per > max_len) max_len = per;
        ib_push(&off, (int32_t)tok.n);
    }
    free(line); fclose(fh);
    if (!tok.v) tok.v = (int32_t *)calloc(1, sizeof(int32_t));
    out->tok = tok.v; out->off = off.v; out->Q = (int32_t)off.n - 1; out->T = (int32_t)tok.n; out->max_len = max_len;
    return 0;
}

void cgxh_queries_free(cgxh_queries_t *q) { free(q->tok); free(q->off); memset(q, 0, sizeof(*q)); }
