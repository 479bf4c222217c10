/* cgx-b200 host: text loaders producing exactly the integer layouts of the reference's loaders, so that
 * the suffix array and every downstream id agree bit-for-bit. */
#define _GNU_SOURCE
#include "cgx_host.h"
#include <ctype.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- vocabulary: open-addressing string -> id map (replaces uthash; ids = 2 + first appearance) ----
 * A slot keeps the upper half of the key's hash next to the key pointer, so a probe that passes another word does not touch that
 * word's characters, and a lookup takes (pointer, length, hash) -- the loaders hash a token while they find its end. */
struct cgxh_vocab {
    char **slot_key;
    int32_t *slot_val;
    uint32_t *slot_tag;  /* hash >> 32 of the key */
    int64_t cap, count;
    char **names;        /* id -> name */
    int64_t names_cap;
};

#define FNV_BASIS 1469598103934665603ULL
#define FNV_PRIME 1099511628211ULL
static uint64_t hash_str(const char *s, size_t *len) {
    uint64_t h = FNV_BASIS;
    const char *p = s;
    while (*p) { h ^= (unsigned char)*p++; h *= FNV_PRIME; }
    *len = (size_t)(p - s);
    return h;
}

static cgxh_vocab_t *vocab_new(void) {
    cgxh_vocab_t *v = (cgxh_vocab_t *)calloc(1, sizeof(*v));
    v->cap = 1 << 16;
    v->slot_key = (char **)calloc((size_t)v->cap, sizeof(char *));
    v->slot_val = (int32_t *)malloc(sizeof(int32_t) * (size_t)v->cap);
    v->slot_tag = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)v->cap);
    v->names_cap = 1 << 16;
    v->names = (char **)calloc((size_t)v->names_cap, sizeof(char *));
    return v;
}

static void vocab_insert_raw(cgxh_vocab_t *v, char *key, uint64_t hash, int32_t val) {
    uint64_t h = hash & (uint64_t)(v->cap - 1);
    while (v->slot_key[h]) h = (h + 1) & (uint64_t)(v->cap - 1);
    v->slot_key[h] = key;
    v->slot_val[h] = val;
    v->slot_tag[h] = (uint32_t)(hash >> 32);
}

static inline int32_t vocab_find(const cgxh_vocab_t *v, const char *name, size_t len, uint64_t hash) {
    uint64_t h = hash & (uint64_t)(v->cap - 1);
    const uint32_t tag = (uint32_t)(hash >> 32);
    while (v->slot_key[h]) {
        if (v->slot_tag[h] == tag && memcmp(v->slot_key[h], name, len) == 0 && v->slot_key[h][len] == '\0') return v->slot_val[h];
        h = (h + 1) & (uint64_t)(v->cap - 1);
    }
    return -1;
}

int32_t cgxh_vocab_id(const cgxh_vocab_t *v, const char *name) {
    size_t len;
    const uint64_t hash = hash_str(name, &len);
    return vocab_find(v, name, len, hash);
}

static int32_t vocab_add(cgxh_vocab_t *v, const char *name, size_t len, uint64_t hash) {
    if (v->count * 2 >= v->cap) {
        char **ok = v->slot_key; int32_t *ov = v->slot_val; uint32_t *ot = v->slot_tag; int64_t oc = v->cap;
        v->cap *= 2;
        v->slot_key = (char **)calloc((size_t)v->cap, sizeof(char *));
        v->slot_val = (int32_t *)malloc(sizeof(int32_t) * (size_t)v->cap);
        v->slot_tag = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)v->cap);
        for (int64_t i = 0; i < oc; i++)
            if (ok[i]) { size_t l; vocab_insert_raw(v, ok[i], hash_str(ok[i], &l), ov[i]); }
        free(ok); free(ov); free(ot);
    }
    int32_t id = (int32_t)v->count + 2;                         /* Start.cu:288 HASH_COUNT + 2 */
    char *cp = (char *)malloc(len + 1);
    memcpy(cp, name, len);
    cp[len] = '\0';
    vocab_insert_raw(v, cp, hash, id);
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
    free(v->slot_key); free(v->slot_val); free(v->slot_tag); free(v->names); free(v);
}

/* ---- growable int buffer ---- */
typedef struct { int32_t *v; int64_t n, cap; } ibuf;
static inline void ib_push(ibuf *b, int32_t x) {
    if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 1 << 16; b->v = (int32_t *)realloc(b->v, sizeof(int32_t) * (size_t)b->cap); }
    b->v[b->n++] = x;
}

/* Tokenisation of the reference (Start.cu:270-310): getline, strip one trailing '\n', strtok on ' ' -- tokens are the maximal
 * runs of non-blank characters, a tab or a carriage return is part of its token -- and the walk over a line stops at the first
 * token that begins with white space.  next_token is that walk without strtok (no hidden state: the two sides of the corpus load
 * on two threads), hashing the token while it looks for its end.  NULL at the end of the line. */
static inline const char *next_token(const char **cursor, size_t *len, uint64_t *hash) {
    const char *p = *cursor;
    while (*p == ' ') p++;
    if (*p == '\0' || isspace((unsigned char)*p)) return NULL;
    const char *s = p;
    uint64_t h = FNV_BASIS;
    while (*p != ' ' && *p != '\0') { h ^= (unsigned char)*p++; h *= FNV_PRIME; }
    *len = (size_t)(p - s);
    *hash = h;
    *cursor = p;
    return s;
}
static inline void strip_newline(char *line, ssize_t got) {
    if (got > 0 && line[got - 1] == '\n') line[got - 1] = '\0';
}

int cgxh_corpus_load(const char *path, int want_P, cgxh_side_t *out) {
    memset(out, 0, sizeof(*out));
    FILE *fh = fopen(path, "r");
    if (!fh) { fprintf(stderr, "Can not open reference file \"%s\"\n", path); return 1; }
    cgxh_vocab_t *v = vocab_new();
    ibuf tok = {0}, sent = {0};
    char *line = NULL; size_t cap = 0;
    ssize_t got;
    int32_t last = -1;
    ib_push(&sent, 0);
    while ((got = getline(&line, &cap, fh)) != -1) {
        strip_newline(line, got);
        const char *cur = line, *t;
        size_t tl;
        uint64_t th;
        while ((t = next_token(&cur, &tl, &th)) != NULL) {
            int32_t id = vocab_find(v, t, tl, th);
            if (id < 0) { id = vocab_add(v, t, tl, th); last = id; }
            ib_push(&tok, id);
        }
        ib_push(&tok, 1);                                        /* EOS, Start.cu:306 */
        ib_push(&sent, (int32_t)tok.n);
    }
    free(line); fclose(fh);
    ib_push(&tok, 1);                                            /* Start.cu:321-326 */
    last++;
    ib_push(&tok, last);
    out->n = tok.n;
    ib_push(&tok, 0); ib_push(&tok, 0); ib_push(&tok, 0);        /* Start.cu:354 */
    out->tok = tok.v;
    if (want_P) {                                                /* uint8_t localcount, Start.cu:300; 0 at an EOS and in the trailer */
        out->P = (uint8_t *)calloc((size_t)out->n, 1);
        for (int64_t q = 0; q + 1 < sent.n; q++)
            for (int64_t i = sent.v[q], e = (int64_t)sent.v[q + 1] - 1; i < e; i++) out->P[i] = (uint8_t)((i - sent.v[q]) & 0xFF);
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

/* atoi of the token at *cursor (a run of characters other than ' ' and '-'): leading white space, an optional '+', digits;
 * anything after the digits is ignored.  Leaves *cursor at the end of the token.  Saturates (any value >= 65535 is refused). */
static inline int link_number(const char **cursor) {
    const char *p = *cursor;
    while (*p != '\0' && *p != ' ' && *p != '-' && isspace((unsigned char)*p)) p++;
    if (*p == '+') p++;
    int v = 0;
    while (*p >= '0' && *p <= '9') { if (v < 100000000) v = v * 10 + (*p - '0'); p++; }
    while (*p != '\0' && *p != ' ' && *p != '-') p++;
    *cursor = p;
    return v;
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
    ssize_t got;
    while ((got = getline(&line, &cap, fh)) != -1) {
        qcount++;
        if (qcount >= src->n_sent || qcount >= tgt->n_sent) { fprintf(stderr, "alignment file has more lines than the corpus; ignoring the rest\n"); break; }
        strip_newline(line, got);
        /* the reference's walk: strtok(line, " -"), atoi of the token, again for the target index; a line ends at its end or at a
         * source token that begins with white space */
        const char *p = line;
        for (;;) {
            while (*p == ' ' || *p == '-') p++;
            if (*p == '\0' || isspace((unsigned char)*p)) break;
            const int s_no = link_number(&p);
            while (*p == ' ' || *p == '-') p++;
            if (*p == '\0') { fprintf(stderr, "Not possible!\n"); rc = 2; goto done; }
            const int t_no = link_number(&p);
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
    FILE *fh = fopen(path, "rb");
    if (!fh) { fprintf(stderr, "The Word Possibility File is not Found!\n"); return 1; }
    /* the reference reads the file as one stream of white-space separated fields (fscanf "%s %s %f %f", :2478): so does this, over
     * the whole file in memory, fields terminated in place.  A field that is not a number ends the table, as a failed %f does. */
    size_t size = 0, room = 1 << 20;
    char *buf = (char *)malloc(room + 1);
    for (size_t r; (r = fread(buf + size, 1, room - size, fh)) > 0;) {
        size += r;
        if (size == room) { room *= 2; buf = (char *)realloc(buf, room + 1); }
    }
    fclose(fh);
    buf[size] = '\0';
    int64_t cap = 1 << 16, cnt = 0;
    int32_t *f = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap), *e = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
    float *v1 = (float *)malloc(sizeof(float) * (size_t)cap), *v2 = (float *)malloc(sizeof(float) * (size_t)cap);
    char *p = buf, *const end = buf + size;
    for (;;) {
        char *field[4];
        int k = 0;
        for (; k < 4; k++) {
            while (p < end && isspace((unsigned char)*p)) p++;
            if (p >= end) break;
            field[k] = p;
            while (p < end && !isspace((unsigned char)*p)) p++;
            if (p < end) *p++ = '\0';
        }
        if (k < 4) break;
        char *stop;
        const float x = strtof(field[2], &stop);
        if (stop == field[2] || *stop != '\0') break;
        const float y = strtof(field[3], &stop);
        if (stop == field[3] || *stop != '\0') break;
        const int32_t fi = cgxh_vocab_id(src->vocab, field[0]), ei = cgxh_vocab_id(tgt->vocab, field[1]);
        if (fi < 0 && strcmp(field[0], "NULL") != 0) continue;
        if (ei < 0 && strcmp(field[1], "NULL") != 0) continue;
        if (cnt == cap) {
            cap *= 2;
            f = (int32_t *)realloc(f, sizeof(int32_t) * (size_t)cap); e = (int32_t *)realloc(e, sizeof(int32_t) * (size_t)cap);
            v1 = (float *)realloc(v1, sizeof(float) * (size_t)cap); v2 = (float *)realloc(v2, sizeof(float) * (size_t)cap);
        }
        f[cnt] = fi < 0 ? -1 : fi; e[cnt] = ei < 0 ? -1 : ei; v1[cnt] = x; v2[cnt] = y;
        cnt++;
    }
    free(buf);
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
    ssize_t got;
    while ((got = getline(&line, &cap, fh)) != -1) {
        int per = 0;
        strip_newline(line, got);
        const char *cur = line, *t;
        size_t tl;
        uint64_t th;
        while ((t = next_token(&cur, &tl, &th)) != NULL) {
            ib_push(&tok, vocab_find(src->vocab, t, tl, th));
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

/* ---- all the corpus files at once: the two sides on two threads, then the alignment and the lexical file on two threads (each of
 * the latter needs both sentence tables / vocabularies).  align_path / lex_path may be NULL (persisted index: neither is parsed). ---- */
typedef struct {
    int kind, rc;
    const char *path;
    int want_P;
    cgxh_side_t *side;
    const cgxh_side_t *src, *tgt;
    cgxh_align_t *al;
    cgxh_lex_t *lex;
} load_job;
static void *load_job_main(void *arg) {
    load_job *j = (load_job *)arg;
    if (j->kind == 0) j->rc = cgxh_corpus_load(j->path, j->want_P, j->side);
    else if (j->kind == 1) j->rc = cgxh_alignment_load(j->path, j->src, j->tgt, j->al);
    else j->rc = cgxh_lex_load(j->path, j->src, j->tgt, j->lex);
    return NULL;
}
/* runs a on a new thread and b on this one (a alone when b is NULL) */
static void run_pair(load_job *a, load_job *b) {
    pthread_t th;
    const int threaded = b != NULL && pthread_create(&th, NULL, load_job_main, a) == 0;
    if (!threaded) load_job_main(a);
    if (b) load_job_main(b);
    if (threaded) pthread_join(th, NULL);
}
int cgxh_load_files(const char *src_path, const char *tgt_path, const char *align_path, const char *lex_path, cgxh_side_t *src,
                    cgxh_side_t *tgt, cgxh_align_t *al, cgxh_lex_t *lex) {
    memset(al, 0, sizeof(*al));
    memset(lex, 0, sizeof(*lex));
    load_job js = {0, 0, src_path, 1, src, NULL, NULL, NULL, NULL}, jt = {0, 0, tgt_path, 0, tgt, NULL, NULL, NULL, NULL};
    run_pair(&jt, &js);
    if (js.rc || jt.rc) return js.rc ? js.rc : jt.rc;
    load_job ja = {1, 0, align_path, 0, NULL, src, tgt, al, NULL}, jl = {2, 0, lex_path, 0, NULL, src, tgt, NULL, lex};
    if (align_path && lex_path) run_pair(&jl, &ja);
    else if (align_path) run_pair(&ja, NULL);
    else if (lex_path) run_pair(&jl, NULL);
    return jl.rc ? jl.rc : ja.rc;
}
