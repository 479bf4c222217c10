/* TEST INFRASTRUCTURE ONLY -- see cgx_oracle.h.  CPU restatement of the hohoCode/cgx path.
 *
 * Parity pin: outputs of the unmodified reference binary on a B200 (tests/golden/, tools/make_golden.py)
 * -- the reference owns no tests or golden vectors of its own.
 *
 * Plain C, single thread, written for clarity: each stage follows the reference kernel / host
 * function it cites, including the quirks that change results (tight-phrase rules, the
 * first-success state machines, the 512-thread early `return`s, sampling arithmetic).
 * Where the reference is nondeterministic (atomicAdd output order + unstable comparator sorts) the
 * oracle picks the order documented at the site.
 */
#define _GNU_SOURCE
#include "cgx_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

/* Width of the alignment fields.  The reference keeps the aligned span of a token and its position in the sentence in 8 bits
 * each (255 = unaligned; RLP = L << 24 | R << 16 | P << 8) and exits on any sentence of 255 tokens or more
 * (ExtractPair.cu:2683, Start.cu:269).  Built with -DORC_WIDE the SAME algorithm runs on 16-bit fields (65535 = unaligned,
 * RLP = L << 48 | R << 32 | P << 16; the word at an EOS still holds the target sentence offset): SURVEY.md 8f, lifted limit.
 * On a corpus the narrow build accepts, both builds give identical results (tests/test_host_cpu.py). */
#ifdef ORC_WIDE
#define UNAL 65535
#define RLP_L(t) ((lr_t)(((t) >> 48) & 0xFFFF))
#define RLP_R(t) ((lr_t)(((t) >> 32) & 0xFFFF))
#define RLP_P(t) ((int)(((t) >> 16) & 0xFFFF))
#define RLP_PACK(L, R, P) (((rlp_t)(L) << 48) | ((rlp_t)(R) << 32) | ((rlp_t)(P) << 16))
#else
#define UNAL 255
#define RLP_L(t) ((lr_t)(((t) >> 24) & 0xFF))
#define RLP_R(t) ((lr_t)(((t) >> 16) & 0xFF))
#define RLP_P(t) ((int)(((t) >> 8) & 0xFF))
#define RLP_PACK(L, R, P) (((rlp_t)(L) << 24) | ((rlp_t)(R) << 16) | ((rlp_t)(P) << 8))
#endif

/* ------------------------------------------------------------------------------------------ */
/* small utilities                                                                              */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int32_t *v; int64_t n, cap; } ivec;
static void iv_push(ivec *a, int32_t x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 1024; a->v = (int32_t *)realloc(a->v, sizeof(int32_t) * a->cap); }
    a->v[a->n++] = x;
}
static void iv_free(ivec *a) { free(a->v); a->v = NULL; a->n = a->cap = 0; }

/* string -> id open addressing map (replaces uthash, Start.cu:288) */
typedef struct { char **key; int32_t *val; int64_t cap, cnt; } smap;
static uint64_t fnv(const char *s) { uint64_t h = 1469598103934665603ULL; while (*s) { h ^= (unsigned char)*s++; h *= 1099511628211ULL; } return h; }
static void smap_init(smap *m, int64_t cap) { m->cap = cap; m->cnt = 0; m->key = (char **)calloc(cap, sizeof(char *)); m->val = (int32_t *)malloc(sizeof(int32_t) * cap); }
static void smap_grow(smap *m);
static int32_t smap_get(const smap *m, const char *s) {
    uint64_t h = fnv(s) & (m->cap - 1);
    while (m->key[h]) { if (!strcmp(m->key[h], s)) return m->val[h]; h = (h + 1) & (m->cap - 1); }
    return -1;
}
static void smap_put(smap *m, char *s, int32_t v) {
    if (m->cnt * 2 >= m->cap) smap_grow(m);
    uint64_t h = fnv(s) & (m->cap - 1);
    while (m->key[h]) h = (h + 1) & (m->cap - 1);
    m->key[h] = s; m->val[h] = v; m->cnt++;
}
static void smap_grow(smap *m) {
    smap o = *m; smap_init(m, o.cap * 2);
    for (int64_t i = 0; i < o.cap; i++) if (o.key[i]) smap_put(m, o.key[i], o.val[i]);
    free(o.key); free(o.val);
}

/* 64-bit key -> int map */
typedef struct { uint64_t *key; int32_t *val; int64_t cap, cnt; } kmap;
static void kmap_init(kmap *m, int64_t cap) { m->cap = cap; m->cnt = 0; m->key = (uint64_t *)malloc(sizeof(uint64_t) * cap); memset(m->key, 0xff, sizeof(uint64_t) * cap); m->val = (int32_t *)malloc(sizeof(int32_t) * cap); }
static uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
static int32_t *kmap_slot(kmap *m, uint64_t k, int *found) {
    if (m->cnt * 2 >= m->cap) {
        kmap o = *m; kmap_init(m, o.cap * 2);
        for (int64_t i = 0; i < o.cap; i++) if (o.key[i] != ~0ULL) { int f; *kmap_slot(m, o.key[i], &f) = o.val[i]; }
        free(o.key); free(o.val);
    }
    uint64_t h = mix64(k) & (m->cap - 1);
    while (m->key[h] != ~0ULL) { if (m->key[h] == k) { *found = 1; return &m->val[h]; } h = (h + 1) & (m->cap - 1); }
    m->key[h] = k; m->cnt++; *found = 0; return &m->val[h];
}
static void kmap_free(kmap *m) { free(m->key); free(m->val); }

/* ------------------------------------------------------------------------------------------ */
typedef struct { int32_t id, tgt_start, end, gap1, gap1_1, gap2, gap2_1; } rec_t;
typedef struct { rec_t *v; int64_t n, cap; } rvec;
static void rv_push(rvec *a, rec_t r) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 1024; a->v = (rec_t *)realloc(a->v, sizeof(rec_t) * a->cap); }
    a->v[a->n++] = r;
}

struct orc_s {
    int32_t n, m;
    int32_t *str, *tgt;
    rlp_t *RLP; lr_t *L_tar, *R_tar;
    int32_t *sa;
    /* text-mode extras */
    char **svocab, **tvocab; int32_t sv, tv;      /* id -> name */
    smap smap_src, smap_tgt; int have_vocab;
    /* lex table sorted by (f,e) (ExtractPair.cu:2537) */
    int64_t *lex_key; float *lex_v1, *lex_v2; int32_t lex_count;
    /* frequent-pair precomputation (SuffixArray.cu:1132-1340, GappyLook.cu:740-870) */
    int have_precomp;
    int32_t freq[ORC_PRECOMP];
    int32_t pidx[ORC_PRECOMP * ORC_PRECOMP * 2];     /* {start,end}; empty = {1,0} (:1306-1307) */
    int32_t missing[ORC_PRECOMP * ORC_PRECOMP];
    int32_t *plist; int32_t pcount;                  /* {start,len} */
    /* queries */
    int32_t Q, T; int32_t *q; int32_t *qoff; int32_t *tok2q;
    int32_t *longest; int32_t *conn_off; int32_t *iv_up, *iv_down;   /* result_two / result_connect */
    /* blocks */
    int32_t G; int32_t *blocks; ivec *qryglobal;
    /* one gap */
    int32_t enu1; int32_t *g1_start; uint8_t *g1_ls, *g1_le, *g1_gap; int32_t *g1_sorted;
    int32_t D1; int32_t *pat1; int32_t *pat1_pos; int32_t *pat1_rep; ivec *q1;
    int32_t hits1; int32_t *h1;
    /* two gap */
    int32_t enu2; int32_t D2; int32_t *pat2; int32_t *pat2_rep; ivec *q2;
    int32_t hits2; int32_t *h2;
    /* records */
    rvec rec_ab, rec_1, rec_2;
    int32_t sep1, sep2a, sep2b;
    int32_t *rec_flat[3];
    orc_counts_t cnt;
    /* rules */
    orc_rule_t *rules[3]; int32_t nrules[3];
    int32_t *updown[3]; int32_t nid[3];          /* per converted id: {down, up} into rules[kind] */
};

/* ------------------------------------------------------------------------------------------ */
/* loaders (text)                                                                               */
/* ------------------------------------------------------------------------------------------ */
/* Start.cu:240-380 initRefSet / :142-238 initRefTargetSet.  ids = 2 + first appearance,
 * EOS = 1 after every line, trailer "1, last+1", three zeros. */
static int load_side(const char *path, smap *map, char ***vocab_out, int32_t *nvocab, int32_t **tok_out, int32_t *n_out,
                     lr_t **P_out, int32_t **sent_out, int32_t *nsent_out) {
    FILE *fh = fopen(path, "r");
    if (!fh) { fprintf(stderr, "oracle: cannot open %s\n", path); return 0; }
    ivec tok = {0}, sent = {0}; ivec Pv = {0};
    smap_init(map, 1 << 16);
    int64_t vcap = 1 << 16; char **vocab = (char **)calloc(vcap, sizeof(char *));
    char *line = NULL; size_t cap = 0; int32_t last = -1;
    iv_push(&sent, 0);
    while (getline(&line, &cap, fh) != -1) {
        size_t l = strlen(line);
        if (l && line[l - 1] == '\n') line[l - 1] = 0;
        int local = 0;
        char *t = strtok(line, " ");
        while (t != NULL && !isspace((unsigned char)*t)) {      /* Start.cu:279 */
            size_t tl = strlen(t);
            if (tl && t[tl - 1] == '\n') t[tl - 1] = 0;
            int32_t id = smap_get(map, t);
            if (id < 0) {
                id = (int32_t)map->cnt + 2; last = id;
                char *cp = strdup(t);
                smap_put(map, cp, id);
                if (id >= vcap) { vocab = (char **)realloc(vocab, sizeof(char *) * vcap * 2); memset(vocab + vcap, 0, sizeof(char *) * vcap); vcap *= 2; }
                vocab[id] = cp;
            }
            iv_push(&tok, id); iv_push(&Pv, local & UNAL); local++;
            t = strtok(NULL, " ");
        }
        iv_push(&tok, 1); iv_push(&Pv, 0);
        iv_push(&sent, (int32_t)tok.n);
    }
    free(line); fclose(fh);
    iv_push(&tok, 1); iv_push(&Pv, 0);
    last++;
    iv_push(&tok, last); iv_push(&Pv, 0);
    int32_t n = (int32_t)tok.n;
    iv_push(&tok, 0); iv_push(&tok, 0); iv_push(&tok, 0);
    *tok_out = tok.v; *n_out = n;
    if (P_out) { lr_t *P = (lr_t *)malloc(sizeof(lr_t) * (size_t)n); for (int32_t i = 0; i < n; i++) P[i] = (lr_t)Pv.v[i]; *P_out = P; }
    iv_free(&Pv);
    *sent_out = sent.v; *nsent_out = (int32_t)sent.n - 1;
    *vocab_out = vocab; *nvocab = (int32_t)map->cnt + 2;
    return 1;
}

/* ExtractPair.cu:2639-2739 initAlignment */
static int load_alignment(const char *path, int32_t n, int32_t m, const lr_t *P, const int32_t *ssent, int32_t ns,
                          const int32_t *tsent, int32_t nt, rlp_t **RLP_out, lr_t **L_out, lr_t **R_out) {
    FILE *fh = fopen(path, "r");
    if (!fh) { fprintf(stderr, "oracle: cannot open %s\n", path); return 0; }
    lr_t *Lt = (lr_t *)malloc(sizeof(lr_t) * (size_t)m), *Rt = (lr_t *)malloc(sizeof(lr_t) * (size_t)m), *Ls = (lr_t *)malloc(sizeof(lr_t) * (size_t)n), *Rs = (lr_t *)malloc(sizeof(lr_t) * (size_t)n);
    memset(Lt, 255, sizeof(lr_t) * (size_t)m); memset(Rt, 255, sizeof(lr_t) * (size_t)m); memset(Ls, 255, sizeof(lr_t) * (size_t)n); memset(Rs, 255, sizeof(lr_t) * (size_t)n);
    char *line = NULL; size_t cap = 0; int qcount = -1;
    while (getline(&line, &cap, fh) != -1) {
        qcount++;
        if (qcount >= ns || qcount >= nt) break;
        size_t l = strlen(line);
        if (l && line[l - 1] == '\n') line[l - 1] = 0;
        char *t = strtok(line, " -");
        while (t != NULL && !isspace((unsigned char)*t)) {
            int s_no = atoi(t);
            t = strtok(NULL, " -");
            if (!t) { fprintf(stderr, "oracle: odd alignment line %d\n", qcount); return 0; }
            int t_no = atoi(t);
            if (s_no >= UNAL || t_no >= UNAL || s_no < 0 || t_no < 0) { fprintf(stderr, "oracle: sentence too long\n"); return 0; }
            int si = ssent[qcount] + s_no, ti = tsent[qcount] + t_no;
            if (Ls[si] == UNAL || Rs[si] == UNAL) { Ls[si] = (lr_t)t_no; Rs[si] = (lr_t)t_no; }
            else if (t_no > Rs[si]) Rs[si] = (lr_t)t_no;
            else if (t_no < Ls[si]) Ls[si] = (lr_t)t_no;
            if (Lt[ti] == UNAL || Rt[ti] == UNAL) { Lt[ti] = (lr_t)s_no; Rt[ti] = (lr_t)s_no; }
            else if (s_no > Rt[ti]) Rt[ti] = (lr_t)s_no;
            else if (s_no < Lt[ti]) Lt[ti] = (lr_t)s_no;
            t = strtok(NULL, " -");
        }
    }
    free(line); fclose(fh);
    rlp_t *RLP = (rlp_t *)calloc(n, sizeof(rlp_t));
    int q = 1;
    for (int32_t i = 0; i < n - 1; i++) {                         /* :2721 */
        if (q <= ns && i == ssent[q] - 1) { RLP[i] = (rlp_t)(uint32_t)tsent[q]; q++; }
        else RLP[i] = RLP_PACK(Ls[i], Rs[i], P[i]);
    }
    free(Ls); free(Rs);
    *RLP_out = RLP; *L_out = Lt; *R_out = Rt;
    return 1;
}

typedef struct { int64_t k; int32_t i; } ki_t;
static int cmp_ki(const void *x, const void *y) {
    const ki_t *p = (const ki_t *)x, *q = (const ki_t *)y;
    if (p->k != q->k) return p->k < q->k ? -1 : 1;
    return p->i < q->i ? -1 : p->i > q->i;
}
static void set_lex(orc_t *o, const int32_t *f, const int32_t *e, const float *v1, const float *v2, int32_t cnt) {
    /* sort by (ch, eng) -- ExtractPair.cu:28-35,2537 */
    if (cnt < 0) cnt = 0;
    const size_t room = cnt > 0 ? (size_t)cnt : 1;
    ki_t *a = (ki_t *)malloc(sizeof(ki_t) * room);
    for (int32_t i = 0; i < cnt; i++) { a[i].k = ((int64_t)f[i] << 32) + (int64_t)e[i] + (1LL << 31); a[i].i = i; }
    qsort(a, (size_t)cnt, sizeof(ki_t), cmp_ki);
    o->lex_key = (int64_t *)malloc(sizeof(int64_t) * room);
    o->lex_v1 = (float *)malloc(sizeof(float) * room);
    o->lex_v2 = (float *)malloc(sizeof(float) * room);
    for (int32_t i = 0; i < cnt; i++) { o->lex_key[i] = a[i].k; o->lex_v1[i] = v1[a[i].i]; o->lex_v2[i] = v2[a[i].i]; }
    o->lex_count = cnt;
    free(a);
}

/* ExtractPair.cu:2442-2554 initWordPossibilityIntKey */
static int load_lex(orc_t *o, const char *path) {
    FILE *fh = fopen(path, "r");
    if (!fh) { fprintf(stderr, "oracle: cannot open %s\n", path); return 0; }
    ivec F = {0}, E = {0}; int64_t cap = 1024, cnt = 0; float *v1 = (float *)malloc(sizeof(float) * cap), *v2 = (float *)malloc(sizeof(float) * cap);
    char a[4096], b[4096]; float x, y;
    while (fscanf(fh, "%4095s %4095s %f %f", a, b, &x, &y) == 4) {
        int32_t fi = smap_get(&o->smap_src, a), ei = smap_get(&o->smap_tgt, b);
        if (fi < 0 && strcmp(a, "NULL")) continue;               /* :2474-2478 */
        if (ei < 0 && strcmp(b, "NULL")) continue;               /* :2483-2487 */
        if (cnt == cap) { cap *= 2; v1 = (float *)realloc(v1, sizeof(float) * cap); v2 = (float *)realloc(v2, sizeof(float) * cap); }
        iv_push(&F, fi < 0 ? -1 : fi); iv_push(&E, ei < 0 ? -1 : ei); v1[cnt] = x; v2[cnt] = y; cnt++;
    }
    fclose(fh);
    set_lex(o, F.v, E.v, v1, v2, (int32_t)cnt);
    iv_free(&F); iv_free(&E); free(v1); free(v2);
    return 1;
}

/* ExtractPair.cu:2108-2142 searchLexFile: value of key (f,e) or 0 when absent (the reference's
 * inclusive [0,count] range and unsigned wrap are hazards, not semantics). */
static float lex_get(const orc_t *o, int32_t f, int32_t e, int one) {
    int64_t k = ((int64_t)f << 32) + (int64_t)e + (1LL << 31);
    int32_t lo = 0, hi = o->lex_count - 1;
    while (lo <= hi) {
        int32_t mid = lo + (hi - lo) / 2;
        if (k < o->lex_key[mid]) hi = mid - 1; else if (k > o->lex_key[mid]) lo = mid + 1;
        else return one ? o->lex_v1[mid] : o->lex_v2[mid];
    }
    return 0.0f;
}

int orc_is_wide(void) { return UNAL != 255; }

/* ------------------------------------------------------------------------------------------ */
orc_t *orc_create(const int32_t *str, int32_t n, const int32_t *tgt, int32_t m, const rlp_t *RLP, const lr_t *L_tar,
                  const lr_t *R_tar, const int32_t *lex_f, const int32_t *lex_e, const float *lex_v1, const float *lex_v2,
                  int32_t lex_count) {
    orc_t *o = (orc_t *)calloc(1, sizeof(orc_t));
    o->n = n; o->m = m;
    o->str = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 3)); memcpy(o->str, str, sizeof(int32_t) * ((size_t)n + 3));
    o->tgt = (int32_t *)malloc(sizeof(int32_t) * ((size_t)m + 3)); memcpy(o->tgt, tgt, sizeof(int32_t) * ((size_t)m + 3));
    o->RLP = (rlp_t *)malloc(sizeof(rlp_t) * (size_t)n); memcpy(o->RLP, RLP, sizeof(rlp_t) * (size_t)n);
    o->L_tar = (lr_t *)malloc(sizeof(lr_t) * (size_t)m); memcpy(o->L_tar, L_tar, sizeof(lr_t) * (size_t)m);
    o->R_tar = (lr_t *)malloc(sizeof(lr_t) * (size_t)m); memcpy(o->R_tar, R_tar, sizeof(lr_t) * (size_t)m);
    set_lex(o, lex_f, lex_e, lex_v1, lex_v2, lex_count);
    return o;
}

orc_t *orc_create_from_files(const char *src, const char *tgt, const char *align, const char *lex) {
    orc_t *o = (orc_t *)calloc(1, sizeof(orc_t));
    lr_t *P = NULL; int32_t *ssent = NULL, *tsent = NULL; int32_t ns = 0, nt = 0;
    if (!load_side(src, &o->smap_src, &o->svocab, &o->sv, &o->str, &o->n, &P, &ssent, &ns)) return NULL;
    if (!load_side(tgt, &o->smap_tgt, &o->tvocab, &o->tv, &o->tgt, &o->m, NULL, &tsent, &nt)) return NULL;
    o->have_vocab = 1;
    if (!load_lex(o, lex)) return NULL;
    if (!load_alignment(align, o->n, o->m, P, ssent, ns, tsent, nt, &o->RLP, &o->L_tar, &o->R_tar)) return NULL;
    free(P); free(ssent); free(tsent);
    return o;
}

void orc_set_sa(orc_t *o, const int32_t *sa) {
    free(o->sa); o->sa = (int32_t *)malloc(sizeof(int32_t) * (size_t)o->n); memcpy(o->sa, sa, sizeof(int32_t) * (size_t)o->n);
    o->have_precomp = 0;
}

/* Suffix array by prefix doubling with LSD radix passes on (rank[i], rank[i+h]).  Same unique
 * answer as SuffixArray.c:51 suffixArrayInt (plain lexicographic order on ids with 0 padding;
 * the final symbol is unique so all suffixes differ). */
void orc_build_sa(orc_t *o) {
    int32_t n = o->n;
    int32_t *sa = (int32_t *)malloc(sizeof(int32_t) * (size_t)n), *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *rk = (int32_t *)malloc(sizeof(int32_t) * (size_t)n), *nrk = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t maxv = 0;
    for (int32_t i = 0; i < n; i++) { rk[i] = o->str[i]; if (rk[i] > maxv) maxv = rk[i]; }
    int64_t K = (int64_t)(maxv > n ? maxv : n) + 2;
    int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * (size_t)(K + 1));
    for (int32_t i = 0; i < n; i++) sa[i] = i;
    for (int32_t h = 0;; h = h ? h * 2 : 1) {
        /* sort by second key rank[i+h] (0 beyond the end), then stable by first key rank[i] */
        if (h > 0) {
            memset(cnt, 0, sizeof(int32_t) * (size_t)(K + 1));
            for (int32_t i = 0; i < n; i++) { int32_t k2 = (i + h < n) ? rk[i + h] + 1 : 0; cnt[k2 + 1]++; }
            for (int64_t k = 0; k < K; k++) cnt[k + 1] += cnt[k];
            for (int32_t i = 0; i < n; i++) { int32_t p = sa[i]; int32_t k2 = (p + h < n) ? rk[p + h] + 1 : 0; tmp[cnt[k2]++] = p; }
        } else memcpy(tmp, sa, sizeof(int32_t) * (size_t)n);
        memset(cnt, 0, sizeof(int32_t) * (size_t)(K + 1));
        for (int32_t i = 0; i < n; i++) cnt[rk[i] + 1]++;
        for (int64_t k = 0; k < K; k++) cnt[k + 1] += cnt[k];
        for (int32_t i = 0; i < n; i++) { int32_t p = tmp[i]; sa[cnt[rk[p]]++] = p; }
        /* new ranks */
        int32_t r = 0; nrk[sa[0]] = 0;
        for (int32_t i = 1; i < n; i++) {
            int32_t a = sa[i - 1], b = sa[i];
            int32_t a2 = (h && a + h < n) ? rk[a + h] + 1 : 0, b2 = (h && b + h < n) ? rk[b + h] + 1 : 0;
            if (rk[a] != rk[b] || a2 != b2) r++;
            nrk[b] = r;
        }
        int32_t *t = rk; rk = nrk; nrk = t;
        if (r == n - 1) break;
    }
    free(tmp); free(rk); free(nrk); free(cnt);
    free(o->sa); o->sa = sa; o->have_precomp = 0;
}

/* ------------------------------------------------------------------------------------------ */
/* alignment-consistency helpers                                                                */
/* ------------------------------------------------------------------------------------------ */
#define RL(o, k) RLP_L((o)->RLP[k])
#define RR(o, k) RLP_R((o)->RLP[k])
#define RP(o, k) RLP_P((o)->RLP[k])

/* ExtractPair.cu:103-133 consistent */
static int consistent(const orc_t *o, int start, int end, int start_chk, int end_chk, int startpos_source) {
    lr_t min_L = UNAL, max_R = 0;
    for (int k = start; k <= end; k++) {
        lr_t L = o->L_tar[k], R = o->R_tar[k];
        if (L == UNAL || R == UNAL) { }
        else if (k == start) { min_L = L; max_R = R; }
        else { if (min_L > L) min_L = L; if (max_R < R) max_R = R; }
    }
    if (startpos_source + min_L != start_chk || startpos_source + max_R != end_chk) return 0;
    return 1;
}

/* GappyLook.cu:43-126 checkBoundaryGap */
static int checkBoundaryGap(const orc_t *o, unsigned int start, unsigned int ender) {
    lr_t L, R, min_L = UNAL, max_R = 0; int sen_target_begin = -1, tempind = 0;
    for (int k = (int)start; k <= (int)ender; k++) {
        rlp_t temp = o->RLP[k]; L = RLP_L(temp); R = RLP_R(temp);
        if ((L == UNAL || R == UNAL) && (k == (int)start || k == (int)ender)) return 0;
        else if (L == UNAL || R == UNAL) { }
        else if (k == (int)start) {
            tempind = k - RLP_P(temp) - 1;
            sen_target_begin = tempind == -1 ? 0 : (int)o->RLP[tempind];
            min_L = L; max_R = R;
        } else { if (min_L > L) min_L = L; if (max_R < R) max_R = R; }
    }
    if (min_L <= max_R && max_R - min_L < ORC_MAX_RULE_SPAN) {
        tempind++;
        int target_start = min_L + sen_target_begin, target_end = max_R + sen_target_begin;
        min_L = UNAL; max_R = 0;
        for (int k = target_start; k <= target_end; k++) {
            L = o->L_tar[k]; R = o->R_tar[k];
            if (L == UNAL || R == UNAL) { }
            else if (k == target_start) { min_L = L; max_R = R; }
            else { if (min_L > L) min_L = L; if (max_R < R) max_R = R; }
        }
        if (tempind + min_L != (int)start || tempind + max_R != (int)ender) return 0;
        return 1;
    }
    return 0;
}

/* ExtractPair.cu:135-194 checkBoundaryFast */
static int checkBoundaryFast(const orc_t *o, unsigned int start, unsigned int ender, lr_t *min_LL, lr_t *max_RR,
                             int *sen_target_begin, int *tempind) {
    lr_t L, R, min_L = UNAL, max_R = 0;
    *sen_target_begin = -1; *tempind = 0;
    for (int k = (int)start; k <= (int)ender; k++) {
        rlp_t temp = o->RLP[k]; L = RLP_L(temp); R = RLP_R(temp);
        if ((L == UNAL || R == UNAL) && (k == (int)start || k == (int)ender)) return 0;
        else if (L == UNAL || R == UNAL) { }
        else if (k == (int)start) {
            *tempind = k - RLP_P(temp) - 1;
            *sen_target_begin = (*tempind == -1) ? 0 : (int)o->RLP[*tempind];
            min_L = L; max_R = R;
        } else { if (min_L > L) min_L = L; if (max_R < R) max_R = R; }
    }
    if (min_L <= max_R && max_R - min_L < ORC_MAX_RULE_SPAN) { (*tempind)++; *min_LL = min_L; *max_RR = max_R; return 1; }
    return 0;
}

/* ExtractPair.cu:196-250 checkBoundaryFast2 */
static int checkBoundaryFast2(const orc_t *o, unsigned int start, unsigned int ender, unsigned int *target_start, unsigned int *target_end) {
    lr_t L, R, min_L = UNAL, max_R = 0; int sen_target_begin = -1, tempind = 0;
    for (int k = (int)start; k <= (int)ender; k++) {
        rlp_t temp = o->RLP[k]; L = RLP_L(temp); R = RLP_R(temp);
        if ((L == UNAL || R == UNAL) && (k == (int)start || k == (int)ender)) return 0;
        else if (L == UNAL || R == UNAL) { }
        else if (k == (int)start) {
            tempind = k - RLP_P(temp) - 1;
            sen_target_begin = tempind == -1 ? 0 : (int)o->RLP[tempind];
            min_L = L; max_R = R;
        } else { if (min_L > L) min_L = L; if (max_R < R) max_R = R; }
    }
    *target_start = (unsigned int)(min_L + sen_target_begin); *target_end = (unsigned int)(max_R + sen_target_begin);
    if (min_L <= max_R && max_R - min_L < ORC_MAX_RULE_SPAN) return 1;
    return 0;
}

/* ExtractPair.cu:252-342 checkBoundary (error codes 0..4) */
static int checkBoundary(const orc_t *o, unsigned int start, unsigned int ender, unsigned int *target_start, unsigned int *target_end) {
    lr_t L, R, min_L = UNAL, max_R = 0; int sen_target_begin = -1, tempind = 0; int wrong = 0;
    for (int k = (int)start; k <= (int)ender; k++) {
        rlp_t temp = o->RLP[k]; L = RLP_L(temp); R = RLP_R(temp);
        if ((L == UNAL || R == UNAL) && (k == (int)start || k == (int)ender)) {
            if (start == ender && wrong == 0) wrong = 4;
            else if (wrong == 0 && k == (int)start) wrong = 2;
            else if (wrong == 0 && k == (int)ender) wrong = 3;
            else if (wrong != 0) wrong = 4;
            if (k == (int)start) {
                tempind = k - RLP_P(temp) - 1;
                sen_target_begin = tempind == -1 ? 0 : (int)o->RLP[tempind];
            }
        } else if (L == UNAL || R == UNAL) { }
        else if (k == (int)start) {
            tempind = k - RLP_P(temp) - 1;
            sen_target_begin = tempind == -1 ? 0 : (int)o->RLP[tempind];
            min_L = L; max_R = R;
        } else { if (min_L > L) min_L = L; if (max_R < R) max_R = R; }
    }
    *target_start = (unsigned int)(min_L + sen_target_begin); *target_end = (unsigned int)(max_R + sen_target_begin);
    if (wrong) return wrong;
    if (min_L <= max_R && max_R - min_L < ORC_MAX_RULE_SPAN) {
        tempind++;
        if (consistent(o, (int)*target_start, (int)*target_end, (int)start, (int)ender, tempind)) return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* frequent-pair precomputation                                                                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int32_t token, length; } toptok;

static int freq_index(const orc_t *o, int32_t tok) {       /* GappyLook.cu:5-40 existPrecomputation (one side) */
    int lo = 0, hi = ORC_PRECOMP - 1;
    while (hi - lo >= 0) { int mid = (lo + hi) >> 1; if (o->freq[mid] > tok) hi = mid - 1; else if (o->freq[mid] < tok) lo = mid + 1; else return mid; }
    return -1;
}
static int existPrecomputation(const orc_t *o, int32_t a, int32_t b) {
    int ia = freq_index(o, a), ib = freq_index(o, b);
    if (ia >= 0 && ib >= 0) return ia * ORC_PRECOMP + ib;
    return -1;
}

/* SuffixArray.cu:1132-1340 preComputation + GappyLook.cu:740-870 precomp.
 * Top-100 tokens by corpus frequency (descending; ties keep SA order = ascending id -- the intended
 * order of compareUserTotal1 under glibc's stable merge sort), stored ascending by id. Then for every
 * ordered pair the collocations a ... b (gap >= 1, all tokens >= 2, b at offset 2..14), split into the
 * ones whose gap passes checkBoundaryGap (the list, sorted by (pair,start,len)) and the rest (counted
 * in featureMissingCount). Both scan directions of the kernel enumerate the same set. */
static int build_precomp(orc_t *o) {
    int32_t n = o->n;
    int32_t maxtok = 0;
    for (int32_t i = 0; i < n; i++) if (o->str[i] > maxtok) maxtok = o->str[i];
    int32_t *count = (int32_t *)calloc((size_t)maxtok + 1, sizeof(int32_t));
    for (int32_t i = 0; i < n; i++) if (o->str[i] >= 2) count[o->str[i]]++;
    int32_t distinct = 0;
    for (int32_t t = 2; t <= maxtok; t++) if (count[t]) distinct++;
    if (distinct < ORC_PRECOMP) { fprintf(stderr, "oracle: need >= %d distinct source tokens (SuffixArray.cu:1175-1176), have %d\n", ORC_PRECOMP, distinct); free(count); return 0; }
    toptok *tl = (toptok *)malloc(sizeof(toptok) * (size_t)distinct);
    int32_t c = 0;
    for (int32_t t = 2; t <= maxtok; t++) if (count[t]) { tl[c].token = t; tl[c].length = count[t]; c++; }
    /* stable sort by length descending (insertion of top 100 only) */
    int32_t best[ORC_PRECOMP]; int nb = 0;
    {
        /* selection: repeatedly take the max length with the smallest token id not yet taken */
        char *taken = (char *)calloc((size_t)distinct, 1);
        for (nb = 0; nb < ORC_PRECOMP; nb++) {
            int32_t bi = -1;
            for (int32_t i = 0; i < distinct; i++) if (!taken[i] && (bi < 0 || tl[i].length > tl[bi].length)) bi = i;
            taken[bi] = 1; best[nb] = bi;
        }
        free(taken);
    }
    /* ascending token id (compareUserTotal2) */
    for (int i = 0; i < ORC_PRECOMP; i++) o->freq[i] = tl[best[i]].token;
    for (int i = 1; i < ORC_PRECOMP; i++) { int32_t v = o->freq[i]; int j = i - 1; while (j >= 0 && o->freq[j] > v) { o->freq[j + 1] = o->freq[j]; j--; } o->freq[j + 1] = v; }
    free(tl);
    int32_t *fidx = (int32_t *)malloc(sizeof(int32_t) * ((size_t)maxtok + 1));
    for (int32_t t = 0; t <= maxtok; t++) fidx[t] = -1;
    for (int i = 0; i < ORC_PRECOMP; i++) fidx[o->freq[i]] = i;
    memset(o->missing, 0, sizeof(o->missing));
    ivec trip = {0};  /* pair, start, len */
    for (int32_t p = 0; p < n; p++) {
        int32_t ta = o->str[p]; if (ta < 2 || fidx[ta] < 0) continue;
        if (o->str[p + ORC_MIN_GAP] < 2) continue;                        /* GappyLook.cu:788-792 */
        for (int move = 0;; move++) {
            int32_t tb = o->str[p + 1 + ORC_MIN_GAP + move];
            if (tb < 2) break;
            if (fidx[tb] >= 0) {
                int pair = fidx[ta] * ORC_PRECOMP + fidx[tb];
                if (checkBoundaryGap(o, (unsigned)p + 1, (unsigned)(p + move + 1 + ORC_MIN_GAP - 1))) {
                    iv_push(&trip, pair); iv_push(&trip, p); iv_push(&trip, move + 1 + ORC_MIN_GAP);
                } else o->missing[pair]++;
            }
            if (1 + ORC_MIN_GAP + (move + 1) + 1 > ORC_MAX_RULE_SPAN) break;   /* :818-821 */
        }
    }
    int32_t cntp = (int32_t)(trip.n / 3);
    /* counting sort by pair, keeps (start,len) order */
    int32_t *pc = (int32_t *)calloc(ORC_PRECOMP * ORC_PRECOMP + 1, sizeof(int32_t));
    for (int32_t i = 0; i < cntp; i++) pc[trip.v[3 * i] + 1]++;
    for (int i = 0; i < ORC_PRECOMP * ORC_PRECOMP; i++) pc[i + 1] += pc[i];
    free(o->plist); o->plist = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(cntp ? cntp : 1));
    for (int i = 0; i < ORC_PRECOMP * ORC_PRECOMP; i++) {
        if (pc[i + 1] > pc[i]) { o->pidx[2 * i] = pc[i]; o->pidx[2 * i + 1] = pc[i + 1] - 1; }
        else { o->pidx[2 * i] = 1; o->pidx[2 * i + 1] = 0; }                 /* SuffixArray.cu:1306-1307 */
    }
    int32_t *cur = (int32_t *)malloc(sizeof(int32_t) * ORC_PRECOMP * ORC_PRECOMP);
    memcpy(cur, pc, sizeof(int32_t) * ORC_PRECOMP * ORC_PRECOMP);
    for (int32_t i = 0; i < cntp; i++) { int32_t k = cur[trip.v[3 * i]]++; o->plist[2 * k] = trip.v[3 * i + 1]; o->plist[2 * k + 1] = trip.v[3 * i + 2]; }
    o->pcount = cntp;
    free(cur); free(pc); free(count); free(fidx); iv_free(&trip);
    o->have_precomp = 1;
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* contiguous lookup                                                                            */
/* ------------------------------------------------------------------------------------------ */
/* Result of SuffixArray.cu:402-767 (longestmatch, 1-gram interval) and :109-400 (n-gram intervals):
 * longest prefix of q[t..) found in the corpus (stops at query end / OOV -1 / corpus EOS 1) and,
 * for every m <= longest, the inclusive SA interval of q[t..t+m).  Both are search-path independent,
 * so the restatement narrows the interval one token at a time. */
static void lookup(orc_t *o) {
    int32_t T = o->T;
    o->longest = (int32_t *)calloc((size_t)T + 1, sizeof(int32_t));
    o->conn_off = (int32_t *)malloc(sizeof(int32_t) * ((size_t)T + 1));
    ivec up = {0}, down = {0};
    for (int32_t qi = 0; qi < o->Q; qi++) {
        int32_t end = o->qoff[qi + 1];
        for (int32_t t = o->qoff[qi]; t < end; t++) {
            o->conn_off[t] = (int32_t)up.n;
            int32_t lo = 0, hi = o->n - 1, mlen = 0;
            while (t + mlen < end && o->q[t + mlen] != -1) {
                int32_t tok = o->q[t + mlen];
                int32_t a = lo, b = hi + 1;                 /* lower bound of tok at offset mlen */
                while (a < b) { int32_t mid = a + (b - a) / 2; if (o->str[o->sa[mid] + mlen] < tok) a = mid + 1; else b = mid; }
                int32_t l = a; b = hi + 1;
                while (a < b) { int32_t mid = a + (b - a) / 2; if (o->str[o->sa[mid] + mlen] <= tok) a = mid + 1; else b = mid; }
                int32_t r = a - 1;
                if (l > r) break;
                lo = l; hi = r; mlen++;
                iv_push(&up, lo); iv_push(&down, hi);
            }
            o->longest[t] = mlen;
        }
    }
    o->conn_off[T] = (int32_t)up.n;
    o->iv_up = up.v; o->iv_down = down.v;
}
static inline void interval(const orc_t *o, int32_t t, int32_t mlen, int32_t *up, int32_t *down) {
    *up = o->iv_up[o->conn_off[t] + mlen - 1]; *down = o->iv_down[o->conn_off[t] + mlen - 1];
}

/* ExtractPair.cu:2742-2903 GenerateBlocks: distinct (up,down,len<=5), ids in first-appearance order
 * (query, token, length); per-query id lists de-duplicated in first-appearance order. */
static void generate_blocks(orc_t *o) {
    kmap map; kmap_init(&map, 1 << 16);
    ivec blk = {0};
    o->qryglobal = (ivec *)calloc((size_t)o->Q, sizeof(ivec));
    int32_t G = 0;
    ivec stamp = {0};
    for (int32_t qi = 0; qi < o->Q; qi++) {
        for (int32_t j = o->qoff[qi]; j < o->qoff[qi + 1]; j++) {
            for (int32_t ct = 1; ct <= o->longest[j] && ct <= ORC_LONGEST_SRC; ct++) {
                int32_t up, down; interval(o, j, ct, &up, &down);
                int found; int32_t *slot = kmap_slot(&map, ((uint64_t)(uint32_t)up << 8) | (uint64_t)ct, &found);
                if (!found) {
                    *slot = G;
                    iv_push(&blk, up); iv_push(&blk, down); iv_push(&blk, ct); iv_push(&blk, o->sa[up]);
                    iv_push(&stamp, -1);
                    G++;
                }
                int32_t id = *slot;
                if (stamp.v[id] != qi) { stamp.v[id] = qi; iv_push(&o->qryglobal[qi], id); }
            }
        }
    }
    o->G = G; o->blocks = blk.v;
    kmap_free(&map); iv_free(&stamp);
}

/* ------------------------------------------------------------------------------------------ */
/* one-gap enumeration, dedup, lookup                                                           */
/* ------------------------------------------------------------------------------------------ */
static void pattern_of(const orc_t *o, int32_t inst, int32_t pat[ORC_MAX_RULE_SYMBOLS], int *number) {
    int32_t t = o->g1_start[inst]; int ls = o->g1_ls[inst], le = o->g1_le[inst]; int32_t s2 = t + ls + o->g1_gap[inst];
    int k = 0;
    for (int i = 0; i < ls; i++) pat[k++] = o->q[t + i];
    pat[k++] = -1;
    for (int i = 0; i < le; i++) pat[k++] = o->q[s2 + i];
    *number = k;
    while (k < ORC_MAX_RULE_SYMBOLS) pat[k++] = -2;
}

static const orc_t *g_sort_ctx;
static int cmp_inst1(const void *a, const void *b) {           /* SuffixArray.cu:51-67, ties by enumeration order */
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    int32_t px[ORC_MAX_RULE_SYMBOLS], py[ORC_MAX_RULE_SYMBOLS]; int nx, ny;
    pattern_of(g_sort_ctx, x, px, &nx); pattern_of(g_sort_ctx, y, py, &ny);
    if (nx != ny) return nx < ny ? -1 : 1;
    for (int i = 0; i < nx; i++) if (px[i] != py[i]) return px[i] < py[i] ? -1 : 1;
    return x < y ? -1 : x > y;
}

/* SuffixArray.cu:928-1039 oneGapEnumeration (+ :1598 sort, :1041 zeroOneDiff, :1670-1719 host scan) */
static void onegap_enumerate(orc_t *o) {
    ivec st = {0}, ls_v = {0}, le_v = {0}, gp = {0};
    int32_t T = o->T;
    for (int32_t tokindex = 0; tokindex < T - 1; tokindex++) {          /* :945 */
        int32_t qi = o->tok2q[tokindex]; int32_t end = o->qoff[qi + 1];
        if (tokindex == end - 1 || tokindex == end - 2) continue;       /* :957-962 */
        int longest_len_start = o->longest[tokindex];
        for (int ls = 1; ls <= longest_len_start; ls++) {
            for (int32_t s = tokindex + ls + ORC_MIN_GAP; s < end && s - tokindex <= ORC_MAX_RULE_SPAN; s++) {
                if (o->q[s] == -1) continue;
                int longest_len_end = o->longest[s];
                for (int le = 1; ls + 1 + le <= ORC_MAX_RULE_SYMBOLS && le <= longest_len_end && s - tokindex + le - 1 <= ORC_MAX_RULE_SPAN; le++) {
                    iv_push(&st, tokindex); iv_push(&ls_v, ls); iv_push(&le_v, le); iv_push(&gp, s - tokindex - ls);
                }
            }
        }
    }
    int32_t E = (int32_t)st.n;
    o->enu1 = E; o->g1_start = st.v;
    o->g1_ls = (uint8_t *)malloc((size_t)E + 1); o->g1_le = (uint8_t *)malloc((size_t)E + 1); o->g1_gap = (uint8_t *)malloc((size_t)E + 1);
    for (int32_t i = 0; i < E; i++) { o->g1_ls[i] = (uint8_t)ls_v.v[i]; o->g1_le[i] = (uint8_t)le_v.v[i]; o->g1_gap[i] = (uint8_t)gp.v[i]; }
    iv_free(&ls_v); iv_free(&le_v); iv_free(&gp);
    o->g1_sorted = (int32_t *)malloc(sizeof(int32_t) * ((size_t)E + 1));
    for (int32_t i = 0; i < E; i++) o->g1_sorted[i] = i;
    g_sort_ctx = o; qsort(o->g1_sorted, (size_t)E, sizeof(int32_t), cmp_inst1);
    /* distinct patterns */
    ivec pat = {0}, pos = {0}, rep = {0};
    o->q1 = (ivec *)calloc((size_t)o->Q, sizeof(ivec));
    int32_t D = 0; int32_t prev[ORC_MAX_RULE_SYMBOLS]; int pn = -1;
    int32_t *lastq = NULL;
    for (int32_t i = 0; i < E; i++) {
        int32_t inst = o->g1_sorted[i]; int32_t p[ORC_MAX_RULE_SYMBOLS]; int nn;
        pattern_of(o, inst, p, &nn);
        int diff = (i == 0) || nn != pn || memcmp(p, prev, sizeof(int32_t) * (size_t)nn);
        if (diff) {
            for (int k = 0; k < ORC_MAX_RULE_SYMBOLS; k++) iv_push(&pat, p[k]);
            iv_push(&pat, nn); iv_push(&pat, o->g1_ls[inst]); iv_push(&pat, o->g1_le[inst]); iv_push(&pat, -1); iv_push(&pat, -1);
            iv_push(&pos, i); iv_push(&rep, inst);
            memcpy(prev, p, sizeof(prev)); pn = nn; D++;
            lastq = NULL;
        }
        int32_t qid = o->tok2q[o->g1_start[inst]];
        /* per-query list without duplicates (checkDup, :1713-1718) */
        ivec *lst = &o->q1[qid];
        if (lst->n == 0 || lst->v[lst->n - 1] != D - 1) iv_push(lst, D - 1);
        (void)lastq;
    }
    iv_push(&pos, E);
    o->D1 = D; o->pat1 = pat.v; o->pat1_pos = pos.v; o->pat1_rep = rep.v;
}

static int cmp_hit3(const void *a, const void *b) {
    const int32_t *x = (const int32_t *)a, *y = (const int32_t *)b;
    for (int i = 0; i < 3; i++) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return 0;
}
/* Two-gap hits: (pattern, position) as twoGapSACompare (SuffixArray.cu:81-89) orders them; hits that tie on both keep the
 * order in which twoGapLookUpSA's atomicAdd cursor handed out their places (thrust's merge sort is stable).  The parents of
 * one (pattern, position) are adjacent entries of the one-gap list, i.e. adjacent lanes of one warp, and a warp that walks
 * `move` in lock step emits them by (move, lane) = (width of the second gap, length of the parent).  Measured against the
 * reference on a B200 (tools/ref_dump_config.py small: 284 962 hits, 48 535 tie groups): the hit SETS are identical; tie groups
 * whose parents come from the sorted one-gap list (GappyLook.cu:656-737) have exactly this order (530 of 530); groups whose
 * parents come from the frequent-pair list (:575-654) follow the warp scheduler (77 % this order, 63 % the (length, width)
 * order the oracle used before) -- they are the remaining disagreement of the gappy sample sets. */
static int cmp_hit4(const void *a, const void *b) {
    const int32_t *x = (const int32_t *)a, *y = (const int32_t *)b;
    for (int i = 0; i < 2; i++) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    const int mx = x[3] - x[2], my = y[3] - y[2];
    if (mx != my) return mx < my ? -1 : 1;
    for (int i = 2; i < 4; i++) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return 0;
}

/* GappyLook.cu:128-474 oneGapLookUpSA, all three strategies, + SuffixArray.cu:1836 sort
 * (ties on equal (id,pos) are unordered in the reference; the oracle orders them by length)
 * + :1854-1875 start/end_on_salist. */
static void onegap_lookup(orc_t *o) {
    ivec h = {0};
    const int32_t *str = o->str;
    for (int32_t d = 0; d < o->D1; d++) {
        int32_t rep = o->pat1_rep[d];
        int ls = o->g1_ls[rep], le = o->g1_le[rep];
        int32_t tokindex = o->g1_start[rep]; int32_t search_tokindex = tokindex + o->g1_gap[rep] + ls;
        int pre = existPrecomputation(o, o->q[tokindex + ls - 1], o->q[search_tokindex]);
        int32_t tb_start, tb_end, dis; int forward = 1;
        if (pre == -1) {
            int32_t u1, d1, u2, d2; interval(o, tokindex, ls, &u1, &d1); interval(o, search_tokindex, le, &u2, &d2);
            if (d1 - u1 <= d2 - u2) { tb_start = u1; tb_end = d1; dis = d1 - u1; forward = 1; }
            else { tb_start = u2; tb_end = d2; dis = d2 - u2; forward = 0; }
        } else { tb_start = o->pidx[2 * pre]; tb_end = o->pidx[2 * pre + 1]; dis = tb_end - tb_start; }
        (void)tb_end;
        if (pre != -1 && ls == 1 && le == 1 && dis >= 0) {                 /* marker record, :258-272 */
            iv_push(&h, d); iv_push(&h, pre); iv_push(&h, 0);
            continue;
        }
        for (int32_t tx = 0; tx <= dis; tx++) {
            if (pre != -1) {                                                /* :289-334 */
                int32_t ps = o->plist[2 * (tb_start + tx)], pl = o->plist[2 * (tb_start + tx) + 1];
                int ok = 1;
                if (pl + 1 + ls - 1 + le - 1 > ORC_MAX_RULE_SPAN) ok = 0;
                if (ok && ls > 1) {
                    int backoff = 0, stop = 0;
                    while (ok && !stop) {
                        backoff++;
                        if (ps - backoff < 0 || str[ps - backoff] != o->q[tokindex + ls - 1 - backoff]) ok = 0;
                        if (ls - backoff <= 1) stop = 1;
                    }
                }
                if (ok && le > 1) {
                    int fw = 1;
                    while (fw < le && ok) { fw++; if (str[ps + pl + fw - 1] != o->q[search_tokindex + fw - 1]) ok = 0; }
                }
                if (ok) { iv_push(&h, d); iv_push(&h, ps - ls + 1); iv_push(&h, pl + ls - 1 + le - 1); }
            } else if (forward) {                                           /* :335-396 */
                int32_t gostart = o->sa[tx + tb_start]; int flager = 1;
                for (int move = 0; flager; ) {
                    if (move == 0 && str[gostart + ls] < 2) flager = 0;
                    int32_t temp = str[gostart + ls + ORC_MIN_GAP + move];
                    if (temp < 2) flager = 0;
                    else if (flager && temp == o->q[search_tokindex]) {
                        int matchcount = 1, stop = 0;
                        while (!stop && matchcount < le && ls + ORC_MIN_GAP + move + 1 + matchcount <= ORC_MAX_RULE_SPAN) {
                            int32_t bk = str[gostart + ls + ORC_MIN_GAP + move + matchcount];
                            if (bk < 2) { stop = 1; flager = 0; }
                            else if (bk == o->q[search_tokindex + matchcount]) matchcount++;
                            else stop = 1;
                        }
                        if (matchcount == le && checkBoundaryGap(o, (unsigned)(gostart + ls), (unsigned)(gostart + ls + ORC_MIN_GAP + move + le - 1 - le))) {
                            iv_push(&h, d); iv_push(&h, gostart); iv_push(&h, ls + ORC_MIN_GAP + move + le - 1);
                        }
                    }
                    move++;
                    if (ls + ORC_MIN_GAP + move + le > ORC_MAX_RULE_SPAN) flager = 0;
                }
            } else {                                                        /* :397-469 */
                int32_t gostart = o->sa[tx + tb_start]; int flager = 1;
                for (int move = 0; flager; ) {
                    if (move == 0) { int32_t t0 = gostart - 1 >= 0 ? str[gostart - 1] : -1; if (t0 < 2) flager = 0; }
                    int32_t temp = (gostart - 1 - ORC_MIN_GAP - move < 0) ? -1 : str[gostart - 1 - ORC_MIN_GAP - move];
                    if (temp < 2) flager = 0;
                    else if (flager && temp == o->q[tokindex + ls - 1]) {
                        int matchcount = 1, stop = 0;
                        while (!stop && matchcount < ls && le + ORC_MIN_GAP + move + 1 + matchcount <= ORC_MAX_RULE_SPAN) {
                            int32_t bk = (gostart - 1 - ORC_MIN_GAP - move - matchcount < 0) ? -1 : str[gostart - 1 - ORC_MIN_GAP - move - matchcount];
                            if (bk < 2) { stop = 1; flager = 0; }
                            else if (bk == o->q[tokindex + ls - 1 - matchcount]) matchcount++;
                            else stop = 1;
                        }
                        if (matchcount == ls && checkBoundaryGap(o, (unsigned)(gostart - 1 - ORC_MIN_GAP - move + 1), (unsigned)(gostart - 1))) {
                            iv_push(&h, d); iv_push(&h, gostart - 1 - ORC_MIN_GAP - move - ls + 1); iv_push(&h, le + ORC_MIN_GAP + move + ls - 1);
                        }
                    }
                    move++;
                    if (ls + ORC_MIN_GAP + move + le > ORC_MAX_RULE_SPAN) flager = 0;
                }
            }
        }
    }
    o->hits1 = (int32_t)(h.n / 3);
    qsort(h.v, (size_t)o->hits1, sizeof(int32_t) * 3, cmp_hit3);
    o->h1 = h.v;
    for (int32_t i = 0; i < o->hits1; i++) {
        int32_t d = h.v[3 * i]; int32_t *p = &o->pat1[10 * d];
        if (p[8] == -1) p[8] = i;
        p[9] = i;
    }
}

/* SuffixArray.cu:816-926 twoGapEnumeration (+ :1989 sort, :1070 zeroOneDiffTwoGap, :2062-2097 scan).
 * Only patterns with one-token a and b qualify (limit_symbol = 5-2-ls-le >= 1, :840-850), c is one token. */
typedef struct { int32_t blockid, tok, gap2, order; } enu2_t;
static int cmp_enu2(const void *a, const void *b) {
    const enu2_t *x = (const enu2_t *)a, *y = (const enu2_t *)b;
    if (x->blockid != y->blockid) return x->blockid < y->blockid ? -1 : 1;
    if (x->tok != y->tok) return x->tok < y->tok ? -1 : 1;
    return x->order < y->order ? -1 : x->order > y->order;
}
static void twogap_enumerate(orc_t *o) {
    int64_t cap = 1024, cnt = 0; enu2_t *e = (enu2_t *)malloc(sizeof(enu2_t) * cap);
    for (int32_t d = 0; d < o->D1; d++) {
        const int32_t *p = &o->pat1[10 * d];
        if (p[8] == -1 || p[9] == -1) continue;
        int limit_symbol = ORC_MAX_RULE_SYMBOLS - 1 - 1 - p[6] - p[7];
        if (limit_symbol < 1) continue;
        for (int32_t k = o->pat1_pos[d]; k < o->pat1_pos[d + 1]; k++) {
            int32_t inst = o->g1_sorted[k];
            int32_t searchStart = o->g1_start[inst] + o->g1_ls[inst] + o->g1_gap[inst] + o->g1_le[inst] - 1;
            int32_t qi = o->tok2q[searchStart]; int32_t end = o->qoff[qi + 1];
            for (int32_t s = searchStart + ORC_MIN_GAP + 1; s < end; s++) {
                int longest_len_end = o->longest[s];
                for (int lc = 1; lc <= limit_symbol && lc <= longest_len_end && s - o->g1_start[inst] + lc - 1 <= ORC_MAX_RULE_SPAN; lc++) {
                    if (cnt == cap) { cap *= 2; e = (enu2_t *)realloc(e, sizeof(enu2_t) * cap); }
                    e[cnt].blockid = d; e[cnt].tok = o->q[s]; e[cnt].gap2 = s; e[cnt].order = (int32_t)cnt; cnt++;
                }
            }
        }
    }
    o->enu2 = (int32_t)cnt;
    qsort(e, (size_t)cnt, sizeof(enu2_t), cmp_enu2);
    ivec pat = {0}, rep = {0};
    o->q2 = (ivec *)calloc((size_t)o->Q, sizeof(ivec));
    int32_t D = 0;
    for (int64_t i = 0; i < cnt; i++) {
        if (i == 0 || e[i].blockid != e[i - 1].blockid || e[i].tok != e[i - 1].tok) {
            iv_push(&pat, e[i].blockid); iv_push(&pat, e[i].tok); iv_push(&pat, -1); iv_push(&pat, -1);
            iv_push(&rep, e[i].gap2); D++;
        }
        int32_t qid = o->tok2q[e[i].gap2];
        ivec *lst = &o->q2[qid];
        if (lst->n == 0 || lst->v[lst->n - 1] != D - 1) iv_push(lst, D - 1);
    }
    free(e);
    o->D2 = D; o->pat2 = pat.v; o->pat2_rep = rep.v;
}

/* GappyLook.cu:476-737 twoGapLookUpSA + SuffixArray.cu:2205 sort + :2214-2233 ranges */
static void twogap_lookup(orc_t *o) {
    ivec h = {0}; const int32_t *str = o->str;
    for (int32_t d2 = 0; d2 < o->D2; d2++) {
        int32_t d1 = o->pat2[4 * d2]; const int32_t *p1 = &o->pat1[10 * d1];
        int32_t startSA = p1[8], endSA = p1[9];
        if (startSA == -1 && endSA == -1) continue;
        int32_t dis = endSA - startSA + 1;
        int32_t preCache = o->pat2[4 * d2 + 1];
        int marker = (dis == 1 && o->h1[3 * startSA + 2] == 0);
        int32_t tb = 0;
        if (marker) { int32_t pre = o->h1[3 * startSA + 1]; dis = o->pidx[2 * pre + 1] - o->pidx[2 * pre] + 1; tb = o->pidx[2 * pre]; }
        for (int32_t tx = 0; tx < dis; tx++) {
            int32_t ps, pl;
            if (marker) { ps = o->plist[2 * (tb + tx)]; pl = o->plist[2 * (tb + tx) + 1]; }
            else { ps = o->h1[3 * (startSA + tx) + 1]; pl = o->h1[3 * (startSA + tx) + 2]; }
            int32_t gostart = ps + pl; int flager = 1;
            for (int move = 0; flager; move++) {
                if (move == 0 && str[gostart + ORC_MIN_GAP] < 2) flager = 0;
                int32_t temp = str[gostart + 1 + ORC_MIN_GAP + move];
                if (pl + 1 + ORC_MIN_GAP + move + 1 > ORC_MAX_RULE_SPAN) flager = 0;
                if (temp < 2) flager = 0;
                else if (flager && temp == preCache) {
                    /* c is a single token (longest_len_end == 1, :544-547) */
                    if (checkBoundaryGap(o, (unsigned)(ps + pl + 1), (unsigned)(ps + 1 + pl + ORC_MIN_GAP + move - 1))) {
                        iv_push(&h, d2); iv_push(&h, ps); iv_push(&h, pl); iv_push(&h, pl + 1 + ORC_MIN_GAP + move + 1 - 1);
                    }
                }
            }
        }
    }
    o->hits2 = (int32_t)(h.n / 4);
    qsort(h.v, (size_t)o->hits2, sizeof(int32_t) * 4, cmp_hit4);
    o->h2 = h.v;
    for (int32_t i = 0; i < o->hits2; i++) {
        int32_t *p = &o->pat2[4 * h.v[4 * i]];
        if (p[2] == -1) p[2] = i;
        p[3] = i;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* sampling                                                                                     */
/* ------------------------------------------------------------------------------------------ */
/* ExtractPair.cu:1133-1160 / :445-471 / :946-972.  `stepsize = (float)n / (float)S` is compiled under
 * -use_fast_math to a single FMUL by the correctly rounded float reciprocal of the constant S
 * (verified in the sm_100 SASS of the reference build: FMUL.FTZ R, R, 0x3b5a740e | 0x3c7c0fc1 |
 * 0x3c6a0ea1); `ROUND(desci*stepsize)` is a float multiply, then +0.5 and truncation in double. */
static int sampled(int32_t j, int32_t n, int S) {
    if (n <= S) return 1;
    volatile float rcp = 1.0f / (float)S;
    volatile float step = (float)n * rcp;
    for (int d = 0; d < S; d++) {
        volatile float prod = (float)d * step;
        int togo = (int)((double)prod + 0.5);
        if (togo == j) return 1;
        if (togo > j) return 0;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* extraction kernels                                                                           */
/* ------------------------------------------------------------------------------------------ */
#define EMIT1(idv, ts, te, g1s, g1e) do { rec_t r_ = {(idv), (int32_t)(ts), (int32_t)((te) - (ts)), (int32_t)((g1s) - (ts)), (int32_t)((g1e) - (ts)), -1, -1}; rv_push(&o->rec_1, r_); } while (0)
#define EMIT2(idv, ts, te, g1s, g1e, g2s, g2e) do { rec_t r_ = {(idv), (int32_t)(ts), (int32_t)((te) - (ts)), (int32_t)((g1s) - (ts)), (int32_t)((g1e) - (ts)), (int32_t)((g2s) - (ts)), (int32_t)((g2e) - (ts))}; rv_push(&o->rec_2, r_); } while (0)

/* ExtractPair.cu:1055-1795 extractConsistentPairs_Gappy, one sampled occurrence.
 * Returns 0 normally, 1 when the reference thread executes `return` (which drops the rest of that
 * thread's strided occurrences). */
static int gappy_body(orc_t *o, int bnum, int globalc, int current, int longestmatch) {
    const int32_t *refstr = o->str;
    int k, current_str; lr_t L, R; int sen_target_begin = -1; lr_t min_L = UNAL, max_R = 0; int tempind = 0; rlp_t temp;
    lr_t i = 1; unsigned int gap1_start = 0, gap1_end = 0, gap2_start = 0, gap2_end = 0, target_start = 0, target_end = 0;
    int next = 1; int ender;
    int abX = 1, Xab = 1, XabX = 1, ab = 1, XabNoSuccess = 1, abXNoSuccess = 1; uint8_t XabCount = 0, abXCount = 0;
    lr_t min_L_Xab = UNAL, max_R_Xab = 0, min_L_abX = UNAL, max_R_abX = 0, min_L_XabX = UNAL, max_R_XabX = 0;

    current_str = o->sa[current];
    for (k = current_str; k < current_str + longestmatch; k++) {
        temp = o->RLP[k]; L = RLP_L(temp); R = RLP_R(temp);
        if (k == current_str) {
            tempind = k - RLP_P(temp) - 1;
            sen_target_begin = tempind == -1 ? 0 : (int)o->RLP[tempind];
        }
        if ((L == UNAL || R == UNAL) && (k == current_str || k == current_str + longestmatch - 1)) {
            ab = 0;
            if (k == current_str) abXNoSuccess = 0; else XabNoSuccess = 0;
        } else if (L == UNAL || R == UNAL) { }
        else { if (min_L > L) min_L = L; if (max_R < R) max_R = R; }
    }
    if (min_L > max_R || max_R - min_L >= ORC_MAX_RULE_SPAN) { abX = 0; Xab = 0; XabX = 0; ab = 0; }
    tempind++;
    ender = current_str + longestmatch - 1;
    if (ab) {
        if (consistent(o, min_L + sen_target_begin, max_R + sen_target_begin, current_str, ender, tempind)) {
            rec_t r = {bnum, min_L + sen_target_begin, (uint8_t)(max_R - min_L), -1, -1, -1, -1};
            rv_push(&o->rec_ab, r);
        }
    }
    if (longestmatch + 1 > ORC_MAX_RULE_SYMBOLS) { abX = 0; Xab = 0; }
    if (longestmatch + 2 > ORC_MAX_RULE_SYMBOLS) XabX = 0;
    i = 1;
    while (longestmatch + i <= ORC_MAX_RULE_SPAN && (abXNoSuccess || XabNoSuccess || XabX)) {
        /* ---- left X ---- :1282-1398 */
        if (Xab && current_str - i >= 0 && refstr[current_str - i] >= 2) {
            next = 1;
            temp = o->RLP[current_str - i]; L = RLP_L(temp); R = RLP_R(temp);
            if (L == UNAL || R == UNAL) { next = 0; if (i == 1) { Xab = 0; XabX = 0; } }
            else { if (min_L_Xab > L) min_L_Xab = L; if (max_R_Xab < R) max_R_Xab = R; }
            if (next && min_L_Xab > max_R_Xab) return 1;
            if (max_R_Xab - min_L_Xab >= ORC_MAX_RULE_SPAN) { next = 0; Xab = 0; }
            if (next) {
                gap1_start = (unsigned)(sen_target_begin + min_L_Xab); gap1_end = (unsigned)(sen_target_begin + max_R_Xab);
                if (gap1_start > gap1_end) return 1;
                next = consistent(o, (int)gap1_start, (int)gap1_end, current_str - i, current_str - 1, tempind);
                if (next) XabCount = i;
            }
            if (XabNoSuccess && next) {
                target_start = (unsigned)(sen_target_begin + (min_L_Xab < min_L ? min_L_Xab : min_L));
                target_end = (unsigned)(sen_target_begin + (max_R_Xab < max_R ? max_R : max_R_Xab));
                if (next && target_start > target_end) return 1;
                if (target_end - target_start >= ORC_MAX_RULE_SPAN) { next = 0; Xab = 0; }
                if (next) next = consistent(o, (int)target_start, (int)target_end, current_str - i, ender, tempind);
            }
            if (XabNoSuccess && next) { EMIT1(bnum, target_start, target_end, gap1_start, gap1_end); XabNoSuccess = 0; }
        } else Xab = 0;
        /* ---- right X ---- :1403-1509 */
        if (abX && refstr[ender + i] >= 2) {
            next = 1;
            temp = o->RLP[ender + i]; L = RLP_L(temp); R = RLP_R(temp);
            if (L == UNAL || R == UNAL) { next = 0; if (i == 1) { abX = 0; XabX = 0; } }
            else { if (min_L_abX > L) min_L_abX = L; if (max_R_abX < R) max_R_abX = R; }
            if (next && min_L_abX > max_R_abX) return 1;
            if (max_R_abX - min_L_abX >= ORC_MAX_RULE_SPAN) { next = 0; abX = 0; }
            if (next) {
                gap1_start = (unsigned)(sen_target_begin + min_L_abX); gap1_end = (unsigned)(sen_target_begin + max_R_abX);
                if (gap1_start > gap1_end) return 1;
                next = consistent(o, (int)gap1_start, (int)gap1_end, ender + 1, ender + i, tempind);
                if (next) abXCount = i;
            }
            if (abXNoSuccess && next) {
                target_start = (unsigned)(sen_target_begin + (min_L_abX < min_L ? min_L_abX : min_L));
                target_end = (unsigned)(sen_target_begin + (max_R_abX < max_R ? max_R : max_R_abX));
                if (next && target_start > target_end) return 1;
                if (target_end - target_start >= ORC_MAX_RULE_SPAN) { next = 0; abX = 0; }
                if (next) next = consistent(o, (int)target_start, (int)target_end, current_str, ender + i, tempind);
            }
            if (abXNoSuccess && next) { EMIT1(globalc + bnum, target_start, target_end, gap1_start, gap1_end); abXNoSuccess = 0; }
        } else abX = 0;
        /* ---- XabX ---- :1514-1777 */
        if (XabX && (abX || Xab)) {
            if (XabCount == i) {
                min_L_XabX = UNAL; max_R_XabX = 0;
                for (uint8_t icount = 1; XabX && icount <= abXCount; icount++) {
                    next = 1;
                    if (icount + XabCount + longestmatch <= ORC_MAX_RULE_SPAN) {
                        temp = o->RLP[ender + icount]; L = RLP_L(temp); R = RLP_R(temp);
                        if (L == UNAL || R == UNAL) { next = 0; if (i == 1) return 1; }
                        else { if (min_L_XabX > L) min_L_XabX = L; if (max_R_XabX < R) max_R_XabX = R; }
                    } else { next = 0; icount = abXCount + 1; }
                    if (next && max_R_XabX - min_L_XabX >= ORC_MAX_RULE_SPAN) { next = 0; icount = abXCount + 1; }
                    if (next) {
                        gap2_start = (unsigned)(sen_target_begin + min_L_XabX); gap2_end = (unsigned)(sen_target_begin + max_R_XabX);
                        if (min_L_XabX > max_R_XabX) return 1;
                        next = consistent(o, (int)gap2_start, (int)gap2_end, ender + 1, ender + icount, tempind);
                    }
                    if (next) {
                        temp = min_L_XabX < min_L_Xab ? min_L_XabX : min_L_Xab; if (temp > min_L) temp = min_L;
                        target_start = (unsigned)sen_target_begin + temp;
                        temp = max_R_XabX < max_R_Xab ? max_R_Xab : max_R_XabX; if (temp < max_R) temp = max_R;
                        target_end = (unsigned)sen_target_begin + temp;
                        if (target_start > target_end) return 1;
                        if (target_end - target_start >= ORC_MAX_RULE_SPAN) { next = 0; icount = abXCount + 1; }
                        if (next) next = consistent(o, (int)target_start, (int)target_end, current_str - XabCount, ender + icount, tempind);
                        if (next) {
                            gap1_start = (unsigned)(sen_target_begin + min_L_Xab); gap1_end = (unsigned)(sen_target_begin + max_R_Xab);
                            EMIT2(bnum, target_start, target_end, gap1_start, gap1_end, gap2_start, gap2_end);
                            XabX = 0;
                        }
                    }
                }
            }
            if (XabX && abXCount == i) {
                min_L_XabX = UNAL; max_R_XabX = 0;
                for (uint8_t icount = 1; XabX && icount <= XabCount; icount++) {
                    next = 1;
                    if (icount + abXCount + longestmatch <= ORC_MAX_RULE_SPAN) {
                        temp = o->RLP[current_str - icount]; L = RLP_L(temp); R = RLP_R(temp);
                        if (L == UNAL || R == UNAL) { next = 0; if (i == 1) return 1; }
                        else { if (min_L_XabX > L) min_L_XabX = L; if (max_R_XabX < R) max_R_XabX = R; }
                    } else { icount = XabCount + 1; next = 0; }
                    if (next && max_R_XabX - min_L_XabX >= ORC_MAX_RULE_SPAN) { icount = XabCount + 1; next = 0; }
                    if (next) {
                        gap1_start = (unsigned)(sen_target_begin + min_L_XabX); gap1_end = (unsigned)(sen_target_begin + max_R_XabX);
                        if (min_L_XabX > max_R_XabX) return 1;
                        next = consistent(o, (int)gap1_start, (int)gap1_end, current_str - icount, current_str - 1, tempind);
                    }
                    if (next) {
                        temp = min_L_XabX < min_L_abX ? min_L_XabX : min_L_abX; if (temp > min_L) temp = min_L;
                        target_start = (unsigned)sen_target_begin + temp;
                        temp = max_R_XabX < max_R_abX ? max_R_abX : max_R_XabX; if (temp < max_R) temp = max_R;
                        target_end = (unsigned)sen_target_begin + temp;
                        if (target_start > target_end) return 1;
                        if (target_end - target_start >= ORC_MAX_RULE_SPAN) { next = 0; icount = XabCount + 1; }
                        if (next) next = consistent(o, (int)target_start, (int)target_end, current_str - icount, ender + abXCount, tempind);
                        if (next) {
                            gap2_start = (unsigned)(sen_target_begin + min_L_abX); gap2_end = (unsigned)(sen_target_begin + max_R_abX);
                            EMIT2(bnum, target_start, target_end, gap1_start, gap1_end, gap2_start, gap2_end);
                            XabX = 0;
                        }
                    }
                }
            }
        } else XabX = 0;
        if (!XabX) {                                               /* :1782-1789 */
            if (!Xab && XabNoSuccess) XabNoSuccess = 0;
            if (!abX && abXNoSuccess) abXNoSuccess = 0;
        }
        i++;
    }
    return 0;
}

static void extract_gappy(orc_t *o) {
    char dead[ORC_THREADS];
    for (int32_t b = 0; b < o->G; b++) {
        int32_t start = o->blocks[4 * b], end = o->blocks[4 * b + 1], len = o->blocks[4 * b + 2];
        if (len < 1) continue;
        int32_t nocc = 1 + end - start;
        memset(dead, 0, sizeof(dead));
        for (int32_t j = 0; j < nocc; j++) {
            int tid = j % ORC_THREADS;
            if (dead[tid]) continue;
            if (!sampled(j, nocc, ORC_SAMPLER)) continue;
            if (gappy_body(o, b, o->G, start + j, len)) dead[tid] = 1;
        }
    }
}

/* ExtractPair.cu:891-1053 extractConsistentPairs_TwoGap */
static void extract_twogap(orc_t *o) {
    char dead[ORC_THREADS];
    for (int32_t d2 = 0; d2 < o->D2; d2++) {
        int32_t startSA = o->pat2[4 * d2 + 2], endSA = o->pat2[4 * d2 + 3];
        if (startSA == -1 && endSA == -1) continue;
        int32_t dis = endSA - startSA + 1; int32_t d1 = o->pat2[4 * d2];
        int startLen = o->pat1[10 * d1 + 6], endLen = o->pat1[10 * d1 + 7];
        memset(dead, 0, sizeof(dead));
        for (int32_t j = 0; j < dis; j++) {
            int tid = j % ORC_THREADS;
            if (dead[tid]) continue;
            if (!sampled(j, dis, ORC_SAMPLER_TWOGAP)) continue;
            unsigned int current_str = (unsigned)o->h2[4 * (startSA + j) + 1], firstEnd = (unsigned)o->h2[4 * (startSA + j) + 2], secondEnd = (unsigned)o->h2[4 * (startSA + j) + 3];
            unsigned int g1s = 0, g1e = 0, g2s = 0, g2e = 0, ts = 0, te = 0;
            int next = checkBoundaryFast2(o, current_str + startLen, current_str + firstEnd - endLen, &g1s, &g1e);
            if (next) next = checkBoundaryFast2(o, current_str + firstEnd + 1, current_str + secondEnd - 1 /* qryend_len == 1 */, &g2s, &g2e);
            if (!next) { dead[tid] = 1; continue; }
            if (checkBoundary(o, current_str, current_str + secondEnd, &ts, &te) == 1) EMIT2(d2, ts, te, g1s, g1e, g2s, g2e);
        }
    }
}

/* ExtractPair.cu:351-889 extractConsistentPairs_OneGap, one sampled hit; returns 1 on thread `return` */
static int onegap_body(orc_t *o, int oneBlockId, unsigned int current_str, lr_t firstEnd, int startLen, int endLen) {
    const int32_t *refstr = o->str;
    unsigned int gap1_start = 0, gap1_end = 0, target_start = 0, target_end = 0, gap2_start = 0, gap2_end = 0;
    int next = 1, firstGap = 1, left = 1, right = 1; unsigned int ender; uint8_t i = 1; int reNext;
    lr_t min_L = UNAL, max_R = 0; int sen_target_begin = -1, tempind = -1;
    if ((int64_t)current_str + firstEnd - endLen > o->n) return 1;
    ender = current_str + firstEnd;
    firstGap = checkBoundaryFast(o, current_str + (unsigned)startLen, ender - (unsigned)endLen, &min_L, &max_R, &sen_target_begin, &tempind);
    if (!firstGap) return 1;
    if (tempind == -1 || sen_target_begin == -1 || min_L > max_R) return 1;
    gap1_start = (unsigned)(min_L + sen_target_begin); gap1_end = (unsigned)(max_R + sen_target_begin);
    reNext = checkBoundary(o, current_str, ender, &target_start, &target_end);
    min_L = (lr_t)(target_start - (unsigned)sen_target_begin); max_R = (lr_t)(target_end - (unsigned)sen_target_begin);
    if (reNext == 0) next = 0; else if (reNext == 1) next = 1; else if (reNext == 2) { next = 0; right = 0; }
    else if (reNext == 3) { next = 0; left = 0; } else if (reNext == 4) { next = 0; left = 0; right = 0; }
    if ((target_start == 0 && target_end == 0) || (min_L > max_R) || gap1_start < target_start || gap1_end > target_end) return 1;   /* :591-595 */
    if (next && firstGap) EMIT1(oneBlockId, target_start, target_end, gap1_start, gap1_end);
    unsigned int originalGapStart, originalGapEnd; rlp_t temp; lr_t L, R;
    lr_t min_XaXb = UNAL, max_XaXb = 0, min_aXbX = UNAL, max_aXbX = 0;
    if (firstGap && startLen + endLen + 1 + 1 <= ORC_MAX_RULE_SYMBOLS) {
        target_start = 0; target_end = 0; i = 1;
        originalGapStart = gap1_start; originalGapEnd = gap1_end; gap1_start = 0; gap1_end = 0;
        while (firstEnd + 1 + i <= ORC_MAX_RULE_SPAN && (left || right)) {
            if (left && (int)(current_str - i) >= 0 && refstr[current_str - i] >= 2) {
                target_start = 0; target_end = 0; gap1_start = 0; gap1_end = 0; next = 1;
                temp = o->RLP[current_str - i]; L = RLP_L(temp); R = RLP_R(temp);
                if (L == UNAL || R == UNAL) { next = 0; if (i == 1) left = 0; }
                else { if (min_XaXb > L) min_XaXb = L; if (max_XaXb < R) max_XaXb = R; }
                if (next && min_XaXb > max_XaXb) return 1;
                if (max_XaXb - min_XaXb >= ORC_MAX_RULE_SPAN) { next = 0; left = 0; }
                if (next) {
                    gap1_start = (unsigned)(sen_target_begin + min_XaXb); gap1_end = (unsigned)(sen_target_begin + max_XaXb);
                    next = consistent(o, (int)gap1_start, (int)gap1_end, (int)(current_str - i), (int)(current_str - 1), tempind);
                }
                if (next) {
                    target_start = (unsigned)(sen_target_begin + (min_XaXb < min_L ? min_XaXb : min_L));
                    target_end = (unsigned)(sen_target_begin + (max_XaXb < max_R ? max_R : max_XaXb));
                    if (next && target_start > target_end) return 1;
                    if (target_end - target_start >= ORC_MAX_RULE_SPAN) { next = 0; left = 0; }
                    if (next) next = consistent(o, (int)target_start, (int)target_end, (int)(current_str - i), (int)ender, tempind);
                }
                if (next) { EMIT2(oneBlockId, target_start, target_end, gap1_start, gap1_end, originalGapStart, originalGapEnd); left = 0; }
            } else left = 0;
            if (right && refstr[ender + i] >= 2) {
                target_start = 0; target_end = 0; next = 1; gap2_start = 0; gap2_end = 0;
                temp = o->RLP[ender + i]; L = RLP_L(temp); R = RLP_R(temp);
                if (L == UNAL || R == UNAL) { next = 0; if (i == 1) right = 0; }
                else { if (min_aXbX > L) min_aXbX = L; if (max_aXbX < R) max_aXbX = R; }
                if (next && min_aXbX > max_aXbX) return 1;
                if (max_aXbX - min_aXbX >= ORC_MAX_RULE_SPAN) { next = 0; right = 0; }
                if (next) {
                    gap2_start = (unsigned)(sen_target_begin + min_aXbX); gap2_end = (unsigned)(sen_target_begin + max_aXbX);
                    if (gap2_start > gap2_end) return 1;
                    next = consistent(o, (int)gap2_start, (int)gap2_end, (int)(ender + 1), (int)(ender + i), tempind);
                }
                if (next) {
                    target_start = (unsigned)(sen_target_begin + (min_aXbX < min_L ? min_aXbX : min_L));
                    target_end = (unsigned)(sen_target_begin + (max_aXbX < max_R ? max_R : max_aXbX));
                    if (next && target_start > target_end) return 1;
                    if (target_end - target_start >= ORC_MAX_RULE_SPAN) { next = 0; right = 0; }
                    if (next) next = consistent(o, (int)target_start, (int)target_end, (int)current_str, (int)(ender + i), tempind);
                }
                if (next) { EMIT2(o->D1 + oneBlockId, target_start, target_end, originalGapStart, originalGapEnd, gap2_start, gap2_end); right = 0; }
            } else right = 0;
            i++;
        }
    }
    return 0;
}

static void extract_onegap(orc_t *o) {
    char dead[ORC_THREADS];
    for (int32_t d = 0; d < o->D1; d++) {
        int32_t startSA = o->pat1[10 * d + 8], endSA = o->pat1[10 * d + 9];
        if (startSA == -1 && endSA == -1) continue;
        int32_t dis = 1 + endSA - startSA; int startLen = o->pat1[10 * d + 6], endLen = o->pat1[10 * d + 7];
        int pre = 0; int32_t base = startSA;
        if (dis == 1 && o->h1[3 * startSA + 2] == 0) {
            pre = 1; int32_t pi = o->h1[3 * startSA + 1];
            base = o->pidx[2 * pi]; dis = 1 + o->pidx[2 * pi + 1] - o->pidx[2 * pi];
        }
        memset(dead, 0, sizeof(dead));
        for (int32_t j = 0; j < dis; j++) {
            int tid = j % ORC_THREADS;
            if (dead[tid]) continue;
            if (!sampled(j, dis, ORC_SAMPLER_ONEGAP)) continue;
            unsigned int cs; lr_t fe;
            if (pre) { cs = (unsigned)o->plist[2 * (base + j)]; fe = (lr_t)o->plist[2 * (base + j) + 1]; }
            else { cs = (unsigned)o->h1[3 * (base + j) + 1]; fe = (lr_t)o->h1[3 * (base + j) + 2]; }
            if (onegap_body(o, d, cs, fe, startLen, endLen)) dead[tid] = 1;
        }
    }
}

typedef struct { rec_t r; int64_t i; } ri_t;
static int cmp_ri(const void *x, const void *y) {
    const ri_t *p = (const ri_t *)x, *q = (const ri_t *)y;
    if (p->r.id != q->r.id) return p->r.id < q->r.id ? -1 : 1;
    return p->i < q->i ? -1 : p->i > q->i;
}
/* stable sort by id (the reference's thrust sorts key on the id only, ExtractPair.cu:3417-3441) */
static void stable_sort_by_id(rec_t *v, int64_t n) {
    if (n <= 1) return;
    ri_t *a = (ri_t *)malloc(sizeof(ri_t) * (size_t)n);
    for (int64_t i = 0; i < n; i++) { a[i].r = v[i]; a[i].i = i; }
    qsort(a, (size_t)n, sizeof(ri_t), cmp_ri);
    for (int64_t i = 0; i < n; i++) v[i] = a[i].r;
    free(a);
}

/* ------------------------------------------------------------------------------------------ */
/* aggregation (ExtractPair.c:515-1276) and lexical weights (ExtractPair.cu:2144-2432)          */
/* ------------------------------------------------------------------------------------------ */
/* target symbol sequence of a record: tokens outside the gaps, gap1 -> -1, gap2 -> -2 (one symbol each) */
static int target_symbols(const orc_t *o, const rec_t *r, int32_t out[64]) {
    int k = 0; int32_t ts = r->tgt_start, te = r->tgt_start + r->end;
    int32_t g1s = r->gap1 >= 0 ? ts + r->gap1 : -1, g1e = r->gap1 >= 0 ? ts + r->gap1_1 : -2;
    int32_t g2s = r->gap2 >= 0 ? ts + r->gap2 : -1, g2e = r->gap2 >= 0 ? ts + r->gap2_1 : -2;
    for (int32_t jj = ts; jj <= te; jj++) {
        if (jj >= g1s && jj <= g1e) { out[k++] = -1; jj = g1e; }
        else if (jj >= g2s && jj <= g2e) { out[k++] = -2; jj = g2e; }
        else out[k++] = o->tgt[jj];
    }
    return k;
}

/* source terminals of a converted id (the F set of lexicalTaskMaxEF) */
static int source_terminals(const orc_t *o, int kind, int32_t cid, int32_t out[8]) {
    int k = 0; int32_t G = o->G, D1 = o->D1, D2 = o->D2;
    int32_t blk = -1, p1 = -1, p2 = -1;
    if (kind == 0) blk = cid;
    else if (kind == 1) { if (cid < G) blk = cid; else if (cid < 2 * G) blk = cid - G; else p1 = cid - 2 * G; }
    else { if (cid < G) blk = cid; else if (cid < G + D2) p2 = cid - G; else if (cid < G + D2 + D1) p1 = cid - G - D2; else p1 = cid - G - D2 - D1; }
    if (blk >= 0) { int32_t s = o->blocks[4 * blk + 3], l = o->blocks[4 * blk + 2]; for (int i = 0; i < l; i++) out[k++] = o->str[s + i]; }
    if (p2 >= 0) p1 = o->pat2[4 * p2];
    if (p1 >= 0) { const int32_t *p = &o->pat1[10 * p1]; for (int i = 0; i < p[5]; i++) if (p[i] >= 0) out[k++] = p[i]; }
    if (p2 >= 0) out[k++] = o->pat2[4 * p2 + 1];
    return k;
}

/* all_suffix_fsample before the cap (ExtractPair.c:637, :891-908, :1211-1247) */
static int32_t fsample_of(const orc_t *o, int kind, int32_t cid) {
    int32_t G = o->G, D1 = o->D1, D2 = o->D2; int32_t blk = -1, p1 = -1, p2 = -1;
    if (kind == 0) blk = cid;
    else if (kind == 1) { if (cid < G) blk = cid; else if (cid < 2 * G) blk = cid - G; else p1 = cid - 2 * G; }
    else { if (cid < G) blk = cid; else if (cid < G + D2) p2 = cid - G; else if (cid < G + D2 + D1) p1 = cid - G - D2; else p1 = cid - G - D2 - D1; }
    if (blk >= 0) return 1 + o->blocks[4 * blk + 1] - o->blocks[4 * blk];
    if (p2 >= 0) return 1 + o->pat2[4 * p2 + 3] - o->pat2[4 * p2 + 2];
    const int32_t *p = &o->pat1[10 * p1];
    int32_t fs = 1 + p[9] - p[8];
    if (fs == 1 && o->h1[3 * p[8] + 2] == 0) { int32_t pi = o->h1[3 * p[8] + 1]; fs = 1 - o->pidx[2 * pi] + o->pidx[2 * pi + 1] + o->missing[pi]; }
    return fs;
}

/* ExtractPair.cu:2144-2432 lexicalTaskMaxEF */
static void lex_scores(const orc_t *o, int kind, const orc_rule_t *r, float *mlfe, float *mlef) {
    int32_t F[8]; int nf = source_terminals(o, kind, r->id, F);
    int32_t ts = r->rec[0], te = r->rec[0] + r->rec[1];
    int32_t g1s = r->rec[2] >= 0 ? ts + r->rec[2] : 1, g1e = r->rec[2] >= 0 ? ts + r->rec[3] : 0;
    int32_t g2s = r->rec[4] >= 0 ? ts + r->rec[4] : 1, g2e = r->rec[4] >= 0 ? ts + r->rec[5] : 0;
    float fgivene = 0, egivenf = 0;
    for (int j = 0; j < nf; j++) {
        float mx = 0; int first = 1;
        for (int32_t jj = ts; jj <= te; jj++) {
            if ((jj < g1s || jj > g1e) && (jj < g2s || jj > g2e)) {
                if (first) { float v = lex_get(o, F[j], -1, 0); if (v > mx) mx = v; first = 0; }
                float v = lex_get(o, F[j], o->tgt[jj], 0); if (v > mx) mx = v;
            }
        }
        if (mx > 0) fgivene += -log10f(mx); else fgivene += ORC_MAXSCORE;
    }
    for (int32_t jj = ts; jj <= te; jj++) {
        if ((jj < g1s || jj > g1e) && (jj < g2s || jj > g2e)) {
            float mx = 0; int first = 1;
            for (int j = 0; j < nf; j++) {
                if (first) { float v = lex_get(o, -1, o->tgt[jj], 1); if (v > mx) mx = v; first = 0; }
                float v = lex_get(o, F[j], o->tgt[jj], 1); if (v > mx) mx = v;
            }
            if (mx > 0) egivenf += -log10f(mx); else egivenf += ORC_MAXSCORE;
        }
    }
    *mlfe = fgivene; *mlef = egivenf;
}

/* createLexiconFast / createLexiconGappyFast / createLexiconTwoGapFast.  `recs` are sorted by
 * converted id within each separator group; de-duplication is per id run (and per separator group),
 * rule order = first sighting. */
static void aggregate(orc_t *o, int kind, const rec_t *recs, int64_t n, const int32_t *seps, int nseps, const int32_t *offsets, int32_t nids) {
    int32_t *fcount = (int32_t *)calloc((size_t)nids + 1, sizeof(int32_t));
    /* convert ids */
    int32_t *cid = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
    for (int64_t i = 0; i < n; i++) {
        int g = 0; while (g < nseps && i >= seps[g]) g++;
        cid[i] = recs[i].id + offsets[g];
        fcount[cid[i]]++;
    }
    orc_rule_t *rules = (orc_rule_t *)malloc(sizeof(orc_rule_t) * (size_t)(n ? n : 1)); int32_t nr = 0;
    int32_t run_begin_rule = 0;
    /* per-run table of (hash -> rule index) */
    for (int64_t i = 0; i < n; i++) {
        int newrun = (i == 0) || cid[i] != cid[i - 1];
        for (int g = 0; g < nseps; g++) if (i == seps[g]) newrun = 1;
        if (newrun) run_begin_rule = nr;
        int32_t sym[64]; int ns = target_symbols(o, &recs[i], sym);
        int found = -1;
        for (int32_t r = run_begin_rule; r < nr && found < 0; r++) {
            rec_t rr = {0, rules[r].rec[0], rules[r].rec[1], rules[r].rec[2], rules[r].rec[3], rules[r].rec[4], rules[r].rec[5]};
            int32_t s2[64]; int n2 = target_symbols(o, &rr, s2);
            if (n2 == ns && !memcmp(sym, s2, sizeof(int32_t) * (size_t)ns)) found = r;
        }
        if (found >= 0) { rules[found].pc++; continue; }
        orc_rule_t *r = &rules[nr++];
        memset(r, 0, sizeof(*r));
        r->id = cid[i];
        r->rec[0] = recs[i].tgt_start; r->rec[1] = recs[i].end; r->rec[2] = recs[i].gap1; r->rec[3] = recs[i].gap1_1; r->rec[4] = recs[i].gap2; r->rec[5] = recs[i].gap2_1;
        r->pc = 1; r->f = fcount[cid[i]];
        int32_t fs = fsample_of(o, kind, cid[i]);
        if (fs > ORC_SAMPLER) fs = ORC_SAMPLER;
        r->fs = fs;
    }
    int32_t *ud = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(nids + 1));
    for (int32_t i = 0; i < 2 * (nids + 1); i++) ud[i] = -1;
    for (int32_t r = 0; r < nr; r++) {
        orc_rule_t *x = &rules[r];
        x->aa = -log10f((float)x->pc / (float)x->fs);             /* ExtractPair.c:653-655 (float overload) */
        x->bb = (float)log10((double)(1 + x->pc));
        x->score = (float)log10((double)(1 + x->fs));
        lex_scores(o, kind, x, &x->mlfe, &x->mlef);
        if (r == 0 || rules[r - 1].id != x->id) ud[2 * x->id] = r;
        ud[2 * x->id + 1] = r;
    }
    o->rules[kind] = rules; o->nrules[kind] = nr; o->updown[kind] = ud; o->nid[kind] = nids;
    free(fcount); free(cid);
}

/* ------------------------------------------------------------------------------------------ */
static void free_run(orc_t *o) {
    free(o->q); free(o->qoff); free(o->tok2q); free(o->longest); free(o->conn_off); free(o->iv_up); free(o->iv_down);
    free(o->blocks);
    if (o->qryglobal) { for (int32_t i = 0; i < o->Q; i++) iv_free(&o->qryglobal[i]); free(o->qryglobal); }
    if (o->q1) { for (int32_t i = 0; i < o->Q; i++) iv_free(&o->q1[i]); free(o->q1); }
    if (o->q2) { for (int32_t i = 0; i < o->Q; i++) iv_free(&o->q2[i]); free(o->q2); }
    free(o->g1_start); free(o->g1_ls); free(o->g1_le); free(o->g1_gap); free(o->g1_sorted);
    free(o->pat1); free(o->pat1_pos); free(o->pat1_rep); free(o->h1); free(o->pat2); free(o->pat2_rep); free(o->h2);
    free(o->rec_ab.v); free(o->rec_1.v); free(o->rec_2.v);
    for (int k = 0; k < 3; k++) { free(o->rules[k]); free(o->updown[k]); free(o->rec_flat[k]); o->rules[k] = NULL; o->updown[k] = NULL; o->rec_flat[k] = NULL; }
    o->q = o->qoff = o->tok2q = o->longest = o->conn_off = o->iv_up = o->iv_down = o->blocks = NULL;
    o->qryglobal = o->q1 = o->q2 = NULL;
    o->g1_start = o->g1_sorted = o->pat1 = o->pat1_pos = o->pat1_rep = o->h1 = o->pat2 = o->pat2_rep = o->h2 = NULL;
    o->g1_ls = o->g1_le = o->g1_gap = NULL;
    memset(&o->rec_ab, 0, sizeof(rvec)); memset(&o->rec_1, 0, sizeof(rvec)); memset(&o->rec_2, 0, sizeof(rvec));
}

int orc_run(orc_t *o, const int32_t *qry_tok, const int32_t *qry_off, int32_t Q) {
    if (!o->sa) orc_build_sa(o);
    if (!o->have_precomp && !build_precomp(o)) return 1;
    free_run(o);
    int32_t T = qry_off[Q];
    o->Q = Q; o->T = T;
    o->q = (int32_t *)malloc(sizeof(int32_t) * ((size_t)T + 8)); memcpy(o->q, qry_tok, sizeof(int32_t) * (size_t)T);
    for (int i = 0; i < 8; i++) o->q[T + i] = -1;
    o->qoff = (int32_t *)malloc(sizeof(int32_t) * ((size_t)Q + 1)); memcpy(o->qoff, qry_off, sizeof(int32_t) * ((size_t)Q + 1));
    o->tok2q = (int32_t *)malloc(sizeof(int32_t) * ((size_t)T + 1));
    for (int32_t qi = 0; qi < Q; qi++) for (int32_t t = qry_off[qi]; t < qry_off[qi + 1]; t++) o->tok2q[t] = qi;
    lookup(o);
    onegap_enumerate(o);
    onegap_lookup(o);
    twogap_enumerate(o);
    twogap_lookup(o);
    generate_blocks(o);
    extract_gappy(o);
    stable_sort_by_id(o->rec_ab.v, o->rec_ab.n);
    stable_sort_by_id(o->rec_1.v, o->rec_1.n);
    stable_sort_by_id(o->rec_2.v, o->rec_2.n);
    o->cnt.n_ab = (int32_t)o->rec_ab.n; o->cnt.n_1gap_contig = (int32_t)o->rec_1.n; o->cnt.n_2gap_contig = (int32_t)o->rec_2.n;
    o->sep1 = (int32_t)o->rec_1.n; o->sep2a = (int32_t)o->rec_2.n;
    extract_twogap(o);
    stable_sort_by_id(o->rec_2.v + o->sep2a, o->rec_2.n - o->sep2a);
    o->cnt.n_axbxc = (int32_t)o->rec_2.n - o->sep2a; o->sep2b = (int32_t)o->rec_2.n;
    extract_onegap(o);
    stable_sort_by_id(o->rec_1.v + o->sep1, o->rec_1.n - o->sep1);
    stable_sort_by_id(o->rec_2.v + o->sep2b, o->rec_2.n - o->sep2b);
    o->cnt.n_axb = (int32_t)o->rec_1.n - o->sep1; o->cnt.n_2gap_from1 = (int32_t)o->rec_2.n - o->sep2b;
    {   /* ids: ExtractPair.c:723-729 / :999-1006 */
        int32_t seps1[1] = {o->sep1}; int32_t off1[2] = {0, 2 * o->G};
        aggregate(o, 1, o->rec_1.v, o->rec_1.n, seps1, 1, off1, 2 * o->G + o->D1);
        int32_t seps2[2] = {o->sep2a, o->sep2b}; int32_t off2[3] = {0, o->G, o->G + o->D2};
        aggregate(o, 2, o->rec_2.v, o->rec_2.n, seps2, 2, off2, o->G + o->D2 + 2 * o->D1);
        int32_t off0[1] = {0};
        aggregate(o, 0, o->rec_ab.v, o->rec_ab.n, NULL, 0, off0, o->G);
        /* expose converted ids */
        for (int64_t i = o->sep1; i < o->rec_1.n; i++) o->rec_1.v[i].id += 2 * o->G;
        for (int64_t i = o->sep2a; i < o->rec_2.n; i++) o->rec_2.v[i].id += (i < o->sep2b) ? o->G : o->G + o->D2;
    }
    o->cnt.n = o->n; o->cnt.m = o->m; o->cnt.Q = Q; o->cnt.T = T; o->cnt.G = o->G; o->cnt.enu1 = o->enu1; o->cnt.D1 = o->D1; o->cnt.hits1 = o->hits1;
    o->cnt.enu2 = o->enu2; o->cnt.D2 = o->D2; o->cnt.hits2 = o->hits2; o->cnt.precomp_count = o->pcount;
    o->cnt.lex_1gap = o->nrules[1]; o->cnt.lex_2gap = o->nrules[2]; o->cnt.lex_ab = o->nrules[0];
    return 0;
}

/* Start.cu:50-132 constructQryIndex */
int orc_run_query_file(orc_t *o, const char *path) {
    if (!o->have_vocab) return 2;
    FILE *fh = fopen(path, "r");
    if (!fh) { fprintf(stderr, "oracle: cannot open %s\n", path); return 3; }
    ivec tok = {0}, off = {0}; char *line = NULL; size_t cap = 0;
    iv_push(&off, 0);
    while (getline(&line, &cap, fh) != -1) {
        char *t = strtok(line, " ");
        while (t != NULL && !isspace((lr_t)*t)) {
            size_t tl = strlen(t);
            if (tl && t[tl - 1] == '\n') t[tl - 1] = 0;
            iv_push(&tok, smap_get(&o->smap_src, t));
            t = strtok(NULL, " ");
        }
        iv_push(&off, (int32_t)tok.n);
    }
    free(line); fclose(fh);
    if (!tok.v) iv_push(&tok, -1), tok.n = 0;
    int rc = orc_run(o, tok.v, off.v, (int32_t)off.n - 1);
    iv_free(&tok); iv_free(&off);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* grammar writer (PrintResults.c:339-577, strings of ExtractPair.c:743-796,1021-1123)          */
/* ------------------------------------------------------------------------------------------ */
static void source_string(const orc_t *o, int kind, int32_t cid, char *out) {
    int32_t G = o->G, D1 = o->D1, D2 = o->D2; char *w = out; *w = 0;
#define APP(s) do { w += sprintf(w, "%s", (s)); } while (0)
    if (kind == 0 || (kind == 1 && cid < 2 * G) || (kind == 2 && cid < G)) {
        int32_t blk = (kind == 1 && cid >= G) ? cid - G : cid;
        if ((kind == 1 && cid < G) || kind == 2) APP("[X,1] ");
        int32_t s = o->blocks[4 * blk + 3], l = o->blocks[4 * blk + 2];
        for (int i = 0; i < l; i++) { if (i) APP(" "); APP(o->svocab[o->str[s + i]]); }
        if (kind == 1 && cid >= G) APP(" [X,1]");
        if (kind == 2) APP(" [X,2]");
        return;
    }
    if (kind == 1) {                                             /* aXb */
        const int32_t *p = &o->pat1[10 * (cid - 2 * G)];
        for (int i = 0; i < p[5]; i++) { if (i) APP(" "); APP(p[i] >= 0 ? o->svocab[p[i]] : "[X,1]"); }
        return;
    }
    if (cid < G + D2) {                                          /* aXbXc */
        int32_t d2 = cid - G; const int32_t *p = &o->pat1[10 * o->pat2[4 * d2]];
        for (int i = 0; i < p[5]; i++) { if (i) APP(" "); APP(p[i] >= 0 ? o->svocab[p[i]] : "[X,1]"); }
        APP(" [X,2] "); APP(o->svocab[o->pat2[4 * d2 + 1]]);
        return;
    }
    if (cid < G + D2 + D1) {                                     /* XaXb */
        const int32_t *p = &o->pat1[10 * (cid - G - D2)];
        APP("[X,1]");
        for (int i = 0; i < p[5]; i++) { APP(" "); APP(p[i] >= 0 ? o->svocab[p[i]] : "[X,2]"); }
        return;
    }
    {                                                            /* aXbX */
        const int32_t *p = &o->pat1[10 * (cid - G - D2 - D1)];
        for (int i = 0; i < p[5]; i++) { if (i) APP(" "); APP(p[i] >= 0 ? o->svocab[p[i]] : "[X,1]"); }
        APP(" [X,2]");
    }
#undef APP
}

static void print_group(const orc_t *o, FILE *fp, int kind, int32_t cid) {
    if (cid < 0 || cid >= o->nid[kind]) return;
    int32_t down = o->updown[kind][2 * cid], up = o->updown[kind][2 * cid + 1];
    if (down == -1 || up == -1) return;
    char src[1024]; source_string(o, kind, cid, src);
    for (int32_t i = down; i <= up; i++) {
        const orc_rule_t *r = &o->rules[kind][i];
        rec_t rr = {0, r->rec[0], r->rec[1], r->rec[2], r->rec[3], r->rec[4], r->rec[5]};
        int32_t sym[64]; int ns = target_symbols(o, &rr, sym);
        fprintf(fp, "[X] ||| %s ||| ", src);
        for (int k = 0; k < ns; k++) { if (k) fputc(' ', fp); fputs(sym[k] == -1 ? "[X,1]" : sym[k] == -2 ? "[X,2]" : o->tvocab[sym[k]], fp); }
        fprintf(fp, " ||| EgivenFCoherent=%f SampleCountF=%f CountEF=%f MaxLexFgivenE=%f MaxLexEgivenF=%f IsSingletonF=%d IsSingletonFE=%d\n",
                r->aa, r->score, r->bb, r->mlfe, r->mlef, r->f == 1, r->pc == 1);
    }
}

/* contexts created from arrays carry no vocabulary strings: name every id "w<id>" (tests rename through their own tables) */
static char **synthetic_vocab(int32_t count) {
    char **v = (char **)calloc((size_t)count + 1, sizeof(char *));
    for (int32_t i = 0; i < count; i++) { char b[32]; snprintf(b, sizeof b, "w%d", i); v[i] = strdup(b); }
    return v;
}

int orc_write_grammars(orc_t *o, const char *outdir) {
    if (!o->svocab) { o->sv = o->str[o->n - 1] + 1; o->svocab = synthetic_vocab(o->sv); }
    if (!o->tvocab) { o->tv = o->tgt[o->m - 1] + 1; o->tvocab = synthetic_vocab(o->tv); }
    char fn[4096];
    for (int32_t qi = 0; qi < o->Q; qi++) {
        snprintf(fn, sizeof fn, "%s/grammar.%d.s", outdir, qi);
        FILE *fp = fopen(fn, "w");
        if (!fp) { fprintf(stderr, "oracle: cannot write %s\n", fn); return 1; }
        for (int64_t ic = 0; ic < o->qryglobal[qi].n; ic++) {
            int32_t p = o->qryglobal[qi].v[ic];
            print_group(o, fp, 1, p + o->G);   /* abX */
            print_group(o, fp, 1, p);          /* Xab */
            print_group(o, fp, 2, p);          /* XabX */
            print_group(o, fp, 0, p);          /* ab */
        }
        for (int64_t si = 0; si < o->q1[qi].n; si++) {
            int32_t g = o->q1[qi].v[si];
            print_group(o, fp, 1, 2 * o->G + g);                 /* aXb */
            print_group(o, fp, 2, o->G + o->D2 + g);             /* XaXb */
            print_group(o, fp, 2, o->G + o->D2 + o->D1 + g);     /* aXbX */
        }
        for (int64_t si = 0; si < o->q2[qi].n; si++) print_group(o, fp, 2, o->G + o->q2[qi].v[si]);   /* aXbXc */
        fclose(fp);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
void orc_get_counts(const orc_t *o, orc_counts_t *c) { *c = o->cnt; }
const int32_t *orc_sa(const orc_t *o) { return o->sa; }
const int32_t *orc_longest(const orc_t *o) { return o->longest; }
int orc_interval(const orc_t *o, int32_t t, int32_t mlen, int32_t *up, int32_t *down) {
    if (t < 0 || t >= o->T || mlen < 1 || mlen > o->longest[t]) return 0;
    interval(o, t, mlen, up, down); return 1;
}
const int32_t *orc_blocks(const orc_t *o) { return o->blocks; }
const int32_t *orc_onegap_patterns(const orc_t *o) { return o->pat1; }
const int32_t *orc_onegap_hits(const orc_t *o) { return o->h1; }
const int32_t *orc_twogap_patterns(const orc_t *o) { return o->pat2; }
const int32_t *orc_twogap_hits(const orc_t *o) { return o->h2; }
const int32_t *orc_frequent(const orc_t *o) { return o->freq; }
const int32_t *orc_feature_missing(const orc_t *o) { return o->missing; }
const int32_t *orc_precomp_index(const orc_t *o) { return o->pidx; }
const int32_t *orc_precomp_list(const orc_t *o) { return o->plist; }
int32_t orc_query_list(const orc_t *o, int which, int32_t qi, const int32_t **out) {
    if (qi < 0 || qi >= o->Q) { *out = NULL; return 0; }
    const ivec *v = which == 0 ? &o->qryglobal[qi] : which == 1 ? &o->q1[qi] : &o->q2[qi];
    *out = v->v; return (int32_t)v->n;
}
int32_t orc_records(const orc_t *o, int kind, const int32_t **out) {
    const rvec *v = kind == 0 ? &o->rec_ab : kind == 1 ? &o->rec_1 : &o->rec_2;
    *out = (const int32_t *)v->v; return (int32_t)v->n;
}
int32_t orc_rules(const orc_t *o, int kind, const orc_rule_t **out) { *out = o->rules[kind]; return o->nrules[kind]; }

void orc_destroy(orc_t *o) {
    if (!o) return;
    free_run(o);
    free(o->str); free(o->tgt); free(o->RLP); free(o->L_tar); free(o->R_tar); free(o->sa);
    free(o->lex_key); free(o->lex_v1); free(o->lex_v2); free(o->plist);
    free(o);
}
