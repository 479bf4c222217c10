// cgx-b200: extern "C" entry points (include/cgx_b200.h).  Thin: argument checks, H2D/D2H, stage
// sequencing and CUDA-event timing.  No CPU fallback anywhere: every call needs the CUDA device.
#include "context.h"
#include <algorithm>
#include <new>

using namespace cgx;

#define CGX_TRY(ctx, ...)                                   \
    try {                                                   \
        __VA_ARGS__;                                        \
        return 0;                                           \
    } catch (const CgxError &e) {                           \
        if (ctx) (ctx)->err = e.msg;                        \
        return e.code;                                      \
    } catch (const std::exception &e) {                     \
        if (ctx) (ctx)->err = e.what();                     \
        return 2;                                           \
    }

static thread_local std::string g_create_err;
namespace cgx { thread_local Prof *g_prof = nullptr; }

extern "C" int cgx_version(void) { return 100; }

extern "C" int cgx_create(int device, cgx_ctx_t **out) {
    *out = nullptr;
    try {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) throw CgxError{std::string("no CUDA device: ") + cudaGetErrorString(e)};
        CGX_REQUIRE(device >= 0 && device < count, "device %d out of range (have %d)", device, count);
        CUDA_CHECK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        CGX_REQUIRE(prop.major >= 10, "cgx_b200 kernels are built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
        cgx_ctx *c = new cgx_ctx();
        c->device = device;
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        for (auto &ev : c->batch.ev) CUDA_CHECK(cudaEventCreate(&ev));
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->batch.copy_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->batch.copy_ev, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->batch.done_ev, cudaEventDisableTiming));
        for (auto &r : c->batch.parked) CUDA_CHECK(cudaEventCreateWithFlags(&r.done, cudaEventDisableTiming));
        memset(&c->batch.info, 0, sizeof(c->batch.info));
        *out = c;
        return 0;
    } catch (const CgxError &e) {
        g_create_err = e.msg;
        fprintf(stderr, "cgx_create: %s\n", e.msg.c_str());
        return 1;
    }
}

extern "C" void cgx_destroy(cgx_ctx_t *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    Index &ix = c->ix;
    DevBuf *ib[] = {&ix.str, &ix.sa, &ix.inv[0], &ix.inv[1], &ix.inv[2], &ix.bkt[0], &ix.bkt[1], &ix.bkt[2], &ix.jwin, &ix.xw, &ix.lr, &ix.tok_start, &ix.RLP, &ix.L_tar, &ix.R_tar, &ix.tgt, &ix.freq_flag, &ix.gapw,
                    &ix.lex_key, &ix.lex_v1, &ix.lex_v2, &ix.lex_hash};
    for (auto *b : ib) b->release();
    c->ws.release();
    Batch &b = c->batch;
    DevBuf *bb[] = {&b.q_tok, &b.q_off, &b.tok2q, &b.longest, &b.iv, &b.ph_keys, &b.ph_keys_tmp, &b.ph_vals, &b.ph_vals_tmp, &b.ph_flags, &b.phrase_id,
                    &b.phrases, &b.e1_count, &b.e1_inst, &b.e1_keys, &b.e1_keys_tmp, &b.e1_vals, &b.e1_vals_tmp, &b.e1_flags, &b.e1_pid, &b.pat1,
                    &b.pat1_dev, &b.pat1_pos, &b.ql_keys, &b.ql_keys_tmp, &b.q1_off, &b.q1_ids, &b.q2_off, &b.q2_ids, &b.j_tiles, &b.j_bitmaps, &b.j_aflag, &b.j_aid, &b.j_hash, &b.j_status, &b.j_segcnt, &b.j_flags, &b.pat1_ga, &b.hit_keys,
                    &b.hit_keys_tmp, &b.counters, &b.missing, &b.hits1_sorted, &b.hits2_sorted, &b.e2_count, &b.e2_keys, &b.e2_keys_tmp, &b.e2_vals,
                    &b.e2_vals_tmp, &b.e2_flags, &b.pat2, &b.rec_hash, &b.rec_tag, &b.rec_live, &b.rec_list, &b.slot_hint, &b.rec_flags, &b.rec_meta, &b.rec_cnt,
                    &b.scratch, &b.scratch2, &b.rule_head, &b.rule_id, &b.radix.hist, &b.radix.status, &b.radix.counters};
    for (auto *x : bb) x->release();
    for (int k = 0; k < 3; k++) { b.slot_off[k].release(); b.rec[k].release(); b.rules[k].release(); b.updown[k].release(); b.idinfo[k].release(); b.id_count[k].release(); }
    for (auto &l : b.scan.level) l.release();
    if (b.h_pinned) cudaFreeHost(b.h_pinned);
    PinnedBuf *pb[] = {&b.h_phrase_id, &b.h_phrases, &b.h_pat1, &b.h_pat2, &b.h_q1_off, &b.h_q1_ids, &b.h_q2_off, &b.h_q2_ids};
    for (auto *x : pb) x->release();
    for (int k = 0; k < 3; k++) { b.h_rules[k].release(); b.h_updown[k].release(); b.h_idinfo[k].release(); }
    c->prof.destroy();
    for (auto &ev : b.ev) if (ev) cudaEventDestroy(ev);
    if (b.copy_ev) cudaEventDestroy(b.copy_ev);
    if (b.done_ev) cudaEventDestroy(b.done_ev);
    for (auto &r : b.parked) r.release();
    if (b.copy_stream) cudaStreamDestroy(b.copy_stream);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" const char *cgx_last_error(const cgx_ctx_t *c) { return c ? c->err.c_str() : g_create_err.c_str(); }

// ------------------------------------------------------------------------------------------------
// index
// ------------------------------------------------------------------------------------------------
static int32_t max_token_of(const int32_t *src, int64_t n) {
    int32_t mx = 0;
    for (int64_t i = 0; i < n; i++) mx = std::max(mx, src[i]);
    return mx;
}

static void build_index_device(cgx_ctx *c) {
    Index &ix = c->ix;
    cudaEvent_t e0, e1, e2;
    CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1)); CUDA_CHECK(cudaEventCreate(&e2));
    c->prof.stream = c->stream;
    g_prof = &c->prof;
    CUDA_CHECK(cudaEventRecord(e0, c->stream));
    build_suffix_array(ix.str.ptr<int32_t>(), ix.n, ix.maxtok, ix.sa.get<int32_t>(ix.n), c->ws, c->stream, &ix.sa_stats);
    CUDA_CHECK(cudaEventRecord(e1, c->stream));
    int launches = 0;
    build_index_aux(ix, c->ws, c->stream, &launches);
    CUDA_CHECK(cudaEventRecord(e2, c->stream));
    CUDA_CHECK(cudaEventSynchronize(e2));
    c->prof.resolve();
    g_prof = nullptr;
    CUDA_CHECK(cudaEventElapsedTime(&ix.sa_stats.ms, e0, e1));
    CUDA_CHECK(cudaEventElapsedTime(&c->aux_ms, e1, e2));
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    c->ws.release();                                  // the build workspace (32 B/token) is not needed at query time
    ix.built = true;
    c->batch.adv_refused_q = c->batch.adv_ok_q = 0;   // batch-size advice belongs to the corpus
}

static uint64_t fnv1a_words(const int32_t *w, size_t count, uint64_t h = 1469598103934665603ull) {
    for (size_t i = 0; i < count; i++) { h ^= (uint32_t)w[i]; h *= 1099511628211ull; }
    return h ? h : 1;
}

static void index_build_host(cgx_ctx *c, const int32_t *src, int64_t n, const int32_t *tgt, int64_t m, const void *RLP, const void *L_tar, const void *R_tar,
                             bool wide) {
    CGX_REQUIRE(c && src && tgt && RLP && L_tar && R_tar, "null argument");
    CGX_REQUIRE(n >= 4 && m >= 1, "empty corpus");
    CUDA_CHECK(cudaSetDevice(c->device));
    Index &ix = c->ix;
    ix.n = (size_t)n; ix.m = (size_t)m; ix.wide = wide;
    ix.maxtok = max_token_of(src, n);
    CGX_REQUIRE(src[n] == 0 && src[n + 1] == 0 && src[n + 2] == 0, "source text must be followed by three zeros (Start.cu:354)");
    const size_t wb = wide ? 8 : 4, lb = wide ? 2 : 1;                 // bytes per RLP word / per L_tar, R_tar entry (align_fields.cuh)
    CUDA_CHECK(cudaMemcpyAsync(ix.str.get<int32_t>(ix.n + 3), src, sizeof(int32_t) * (ix.n + 3), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(ix.tgt.get<int32_t>(ix.m + 3), tgt, sizeof(int32_t) * (ix.m + 3), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(ix.RLP.get<uint8_t>(ix.n * wb), RLP, ix.n * wb, cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(ix.L_tar.get<uint8_t>(ix.m * lb), L_tar, ix.m * lb, cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(ix.R_tar.get<uint8_t>(ix.m * lb), R_tar, ix.m * lb, cudaMemcpyHostToDevice, c->stream));
    build_index_device(c);
    ix.src_sum = fnv1a_words(src, ix.n + 3);
    ix.tgt_sum = fnv1a_words(tgt, ix.m + 3);
}

extern "C" int cgx_index_build_wide(cgx_ctx_t *c, const int32_t *src, int64_t n, const int32_t *tgt, int64_t m, const uint64_t *RLP64,
                                    const uint16_t *L_tar16, const uint16_t *R_tar16) {
    CGX_TRY(c, index_build_host(c, src, n, tgt, m, RLP64, L_tar16, R_tar16, true));
}

extern "C" int cgx_index_build(cgx_ctx_t *c, const int32_t *src, int64_t n, const int32_t *tgt, int64_t m, const uint32_t *RLP,
                               const uint8_t *L_tar, const uint8_t *R_tar) {
    CGX_TRY(c, {
        const char *force = getenv("CGX_FORCE_WIDE");
        if (force && force[0] == '1' && c && src && tgt && RLP && L_tar && R_tar && n >= 4 && m >= 1) {     // tests: the 16-bit layout on an 8-bit corpus
            std::vector<uint64_t> w((size_t)n);
            std::vector<uint16_t> l((size_t)m), r((size_t)m);
            for (int64_t i = 0; i < n; i++) {
                const uint32_t x = RLP[i];
                const bool eos = src[i] < 2;                            // the word at an EOS is the target sentence offset
                const uint64_t L = (x >> 24) & 0xFF, R = (x >> 16) & 0xFF, P = (x >> 8) & 0xFF;
                w[(size_t)i] = eos ? (uint64_t)x : (((L == 255 ? 65535ull : L) << 48) | ((R == 255 ? 65535ull : R) << 32) | (P << 16));
            }
            for (int64_t i = 0; i < m; i++) { l[(size_t)i] = L_tar[i] == 255 ? 65535 : L_tar[i]; r[(size_t)i] = R_tar[i] == 255 ? 65535 : R_tar[i]; }
            index_build_host(c, src, n, tgt, m, w.data(), l.data(), r.data(), true);
        } else index_build_host(c, src, n, tgt, m, RLP, L_tar, R_tar, false);
    });
}

extern "C" int cgx_sa_build_dev(cgx_ctx_t *c, const int32_t *str_dev, int64_t n, int32_t max_token, int32_t *sa_dev, int32_t *rounds_out,
                                float *ms_out) {
    CGX_TRY(c, {
        CGX_REQUIRE(c && str_dev && sa_dev && n >= 2, "bad argument");
        CUDA_CHECK(cudaSetDevice(c->device));
        cudaEvent_t e0, e1;
        CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
        SaStats st;
        CUDA_CHECK(cudaEventRecord(e0, c->stream));
        build_suffix_array(str_dev, (size_t)n, max_token, sa_dev, c->ws, c->stream, &st);
        CUDA_CHECK(cudaEventRecord(e1, c->stream));
        CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (rounds_out) *rounds_out = st.rounds;
        if (ms_out) *ms_out = ms;
        c->ix.sa_stats = st;
        c->ix.sa_stats.ms = ms;
    });
}

__global__ void lex_keys_kernel(const int32_t *__restrict__ f, const int32_t *__restrict__ e, size_t n, uint64_t *__restrict__ keys, uint32_t *__restrict__ idx) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = ((uint64_t)(uint32_t)(f[i] + 1) << 32) | (uint64_t)(uint32_t)(e[i] + 1);
    idx[i] = (uint32_t)i;
}
__global__ void lex_gather_kernel(const uint32_t *__restrict__ idx, const float *__restrict__ v1, const float *__restrict__ v2, size_t n,
                                  float *__restrict__ o1, float *__restrict__ o2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    o1[i] = v1[idx[i]];
    o2[i] = v2[idx[i]];
}

extern "C" int cgx_lex_load(cgx_ctx_t *c, const int32_t *f, const int32_t *e, const float *v1, const float *v2, int64_t count) {
    CGX_TRY(c, {
        CGX_REQUIRE(c && (count == 0 || (f && e && v1 && v2)), "null argument");
        CUDA_CHECK(cudaSetDevice(c->device));
        Index &ix = c->ix;
        size_t n = (size_t)count;
        ix.lex_count = n;
        uint64_t *keys = ix.lex_key.get<uint64_t>(n + 1);
        float *o1 = ix.lex_v1.get<float>(n + 1), *o2 = ix.lex_v2.get<float>(n + 1);
        if (n == 0) { build_lex_hash(ix, c->stream); return 0; }
        DevBuf df, de, d1, d2, kt, ix0, ix1, ksrc;
        RadixTemp rt;
        int32_t *pf = df.get<int32_t>(n), *pe = de.get<int32_t>(n);
        float *p1 = d1.get<float>(n), *p2 = d2.get<float>(n);
        CUDA_CHECK(cudaMemcpyAsync(pf, f, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(pe, e, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(p1, v1, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(p2, v2, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
        uint64_t *k0 = ksrc.get<uint64_t>(n), *k1 = kt.get<uint64_t>(n);
        uint32_t *i0 = ix0.get<uint32_t>(n), *i1 = ix1.get<uint32_t>(n);
        lex_keys_kernel<<<cgx_div_up(n, 256), 256, 0, c->stream>>>(pf, pe, n, k0, i0);
        uint64_t *ks;
        uint32_t *is;
        radix_sort<uint64_t>(k0, k1, i0, i1, n, 0, 64, c->stream, rt, &ks, &is);
        CUDA_CHECK(cudaMemcpyAsync(keys, ks, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, c->stream));
        lex_gather_kernel<<<cgx_div_up(n, 256), 256, 0, c->stream>>>(is, p1, p2, n, o1, o2);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        build_lex_hash(ix, c->stream);
        DevBuf *tmp[] = {&df, &de, &d1, &d2, &kt, &ix0, &ix1, &ksrc, &rt.hist, &rt.status, &rt.counters};
        for (auto *t : tmp) t->release();
    });
}

extern "C" int cgx_index_info(const cgx_ctx_t *c, cgx_index_info_t *out) {
    if (!c || !out) return 1;
    const Index &ix = c->ix;
    out->n = (int64_t)ix.n; out->m = (int64_t)ix.m;
    out->sa_rounds = ix.sa_stats.rounds; out->sa_key_bits = ix.sa_stats.key_bits; out->sa_launches = ix.sa_stats.launches;
    out->sa_build_ms = ix.sa_stats.ms; out->aux_build_ms = c->aux_ms;
    out->index_bytes = (int64_t)(ix.str.cap + ix.sa.cap + ix.inv[0].cap + ix.inv[1].cap + ix.inv[2].cap + ix.bkt[0].cap + ix.bkt[1].cap + ix.bkt[2].cap + ix.jwin.cap + ix.xw.cap + ix.lr.cap + ix.tok_start.cap + ix.RLP.cap + ix.L_tar.cap +
                                 ix.R_tar.cap + ix.tgt.cap + ix.freq_flag.cap + ix.gapw.cap + ix.lex_key.cap + ix.lex_v1.cap + ix.lex_v2.cap + ix.lex_hash.cap);
    return 0;
}

extern "C" int cgx_index_export(cgx_ctx_t *c, cgx_index_arrays_t *o) {
    CGX_TRY(c, {
        CGX_REQUIRE(c && o && c->ix.built, "index not built");
        Index &ix = c->ix;
        o->n = (int64_t)ix.n; o->m = (int64_t)ix.m; o->lex_count = (int64_t)ix.lex_count; o->max_token = ix.maxtok; o->wide = ix.wide ? 1 : 0;
        memcpy(o->freq_list, ix.freq_list, sizeof(ix.freq_list));
        o->str = ix.str.p; o->sa = ix.sa.p; o->inv1 = ix.inv[0].p; o->inv2 = ix.inv[1].p; o->inv3 = ix.inv[2].p; o->bkt1 = ix.bkt[0].p; o->bkt2 = ix.bkt[1].p; o->bkt3 = ix.bkt[2].p; o->tok_start = ix.tok_start.p;
        o->RLP = ix.RLP.p; o->L_tar = ix.L_tar.p; o->R_tar = ix.R_tar.p; o->tgt = ix.tgt.p; o->freq_flag = ix.freq_flag.p; o->gapw = ix.gapw.p;
        o->lex_key = ix.lex_key.p; o->lex_v1 = ix.lex_v1.p; o->lex_v2 = ix.lex_v2.p;
    });
}

extern "C" int cgx_index_alloc(cgx_ctx_t *c, const cgx_index_arrays_t *s, cgx_index_arrays_t *o) {
    CGX_TRY(c, {
        CGX_REQUIRE(c && s && o, "null argument");
        CUDA_CHECK(cudaSetDevice(c->device));
        Index &ix = c->ix;
        ix.n = (size_t)s->n; ix.m = (size_t)s->m; ix.lex_count = (size_t)s->lex_count; ix.maxtok = s->max_token; ix.wide = s->wide != 0;
        memcpy(ix.freq_list, s->freq_list, sizeof(ix.freq_list));
        size_t nt = (size_t)ix.maxtok + 2;
        ix.str.get<int32_t>(ix.n + 3); ix.sa.get<int32_t>(ix.n);
        for (int k = 0; k < 3; k++) { ix.inv[k].get<int32_t>(ix.n); ix.bkt[k].get<int32_t>(ix.n); }
        ix.tok_start.get<int32_t>(nt); ix.RLP.get<uint8_t>(ix.n * (ix.wide ? 8 : 4)); ix.L_tar.get<uint8_t>(ix.m * (ix.wide ? 2 : 1)); ix.R_tar.get<uint8_t>(ix.m * (ix.wide ? 2 : 1));
        ix.tgt.get<int32_t>(ix.m + 3); ix.freq_flag.get<uint8_t>(nt); ix.gapw.get<uint32_t>(ix.n);
        ix.lex_key.get<uint64_t>(ix.lex_count + 1); ix.lex_v1.get<float>(ix.lex_count + 1); ix.lex_v2.get<float>(ix.lex_count + 1);
        ix.built = false;
        *o = *s;
        o->str = ix.str.p; o->sa = ix.sa.p; o->inv1 = ix.inv[0].p; o->inv2 = ix.inv[1].p; o->inv3 = ix.inv[2].p; o->bkt1 = ix.bkt[0].p; o->bkt2 = ix.bkt[1].p; o->bkt3 = ix.bkt[2].p; o->tok_start = ix.tok_start.p;
        o->RLP = ix.RLP.p; o->L_tar = ix.L_tar.p; o->R_tar = ix.R_tar.p; o->tgt = ix.tgt.p; o->freq_flag = ix.freq_flag.p; o->gapw = ix.gapw.p;
        o->lex_key = ix.lex_key.p; o->lex_v1 = ix.lex_v1.p; o->lex_v2 = ix.lex_v2.p;
    });
}

extern "C" int cgx_index_commit(cgx_ctx_t *c) {
    CGX_TRY(c, {
        CGX_REQUIRE(c, "null context");
        CUDA_CHECK(cudaSetDevice(c->device));
        build_lex_hash(c->ix, c->stream);          // derived from the (broadcast) sorted lexical arrays
        build_jwin(c->ix, c->stream);              // derived from the (broadcast) bucket arrays, gap words, text and alignment arrays
        c->ix.built = true;
        c->batch.adv_refused_q = c->batch.adv_ok_q = 0;
    });
}

// ---- persisted index (the reference only has the dead sa_precomp.txt stub, SuffixArray.c:208-230) --------------
// The file holds what cannot be recomputed on the device -- the two token arrays, the alignment arrays and the sorted lexical table
// (18 bytes per token pair at C2) -- and a load rebuilds the suffix array and the auxiliary arrays from them (30 ms at C2).  Round 2
// first saved every resident array (88 bytes per token pair): reading 2.3 GB took 1.3 s from the page cache and 3-5 s from disk,
// longer than parsing the text files and building from scratch.
struct IndexFileHeader {
    char magic[8];                 // "CGXIDX03"
    int64_t n, m, lex_count;
    int32_t max_token, sa_rounds, sa_key_bits, wide;      // wide: 16-bit alignment fields (0 in files written before the field existed)
    int32_t freq_list[CGX_PRECOMP];
    uint64_t src_sum, tgt_sum;     // FNV-1a of the source / target token arrays (0 = not recorded)
};
struct IndexArrayRef { void *p; size_t bytes; };
static int index_array_refs(const cgx_index_arrays_t &a, IndexArrayRef out[18]) {
    const size_t n = (size_t)a.n, m = (size_t)a.m, nt = (size_t)a.max_token + 2, lx = (size_t)a.lex_count + 1;
    const IndexArrayRef r[18] = {{a.str, 4 * (n + 3)}, {a.sa, 4 * n}, {a.inv1, 4 * n}, {a.inv2, 4 * n}, {a.inv3, 4 * n}, {a.bkt1, 4 * n}, {a.bkt2, 4 * n}, {a.bkt3, 4 * n},
                                 {a.tok_start, 4 * nt}, {a.RLP, (a.wide ? 8 : 4) * n}, {a.L_tar, (a.wide ? 2 : 1) * m}, {a.R_tar, (a.wide ? 2 : 1) * m}, {a.tgt, 4 * (m + 3)}, {a.freq_flag, nt}, {a.gapw, 4 * n},
                                 {a.lex_key, 8 * lx}, {a.lex_v1, 4 * lx}, {a.lex_v2, 4 * lx}};
    for (int i = 0; i < 18; i++) out[i] = r[i];
    return 18;
}
// the arrays of index_array_refs that a file holds: str, RLP, L_tar, R_tar, tgt, lex_key, lex_v1, lex_v2
static const int kPersisted[8] = {0, 9, 10, 11, 12, 15, 16, 17};

extern "C" int cgx_index_save(cgx_ctx_t *c, const char *path) {
    FILE *fh = nullptr;
    void *stage = nullptr;
    try {
        CGX_REQUIRE(c && path && c->ix.built, "index not built");
        CUDA_CHECK(cudaSetDevice(c->device));
        cgx_index_arrays_t a;
        CGX_REQUIRE(cgx_index_export(c, &a) == 0, "export failed");
        const std::string tmp_path = std::string(path) + ".tmp";       // written beside the target and renamed when complete: a failed save leaves no truncated index behind
        fh = fopen(tmp_path.c_str(), "wb");
        CGX_REQUIRE(fh, "cannot open %s for writing", tmp_path.c_str());
        IndexFileHeader h;
        memset(&h, 0, sizeof h);
        memcpy(h.magic, "CGXIDX03", 8);
        h.n = a.n; h.m = a.m; h.lex_count = a.lex_count; h.max_token = a.max_token; h.sa_rounds = c->ix.sa_stats.rounds; h.sa_key_bits = c->ix.sa_stats.key_bits; h.wide = a.wide;
        h.src_sum = c->ix.src_sum; h.tgt_sum = c->ix.tgt_sum;
        memcpy(h.freq_list, a.freq_list, sizeof h.freq_list);
        CGX_REQUIRE(fwrite(&h, sizeof h, 1, fh) == 1, "write failed");
        const size_t piece = (size_t)64 << 20;
        CUDA_CHECK(cudaMallocHost(&stage, piece));
        IndexArrayRef r[18];
        index_array_refs(a, r);
        for (int i : kPersisted)
            for (size_t o = 0; o < r[i].bytes; o += piece) {
                const size_t len = std::min(piece, r[i].bytes - o);
                CUDA_CHECK(cudaMemcpy(stage, (const char *)r[i].p + o, len, cudaMemcpyDeviceToHost));
                CGX_REQUIRE(fwrite(stage, 1, len, fh) == len, "write failed (disk full?)");
            }
        cudaFreeHost(stage);
        stage = nullptr;
        FILE *done = fh;
        fh = nullptr;
        CGX_REQUIRE(fclose(done) == 0, "close failed");
        CGX_REQUIRE(rename(tmp_path.c_str(), path) == 0, "cannot rename %s to %s", tmp_path.c_str(), path);
        return 0;
    } catch (const CgxError &e) {
        if (fh) fclose(fh);
        if (stage) cudaFreeHost(stage);
        if (c) c->err = e.msg;
        return e.code;
    }
}

extern "C" int cgx_index_load(cgx_ctx_t *c, const char *path) {
    FILE *fh = nullptr;
    void *stage = nullptr;
    try {
        CGX_REQUIRE(c && path, "null argument");
        CUDA_CHECK(cudaSetDevice(c->device));
        fh = fopen(path, "rb");
        CGX_REQUIRE(fh, "cannot open %s", path);
        IndexFileHeader h;
        CGX_REQUIRE(fread(&h, sizeof h, 1, fh) == 1 && memcmp(h.magic, "CGXIDX03", 8) == 0, "%s is not a cgx-b200 index file of this version", path);
        CGX_REQUIRE(h.n >= 4 && h.m >= 1 && h.lex_count >= 0 && h.max_token >= 1, "%s: corrupt header", path);
        cgx_index_arrays_t shape, a;
        memset(&shape, 0, sizeof shape);
        shape.n = h.n; shape.m = h.m; shape.lex_count = h.lex_count; shape.max_token = h.max_token; shape.wide = h.wide;
        memcpy(shape.freq_list, h.freq_list, sizeof h.freq_list);
        CGX_REQUIRE(cgx_index_alloc(c, &shape, &a) == 0, "%s", c->err.c_str());
        const size_t piece = (size_t)64 << 20;
        CUDA_CHECK(cudaMallocHost(&stage, piece));
        IndexArrayRef r[18];
        index_array_refs(a, r);
        for (int i : kPersisted)
            for (size_t o = 0; o < r[i].bytes; o += piece) {
                const size_t len = std::min(piece, r[i].bytes - o);
                CGX_REQUIRE(fread(stage, 1, len, fh) == len, "%s: truncated", path);
                CUDA_CHECK(cudaMemcpy((char *)r[i].p + o, stage, len, cudaMemcpyHostToDevice));
            }
        CGX_REQUIRE(fgetc(fh) == EOF, "%s: longer than its header says", path);
        cudaFreeHost(stage);
        stage = nullptr;
        fclose(fh);
        fh = nullptr;
        build_index_device(c);                     // suffix array + auxiliary arrays from the token and alignment arrays just read
        build_lex_hash(c->ix, c->stream);          // derived from the sorted lexical arrays
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->ix.src_sum = h.src_sum; c->ix.tgt_sum = h.tgt_sum;
        return 0;
    } catch (const CgxError &e) {
        if (fh) fclose(fh);
        if (stage) cudaFreeHost(stage);
        if (c) c->err = e.msg;
        return e.code;
    }
}

extern "C" int cgx_index_matches(const cgx_ctx_t *c, const int32_t *src, int64_t n, const int32_t *tgt, int64_t m) {
    if (!c || !src || !tgt || !c->ix.built) return 0;
    const Index &ix = c->ix;
    if ((int64_t)ix.n != n || (int64_t)ix.m != m) return 0;
    if (ix.src_sum == 0 || ix.tgt_sum == 0) return 1;                     // (an index adopted from a broadcast carries no checksums)
    return fnv1a_words(src, (size_t)n + 3) == ix.src_sum && fnv1a_words(tgt, (size_t)m + 3) == ix.tgt_sum;
}

extern "C" int cgx_index_copy_sa(cgx_ctx_t *c, int32_t *out) {
    CGX_TRY(c, {
        CGX_REQUIRE(c && out && c->ix.n, "no index");
        CUDA_CHECK(cudaSetDevice(c->device));
        CUDA_CHECK(cudaMemcpy(out, c->ix.sa.p, sizeof(int32_t) * c->ix.n, cudaMemcpyDeviceToHost));
    });
}
extern "C" int cgx_index_copy_inv(cgx_ctx_t *c, int which, int32_t *out) {
    CGX_TRY(c, {
        CGX_REQUIRE(c && out && c->ix.built && which >= 1 && which <= 3, "bad argument");
        CUDA_CHECK(cudaSetDevice(c->device));
        CUDA_CHECK(cudaMemcpy(out, c->ix.inv[which - 1].p, sizeof(int32_t) * c->ix.n, cudaMemcpyDeviceToHost));
    });
}
extern "C" int cgx_index_copy_frequent(cgx_ctx_t *c, int32_t *out) {
    if (!c || !out || !c->ix.built) return 1;
    memcpy(out, c->ix.freq_list, sizeof(int32_t) * CGX_PRECOMP);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// query batch
// ------------------------------------------------------------------------------------------------

// Runs the stages on queries already resident in b.q_tok / b.q_off / b.tok2q.
// wait_copies = false (cgx_extract_begin): returns when the kernels are done; the tail of the D2H is still travelling.
static void run_batch(cgx_ctx *c, int32_t Q, int32_t T, bool fetch, bool wait_copies = true) {
    Batch &b = c->batch;
    const Index &ix = c->ix;
    cudaStream_t s = c->stream;
    b.fetch_results = fetch;
    if (!ix.lex_hash.p) build_lex_hash(c->ix, s);      // no lexical table loaded: an empty one (every weight = MAXSCORE)
    c->prof.stream = s;
    g_prof = &c->prof;
    stage_lookup(ix, b, s);
    CUDA_CHECK(cudaEventRecord(b.ev[1], s));
    stage_phrases(ix, b, s);
    stage_onegap_enumerate(ix, b, s);
    CUDA_CHECK(cudaEventRecord(b.ev[2], s));
    try {
        stage_onegap_join(ix, b, s);
        CUDA_CHECK(cudaEventRecord(b.ev[3], s));
        stage_twogap_enumerate(ix, b, s);
        CUDA_CHECK(cudaEventRecord(b.ev[4], s));
        stage_twogap_join(ix, b, s);
    CUDA_CHECK(cudaEventRecord(b.ev[5], s));
    if (fetch) {   // the pattern tables and phrase ids are final here: they travel while extraction and aggregation run
        int32_t *hp = b.h_phrase_id.get<int32_t>((size_t)T * CGX_LONGEST_SRC + 1);
        int32_t *hph = b.h_phrases.get<int32_t>((size_t)b.G * 4 + 1);
        int32_t *h1 = b.h_pat1.get<int32_t>((size_t)b.D1 * 4 + 1);       // host view: {a_pos, ls, b_pos, le} of the 32-byte device record
        int32_t *h2 = b.h_pat2.get<int32_t>((size_t)b.D2 * 2 + 1);       // host view: {pat1, ctok} of the 16-byte device record
        fetch_async(b, hp, b.phrase_id.p, sizeof(int32_t) * (size_t)T * CGX_LONGEST_SRC, s);
        fetch_async(b, hph, b.phrases.p, sizeof(int32_t) * (size_t)b.G * 4, s);
        fetch_async_2d(b, h1, 16, b.pat1.p, sizeof(Pat1), (size_t)b.D1, s);      // the hit ranges, marker and featureMissingCount stay on the device
        fetch_async_2d(b, h2, 8, b.pat2.p, sizeof(Pat2), (size_t)b.D2, s);
    }
    stage_extract(ix, b, s);
    CUDA_CHECK(cudaEventRecord(b.ev[6], s));
    stage_aggregate(ix, b, s);
    } catch (const CgxError &e) {                       // refused (hit / cell indices, or device memory): remembered for cgx_batch_advice
        if (e.code == 3 && (b.adv_refused_q == 0 || Q < b.adv_refused_q)) b.adv_refused_q = Q;
        throw;
    }
    b.adv_ok_q = Q;
    b.adv_ok_hits = (double)std::max(b.hits1, b.hits2);
    if (fetch) {   // the batch ends when the last result byte is on the host
        CUDA_CHECK(cudaEventRecord(b.done_ev, b.copy_stream));
        if (wait_copies) CUDA_CHECK(cudaStreamWaitEvent(s, b.done_ev, 0));
    }
    b.valid = true;
    CUDA_CHECK(cudaEventRecord(b.ev[7], s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    c->prof.resolve();
    g_prof = nullptr;
    cgx_batch_info_t &in = b.info;
    in.Q = Q; in.T = T; in.G = b.G; in.enu1 = b.enu1; in.D1 = b.D1; in.hits1 = b.hits1; in.enu2 = b.enu2; in.D2 = b.D2; in.hits2 = b.hits2;
    in.samples = b.samples; in.n_ab = b.n_rec[0]; in.n_1gap = b.n_rec[1]; in.n_2gap = b.n_rec[2];
    for (int k = 0; k < 3; k++) in.rules[k] = b.n_rules[k];
    in.launches = b.launches;
    float t1, t2, t3, t4;
    CUDA_CHECK(cudaEventElapsedTime(&in.ms_total, b.ev[0], b.ev[7]));
    CUDA_CHECK(cudaEventElapsedTime(&in.ms_lookup, b.ev[0], b.ev[1]));
    CUDA_CHECK(cudaEventElapsedTime(&t1, b.ev[1], b.ev[2]));
    CUDA_CHECK(cudaEventElapsedTime(&t2, b.ev[3], b.ev[4]));
    in.ms_enum = t1 + t2;
    CUDA_CHECK(cudaEventElapsedTime(&t3, b.ev[2], b.ev[3]));
    CUDA_CHECK(cudaEventElapsedTime(&t4, b.ev[4], b.ev[5]));
    in.ms_join = t3 + t4;
    CUDA_CHECK(cudaEventElapsedTime(&in.ms_extract, b.ev[5], b.ev[6]));
    CUDA_CHECK(cudaEventElapsedTime(&in.ms_aggregate, b.ev[6], b.ev[7]));
}

static void extract_host_run(cgx_ctx *c, const int32_t *qry_tok, const int32_t *qry_off, int32_t Q, int32_t T, bool pipelined);

static void extract_host(cgx_ctx *c, const int32_t *qry_tok, const int32_t *qry_off, int32_t Q, bool pipelined) {
    CGX_REQUIRE(c && qry_off && Q >= 0, "bad argument");
    CGX_REQUIRE(c->ix.built, "index not built");
    CUDA_CHECK(cudaSetDevice(c->device));
    Batch &b = c->batch;
    const int32_t T = qry_off[Q];
    CGX_REQUIRE(T == 0 || qry_tok, "null query tokens");
    if (pipelined) rotate_results(b);
    try {
        extract_host_run(c, qry_tok, qry_off, Q, T, pipelined);
    } catch (...) {
        if (pipelined) { swap_results(b, b.parked[1]); swap_results(b, b.parked[0]); }      // a failed batch leaves the pipeline as it was
        throw;
    }
}

static void extract_host_run(cgx_ctx *c, const int32_t *qry_tok, const int32_t *qry_off, int32_t Q, int32_t T, bool pipelined) {
    Batch &b = c->batch;
    cudaStream_t s = c->stream;
    b.valid = false;
    b.Q = Q; b.T = T; b.launches = 0;
    memset(&b.info, 0, sizeof(b.info));
    // stage the inputs through pinned memory: [tok T][off Q+1][tok2q T]
    size_t need = (size_t)T * 2 + (size_t)Q + 1;
    if (need > b.h_pinned_cap) {
        if (b.h_pinned) CUDA_CHECK(cudaFreeHost(b.h_pinned));
        CUDA_CHECK(cudaMallocHost((void **)&b.h_pinned, sizeof(int32_t) * (need + need / 4 + 64)));
        b.h_pinned_cap = need + need / 4 + 64;
    }
    int32_t *hp = b.h_pinned;
    if (T) memcpy(hp, qry_tok, sizeof(int32_t) * (size_t)T);
    memcpy(hp + T, qry_off, sizeof(int32_t) * ((size_t)Q + 1));
    int32_t *t2q = hp + T + Q + 1;
    for (int32_t q = 0; q < Q; q++) {
        CGX_REQUIRE(qry_off[q + 1] >= qry_off[q], "query offsets must be non-decreasing");
        for (int32_t t = qry_off[q]; t < qry_off[q + 1]; t++) t2q[t] = q;
    }
    CUDA_CHECK(cudaEventRecord(b.ev[0], s));
    int32_t *d_tok = b.q_tok.get<int32_t>((size_t)T + 8);
    int32_t *d_off = b.q_off.get<int32_t>((size_t)Q + 1);
    int32_t *d_t2q = b.tok2q.get<int32_t>((size_t)T + 1);
    if (T) CUDA_CHECK(cudaMemcpyAsync(d_tok, hp, sizeof(int32_t) * (size_t)T, cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(d_off, hp + T, sizeof(int32_t) * ((size_t)Q + 1), cudaMemcpyHostToDevice, s));
    if (T) CUDA_CHECK(cudaMemcpyAsync(d_t2q, t2q, sizeof(int32_t) * (size_t)T, cudaMemcpyHostToDevice, s));
    run_batch(c, Q, T, true, !pipelined);
}

extern "C" int cgx_extract(cgx_ctx_t *c, const int32_t *qry_tok, const int32_t *qry_off, int32_t Q) {
    CGX_TRY(c, extract_host(c, qry_tok, qry_off, Q, false));
}

extern "C" int cgx_extract_begin(cgx_ctx_t *c, const int32_t *qry_tok, const int32_t *qry_off, int32_t Q) {
    CGX_TRY(c, extract_host(c, qry_tok, qry_off, Q, true));
}

extern "C" int cgx_extract_dev(cgx_ctx_t *c, const int32_t *qry_tok_dev, const int32_t *qry_off_dev, const int32_t *tok2q_dev, int32_t Q, int32_t T) {
    CGX_TRY(c, {
        CGX_REQUIRE(c && qry_off_dev && Q >= 0 && T >= 0 && (T == 0 || (qry_tok_dev && tok2q_dev)), "bad argument");
        CGX_REQUIRE(c->ix.built, "index not built");
        CUDA_CHECK(cudaSetDevice(c->device));
        Batch &b = c->batch;
        cudaStream_t s = c->stream;
        b.valid = false;
        b.Q = Q; b.T = T; b.launches = 0;
        memset(&b.info, 0, sizeof(b.info));
        int32_t *d_tok = b.q_tok.get<int32_t>((size_t)T + 8);
        int32_t *d_off = b.q_off.get<int32_t>((size_t)Q + 1);
        int32_t *d_t2q = b.tok2q.get<int32_t>((size_t)T + 1);
        CUDA_CHECK(cudaEventRecord(b.ev[0], s));
        if (T) CUDA_CHECK(cudaMemcpyAsync(d_tok, qry_tok_dev, sizeof(int32_t) * (size_t)T, cudaMemcpyDeviceToDevice, s));
        CUDA_CHECK(cudaMemcpyAsync(d_off, qry_off_dev, sizeof(int32_t) * ((size_t)Q + 1), cudaMemcpyDeviceToDevice, s));
        if (T) CUDA_CHECK(cudaMemcpyAsync(d_t2q, tok2q_dev, sizeof(int32_t) * (size_t)T, cudaMemcpyDeviceToDevice, s));
        run_batch(c, Q, T, false);
    });
}

extern "C" int cgx_profile_enable(cgx_ctx_t *c, int on) {
    if (!c) return 1;
    c->prof.enabled = on != 0;
    c->prof.reset();
    return 0;
}

extern "C" const char *cgx_profile_report(cgx_ctx_t *c) {
    if (!c) return "";
    std::string j = "{";
    bool first = true;
    for (auto &kv : c->prof.table) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %ld, \"ms\": %.6f, \"bytes\": %.0f}", first ? "" : ", ", kv.first.c_str(), kv.second.launches,
                 kv.second.ms, kv.second.bytes);
        j += buf;
        first = false;
    }
    j += "}";
    c->prof_json = j;
    return c->prof_json.c_str();
}

extern "C" int cgx_batch_info(const cgx_ctx_t *c, cgx_batch_info_t *out) {
    if (!c || !out) return 1;
    *out = c->batch.info;
    return 0;
}

extern "C" int32_t cgx_batch_advice(const cgx_ctx_t *c, int32_t wanted) {
    if (!c || wanted <= 1) return wanted;
    const Batch &b = c->batch;
    if (b.adv_refused_q == 0) return wanted;                                   // nothing has been refused yet
    // Hits grow sublinearly with the batch (queries share patterns), so scaling the last finished batch linearly up to 90 % of
    // the limit over-estimates: that size is safe.  Never less than half of the smallest refused batch -- the classic split.
    double fit = (double)(b.adv_refused_q / 2);
    if (b.adv_ok_q > 0 && b.adv_ok_hits > 0.0) fit = std::max(fit, (double)b.adv_ok_q * 0.9 * (double)hit_limit() / b.adv_ok_hits);
    fit = std::min(fit, (double)(b.adv_refused_q - 1));
    return fit < 1.0 ? 1 : (fit >= (double)wanted ? wanted : (int32_t)fit);
}

extern "C" int cgx_result_at(cgx_ctx_t *c, int age, cgx_result_t *o) {
    CGX_TRY(c, {
        CGX_REQUIRE(c && o && age >= 0 && age < CGX_RESULT_SETS, "bad argument (age must be 0..%d)", CGX_RESULT_SETS - 1);
        Batch &b = c->batch;
        CUDA_CHECK(cudaSetDevice(c->device));
        if (age == 0) {
            CGX_REQUIRE(b.valid, "no finished batch");
            CGX_REQUIRE(b.fetch_results, "the last batch kept its results on the device (cgx_extract_dev)");
            CUDA_CHECK(cudaEventSynchronize(b.done_ev));
            o->Q = b.Q; o->T = b.T; o->G = b.G; o->D1 = b.D1; o->D2 = b.D2;
            o->phrase_id = b.h_phrase_id.ptr<int32_t>(); o->phrases = b.h_phrases.ptr<int32_t>(); o->pat1 = b.h_pat1.ptr<int32_t>(); o->pat2 = b.h_pat2.ptr<int32_t>();
            o->q1_off = b.h_q1_off.ptr<int32_t>(); o->q1_ids = b.h_q1_ids.ptr<int32_t>(); o->q2_off = b.h_q2_off.ptr<int32_t>(); o->q2_ids = b.h_q2_ids.ptr<int32_t>();
            for (int k = 0; k < 3; k++) {
                o->rules[k] = b.h_rules[k].ptr<cgx_rule_t>(); o->n_rules[k] = b.n_rules[k];
                o->first[k] = b.h_updown[k].ptr<int32_t>(); o->n_ids[k] = b.n_ids[k];
                o->idinfo[k] = b.h_idinfo[k].ptr<uint32_t>();
            }
        } else {
            ResultSet &r = b.parked[age - 1];
            CGX_REQUIRE(r.valid && r.fetched, "no batch of that age (cgx_extract_begin keeps the last %d)", CGX_RESULT_SETS);
            CUDA_CHECK(cudaEventSynchronize(r.done));
            o->Q = r.Q; o->T = r.T; o->G = r.G; o->D1 = r.D1; o->D2 = r.D2;
            o->phrase_id = r.h_phrase_id.ptr<int32_t>(); o->phrases = r.h_phrases.ptr<int32_t>(); o->pat1 = r.h_pat1.ptr<int32_t>(); o->pat2 = r.h_pat2.ptr<int32_t>();
            o->q1_off = r.h_q1_off.ptr<int32_t>(); o->q1_ids = r.h_q1_ids.ptr<int32_t>(); o->q2_off = r.h_q2_off.ptr<int32_t>(); o->q2_ids = r.h_q2_ids.ptr<int32_t>();
            for (int k = 0; k < 3; k++) {
                o->rules[k] = r.h_rules[k].ptr<cgx_rule_t>(); o->n_rules[k] = r.n_rules[k];
                o->first[k] = r.h_updown[k].ptr<int32_t>(); o->n_ids[k] = r.n_ids[k];
                o->idinfo[k] = r.h_idinfo[k].ptr<uint32_t>();
            }
        }
    });
}

extern "C" int cgx_result(cgx_ctx_t *c, cgx_result_t *o) { return cgx_result_at(c, 0, o); }

extern "C" int cgx_debug_sort_u64(cgx_ctx_t *c, uint64_t *keys_dev, uint32_t *vals_dev, int64_t n, int begin_bit, int end_bit, float *ms_out, int *passes_out) {
    CGX_TRY(c, {
        CGX_REQUIRE(c && keys_dev && n >= 0 && begin_bit >= 0 && end_bit <= 64 && begin_bit < end_bit, "bad argument");
        CUDA_CHECK(cudaSetDevice(c->device));
        Batch &b = c->batch;
        uint64_t *tmp = b.hit_keys_tmp.get<uint64_t>((size_t)n + 1);
        uint32_t *vtmp = vals_dev ? b.scratch.get<uint32_t>((size_t)n + 1) : nullptr;
        cudaEvent_t e0, e1;
        CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
        uint64_t *ks;
        uint32_t *vs;
        int launches = 0;
        CUDA_CHECK(cudaEventRecord(e0, c->stream));
        radix_sort<uint64_t>(keys_dev, tmp, vals_dev, vtmp, (size_t)n, begin_bit, end_bit, c->stream, b.radix, &ks, &vs, &launches);
        CUDA_CHECK(cudaEventRecord(e1, c->stream));
        if (ks != keys_dev) CUDA_CHECK(cudaMemcpyAsync(keys_dev, ks, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
        if (vals_dev && vs != vals_dev) CUDA_CHECK(cudaMemcpyAsync(vals_dev, vs, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (ms_out) *ms_out = ms;
        if (passes_out) *passes_out = launches - 2;
    });
}

extern "C" int64_t cgx_debug_fetch(cgx_ctx_t *c, const char *what, int32_t *out, int64_t cap) {
    if (!c || !what || !out) return -1;
    try {
        CUDA_CHECK(cudaSetDevice(c->device));
        Batch &b = c->batch;
        std::string w(what);
        auto copy_i32 = [&](const void *dev, size_t count) -> int64_t {
            if ((int64_t)count > cap) return -2;
            if (count) CUDA_CHECK(cudaMemcpy(out, dev, sizeof(int32_t) * count, cudaMemcpyDeviceToHost));
            return (int64_t)count;
        };
        if (w == "longest") return copy_i32(b.longest.p, (size_t)b.T);
        if (w == "intervals") return copy_i32(b.iv.p, (size_t)b.T * CGX_LONGEST_SRC * 2);
        if (w == "pat1_dev") return copy_i32(b.pat1_dev.p, (size_t)b.D1 * 4);
        if (w == "pat1_full") return copy_i32(b.pat1.p, (size_t)b.D1 * 8);      // {a_pos, ls, b_pos, le, hit_start, hit_count, marker_pair, fs_extra}
        if (w == "pat2_full") return copy_i32(b.pat2.p, (size_t)b.D2 * 4);      // {pat1, ctok, hit_start, hit_count}
        if (w == "hits1" || w == "hits2") {
            bool two = w == "hits2";
            size_t H = (size_t)(two ? b.hits2 : b.hits1);
            int cols = two ? 4 : 3;
            if ((int64_t)(H * cols) > cap) return -2;
            std::vector<uint64_t> h(H);
            if (H) CUDA_CHECK(cudaMemcpy(h.data(), two ? b.hits2_sorted.p : b.hits1_sorted.p, sizeof(uint64_t) * H, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < H; i++) {
                const uint64_t pm = (1ull << b.pbits) - 1;
                if (two) { out[4 * i] = (int32_t)(h[i] >> (b.pbits + 8)); out[4 * i + 1] = (int32_t)((h[i] >> 8) & pm); out[4 * i + 2] = (int32_t)(h[i] & 15); out[4 * i + 3] = (int32_t)((h[i] & 15) + 1 + ((h[i] >> 4) & 15)); }
                else { out[3 * i] = (int32_t)((h[i] >> (b.pbits + 4)) & ((1ull << cgx_bits_for((uint64_t)b.D1)) - 1ull)); out[3 * i + 1] = (int32_t)((h[i] >> 4) & pm); out[3 * i + 2] = (int32_t)(h[i] & 15); }
            }
            return (int64_t)(H * cols);
        }
        if (w == "rec_ab" || w == "rec_1" || w == "rec_2") {
            int k = w == "rec_ab" ? 0 : w == "rec_1" ? 1 : 2;
            size_t cells = b.rec_cells[k], N = 0;                // slot-indexed cells: keep the non-empty ones
            std::vector<RuleRec> r(cells);
            if (cells) CUDA_CHECK(cudaMemcpy(r.data(), b.rec[k].p, sizeof(RuleRec) * cells, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < cells; i++) if (r[i].id >= 0) r[N++] = r[i];
            if ((int64_t)(N * 7) > cap) return -2;
            for (size_t i = 0; i < N; i++) {
                out[7 * i] = r[i].id; out[7 * i + 1] = r[i].tgt_start; out[7 * i + 2] = r[i].end;
                out[7 * i + 3] = r[i].gap1 == 255 ? -1 : r[i].gap1; out[7 * i + 4] = r[i].gap1 == 255 ? -1 : r[i].gap1_1;
                out[7 * i + 5] = r[i].gap2 == 255 ? -1 : r[i].gap2; out[7 * i + 6] = r[i].gap2 == 255 ? -1 : r[i].gap2_1;
            }
            return (int64_t)(N * 7);
        }
        c->err = "unknown debug array";
        return -3;
    } catch (const CgxError &e) {
        c->err = e.msg;
        return -4;
    }
}
