// cgx-b200: auxiliary index arrays built once per corpus, on the GPU, after the suffix array.
#include "index.h"
#include "align_fields.cuh"
#include "hash.cuh"
#include <algorithm>
#include <vector>

namespace cgx {

__global__ void ix_token_hist_kernel(const int32_t *__restrict__ str, size_t n, uint32_t *__restrict__ counts) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&counts[str[i]], 1u);
}

// key = first m tokens of the suffix at p (token ids packed, tokbits each); value = p
__global__ void ix_ngram_keys_kernel(const int32_t *__restrict__ str, size_t n, int mlen, int tokbits, uint64_t *__restrict__ keys,
                                     uint32_t *__restrict__ vals) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint64_t k = 0;
    for (int j = 0; j < mlen; j++) k = (k << tokbits) | (uint64_t)(uint32_t)str[p + j];
    keys[p] = k;
    vals[p] = (uint32_t)p;
}
// the same order from the bucket id of the (m-1)-gram at p and the m-th token: used when m tokens do not fit 64 bits
// (vocabularies of 2^21 types and more: 3 * tokbits > 64)
__global__ void ix_ngram_keys_from_bucket_kernel(const int32_t *__restrict__ str, const int32_t *__restrict__ bkt_prev, size_t n, int mlen, int tokbits,
                                                 uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    keys[p] = ((uint64_t)(uint32_t)bkt_prev[p] << tokbits) | (uint64_t)(uint32_t)str[p + mlen - 1];
    vals[p] = (uint32_t)p;
}

// bucket start of every m-gram: heads of the sorted key array mark the bucket starts; gid = dense bucket number
__global__ void ix_bucket_heads_kernel(const uint64_t *__restrict__ keys, size_t n, uint32_t *__restrict__ flags) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) flags[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1u : 0u;
}
__global__ void ix_bucket_starts_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ excl, size_t n, uint32_t *__restrict__ start_of) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && (k == 0 || keys[k] != keys[k - 1])) start_of[excl[k]] = (uint32_t)k;
}
__global__ void ix_bucket_scatter_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pos, const uint32_t *__restrict__ excl,
                                         const uint32_t *__restrict__ start_of, size_t n, int32_t *__restrict__ bkt) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t head = (k == 0 || keys[k] != keys[k - 1]) ? 1u : 0u;
    bkt[pos[k]] = (int32_t)start_of[excl[k] + head - 1];
}

// Gap-consistency words.  For a span [i, i+g-1] of source tokens (a gap of a hierarchical phrase):
// GappyLook.cu:43-126 checkBoundaryGap = first and last token aligned; target span [min L, max R] of the
// aligned tokens narrower than 15; and, over that target span, min L_tar / max R_tar map back exactly onto the
// source span.  The source-side min/max is maintained incrementally while g grows.
template <class A>
__global__ void ix_gap_words_kernel(const int32_t *__restrict__ str, const typename A::word_t *__restrict__ RLP, const typename A::lr_t *__restrict__ L_tar,
                                    const typename A::lr_t *__restrict__ R_tar, size_t n, size_t m, uint32_t *__restrict__ gapw) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t word = 0;
    if (str[i] >= 2) {
        int run = 1;
        while (run < 15 && str[i + run] >= 2) run++;
        word = (uint32_t)run << 16;
        const typename A::word_t w0 = RLP[i];
        const unsigned L0 = A::L(w0), R0 = A::R(w0);
        // the unique final symbol at n-1 (Start.cu:324-326) has no alignment record and is never matched by a query
        if (L0 != A::UNAL && R0 != A::UNAL && i + 1 < n) {
            const int eos_prev = (int)i - (int)A::P(w0) - 1;
            const int tgt_base = eos_prev < 0 ? 0 : (int)(uint32_t)RLP[eos_prev];
            const int src_base = eos_prev + 1;
            unsigned mn = L0, mx = R0;
            const int gmax = run < 13 ? run : 13;
            for (int g = 1; g <= gmax; g++) {
                if (g > 1) {
                    const typename A::word_t w = RLP[i + g - 1];
                    const unsigned L = A::L(w), R = A::R(w);
                    if (L == A::UNAL || R == A::UNAL) continue;          // last token unaligned: this g fails, larger g may pass
                    mn = min(mn, L); mx = max(mx, R);
                }
                if (mx - mn >= CGX_MAX_RULE_SPAN) break;           // the span only grows with g
                if (tgt_base < 0 || (size_t)tgt_base + mx >= m) break;   // malformed alignment record: never consistent
                unsigned tmn = A::UNAL, tmx = 0;
                for (int k = tgt_base + (int)mn; k <= tgt_base + (int)mx; k++) {
                    const unsigned L = L_tar[k], R = R_tar[k];
                    if (L == A::UNAL || R == A::UNAL) continue;
                    tmn = min(tmn, L); tmx = max(tmx, R);
                }
                if (src_base + (int)tmn == (int)i && src_base + (int)tmx == (int)i + g - 1) word |= 1u << (g - 1);
            }
        }
    }
    gapw[i] = word;
}

__global__ void ix_jwin_kernel(const int32_t *__restrict__ b1, const int32_t *__restrict__ b2, const int32_t *__restrict__ b3,
                               const uint32_t *__restrict__ gapw, size_t n, int4 *__restrict__ jwin) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) jwin[p] = make_int4(b1[p], b2[p], b3[p], (int)gapw[p]);
}

// extraction views: xw[k] = RLP[k] | 1 where the source token is a word (>= 2), else 0 (EOS, padding; n+3 entries like str);
// lrq[j] = range-minimum table of the target side for consistent() (ExtractPair.cu:103-133), which needs min L_tar / max R_tar
// over a target window of at most 15 tokens, unaligned tokens skipped: level k holds {min L, max R} over tokens j .. j+2^k-1,
// k = 0..3 (packing: align_fields.cuh), with an unaligned token entered as {UNAL, 0} (the identities of min and max).  Any window is
// the union of two level-floor(log2 len) entries: two independent 8-byte loads instead of a loop of <= 15 dependent 2-byte ones.
template <class A>
__global__ void ix_extract_views_kernel(const int32_t *__restrict__ str, const typename A::word_t *__restrict__ RLP, size_t n,
                                        const typename A::lr_t *__restrict__ L_tar, const typename A::lr_t *__restrict__ R_tar, size_t m,
                                        typename A::word_t *__restrict__ xw, typename A::lrq_t *__restrict__ lrq) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n + 3) xw[i] = (i < n && str[i] >= 2) ? (RLP[i] | 1u) : 0u;
    if (i < m) {
        unsigned mn = A::UNAL, mx = 0, lmn[4], lmx[4];
        int t = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            for (; t < (1 << k); t++) {
                if (i + t >= m) continue;
                const unsigned L = L_tar[i + t], R = R_tar[i + t];
                if (L == A::UNAL || R == A::UNAL) continue;
                mn = min(mn, L); mx = max(mx, R);
            }
            lmn[k] = mn; lmx[k] = mx;
        }
        lrq[i] = A::lrq_make(lmn, lmx);
    }
}

template <class A>
static void build_views_t(Index &ix, cudaStream_t stream) {
    using W = typename A::word_t;
    using LR = typename A::lr_t;
    const size_t cnt = ix.n + 3 > ix.m ? ix.n + 3 : ix.m;
    ix_extract_views_kernel<A><<<cgx_div_up(cnt, 256), 256, 0, stream>>>(ix.str.ptr<int32_t>(), ix.RLP.ptr<W>(), ix.n, ix.L_tar.ptr<LR>(), ix.R_tar.ptr<LR>(), ix.m,
                                                                          ix.xw.get<W>(ix.n + 3), ix.lr.get<typename A::lrq_t>(ix.m));
}

void build_jwin(Index &ix, cudaStream_t stream) {
    ix_jwin_kernel<<<cgx_div_up(ix.n, 256), 256, 0, stream>>>(ix.bkt[0].ptr<int32_t>(), ix.bkt[1].ptr<int32_t>(), ix.bkt[2].ptr<int32_t>(),
                                                            ix.gapw.ptr<uint32_t>(), ix.n, ix.jwin.get<int4>(ix.n));
    if (ix.wide) build_views_t<AlignWide>(ix, stream); else build_views_t<AlignNarrow>(ix, stream);
    CUDA_CHECK(cudaStreamSynchronize(stream));
}

// lexical table -> hash.  Duplicate (f,e) rows: the first of the (stably) sorted run wins, like the binary search it replaces.
__global__ void ix_lex_hash_kernel(const uint64_t *__restrict__ keys, const float *__restrict__ v1, const float *__restrict__ v2, size_t n,
                                   ulonglong2 *__restrict__ slots, uint32_t mask) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (i > 0 && keys[i] == keys[i - 1])) return;
    ht_insert(slots, mask, keys[i], (uint64_t)__float_as_uint(v1[i]) | ((uint64_t)__float_as_uint(v2[i]) << 32));
}

void build_lex_hash(Index &ix, cudaStream_t stream) {
    const uint32_t slots_n = ht_slots_for(ix.lex_count);
    ulonglong2 *slots = ix.lex_hash.get<ulonglong2>(slots_n);
    ix.lex_hash_mask = slots_n - 1;
    CUDA_CHECK(cudaMemsetAsync(slots, 0xff, sizeof(ulonglong2) * (size_t)slots_n, stream));
    if (ix.lex_count)
        ix_lex_hash_kernel<<<cgx_div_up(ix.lex_count, 256), 256, 0, stream>>>(ix.lex_key.ptr<uint64_t>(), ix.lex_v1.ptr<float>(), ix.lex_v2.ptr<float>(), ix.lex_count, slots,
                                                                            ix.lex_hash_mask);
    CUDA_CHECK(cudaStreamSynchronize(stream));
}

// SuffixArray.cu:1148-1198: the PRECOMPUTECOUNT most frequent tokens (frequency descending, ties by
// ascending id -- the intended order of compareUserTotal1 under a stable sort), stored ascending by id.
static void pick_frequent(const std::vector<uint32_t> &counts, int32_t maxtok, int32_t *freq_list, std::vector<uint8_t> &flag) {
    std::vector<int32_t> toks;
    for (int32_t t = 2; t <= maxtok; t++) if (counts[t]) toks.push_back(t);
    CGX_REQUIRE((int)toks.size() >= CGX_PRECOMP, "corpus has %d distinct source tokens; the frequent-pair table needs >= %d (SuffixArray.cu:1175-1176)",
                (int)toks.size(), CGX_PRECOMP);
    std::stable_sort(toks.begin(), toks.end(), [&](int32_t a, int32_t b) { return counts[a] > counts[b]; });
    toks.resize(CGX_PRECOMP);
    std::sort(toks.begin(), toks.end());
    flag.assign((size_t)maxtok + 2, 0);
    for (int i = 0; i < CGX_PRECOMP; i++) { freq_list[i] = toks[i]; flag[toks[i]] = (uint8_t)(i + 1); }   // rank+1 in the id-ascending list
}

void build_index_aux(Index &ix, SaWorkspace &ws, cudaStream_t stream, int *launches) {
    const size_t n = ix.n;
    const int32_t *str = ix.str.ptr<int32_t>();
    const size_t nt = (size_t)ix.maxtok + 2;
    uint32_t *ts = (uint32_t *)ix.tok_start.get<int32_t>(nt);
    CUDA_CHECK(cudaMemsetAsync(ts, 0, sizeof(uint32_t) * nt, stream));
    ix_token_hist_kernel<<<CGX_NUM_SMS * 8, 256, 0, stream>>>(str, n, ts);
    if (launches) *launches += 1;
    std::vector<uint32_t> counts(nt);
    CUDA_CHECK(cudaMemcpyAsync(counts.data(), ts, sizeof(uint32_t) * nt, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    std::vector<uint8_t> flag;
    pick_frequent(counts, ix.maxtok, ix.freq_list, flag);
    CUDA_CHECK(cudaMemcpyAsync(ix.freq_flag.get<uint8_t>(nt), flag.data(), nt, cudaMemcpyHostToDevice, stream));
    exclusive_scan_u32(ts, ts, nt, nullptr, stream, ws.scan, 0, launches);
    if (ix.wide) ix_gap_words_kernel<AlignWide><<<cgx_div_up(n, 128), 128, 0, stream>>>(str, ix.RLP.ptr<uint64_t>(), ix.L_tar.ptr<uint16_t>(), ix.R_tar.ptr<uint16_t>(), n, ix.m,
                                                              ix.gapw.get<uint32_t>(n));
    else ix_gap_words_kernel<AlignNarrow><<<cgx_div_up(n, 128), 128, 0, stream>>>(str, ix.RLP.ptr<uint32_t>(), ix.L_tar.ptr<uint8_t>(), ix.R_tar.ptr<uint8_t>(), n, ix.m,
                                                              ix.gapw.get<uint32_t>(n));
    if (launches) *launches += 1;
    // position-sorted occurrence lists of every 1-, 2- and 3-gram
    const int tokbits = cgx_bits_for((uint64_t)ix.maxtok);
    uint64_t *keys = ws.keys.get<uint64_t>(n), *keys_tmp = ws.keys_tmp.get<uint64_t>(n);
    uint32_t *vals = ws.vals.get<uint32_t>(n), *vals_tmp = ws.vals_tmp.get<uint32_t>(n);
    for (int mlen = 1; mlen <= 3; mlen++) {
        // m packed token ids while they fit 64 bits, else (bucket of the (m-1)-gram, m-th token): same order, always <= 61 bits
        int key_bits = mlen * tokbits;
        const bool force_bucket = getenv("CGX_FORCE_BUCKET_KEYS") != nullptr;      // tests: exercise the wide-vocabulary form
        if (key_bits <= 64 && !(force_bucket && mlen > 1)) ix_ngram_keys_kernel<<<cgx_div_up(n, 256), 256, 0, stream>>>(str, n, mlen, tokbits, keys, vals);
        else {
            key_bits = cgx_bits_for((uint64_t)n) + tokbits;
            CGX_REQUIRE(key_bits <= 64, "index: %d-gram keys need %d bits", mlen, key_bits);
            ix_ngram_keys_from_bucket_kernel<<<cgx_div_up(n, 256), 256, 0, stream>>>(str, ix.bkt[mlen - 2].ptr<int32_t>(), n, mlen, tokbits, keys, vals);
        }
        if (launches) *launches += 1;
        uint64_t *ks;
        uint32_t *vs;
        radix_sort<uint64_t>(keys, keys_tmp, vals, vals_tmp, n, 0, key_bits, stream, ws.radix, &ks, &vs, launches);
        CUDA_CHECK(cudaMemcpyAsync(ix.inv[mlen - 1].get<int32_t>(n), vs, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, stream));
        uint32_t *flags = ws.flags.get<uint32_t>(n), *start_of = ws.rank.get<uint32_t>(n);
        ix_bucket_heads_kernel<<<cgx_div_up(n, 256), 256, 0, stream>>>(ks, n, flags);
        exclusive_scan_u32(flags, flags, n, nullptr, stream, ws.scan, 0, launches);
        ix_bucket_starts_kernel<<<cgx_div_up(n, 256), 256, 0, stream>>>(ks, flags, n, start_of);
        ix_bucket_scatter_kernel<<<cgx_div_up(n, 256), 256, 0, stream>>>(ks, vs, flags, start_of, n, ix.bkt[mlen - 1].get<int32_t>(n));
        if (launches) *launches += 3;
    }
    build_jwin(ix, stream);
    if (launches) *launches += 1;
}

}  // namespace cgx
