// cgx-b200: on-device rule aggregation and lexical scoring.
//
// Replaces the single-threaded host loops createLexiconFast / createLexiconGappyFast /
// createLexiconTwoGapFast (ExtractPair.c:515-1276: sprintf/strcat a "src ||| tgt" string per extracted
// record and de-duplicate it in uthash / std::map), extractGlobalPairsUpDown (ExtractPair.cu:2082-2106)
// and lexicalTaskMaxEF + searchLexFile (:2108-2432).
//
// Rule identity = (converted source id, target symbol sequence with each gap collapsed to one marker)
// (ExtractPair.c:813-848, :1141-1173).  Every record gets a 64-bit hash of its target symbol sequence;
// records are ordered by (id, hash) with two stable onesweep radix sorts of (key, record index) pairs;
// run heads are flagged, prefix-summed and compacted into rules: paircount = run length, f = records per
// id, all_suffix_fsample from the pattern tables.  Exactness does not rest on the hash: every record is
// compared symbol-by-symbol with its predecessor in the run, and a mismatch raises a flag on which the
// host re-runs the aggregation with another hash seed.
// Lexical weights: one thread per distinct rule; every (f,e) pair is ONE probe of the lexical hash table (16-byte
// slots holding both directions' values; 2^21 slots = 32 MB at C2, L2-resident) instead of a 20-step binary search;
// -log10 through the same lg2.approx path the reference's -use_fast_math build takes.
#include "batch.h"
#include "hash.cuh"
#include "prof.h"

namespace cgx {

struct AggIdx {
    const int32_t *str, *tgt;
    const int32_t *phrases;
    const Pat1 *pat1;
    const Pat2 *pat2;
    int G, D1, D2;
};

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 33; z *= 0xff51afd7ed558ccdULL; z ^= z >> 33; z *= 0xc4ceb9fe1a85ec53ULL; z ^= z >> 33;
    return z;
}

// target symbol sequence: tokens outside the gaps, gap1 -> 0xFFFFFFFF, gap2 -> 0xFFFFFFFE
__device__ __forceinline__ int target_symbols(const int32_t *__restrict__ tgt, const RuleRec &r, uint32_t sym[16]) {
    int n = 0;
    for (int j = 0; j <= (int)r.end; j++) {
        if (r.gap1 != 255 && j >= (int)r.gap1 && j <= (int)r.gap1_1) { sym[n++] = 0xFFFFFFFFu; j = r.gap1_1; }
        else if (r.gap2 != 255 && j >= (int)r.gap2 && j <= (int)r.gap2_1) { sym[n++] = 0xFFFFFFFEu; j = r.gap2_1; }
        else sym[n++] = (uint32_t)__ldg(&tgt[r.tgt_start + j]);
    }
    return n;
}

__global__ void agg_hash_kernel(const RuleRec *__restrict__ rec, size_t n, const int32_t *__restrict__ tgt, uint64_t seed, uint64_t *__restrict__ keys,
                                uint32_t *__restrict__ idx) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RuleRec r = rec[i];
    uint32_t sym[16];
    int ns = target_symbols(tgt, r, sym);
    uint64_t h = seed ^ (uint64_t)ns;
    for (int k = 0; k < ns; k++) h = mix64(h ^ (uint64_t)sym[k]) + 0x9e3779b97f4a7c15ULL;
    keys[i] = h;
    idx[i] = (uint32_t)i;
}

__global__ void agg_id_keys_kernel(const RuleRec *__restrict__ rec, const uint32_t *__restrict__ idx, size_t n, uint32_t *__restrict__ keys) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = (uint32_t)rec[idx[i]].id;
}

// flags[i] = 1 when record i (in sorted order) starts a new rule; also verifies equal-hash neighbours
__global__ void agg_flags_kernel(const RuleRec *__restrict__ rec, const uint32_t *__restrict__ idx, const uint64_t *__restrict__ hash, size_t n,
                                 const int32_t *__restrict__ tgt, uint32_t *__restrict__ flags, uint32_t *__restrict__ id_count, int *__restrict__ collision) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RuleRec r = rec[idx[i]];
    atomicAdd(&id_count[r.id], 1u);
    if (i == 0) { flags[i] = 1; return; }
    RuleRec q = rec[idx[i - 1]];
    bool head = (q.id != r.id) || (hash[idx[i]] != hash[idx[i - 1]]);
    if (!head) {
        uint32_t a[16], c[16];
        int na = target_symbols(tgt, r, a), nc = target_symbols(tgt, q, c);
        bool same = na == nc;
        for (int k = 0; same && k < na; k++) same = a[k] == c[k];
        if (!same) atomicExch(collision, 1);
    }
    flags[i] = head ? 1u : 0u;
}

// all_suffix_fsample before the cap (ExtractPair.c:637, :891-908, :1211-1247)
__device__ __forceinline__ int fsample_of(const AggIdx &a, int kind, int id) {
    int blk = -1, p1 = -1, p2 = -1;
    if (kind == 0) blk = id;
    else if (kind == 1) { if (id < a.G) blk = id; else if (id < 2 * a.G) blk = id - a.G; else p1 = id - 2 * a.G; }
    else { if (id < a.G) blk = id; else if (id < a.G + a.D2) p2 = id - a.G; else if (id < a.G + a.D2 + a.D1) p1 = id - a.G - a.D2; else p1 = id - a.G - a.D2 - a.D1; }
    if (blk >= 0) return a.phrases[blk * 4 + 1] - a.phrases[blk * 4] + 1;
    if (p2 >= 0) return a.pat2[p2].hit_count;
    Pat1 p = a.pat1[p1];
    return p.hit_count + (p.marker_pair >= 0 ? p.fs_extra : 0);
}

// source terminals of a converted id (the F set of lexicalTaskMaxEF)
__device__ __forceinline__ int source_terminals(const AggIdx &a, int kind, int id, int32_t out[8]) {
    int blk = -1, p1 = -1, p2 = -1, n = 0;
    if (kind == 0) blk = id;
    else if (kind == 1) { if (id < a.G) blk = id; else if (id < 2 * a.G) blk = id - a.G; else p1 = id - 2 * a.G; }
    else { if (id < a.G) blk = id; else if (id < a.G + a.D2) p2 = id - a.G; else if (id < a.G + a.D2 + a.D1) p1 = id - a.G - a.D2; else p1 = id - a.G - a.D2 - a.D1; }
    if (blk >= 0) { int s = a.phrases[blk * 4 + 3], l = a.phrases[blk * 4 + 2]; for (int i = 0; i < l; i++) out[n++] = __ldg(&a.str[s + i]); }
    if (p2 >= 0) p1 = a.pat2[p2].pat1;
    if (p1 >= 0) {
        Pat1 p = a.pat1[p1];
        for (int i = 0; i < p.ls; i++) out[n++] = __ldg(&a.str[p.a_pos + i]);
        for (int i = 0; i < p.le; i++) out[n++] = __ldg(&a.str[p.b_pos + i]);
    }
    if (p2 >= 0) out[n++] = a.pat2[p2].ctok;
    return n;
}

// ExtractPair.cu:2108-2142 searchLexFile: both values of (f,e) in one probe of the lexical hash table; absent -> 0
__device__ __forceinline__ void lex_get(const ulonglong2 *__restrict__ slots, uint32_t mask, int f, int e, float *v1, float *v2) {
    const uint64_t k = ((uint64_t)(uint32_t)(f + 1) << 32) | (uint64_t)(uint32_t)(e + 1);
    uint64_t pay;
    if (ht_find(slots, mask, k, &pay)) { *v1 = __uint_as_float((uint32_t)pay); *v2 = __uint_as_float((uint32_t)(pay >> 32)); }
    else { *v1 = 0.f; *v2 = 0.f; }
}

// head_pos[r] = sorted index of the first record of rule r (excl = exclusive scan of the head flags); head_pos[R] = n
__global__ void agg_head_pos_kernel(const uint32_t *__restrict__ excl, size_t n, uint32_t n_rules, uint32_t *__restrict__ head_pos) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t e = excl[i];
    const uint32_t nxt = (i + 1 < n) ? excl[i + 1] : n_rules;
    if (nxt != e) head_pos[e] = (uint32_t)i;
    if (i == 0) head_pos[n_rules] = (uint32_t)n;
}

// One thread per distinct rule: paircount, f, fs, representative record, lexical weights.
__global__ void __launch_bounds__(128) agg_rules_kernel(AggIdx a, int kind, const RuleRec *__restrict__ rec, const uint32_t *__restrict__ idx,
                                                        const uint32_t *__restrict__ head_pos, uint32_t n_rules,
                                                        const uint32_t *__restrict__ id_count, const ulonglong2 *__restrict__ lex, uint32_t lex_mask,
                                                        cgx_rule_t *__restrict__ rules) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rules) return;
    const uint32_t head = head_pos[r], next_head = head_pos[r + 1];
    const int pc = (int)(next_head - head);
    // deterministic representative: smallest tgt_start of the run
    RuleRec best = rec[idx[head]];
    for (uint32_t i = head + 1; i < next_head; i++) {
        RuleRec c = rec[idx[i]];
        if (c.tgt_start < best.tgt_start) best = c;
    }
    cgx_rule_t out;
    out.id = best.id; out.tgt_start = best.tgt_start; out.end = best.end;
    out.gap1 = best.gap1; out.gap1_1 = best.gap1_1; out.gap2 = best.gap2; out.gap2_1 = best.gap2_1;
    out.pad[0] = out.pad[1] = out.pad[2] = 0;
    out.pc = pc;
    out.f = (int)id_count[best.id];
    int fs = fsample_of(a, kind, best.id);
    out.fs = fs > CGX_SAMPLER ? CGX_SAMPLER : fs;                          // ExtractPair.c:638,910,1249
    // ---- lexicalTaskMaxEF (ExtractPair.cu:2144-2432): for every source terminal the best MaxLexFgivenE over the target
    // terminals (and NULL), for every target terminal the best MaxLexEgivenF over the source terminals (and NULL).  One
    // table probe serves both directions of a (f, e) pair.
    int32_t F[8];
    const int nf = source_terminals(a, kind, best.id, F);
    float mxf[8];
#pragma unroll
    for (int j = 0; j < 8; j++) mxf[j] = 0.f;
    float egivenf = 0.f, v1, v2;
    bool any_e = false;
    const int ts = best.tgt_start;
    for (int jj = 0; jj <= (int)best.end; jj++) {
        if (best.gap1 != 255 && jj >= (int)best.gap1 && jj <= (int)best.gap1_1) continue;
        if (best.gap2 != 255 && jj >= (int)best.gap2 && jj <= (int)best.gap2_1) continue;
        any_e = true;
        const int e = __ldg(&a.tgt[ts + jj]);
        float mx = 0.f;
        if (nf > 0) { lex_get(lex, lex_mask, -1, e, &v1, &v2); mx = fmaxf(mx, v1); }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j < nf) {
                lex_get(lex, lex_mask, F[j], e, &v1, &v2);
                mx = fmaxf(mx, v1);
                mxf[j] = fmaxf(mxf[j], v2);
            }
        }
        egivenf += mx > 0.f ? -__log10f(mx) : CGX_MAXSCORE;
    }
    float fgivene = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j < nf) {
            float mx = mxf[j];
            if (any_e) { lex_get(lex, lex_mask, F[j], -1, &v1, &v2); mx = fmaxf(mx, v2); }
            fgivene += mx > 0.f ? -__log10f(mx) : CGX_MAXSCORE;
        }
    }
    out.max_lex_f_given_e = fgivene;
    out.max_lex_e_given_f = egivenf;
    rules[r] = out;
}

// per converted id: [first rule, last rule]  (globalOnPairsUpDown*, ExtractPair.cu:3745-3756, :3805-3816, :2082)
__global__ void agg_updown_kernel(const cgx_rule_t *__restrict__ rules, uint32_t n_rules, int32_t *__restrict__ updown) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rules) return;
    int id = rules[r].id;
    if (r == 0 || rules[r - 1].id != id) updown[2 * id] = (int32_t)r;
    if (r == n_rules - 1 || rules[r + 1].id != id) updown[2 * id + 1] = (int32_t)r;
}

static uint32_t read_u32(const uint32_t *d, cudaStream_t stream) {
    uint32_t v = 0;
    CUDA_CHECK(cudaMemcpyAsync(&v, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    return v;
}

void stage_aggregate(const Index &ix, Batch &b, cudaStream_t stream) {
    AggIdx a{ix.str.ptr<int32_t>(), ix.tgt.ptr<int32_t>(), b.phrases.ptr<int32_t>(), b.pat1.ptr<Pat1>(), b.pat2.ptr<Pat2>(), b.G, b.D1, b.D2};
    const int nids[3] = {b.G, 2 * b.G + b.D1, b.G + b.D2 + 2 * b.D1};
    uint32_t *tot = b.counters.get<uint32_t>(16);
    int *collision = (int *)(tot + 14);
    for (int kind = 0; kind < 3; kind++) {
        const size_t N = (size_t)b.n_rec[kind];
        b.n_ids[kind] = nids[kind];
        b.n_rules[kind] = 0;
        int32_t *h_ud = b.h_updown[kind].get<int32_t>((size_t)2 * nids[kind] + 2);
        memset(h_ud, 0xff, sizeof(int32_t) * 2 * (size_t)nids[kind]);
        b.h_rules[kind].get<cgx_rule_t>(1);
        if (N == 0 || nids[kind] == 0) continue;
        const RuleRec *rec = b.rec[kind].ptr<RuleRec>();
        uint64_t *hk = b.rec_keys.get<uint64_t>(N), *hk_tmp = b.rec_keys_tmp.get<uint64_t>(N);
        uint32_t *idx = b.rec_idx.get<uint32_t>(N), *idx_tmp = b.rec_idx_tmp.get<uint32_t>(N);
        uint64_t *hash_by_rec = b.rec_hash.get<uint64_t>(N);
        uint32_t *idk = b.scratch.get<uint32_t>(2 * N + 2), *idk_tmp = idk + N;
        uint32_t *flags = b.rec_flags.get<uint32_t>(N + 2);
        uint32_t *id_count = b.id_count[kind].get<uint32_t>((size_t)nids[kind]);
        int32_t *updown = b.updown[kind].get<int32_t>((size_t)2 * nids[kind]);
        uint64_t seed = 0x243f6a8885a308d3ULL;
        uint32_t R = 0;
        for (int attempt = 0; attempt < 8; attempt++, seed = seed * 6364136223846793005ULL + 1442695040888963407ULL) {
            PROF("agg_hash", (double)N * 28, (agg_hash_kernel<<<cgx_div_up(N, 256), 256, 0, stream>>>(rec, N, ix.tgt.ptr<int32_t>(), seed, hk, idx)));
            CUDA_CHECK(cudaMemcpyAsync(hash_by_rec, hk, sizeof(uint64_t) * N, cudaMemcpyDeviceToDevice, stream));
            uint64_t *ks;
            uint32_t *is;
            radix_sort<uint64_t>(hk, hk_tmp, idx, idx_tmp, N, 0, 64, stream, b.radix, &ks, &is, &b.launches);
            agg_id_keys_kernel<<<cgx_div_up(N, 256), 256, 0, stream>>>(rec, is, N, idk);
            uint32_t *ids_sorted, *is2;
            uint32_t *other = (is == idx) ? idx_tmp : idx;
            radix_sort<uint32_t>(idk, idk_tmp, is, other, N, 0, cgx_bits_for((uint64_t)nids[kind]), stream, b.radix, &ids_sorted, &is2, &b.launches);
            CUDA_CHECK(cudaMemsetAsync(id_count, 0, sizeof(uint32_t) * (size_t)nids[kind], stream));
            CUDA_CHECK(cudaMemsetAsync(collision, 0, sizeof(int), stream));
            PROF("agg_flags", (double)N * 48, (agg_flags_kernel<<<cgx_div_up(N, 256), 256, 0, stream>>>(rec, is2, hash_by_rec, N, ix.tgt.ptr<int32_t>(), flags, id_count, collision)));
            exclusive_scan_u32(flags, flags, N, tot, stream, b.scan, 0, &b.launches);
            b.launches += 3;
            uint32_t hostv[16];
            CUDA_CHECK(cudaMemcpyAsync(hostv, tot, sizeof(hostv), cudaMemcpyDeviceToHost, stream));
            CUDA_CHECK(cudaStreamSynchronize(stream));
            if (hostv[14] == 0) {
                R = hostv[0];
                cgx_rule_t *rules = b.rules[kind].get<cgx_rule_t>(R);
                uint32_t *head_pos = b.rule_head.get<uint32_t>((size_t)R + 2);
                agg_head_pos_kernel<<<cgx_div_up(N, 256), 256, 0, stream>>>(flags, N, R, head_pos);
                PROF("agg_rules", (double)R * 36 + (double)N * 20, (agg_rules_kernel<<<cgx_div_up(R, 128), 128, 0, stream>>>(a, kind, rec, is2, head_pos, R, id_count,
                                                                       ix.lex_hash.ptr<ulonglong2>(), ix.lex_hash_mask, rules)));
                b.launches++;
                CUDA_CHECK(cudaMemsetAsync(updown, 0xff, sizeof(int32_t) * 2 * (size_t)nids[kind], stream));
                agg_updown_kernel<<<cgx_div_up(R, 256), 256, 0, stream>>>(rules, R, updown);
                b.launches += 2;
                break;
            }
            CGX_REQUIRE(attempt < 7, "aggregation: hash collisions persisted over 8 seeds");
        }
        b.n_rules[kind] = (int32_t)R;
        if (b.fetch_results) {
            cgx_rule_t *h_r = b.h_rules[kind].get<cgx_rule_t>((size_t)R + 1);
            if (R) CUDA_CHECK(cudaMemcpyAsync(h_r, b.rules[kind].ptr<cgx_rule_t>(), sizeof(cgx_rule_t) * R, cudaMemcpyDeviceToHost, stream));
            CUDA_CHECK(cudaMemcpyAsync(h_ud, updown, sizeof(int32_t) * 2 * (size_t)nids[kind], cudaMemcpyDeviceToHost, stream));
        }
    }
    (void)read_u32;
}

}  // namespace cgx
