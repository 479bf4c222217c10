// cgx-b200: device-side enumeration and de-duplication of the source patterns present in the queries.
//
// Replaces, with no host round trips:
//   * GenerateBlocks (ExtractPair.cu:2742-2903): distinct contiguous phrases (up,down,len<=5)  [std::map on the host]
//   * oneGapEnumeration (SuffixArray.cu:928-1039) + thrust::sort_by_key (:1598) + zeroOneDiff (:1041)
//     + the sequential host scan (:1670-1719): distinct aXb patterns, per-query id lists
//   * twoGapEnumeration (:816-926) + sort (:1989) + zeroOneDiffTwoGap (:1070) + host scan (:2062-2097)
//
// Patterns are identified by packed integer keys built from SA-interval starts instead of 21-byte
// token tuples: for phrases of equal length the interval start orders them exactly like their token
// sequences, and a shorter `a` that is a prefix of a longer one sorts first -- i.e. the key order is
// the reference comparator's order (number, then tokens with the gap marker -1 below every token,
// SuffixArray.cu:51-67), so distinct-pattern ids agree with the reference's.
// Every enumeration is count -> prefix sum -> fill (deterministic slots, no atomics), sorted with the
// onesweep radix sort, then flag / scan / compact.
#include "batch.h"
#include "prof.h"

namespace cgx {

// ------------------------------------------------------------------------------------------------
// generic helpers
// ------------------------------------------------------------------------------------------------
__global__ void head_flags_u64_kernel(const uint64_t *__restrict__ keys, size_t n, uint32_t *__restrict__ flags) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    flags[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1u : 0u;
}

static uint32_t read_u32(const uint32_t *d, cudaStream_t stream) {
    uint32_t v = 0;
    cgx_read_back(&v, d, sizeof(uint32_t), stream);
    return v;
}

// offsets[q] = first index k with (keys[k] >> 32) >= q, for q = 0..Q
__global__ void query_offsets_kernel(const uint64_t *__restrict__ keys, int n, int Q, int32_t *__restrict__ off) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > Q) return;
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if ((int)(keys[mid] >> 32) < q) lo = mid + 1; else hi = mid;
    }
    off[q] = lo;
}

__global__ void low32_kernel(const uint64_t *__restrict__ keys, int n, int32_t *__restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (int32_t)(uint32_t)keys[k];
}

// ------------------------------------------------------------------------------------------------
// contiguous phrases
// ------------------------------------------------------------------------------------------------
__global__ void ph_keys_kernel(const int32_t *__restrict__ longest, const uint32_t *__restrict__ off, const int32_t *__restrict__ iv, int T,
                               uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int L = longest[t];
    uint32_t o = off[t];
    for (int m = 1; m <= L; m++) {
        uint32_t up = (uint32_t)iv[((size_t)t * CGX_LONGEST_SRC + (m - 1)) * 2];
        keys[o + m - 1] = ((uint64_t)up << 3) | (uint64_t)m;
        vals[o + m - 1] = (uint32_t)(t * CGX_LONGEST_SRC + (m - 1));
    }
}

__global__ void ph_assign_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ excl, int n,
                                 const int32_t *__restrict__ iv, const int32_t *__restrict__ sa, int32_t *__restrict__ phrase_id,
                                 int32_t *__restrict__ phrases) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    bool head = (k == 0 || keys[k] != keys[k - 1]);
    int idx = (int)excl[k] + (head ? 1 : 0) - 1;
    uint32_t v = vals[k];
    phrase_id[v] = idx;
    if (head) {
        int up = iv[(size_t)v * 2], down = iv[(size_t)v * 2 + 1];
        phrases[idx * 4 + 0] = up;
        phrases[idx * 4 + 1] = down;
        phrases[idx * 4 + 2] = (int)(keys[k] & 7);
        phrases[idx * 4 + 3] = sa[up];
    }
}

void stage_phrases(const Index &ix, Batch &b, cudaStream_t stream) {
    const int T = b.T;
    b.G = 0;
    int32_t *phrase_id = b.phrase_id.get<int32_t>((size_t)T * CGX_LONGEST_SRC + 1);
    if (T == 0) return;
    CUDA_CHECK(cudaMemsetAsync(phrase_id, 0xff, sizeof(int32_t) * (size_t)T * CGX_LONGEST_SRC, stream));
    uint32_t *off = b.scratch.get<uint32_t>((size_t)T + 2);
    uint32_t *tot = b.counters.get<uint32_t>(16);
    exclusive_scan_u32((const uint32_t *)b.longest.ptr<int32_t>(), off, (size_t)T, tot, stream, b.scan, 0, &b.launches);
    uint32_t N = read_u32(tot, stream);
    if (N == 0) return;
    uint64_t *keys = b.ph_keys.get<uint64_t>(N), *keys_tmp = b.ph_keys_tmp.get<uint64_t>(N);
    uint32_t *vals = b.ph_vals.get<uint32_t>(N), *vals_tmp = b.ph_vals_tmp.get<uint32_t>(N);
    ph_keys_kernel<<<cgx_div_up(T, 256), 256, 0, stream>>>(b.longest.ptr<int32_t>(), off, b.iv.ptr<int32_t>(), T, keys, vals);
    b.launches++;
    uint64_t *ks;
    uint32_t *vs;
    radix_sort<uint64_t>(keys, keys_tmp, vals, vals_tmp, N, 0, 3 + cgx_bits_for(ix.n), stream, b.radix, &ks, &vs, &b.launches);
    uint32_t *flags = b.ph_flags.get<uint32_t>(N);
    head_flags_u64_kernel<<<cgx_div_up(N, 256), 256, 0, stream>>>(ks, N, flags);
    exclusive_scan_u32(flags, flags, N, tot, stream, b.scan, 0, &b.launches);
    b.G = (int32_t)read_u32(tot, stream);
    int32_t *phrases = b.phrases.get<int32_t>((size_t)b.G * 4);
    ph_assign_kernel<<<cgx_div_up(N, 256), 256, 0, stream>>>(ks, vs, flags, (int)N, b.iv.ptr<int32_t>(), ix.sa.ptr<int32_t>(), phrase_id, phrases);
    b.launches += 2;
}

// ------------------------------------------------------------------------------------------------
// one-gap patterns aXb
// ------------------------------------------------------------------------------------------------
// instance info word: ls | gap << 8 | le << 16
template <bool FILL>
__global__ void e1_enum_kernel(const int32_t *__restrict__ q_tok, const int32_t *__restrict__ q_off, const int32_t *__restrict__ tok2q, int T,
                               const int32_t *__restrict__ longest, const int32_t *__restrict__ iv, uint32_t *__restrict__ count,
                               const uint32_t *__restrict__ off, int32_t *__restrict__ inst_t, uint32_t *__restrict__ inst_info,
                               uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    uint32_t c = 0;
    uint32_t o = FILL ? off[t] : 0;
    const int end = q_off[tok2q[t] + 1];
    // SuffixArray.cu:945,957-962: not the last token of the batch, not the last two tokens of a query
    if (t < T - 1 && t != end - 1 && t != end - 2) {
        const int Ls = min(longest[t], CGX_MAX_RULE_SYMBOLS - 2);
        for (int ls = 1; ls <= Ls; ls++) {
            const int smax = min(end - 1, t + CGX_MAX_RULE_SPAN);
            for (int s = t + ls + 1; s <= smax; s++) {
                if (q_tok[s] == -1) continue;
                int le_max = min(min(longest[s], CGX_MAX_RULE_SYMBOLS - 1 - ls), CGX_MAX_RULE_SPAN + 1 - (s - t));
                if (le_max <= 0) continue;
                if (FILL) {
                    uint32_t up_a = (uint32_t)iv[((size_t)t * CGX_LONGEST_SRC + (ls - 1)) * 2];
                    for (int le = 1; le <= le_max; le++) {
                        uint32_t up_b = (uint32_t)iv[((size_t)s * CGX_LONGEST_SRC + (le - 1)) * 2];
                        uint32_t i = o + c + (le - 1);
                        inst_t[i] = t;
                        inst_info[i] = (uint32_t)ls | ((uint32_t)(s - t - ls) << 8) | ((uint32_t)le << 16);
                        keys[i] = ((uint64_t)(ls + 1 + le - 3) << 62) | ((uint64_t)up_a << 32) | ((uint64_t)(ls - 1) << 30) | (uint64_t)up_b;
                        vals[i] = i;
                    }
                }
                c += (uint32_t)le_max;
            }
        }
    }
    if (!FILL) count[t] = c;
}

__global__ void e1_patterns_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ excl, int n,
                                   const int32_t *__restrict__ inst_t, const uint32_t *__restrict__ inst_info, const int32_t *__restrict__ iv,
                                   const int32_t *__restrict__ sa, const int32_t *__restrict__ str, const uint8_t *__restrict__ freq_rank,
                                   const int32_t *__restrict__ phrase_id, int32_t *__restrict__ pid, Pat1 *__restrict__ pat, Pat1Dev *__restrict__ patd,
                                   int32_t *__restrict__ pat_pos, int32_t *__restrict__ pat_ga) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    bool head = (k == 0 || keys[k] != keys[k - 1]);
    int d = (int)excl[k] + (head ? 1 : 0) - 1;
    pid[k] = d;
    if (!head) return;
    uint32_t inst = vals[k];
    int t = inst_t[inst];
    uint32_t info = inst_info[inst];
    int ls = info & 0xff, gap = (info >> 8) & 0xff, le = (info >> 16) & 0xff;
    int s = t + ls + gap;
    Pat1Dev pd;
    pd.up_a = iv[((size_t)t * CGX_LONGEST_SRC + (ls - 1)) * 2];
    pd.down_a = iv[((size_t)t * CGX_LONGEST_SRC + (ls - 1)) * 2 + 1];
    pd.up_b = iv[((size_t)s * CGX_LONGEST_SRC + (le - 1)) * 2];
    pd.down_b = iv[((size_t)s * CGX_LONGEST_SRC + (le - 1)) * 2 + 1];
    patd[d] = pd;
    Pat1 p;
    p.a_pos = sa[pd.up_a]; p.ls = ls; p.b_pos = sa[pd.up_b]; p.le = le;
    p.hit_start = -1; p.hit_count = 0; p.marker_pair = -1; p.fs_extra = 0;
    if (ls == 1 && le == 1) {
        int ra = freq_rank[str[p.a_pos]], rb = freq_rank[str[p.b_pos]];
        if (ra && rb) p.marker_pair = (ra - 1) * CGX_PRECOMP + (rb - 1);
    }
    pat[d] = p;
    pat_pos[d] = k;
    pat_ga[d] = phrase_id[(size_t)t * CGX_LONGEST_SRC + (ls - 1)];      // distinct-phrase id of a (drives the join scan)
}

// (query, pattern) pairs without duplicates: instances of one pattern are sorted by instance index,
// hence by query, so duplicates are adjacent.
__global__ void qlist_flags_kernel(const int32_t *__restrict__ pid, const int32_t *__restrict__ qid_of_sorted, int n, uint32_t *__restrict__ flags) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    flags[k] = (k == 0 || pid[k] != pid[k - 1] || qid_of_sorted[k] != qid_of_sorted[k - 1]) ? 1u : 0u;
}
__global__ void e1_sorted_qid_kernel(const uint32_t *__restrict__ vals, const int32_t *__restrict__ inst_t, const int32_t *__restrict__ tok2q, int n,
                                     int32_t *__restrict__ qid) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) qid[k] = tok2q[inst_t[vals[k]]];
}
__global__ void qlist_compact_kernel(const int32_t *__restrict__ pid, const int32_t *__restrict__ qid, const uint32_t *__restrict__ excl, int n,
                                     uint64_t *__restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    bool f = (k == 0 || pid[k] != pid[k - 1] || qid[k] != qid[k - 1]);
    if (f) out[excl[k]] = ((uint64_t)(uint32_t)qid[k] << 32) | (uint64_t)(uint32_t)pid[k];
}

// builds per-query sorted id lists from (pid,qid) of the pattern-sorted instances
static void build_query_lists(Batch &b, const int32_t *pid, const int32_t *qid, int n, int idbits, DevBuf &off_buf, DevBuf &ids_buf,
                              PinnedBuf &h_off, PinnedBuf &h_ids, cudaStream_t stream) {
    const int Q = b.Q;
    int32_t *ho = h_off.get<int32_t>((size_t)Q + 1);
    memset(ho, 0, sizeof(int32_t) * ((size_t)Q + 1));
    h_ids.get<int32_t>(1);
    if (n == 0) return;
    uint32_t *flags = b.scratch.get<uint32_t>((size_t)n + 2);
    uint32_t *tot = b.counters.get<uint32_t>(16);
    qlist_flags_kernel<<<cgx_div_up(n, 256), 256, 0, stream>>>(pid, qid, n, flags);
    exclusive_scan_u32(flags, flags, (size_t)n, tot, stream, b.scan, 0, &b.launches);
    uint32_t M = read_u32(tot, stream);
    uint64_t *keys = b.ql_keys.get<uint64_t>(M), *keys_tmp = b.ql_keys_tmp.get<uint64_t>(M);
    qlist_compact_kernel<<<cgx_div_up(n, 256), 256, 0, stream>>>(pid, qid, flags, n, keys);
    uint64_t *ks;
    radix_sort<uint64_t>(keys, keys_tmp, nullptr, nullptr, M, 0, 32 + cgx_bits_for((uint64_t)Q), stream, b.radix, &ks, nullptr, &b.launches);
    (void)idbits;
    int32_t *off = off_buf.get<int32_t>((size_t)Q + 1);
    int32_t *ids = ids_buf.get<int32_t>(M);
    query_offsets_kernel<<<cgx_div_up(Q + 1, 256), 256, 0, stream>>>(ks, (int)M, Q, off);
    low32_kernel<<<cgx_div_up(M, 256), 256, 0, stream>>>(ks, (int)M, ids);
    b.launches += 4;
    if (!b.fetch_results) return;
    int32_t *hi = h_ids.get<int32_t>((size_t)M + 1);
    fetch_async(b, ho, off, sizeof(int32_t) * ((size_t)Q + 1), stream);
    fetch_async(b, hi, ids, sizeof(int32_t) * (size_t)M, stream);
}

void stage_onegap_enumerate(const Index &ix, Batch &b, cudaStream_t stream) {
    const int T = b.T;
    b.enu1 = 0; b.D1 = 0;
    memset(b.h_q1_off.get<int32_t>((size_t)b.Q + 1), 0, sizeof(int32_t) * ((size_t)b.Q + 1));
    b.h_q1_ids.get<int32_t>(1);
    if (T == 0) return;
    uint32_t *cnt = b.e1_count.get<uint32_t>((size_t)T + 2);
    uint32_t *tot = b.counters.get<uint32_t>(16);
    const int32_t *q_tok = b.q_tok.ptr<int32_t>(), *q_off = b.q_off.ptr<int32_t>(), *tok2q = b.tok2q.ptr<int32_t>();
    const int32_t *longest = b.longest.ptr<int32_t>(), *iv = b.iv.ptr<int32_t>();
    e1_enum_kernel<false><<<cgx_div_up(T, 128), 128, 0, stream>>>(q_tok, q_off, tok2q, T, longest, iv, cnt, nullptr, nullptr, nullptr, nullptr, nullptr);
    exclusive_scan_u32(cnt, cnt, (size_t)T, tot, stream, b.scan, 0, &b.launches);
    uint32_t E = read_u32(tot, stream);
    b.launches++;
    b.enu1 = (int32_t)E;
    if (E == 0) return;
    int32_t *inst_t = b.e1_inst.get<int32_t>((size_t)E * 2);
    uint32_t *inst_info = (uint32_t *)(inst_t + E);
    uint64_t *keys = b.e1_keys.get<uint64_t>(E), *keys_tmp = b.e1_keys_tmp.get<uint64_t>(E);
    uint32_t *vals = b.e1_vals.get<uint32_t>(E), *vals_tmp = b.e1_vals_tmp.get<uint32_t>(E);
    PROF("enum_onegap", (double)E * 24, (e1_enum_kernel<true><<<cgx_div_up(T, 128), 128, 0, stream>>>(q_tok, q_off, tok2q, T, longest, iv, nullptr, cnt, inst_t, inst_info, keys, vals)));
    b.launches++;
    uint64_t *ks;
    uint32_t *vs;
    radix_sort<uint64_t>(keys, keys_tmp, vals, vals_tmp, E, 0, 64, stream, b.radix, &ks, &vs, &b.launches);
    // keep the sorted instance order (two-gap enumeration walks the instances of a pattern)
    if (vs != vals) CUDA_CHECK(cudaMemcpyAsync(vals, vs, sizeof(uint32_t) * E, cudaMemcpyDeviceToDevice, stream));
    uint32_t *flags = b.e1_flags.get<uint32_t>((size_t)E + 2);
    head_flags_u64_kernel<<<cgx_div_up(E, 256), 256, 0, stream>>>(ks, E, flags);
    exclusive_scan_u32(flags, flags, E, tot, stream, b.scan, 0, &b.launches);
    b.D1 = (int32_t)read_u32(tot, stream);
    b.launches++;
    Pat1 *pat = b.pat1.get<Pat1>((size_t)b.D1);
    Pat1Dev *patd = b.pat1_dev.get<Pat1Dev>((size_t)b.D1);
    int32_t *pat_pos = b.pat1_pos.get<int32_t>((size_t)b.D1 + 1);
    int32_t *pat_ga = b.pat1_ga.get<int32_t>((size_t)b.D1 + 1);
    int32_t *pid = b.e1_pid.get<int32_t>((size_t)E * 2);
    int32_t *qid = pid + E;
    e1_patterns_kernel<<<cgx_div_up(E, 256), 256, 0, stream>>>(ks, vals, flags, (int)E, inst_t, inst_info, iv, ix.sa.ptr<int32_t>(), ix.str.ptr<int32_t>(),
                                                             ix.freq_flag.ptr<uint8_t>(), b.phrase_id.ptr<int32_t>(), pid, pat, patd, pat_pos, pat_ga);
    int32_t e_i = (int32_t)E;
    CUDA_CHECK(cudaMemcpyAsync(pat_pos + b.D1, &e_i, sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    e1_sorted_qid_kernel<<<cgx_div_up(E, 256), 256, 0, stream>>>(vals, inst_t, tok2q, (int)E, qid);
    b.launches += 2;
    build_query_lists(b, pid, qid, (int)E, 0, b.q1_off, b.q1_ids, b.h_q1_off, b.h_q1_ids, stream);
    CUDA_CHECK(cudaStreamSynchronize(stream));
}

// ------------------------------------------------------------------------------------------------
// two-gap patterns aXbXc  (a, b, c single tokens: limit_symbol = 5-2-ls-le >= 1, SuffixArray.cu:840-850)
// ------------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void e2_enum_kernel(const int32_t *__restrict__ pid, const uint32_t *__restrict__ sorted_inst, int E, const Pat1 *__restrict__ pat,
                               const int32_t *__restrict__ inst_t, const uint32_t *__restrict__ inst_info, const int32_t *__restrict__ q_tok,
                               const int32_t *__restrict__ q_off, const int32_t *__restrict__ tok2q, const int32_t *__restrict__ longest,
                               uint32_t *__restrict__ count, const uint32_t *__restrict__ off, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= E) return;
    uint32_t c = 0;
    const int d = pid[k];
    const Pat1 p = pat[d];
    if (p.hit_count > 0 && p.ls == 1 && p.le == 1) {
        uint32_t inst = sorted_inst[k];
        int t = inst_t[inst];
        int gap = (inst_info[inst] >> 8) & 0xff;
        int search_start = t + 1 + gap + 1 - 1;                     // last token of b
        int end = q_off[tok2q[search_start] + 1];
        int smax = min(end - 1, t + CGX_MAX_RULE_SPAN);
        uint32_t o = FILL ? off[k] : 0;
        for (int s = search_start + 2; s <= smax; s++) {
            if (longest[s] >= 1) {
                if (FILL) {
                    keys[o + c] = ((uint64_t)(uint32_t)d << 32) | (uint64_t)(uint32_t)q_tok[s];
                    vals[o + c] = (uint32_t)s;
                }
                c++;
            }
        }
    }
    if (!FILL) count[k] = c;
}

__global__ void e2_patterns_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ excl, int n,
                                   const int32_t *__restrict__ tok2q, int32_t *__restrict__ pid2, int32_t *__restrict__ qid, Pat2 *__restrict__ pat2) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    bool head = (k == 0 || keys[k] != keys[k - 1]);
    int d = (int)excl[k] + (head ? 1 : 0) - 1;
    pid2[k] = d;
    qid[k] = tok2q[vals[k]];
    if (head) {
        Pat2 p;
        p.pat1 = (int32_t)(keys[k] >> 32); p.ctok = (int32_t)(uint32_t)keys[k]; p.hit_start = -1; p.hit_count = 0;
        pat2[d] = p;
    }
}

void stage_twogap_enumerate(const Index &ix, Batch &b, cudaStream_t stream) {
    b.enu2 = 0; b.D2 = 0;
    memset(b.h_q2_off.get<int32_t>((size_t)b.Q + 1), 0, sizeof(int32_t) * ((size_t)b.Q + 1));
    b.h_q2_ids.get<int32_t>(1);
    const int E = b.enu1;
    if (E == 0 || b.D1 == 0) return;
    uint32_t *cnt = b.e2_count.get<uint32_t>((size_t)E + 2);
    uint32_t *tot = b.counters.get<uint32_t>(16);
    const int32_t *pid = b.e1_pid.ptr<int32_t>();
    const uint32_t *sorted_inst = b.e1_vals.ptr<uint32_t>();
    const int32_t *inst_t = b.e1_inst.ptr<int32_t>();
    const uint32_t *inst_info = (const uint32_t *)(inst_t + E);
    e2_enum_kernel<false><<<cgx_div_up(E, 256), 256, 0, stream>>>(pid, sorted_inst, E, b.pat1.ptr<Pat1>(), inst_t, inst_info, b.q_tok.ptr<int32_t>(),
                                                                 b.q_off.ptr<int32_t>(), b.tok2q.ptr<int32_t>(), b.longest.ptr<int32_t>(), cnt, nullptr, nullptr, nullptr);
    exclusive_scan_u32(cnt, cnt, (size_t)E, tot, stream, b.scan, 0, &b.launches);
    uint32_t E2 = read_u32(tot, stream);
    b.launches++;
    b.enu2 = (int32_t)E2;
    if (E2 == 0) return;
    uint64_t *keys = b.e2_keys.get<uint64_t>(E2), *keys_tmp = b.e2_keys_tmp.get<uint64_t>(E2);
    uint32_t *vals = b.e2_vals.get<uint32_t>(E2), *vals_tmp = b.e2_vals_tmp.get<uint32_t>(E2);
    PROF("enum_twogap", (double)E2 * 12, (e2_enum_kernel<true><<<cgx_div_up(E, 256), 256, 0, stream>>>(pid, sorted_inst, E, b.pat1.ptr<Pat1>(), inst_t, inst_info, b.q_tok.ptr<int32_t>(),
                                                                b.q_off.ptr<int32_t>(), b.tok2q.ptr<int32_t>(), b.longest.ptr<int32_t>(), nullptr, cnt, keys, vals)));
    b.launches++;
    uint64_t *ks;
    uint32_t *vs;
    radix_sort<uint64_t>(keys, keys_tmp, vals, vals_tmp, E2, 0, 32 + cgx_bits_for((uint64_t)b.D1), stream, b.radix, &ks, &vs, &b.launches);
    uint32_t *flags = b.e2_flags.get<uint32_t>((size_t)E2 + 2);
    head_flags_u64_kernel<<<cgx_div_up(E2, 256), 256, 0, stream>>>(ks, E2, flags);
    exclusive_scan_u32(flags, flags, E2, tot, stream, b.scan, 0, &b.launches);
    b.D2 = (int32_t)read_u32(tot, stream);
    b.launches++;
    Pat2 *pat2 = b.pat2.get<Pat2>((size_t)b.D2);
    int32_t *pid2 = b.scratch2.get<int32_t>((size_t)E2 * 2);
    int32_t *qid = pid2 + E2;
    e2_patterns_kernel<<<cgx_div_up(E2, 256), 256, 0, stream>>>(ks, vs, flags, (int)E2, b.tok2q.ptr<int32_t>(), pid2, qid, pat2);
    b.launches++;
    build_query_lists(b, pid2, qid, (int)E2, 0, b.q2_off, b.q2_ids, b.h_q2_off, b.h_q2_ids, stream);
    CUDA_CHECK(cudaStreamSynchronize(stream));
    (void)ix;
}

}  // namespace cgx
