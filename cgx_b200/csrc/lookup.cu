// cgx-b200: contiguous-phrase lookup -- warp-cooperative suffix-array search.
//
// Replaces suffixArrayFindLwRwKernelTwoWayTDI (SuffixArray.cu:402-767: longest match + 1-gram
// interval per query token) and suffixArrayFindConnectionTwoWayTDI (:109-400: interval of every
// n-gram), plus the host prefix sum and two round trips between them (:1455-1484).
//
// Result per query token t (identical to the reference's, which is search-path independent):
//   longest[t] = min(5, longest prefix of q[t..] present in the corpus; stops at query end / OOV)
//   iv[t][m-1] = inclusive SA interval of q[t..t+m) for m <= longest[t]
// (Only lengths <= 5 = LONGESTCHSOURCE are consumed downstream: ExtractPair.cu:2832,
//  SuffixArray.cu:985 -- the reference's longer intervals are never used.)
//
// One warp per query token.  The 1-gram interval is a table lookup (tok_start).  Each further
// token narrows the interval with two 32-ary searches: the 32 lanes probe 32 evenly spaced suffixes
// of the current range, compare the token at offset m (sa[k] then str[sa[k]+m]: two dependent 4-byte
// gathers), and a ballot picks the sub-range -- log33(range) dependent rounds instead of log2(range).
// The first rounds of every warp hit the same top-of-tree probes, which therefore stay L1/L2 resident.
#include "batch.h"
#include "prof.h"

namespace cgx {

constexpr int LK_WARPS = 8;

// first k in [lo, hi) with str[sa[k]+off] >= x (UPPER=false) or > x (UPPER=true); all lanes return the same value
template <bool UPPER>
__device__ __forceinline__ int warp_bound(const int32_t *__restrict__ sa, const int32_t *__restrict__ str, int lo, int hi, int off, int x, unsigned &probes) {
    const unsigned lane = threadIdx.x & 31;
    while (hi > lo) {
        int len = hi - lo;
        int stride = (len + 32) / 33;
        long long idx = (long long)lo + (long long)(lane + 1) * stride - 1;
        bool inr = idx < hi;
        probes += __popc(__ballot_sync(0xffffffffu, inr));      // suffixes examined (the algorithmic-byte account: 8 B each)
        int v = 0x7fffffff;
        if (inr) v = __ldg(&str[__ldg(&sa[idx]) + off]);
        bool ge = UPPER ? (v > x) : (v >= x);
        unsigned b = __ballot_sync(0xffffffffu, ge);
        if (b == 0) { lo = lo + 32 * stride; continue; }     // all 32 probes in range and below x
        int f = __ffs(b) - 1;
        long long idx_f = (long long)lo + (long long)(f + 1) * stride - 1;
        int new_hi = idx_f < hi ? (int)idx_f : hi;           // answer <= idx_f
        int new_lo = f == 0 ? lo : (int)(lo + (long long)f * stride);   // answer > idx_{f-1}
        lo = new_lo;
        hi = new_hi;
    }
    return lo;
}

__global__ void __launch_bounds__(LK_WARPS * 32) lookup_kernel(const int32_t *__restrict__ sa, const int32_t *__restrict__ str,
                                                               const int32_t *__restrict__ tok_start, int32_t maxtok,
                                                               const int32_t *__restrict__ q_tok, const int32_t *__restrict__ q_off,
                                                               const int32_t *__restrict__ tok2q, int T, int32_t *__restrict__ longest,
                                                               int32_t *__restrict__ iv, unsigned long long *__restrict__ probe_count) {
    const int t = blockIdx.x * LK_WARPS + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31;
    if (t >= T) return;
    const int qend = q_off[tok2q[t] + 1];
    int mlen = 0, lo = 0, hi = -1;
    int my_up = -1, my_down = -1;                 // lane m-1 keeps the interval of length m
    unsigned probes = 0;
    int x = q_tok[t];
    if (x >= 2 && x <= maxtok) {
        lo = tok_start[x];
        hi = tok_start[x + 1] - 1;
        if (hi >= lo) {
            mlen = 1;
            if (lane == 0) { my_up = lo; my_down = hi; }
            while (mlen < CGX_LONGEST_SRC && t + mlen < qend) {
                x = q_tok[t + mlen];
                if (x < 2) break;
                int l = warp_bound<false>(sa, str, lo, hi + 1, mlen, x, probes);
                if (l > hi) break;
                int r = warp_bound<true>(sa, str, l, hi + 1, mlen, x, probes) - 1;
                if (r < l) break;
                lo = l; hi = r;
                if ((int)lane == mlen) { my_up = lo; my_down = hi; }
                mlen++;
            }
        }
    }
    if (lane == 0) longest[t] = mlen;
    if (lane == 0 && probe_count && probes) atomicAdd(probe_count, (unsigned long long)probes);
    if (lane < CGX_LONGEST_SRC) {
        iv[((size_t)t * CGX_LONGEST_SRC + lane) * 2 + 0] = my_up;
        iv[((size_t)t * CGX_LONGEST_SRC + lane) * 2 + 1] = my_down;
    }
}

void stage_lookup(const Index &ix, Batch &b, cudaStream_t stream) {
    const int T = b.T;
    int32_t *longest = b.longest.get<int32_t>((size_t)T + 1);
    int32_t *iv = b.iv.get<int32_t>((size_t)T * CGX_LONGEST_SRC * 2 + 2);
    if (T == 0) return;
    // algorithmic bytes = 8 per suffix examined (its sa entry and the token compared: SURVEY.md 8d, B_look = 8 P); the probes are
    // counted by the kernel only while a profile is being taken (one more read-back)
    const bool counting = g_prof && g_prof->enabled;
    unsigned long long *pc = nullptr;
    if (counting) {
        pc = reinterpret_cast<unsigned long long *>(b.counters.get<uint32_t>(32) + 30);
        CUDA_CHECK(cudaMemsetAsync(pc, 0, sizeof(unsigned long long), stream));
    }
    PROF("lookup", 0.0, (lookup_kernel<<<cgx_div_up(T, LK_WARPS), LK_WARPS * 32, 0, stream>>>(ix.sa.ptr<int32_t>(), ix.str.ptr<int32_t>(), ix.tok_start.ptr<int32_t>(),
                                                                        ix.maxtok, b.q_tok.ptr<int32_t>(), b.q_off.ptr<int32_t>(),
                                                                        b.tok2q.ptr<int32_t>(), T, longest, iv, pc)));
    if (counting) {
        unsigned long long probes = 0;
        cgx_read_back(&probes, pc, sizeof(probes), stream);
        g_prof->table["lookup"].bytes += 8.0 * (double)probes;
    }
    b.launches++;
}

}  // namespace cgx
