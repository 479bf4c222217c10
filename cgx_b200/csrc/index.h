// cgx-b200: the resident corpus index (one per GPU) and the per-batch query workspace.
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace cgx {

struct SaStats {
    int rounds = 0;
    int launches = 0;
    int key_bits = 0;
    float ms = 0.f;
};

struct SaWorkspace {
    DevBuf keys, keys_tmp, vals, vals_tmp, rank, flags, total, head, excl, headpos, apos;
    RadixTemp radix;
    ScanTemp scan;
    void release() {
        keys.release(); keys_tmp.release(); vals.release(); vals_tmp.release(); rank.release(); flags.release(); total.release();
        head.release(); excl.release(); headpos.release(); apos.release();
        radix.hist.release(); radix.status.release(); radix.counters.release();
        for (auto &l : scan.level) l.release();
    }
};

void build_suffix_array(const int32_t *d_str, size_t n, int32_t maxtok, int32_t *d_sa_out, SaWorkspace &ws, cudaStream_t stream,
                        SaStats *stats);

// HBM layout of the index (all arrays dense, 16-byte aligned by cudaMalloc):
//   str     int32 [n+3]   source text ids (Start.cu:240-380 layout: EOS=1, trailer, 3 zeros)
//   sa      int32 [n]     suffix array
//   inv[m]  int32 [n]     m = 1..3: positions sorted by (first m tokens, position).  The bucket of an
//                         m-gram occupies the same index range as its SA interval, so inv[m][up..down]
//                         is the *position-sorted* occurrence list of that m-gram (drives the band joins)
//   bkt[m]  int32 [n]     m = 1..3: start of the SA interval (= inv[m] bucket) of the m-gram at every corpus position --
//                         the canonical id of that m-gram; lets the gappy join recognise "b starts at q" with one
//                         coalesced load instead of a search in b's occurrence list
//   jwin    int4  [n]     {bkt[1], bkt[2], bkt[3], gapw} of every position interleaved: the position-major join streams it once
//                         through shared memory (derived locally from bkt[] and gapw, like lex_hash)
//   xw      uint32 [n+3]  extraction view of the source side: RLP | 1 where the token is a word (>= 2), 0 at EOS / padding
//   lr      uint2 [m]     range-minimum table {min L_tar, max R_tar} over 1/2/4/8 target tokens (index.cu; derived locally, like jwin)
//   tok_start int32 [maxtok+2]  SA bucket start of every token id (1-gram intervals in O(1))
//   RLP     uint32 [n]    (L<<24)|(R<<16)|(P<<8) per source token; target sentence offset at EOS
//   L_tar/R_tar uint8 [m] min/max aligned source index per target token (255 = unaligned)
//   tgt     int32 [m+3]   target text ids
//   gapw    uint32 [n]    gap-consistency word of every source position i (built once, on the GPU):
//                         bit g-1 (g = 1..13) = checkBoundaryGap(i, i+g-1) holds (GappyLook.cu:43-126) and every
//                         token of the span is >= 2; bits 16..19 = number of consecutive tokens >= 2 from i
//                         (capped at 15).  Turns the per-candidate gap test of the joins into one word load.
//   freq_flag uint8 [maxtok+2]  rank+1 among the PRECOMPUTECOUNT most frequent source tokens, else 0
//   lex_key uint64 [L] / lex_v1, lex_v2 float [L]  lexical table sorted by (f+1)<<32 | (e+1)
//   lex_hash 16 B x 2^k    the same table as an open-addressing hash (key -> v1 | v2 << 32): one probe per lookup
//                         instead of log2(L) dependent loads (rebuilt locally from the sorted arrays after a broadcast)
struct Index {
    size_t n = 0, m = 0;
    int32_t maxtok = 0;
    DevBuf str, sa, inv[3], bkt[3], jwin, xw, lr, tok_start, RLP, L_tar, R_tar, tgt, freq_flag, gapw;
    DevBuf lex_key, lex_v1, lex_v2, lex_hash;
    size_t lex_count = 0;
    uint32_t lex_hash_mask = 0;
    int32_t freq_list[CGX_PRECOMP];
    bool built = false;
    uint64_t src_sum = 0, tgt_sum = 0;     // FNV-1a of the token arrays the index was built from (cgx_index_matches; 0 = unknown)
    bool wide = false;          // 16-bit alignment fields (align_fields.cuh): RLP / xw are uint64, L_tar / R_tar uint16, lr uint4
    SaStats sa_stats;
};

void build_index_aux(Index &ix, SaWorkspace &ws, cudaStream_t stream, int *launches);
void build_lex_hash(Index &ix, cudaStream_t stream);
void build_jwin(Index &ix, cudaStream_t stream);

}  // namespace cgx
