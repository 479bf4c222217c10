// cgx-b200: width of the alignment fields.
//
// The reference keeps, per source token, the aligned target span [L, R] and the token's position P in its sentence in 8 bits
// each (RLP = L << 24 | R << 16 | P << 8, 255 = unaligned; L_tar / R_tar likewise per target token) and exits on any sentence
// of 255 tokens or more (ExtractPair.cu:2683, Start.cu:269).  The kernels that read these fields -- the gap-word builder, the
// extraction views and the three extraction kernels -- are templated on one of the two layouts below: AlignNarrow is the
// reference's and the default; AlignWide carries the same fields in 16 bits (RLP = L << 48 | R << 32 | P << 16, 65535 =
// unaligned) and is chosen when the corpus has a longer sentence (cgx_index_build_wide; SURVEY.md 8f, lifted limit).  On a
// corpus that fits the narrow layout both give identical results (tests: CGX_FORCE_WIDE=1).  In both, bit 0 of the extraction
// view xw marks a word (token >= 2), and the word at an EOS position holds the target sentence offset in its low 32 bits.
#pragma once
#include "common.cuh"

namespace cgx {

struct AlignNarrow {
    using word_t = uint32_t;           // RLP / xw word
    using lr_t = uint8_t;              // L_tar / R_tar
    using lrq_t = uint2;               // range-minimum entry of a target token: levels 0..3 x {min L : 8, max R : 8}
    static constexpr unsigned UNAL = 255u;
    static __host__ __device__ __forceinline__ unsigned L(word_t w) { return (w >> 24) & 0xFFu; }
    static __host__ __device__ __forceinline__ unsigned R(word_t w) { return (w >> 16) & 0xFFu; }
    static __host__ __device__ __forceinline__ unsigned P(word_t w) { return (w >> 8) & 0xFFu; }
    static __device__ __forceinline__ void lrq_get(lrq_t v, int k, unsigned &mn, unsigned &mx) {
        const unsigned w = ((k & 2) ? v.y : v.x) >> (16 * (k & 1));
        mn = w & 0xFFu; mx = (w >> 8) & 0xFFu;
    }
    static __device__ __forceinline__ lrq_t lrq_make(const unsigned mn[4], const unsigned mx[4]) {
        return make_uint2((mn[0] | (mx[0] << 8)) | ((mn[1] | (mx[1] << 8)) << 16), (mn[2] | (mx[2] << 8)) | ((mn[3] | (mx[3] << 8)) << 16));
    }
};

struct AlignWide {
    using word_t = uint64_t;
    using lr_t = uint16_t;
    using lrq_t = uint4;               // levels 0..3 x {min L : 16, max R : 16}
    static constexpr unsigned UNAL = 65535u;
    static __host__ __device__ __forceinline__ unsigned L(word_t w) { return (unsigned)(w >> 48) & 0xFFFFu; }
    static __host__ __device__ __forceinline__ unsigned R(word_t w) { return (unsigned)(w >> 32) & 0xFFFFu; }
    static __host__ __device__ __forceinline__ unsigned P(word_t w) { return (unsigned)(w >> 16) & 0xFFFFu; }
    static __device__ __forceinline__ void lrq_get(lrq_t v, int k, unsigned &mn, unsigned &mx) {
        const unsigned w = k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w;
        mn = w & 0xFFFFu; mx = w >> 16;
    }
    static __device__ __forceinline__ lrq_t lrq_make(const unsigned mn[4], const unsigned mx[4]) {
        return make_uint4(mn[0] | (mx[0] << 16), mn[1] | (mx[1] << 16), mn[2] | (mx[2] << 16), mn[3] | (mx[3] << 16));
    }
};

}  // namespace cgx
