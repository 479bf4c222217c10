// cgx-b200: open-addressing device hash tables (64-bit key -> 64-bit payload), linear probing, 16-byte slots so
// that key and payload arrive in one 16-byte load (one sector).  Built once per batch (patterns) or once per
// corpus (lexical table); capacity is a power of two with load factor <= 0.5.
#pragma once
#include "common.cuh"

namespace cgx {

constexpr uint64_t HT_EMPTY = ~0ull;

static inline uint32_t ht_slots_for(size_t entries) {
    uint32_t s = 1024;
    while ((size_t)s < 2 * entries) s <<= 1;
    return s;
}

#ifdef __CUDACC__
// 32-bit mixing of a 64-bit key (two odd multipliers + an xorshift-multiply finaliser): ~8 instructions instead of the
// ~25 of a 64-bit murmur finaliser -- the join and scoring kernels are instruction-bound as much as latency-bound
__device__ __forceinline__ uint32_t ht_mix(uint64_t z) {
    uint32_t h = (uint32_t)z * 0x9E3779B1u ^ (uint32_t)(z >> 32) * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}

// first writer of a key wins the slot; later inserts of the same key overwrite the payload (callers insert each key once)
__device__ __forceinline__ void ht_insert(ulonglong2 *__restrict__ slots, uint32_t mask, uint64_t key, uint64_t payload) {
    uint32_t s = ht_mix(key) & mask;
    while (true) {
        unsigned long long prev = atomicCAS(&slots[s].x, (unsigned long long)HT_EMPTY, (unsigned long long)key);
        if (prev == HT_EMPTY || prev == key) { slots[s].y = payload; return; }
        s = (s + 1) & mask;
    }
}

// Two-step lookup for memory-level parallelism: the caller loads the first slot of several keys back to back
// (ht_first), then resolves each (ht_resolve continues along the probe sequence only on a collision).
__device__ __forceinline__ ulonglong2 ht_first(const ulonglong2 *__restrict__ slots, uint32_t mask, uint64_t key, uint32_t *slot) {
    *slot = ht_mix(key) & mask;
    return __ldg(&slots[*slot]);
}
__device__ __forceinline__ bool ht_resolve(const ulonglong2 *__restrict__ slots, uint32_t mask, uint64_t key, uint32_t s, ulonglong2 v, uint64_t *payload) {
    while (true) {
        if (v.x == key) { *payload = v.y; return true; }
        if (v.x == HT_EMPTY) return false;
        s = (s + 1) & mask;
        v = __ldg(&slots[s]);
    }
}

__device__ __forceinline__ bool ht_find(const ulonglong2 *__restrict__ slots, uint32_t mask, uint64_t key, uint64_t *payload) {
    uint32_t s = ht_mix(key) & mask;
    while (true) {
        const ulonglong2 v = __ldg(&slots[s]);
        if (v.x == key) { *payload = v.y; return true; }
        if (v.x == HT_EMPTY) return false;
        s = (s + 1) & mask;
    }
}
// ---- packed 8-byte table: slot = key << vbits | value (key << vbits | value < 2^63, so no slot equals HT_EMPTY), linear
// probing at load factor <= 0.71.  Used where key and value fit 63 bits (two-gap patterns: parent id, token -> id): half the
// bytes of the 16-byte form, 44 MB for the 5.5e6 two-gap patterns of a C2 batch -> L2-resident.  (Measured for the one-gap
// table too: its key needs a second dependent lookup to fit, and the longer latency chain made that kernel 2x slower.)
struct PackTab {
    unsigned long long *slots;
    uint32_t mask;
    int vbits;
};

__device__ __forceinline__ void pt_insert(const PackTab t, uint64_t key, uint64_t val) {
    const unsigned long long w = (key << t.vbits) | val;
    uint32_t s = ht_mix(key) & t.mask;
    while (true) {
        const unsigned long long prev = atomicCAS(&t.slots[s], (unsigned long long)HT_EMPTY, w);
        if (prev == HT_EMPTY || (prev >> t.vbits) == key) return;
        s = (s + 1) & t.mask;
    }
}
__device__ __forceinline__ unsigned long long pt_first(const PackTab t, uint64_t key, uint32_t *slot) {
    *slot = ht_mix(key) & t.mask;
    return __ldg(&t.slots[*slot]);
}
__device__ __forceinline__ bool pt_resolve(const PackTab t, uint64_t key, uint32_t s, unsigned long long w, uint64_t *val) {
    while (true) {
        if ((w >> t.vbits) == key) { *val = w & ((1ull << t.vbits) - 1ull); return true; }
        if (w == HT_EMPTY) return false;
        s = (s + 1) & t.mask;
        w = __ldg(&t.slots[s]);
    }
}
__device__ __forceinline__ bool pt_find(const PackTab t, uint64_t key, uint64_t *val) {
    uint32_t s;
    const unsigned long long w = pt_first(t, key, &s);
    return pt_resolve(t, key, s, w, val);
}
#endif

// ---- quotient table: 8-byte slots in 32-byte buckets (one sector), for keys too wide for PackTab.  A key is a pair (A, B) of
// abits + bbits bits; it goes through a BIJECTION of that many bits built from 32-bit operations (add a multiple of B to A, odd
// multiplies and xor-shifts of A inside abits bits, then xor B with bits of the result; callers put the WIDER half in A, whose
// top bits become the bucket number -- with the narrow half there, keys sharing it pile up in a few buckets); the top bits of the image choose the
// home bucket and only the remaining low bits -- the remainder, <= 27 bits -- are stored, so (bucket, remainder) still
// identifies the key exactly.  An entry that finds its home bucket full moves on to the next one (at most QT_MAX_DISP buckets
// away) and records how far it went, so its remainder is read against the right home.
//   slot = (remainder << 4 | displacement) << 32 | value          (value < 2^31, so no slot equals HT_EMPTY)
// A bucket with an empty slot has never overflowed, which ends an unsuccessful lookup after one sector.  At <= 3 entries per
// 4-slot bucket the one-gap pattern table of a C2 batch (5.9e6 patterns) is 67 MB instead of the 268 MB of the 16-byte form.
struct QTab {
    unsigned long long *slots;
    uint32_t bmask;        // buckets - 1
    int abits, bbits, rb;  // widths of the key halves; remainder bits = abits + bbits - log2(buckets), <= 27 (tag word < 2^31: never an empty slot's)
};
constexpr int QT_MAX_DISP = 15;

static inline int qt_log2(uint32_t pow2) { int l = 0; while ((1u << l) < pow2) l++; return l; }
// power of two; <= 3 entries per 4-slot bucket and enough buckets for the remainder to fit 27 bits
static inline uint32_t qt_buckets_for(size_t entries, int key_bits) {
    uint32_t b = 256;
    while ((size_t)b * 3 < entries || key_bits - qt_log2(b) > 27) b <<= 1;
    return b;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t qt_mix(const QTab t, uint32_t A, uint32_t B) {      // bijection on abits + bbits bits
    const uint32_t MA = (1u << t.abits) - 1u;
    const int s = max(1, t.abits >> 1);
    A = (A + B * 0x9E3779B1u) & MA;
    A = (A * 0x85EBCA77u) & MA;
    A ^= A >> s;
    A = (A * 0xC2B2AE3Du) & MA;
    A ^= A >> s;
    B ^= (A * 0x27D4EB2Fu) >> (32 - t.bbits);
    return ((uint64_t)A << t.bbits) | (uint64_t)B;
}
// false: no free slot within QT_MAX_DISP buckets of the home bucket (the caller rebuilds the table with more buckets)
__device__ __forceinline__ bool qt_insert(const QTab t, uint32_t A, uint32_t B, uint32_t val) {
    const uint64_t x = qt_mix(t, A, B);
    const uint32_t home = (uint32_t)(x >> t.rb);
    const uint32_t tag = (uint32_t)(x & ((1ull << t.rb) - 1ull)) << 4;
    for (int d = 0; d <= QT_MAX_DISP; d++) {
        unsigned long long *bk = t.slots + (size_t)((home + d) & t.bmask) * 4;
        const unsigned long long w = ((unsigned long long)(tag | (uint32_t)d) << 32) | val;
        for (int i = 0; i < 4; i++) {
            const unsigned long long prev = atomicCAS(&bk[i], (unsigned long long)HT_EMPTY, w);
            if (prev == HT_EMPTY || (uint32_t)(prev >> 32) == (tag | (uint32_t)d)) return true;
        }
    }
    return false;
}
// two-step lookup for memory-level parallelism: qt_touch computes the home bucket of a key and asks for its sector
// (prefetch: no destination registers); qt_resolve reads it
__device__ __forceinline__ void qt_touch(const QTab t, uint32_t A, uint32_t B, uint32_t *home, uint32_t *tag) {
    const uint64_t x = qt_mix(t, A, B);
    *home = (uint32_t)(x >> t.rb);
    *tag = (uint32_t)(x & ((1ull << t.rb) - 1ull)) << 4;                // remainder | displacement 0
    asm volatile("prefetch.global.L1 [%0];" ::"l"(t.slots + (size_t)(*home & t.bmask) * 4));
}
__device__ __forceinline__ bool qt_resolve(const QTab t, uint32_t home, uint32_t tag, uint32_t *val) {
    for (int d = 0;; d++) {
        const uint4 *bk = reinterpret_cast<const uint4 *>(t.slots + (size_t)((home + d) & t.bmask) * 4);
        const uint4 lo = __ldg(bk), hi = __ldg(bk + 1);              // slots: (lo.x, lo.y) (lo.z, lo.w) (hi.x, hi.y) (hi.z, hi.w), tag in the odd word
        const uint32_t key = tag | (uint32_t)d;
        if (lo.y == key) { *val = lo.x; return true; }
        if (lo.w == key) { *val = lo.z; return true; }
        if (hi.y == key) { *val = hi.x; return true; }
        if (hi.w == key) { *val = hi.z; return true; }
        // an empty slot has value word 0xFFFFFFFF; a stored value is < 2^31
        const bool full = (int32_t)(lo.x | lo.z | hi.x | hi.z) >= 0;
        if (!full || d == QT_MAX_DISP) return false;
    }
}
#endif

static inline uint32_t pt_slots_for(size_t entries) {       // power of two, load factor <= 0.71
    uint32_t s = 1024;
    while ((double)s * 0.71 < (double)entries) s <<= 1;
    return s;
}

}  // namespace cgx
