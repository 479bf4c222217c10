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
#endif

static inline uint32_t pt_slots_for(size_t entries) {       // power of two, load factor <= 0.71
    uint32_t s = 1024;
    while ((double)s * 0.71 < (double)entries) s <<= 1;
    return s;
}

}  // namespace cgx
