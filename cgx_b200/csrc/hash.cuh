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
__device__ __forceinline__ uint64_t ht_mix(uint64_t z) {
    z ^= z >> 33; z *= 0xff51afd7ed558ccdULL; z ^= z >> 33; z *= 0xc4ceb9fe1a85ec53ULL; z ^= z >> 33;
    return z;
}

// first writer of a key wins the slot; later inserts of the same key overwrite the payload (callers insert each key once)
__device__ __forceinline__ void ht_insert(ulonglong2 *__restrict__ slots, uint32_t mask, uint64_t key, uint64_t payload) {
    uint32_t s = (uint32_t)ht_mix(key) & mask;
    while (true) {
        unsigned long long prev = atomicCAS(&slots[s].x, (unsigned long long)HT_EMPTY, (unsigned long long)key);
        if (prev == HT_EMPTY || prev == key) { slots[s].y = payload; return; }
        s = (s + 1) & mask;
    }
}

__device__ __forceinline__ bool ht_find(const ulonglong2 *__restrict__ slots, uint32_t mask, uint64_t key, uint64_t *payload) {
    uint32_t s = (uint32_t)ht_mix(key) & mask;
    while (true) {
        const ulonglong2 v = __ldg(&slots[s]);
        if (v.x == key) { *payload = v.y; return true; }
        if (v.x == HT_EMPTY) return false;
        s = (s + 1) & mask;
    }
}
#endif

}  // namespace cgx
