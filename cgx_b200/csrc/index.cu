// cgx-b200: auxiliary index arrays built once per corpus, on the GPU, after the suffix array.
#include "index.h"
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
    // position-sorted occurrence lists of every 1-, 2- and 3-gram
    const int tokbits = cgx_bits_for((uint64_t)ix.maxtok);
    uint64_t *keys = ws.keys.get<uint64_t>(n), *keys_tmp = ws.keys_tmp.get<uint64_t>(n);
    uint32_t *vals = ws.vals.get<uint32_t>(n), *vals_tmp = ws.vals_tmp.get<uint32_t>(n);
    for (int mlen = 1; mlen <= 3; mlen++) {
        ix_ngram_keys_kernel<<<cgx_div_up(n, 256), 256, 0, stream>>>(str, n, mlen, tokbits, keys, vals);
        if (launches) *launches += 1;
        uint64_t *ks;
        uint32_t *vs;
        radix_sort<uint64_t>(keys, keys_tmp, vals, vals_tmp, n, 0, mlen * tokbits, stream, ws.radix, &ks, &vs, launches);
        CUDA_CHECK(cudaMemcpyAsync(ix.inv[mlen - 1].get<int32_t>(n), vs, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(stream));
}

}  // namespace cgx
