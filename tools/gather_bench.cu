// GPU box: random-gather roofline of HBM3e -- how many independent 4-byte (one 32-byte sector each) and 16-byte loads per
// second a B200 sustains from a working set far larger than L2.  The join / extraction / scoring kernels are made of such
// gathers; this number, not the streaming copy peak, is their practical ceiling.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_bench tools/gather_bench.cu && /tmp/gather_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int ILP, typename T>
__global__ void gather(const T *__restrict__ a, size_t n_items, int iters, unsigned long long *out) {
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
    unsigned long long acc = 0;
    for (int it = 0; it < iters; it++) {
        T v[ILP];
#pragma unroll
        for (int k = 0; k < ILP; k++) { s = mix(s + 0x9e3779b9u * (k + 1)); v[k] = __ldg(&a[(size_t)s % n_items]); }
#pragma unroll
        for (int k = 0; k < ILP; k++) acc += *reinterpret_cast<const uint32_t *>(&v[k]);
    }
    if (acc == 0x1234567812345678ull) *out = acc;
}

template <int ILP, typename T>
void run(const char *name, const void *buf, size_t bytes, unsigned long long *out) {
    const int iters = 64, block = 256, grid = 148 * 8 * 16;
    size_t n_items = bytes / sizeof(T);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<ILP, T><<<grid, block>>>((const T *)buf, n_items, 4, out);
    cudaEventRecord(e0);
    gather<ILP, T><<<grid, block>>>((const T *)buf, n_items, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double loads = (double)grid * block * iters * ILP;
    printf("%-28s working set %6.1f GB: %7.2f G loads/s = %7.1f GB/s of 32-byte sectors (%5.3f ms)\n", name, bytes / 1e9, loads / ms / 1e6, loads * 32 / ms / 1e6, ms);
}

int main() {
    void *buf;
    unsigned long long *out;
    size_t bytes = (size_t)4 << 30;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&out, 8);
    cudaMemset(buf, 1, bytes);
    for (size_t ws : {(size_t)64 << 20, (size_t)1 << 30, (size_t)4 << 30}) {
        run<1, uint32_t>("4-byte loads, ILP 1", buf, ws, out);
        run<4, uint32_t>("4-byte loads, ILP 4", buf, ws, out);
        run<8, uint32_t>("4-byte loads, ILP 8", buf, ws, out);
        run<4, uint4>("16-byte loads, ILP 4", buf, ws, out);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
