// cgx-b200: one-time broadcast of the resident index from GPU 0 to the other GPUs of the box (NCCL over
// NVLink 5 / NVSwitch).  The reference is single-GPU; this is the only collective on the path -- queries
// are embarrassingly parallel, so after the broadcast every GPU works on its own query shard.
#include "context.h"
#include <dlfcn.h>
#include <nccl.h>

using namespace cgx;

namespace {
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    void load() {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) { h = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (h) break; }
        CGX_REQUIRE(h != nullptr, "cannot dlopen libnccl.so.2: %s", dlerror());
        CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
        Broadcast = (decltype(Broadcast))dlsym(h, "ncclBroadcast");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        CGX_REQUIRE(CommInitAll && CommDestroy && GroupStart && GroupEnd && Broadcast && GetErrorString, "libnccl lacks a required symbol");
    }
};
}  // namespace

#define NCCL_CHECK(api, expr)                                                            \
    do {                                                                                 \
        ncclResult_t r_ = (expr);                                                        \
        CGX_REQUIRE(r_ == ncclSuccess, "%s -> %s", #expr, (api).GetErrorString(r_));     \
    } while (0)

extern "C" int cgx_index_broadcast(cgx_ctx_t **ctxs, int n) {
    if (!ctxs || n < 1 || !ctxs[0]) return 1;
    cgx_ctx *root = ctxs[0];
    try {
        CGX_REQUIRE(root->ix.built, "index not built on the root context");
        if (n == 1) return 0;
        NcclApi api;
        api.load();
        std::vector<int> devs(n);
        for (int i = 0; i < n; i++) devs[i] = ctxs[i]->device;
        std::vector<ncclComm_t> comms(n);
        NCCL_CHECK(api, api.CommInitAll(comms.data(), n, devs.data()));
        cgx_index_arrays_t shape;
        CGX_REQUIRE(cgx_index_export(root, &shape) == 0, "export failed: %s", root->err.c_str());
        std::vector<cgx_index_arrays_t> arr(n);
        arr[0] = shape;
        for (int i = 1; i < n; i++) CGX_REQUIRE(cgx_index_alloc(ctxs[i], &shape, &arr[i]) == 0, "alloc on device %d failed: %s", devs[i], ctxs[i]->err.c_str());
        const size_t nt = (size_t)shape.max_token + 2;
        struct Item { size_t off; size_t bytes; };
        const Item items[] = {
            {offsetof(cgx_index_arrays_t, str), (size_t)(shape.n + 3) * 4}, {offsetof(cgx_index_arrays_t, sa), (size_t)shape.n * 4},
            {offsetof(cgx_index_arrays_t, inv1), (size_t)shape.n * 4}, {offsetof(cgx_index_arrays_t, inv2), (size_t)shape.n * 4},
            {offsetof(cgx_index_arrays_t, inv3), (size_t)shape.n * 4}, {offsetof(cgx_index_arrays_t, bkt1), (size_t)shape.n * 4},
            {offsetof(cgx_index_arrays_t, bkt2), (size_t)shape.n * 4}, {offsetof(cgx_index_arrays_t, bkt3), (size_t)shape.n * 4}, {offsetof(cgx_index_arrays_t, tok_start), nt * 4},
            {offsetof(cgx_index_arrays_t, RLP), (size_t)shape.n * (shape.wide ? 8 : 4)}, {offsetof(cgx_index_arrays_t, L_tar), (size_t)shape.m * (shape.wide ? 2 : 1)},
            {offsetof(cgx_index_arrays_t, R_tar), (size_t)shape.m * (shape.wide ? 2 : 1)}, {offsetof(cgx_index_arrays_t, tgt), (size_t)(shape.m + 3) * 4},
            {offsetof(cgx_index_arrays_t, freq_flag), nt}, {offsetof(cgx_index_arrays_t, gapw), (size_t)shape.n * 4}, {offsetof(cgx_index_arrays_t, lex_key), (size_t)(shape.lex_count + 1) * 8},
            {offsetof(cgx_index_arrays_t, lex_v1), (size_t)(shape.lex_count + 1) * 4}, {offsetof(cgx_index_arrays_t, lex_v2), (size_t)(shape.lex_count + 1) * 4}};
        for (const Item &it : items) {
            NCCL_CHECK(api, api.GroupStart());
            for (int i = 0; i < n; i++) {
                CUDA_CHECK(cudaSetDevice(devs[i]));
                void *buf = *(void **)((char *)&arr[i] + it.off);
                void *src = *(void **)((char *)&arr[0] + it.off);
                NCCL_CHECK(api, api.Broadcast(src, buf, it.bytes, ncclChar, 0, comms[i], ctxs[i]->stream));
            }
            NCCL_CHECK(api, api.GroupEnd());
        }
        for (int i = 0; i < n; i++) {
            CUDA_CHECK(cudaSetDevice(devs[i]));
            CUDA_CHECK(cudaStreamSynchronize(ctxs[i]->stream));
            if (i) cgx_index_commit(ctxs[i]);
        }
        for (int i = 0; i < n; i++) api.CommDestroy(comms[i]);
        return 0;
    } catch (const CgxError &e) {
        root->err = e.msg;
        return 1;
    }
}
