// cgx-b200: per-kernel CUDA-event timing on the launching stream (bench.py roofline, development).
// When enabled, every instrumented launch is bracketed by two events taken from a pool; they are resolved
// after the batch (one cudaEventSynchronize), so the timed region sees two event records per launch and no
// extra synchronisation.
#pragma once
#include "common.cuh"
#include <map>

namespace cgx {

struct ProfEntry { double ms = 0; double bytes = 0; long launches = 0; };

struct Prof {
    bool enabled = false;
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Pending { const char *name; double bytes; size_t e0, e1; };
    std::vector<Pending> pending;
    std::map<std::string, ProfEntry> table;

    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t e;
            CUDA_CHECK(cudaEventCreate(&e));
            pool.push_back(e);
        }
        return pool[used++];
    }
    size_t begin() {
        size_t i = used;
        CUDA_CHECK(cudaEventRecord(get(), stream));
        return i;
    }
    void end(const char *name, double bytes, size_t e0) {
        size_t e1 = used;
        CUDA_CHECK(cudaEventRecord(get(), stream));
        pending.push_back({name, bytes, e0, e1});
    }
    void resolve() {
        if (pending.empty()) { used = 0; return; }
        CUDA_CHECK(cudaEventSynchronize(pool[pending.back().e1]));
        for (auto &p : pending) {
            float ms = 0;
            CUDA_CHECK(cudaEventElapsedTime(&ms, pool[p.e0], pool[p.e1]));
            ProfEntry &t = table[p.name];
            t.ms += ms; t.bytes += p.bytes; t.launches += 1;
        }
        pending.clear();
        used = 0;
    }
    void reset() { table.clear(); pending.clear(); used = 0; }
    void destroy() { for (auto e : pool) cudaEventDestroy(e); pool.clear(); }
};

extern thread_local Prof *g_prof;

// PROF(name, algorithmic_bytes, launch-statement)
#define PROF(name, bytes, ...)                                            \
    do {                                                                  \
        cgx::Prof *p_ = cgx::g_prof;                                      \
        if (p_ && p_->enabled) {                                          \
            size_t e0_ = p_->begin();                                     \
            __VA_ARGS__;                                                  \
            p_->end(name, (double)(bytes), e0_);                          \
        } else {                                                          \
            __VA_ARGS__;                                                  \
        }                                                                 \
    } while (0)

}  // namespace cgx
