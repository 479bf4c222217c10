// cgx-b200: the per-GPU context behind the C ABI.
#pragma once
#include "../../include/cgx_b200.h"
#include "index.h"
#include "batch.h"
#include "prof.h"

struct cgx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    cgx::Index ix;
    cgx::SaWorkspace ws;
    cgx::Batch batch;
    float aux_ms = 0.f;
    cgx::Prof prof;
    std::string prof_json;
};
