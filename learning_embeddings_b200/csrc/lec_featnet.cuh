// Arguments of the FeatNet kernels (lec_featnet.cu), shared with the C ABI layer (lec_api.cu).
#pragma once
#include "lec_common.cuh"

namespace lec {

struct FeatArgs {
    const float* features; int64_t n_pool; int F;
    const void* sel; int sel_bytes; int64_t m;
    const float* weight; const float* bias; int D;
    float* Y;
    const float* gY; float* grad; int replicas; int64_t grad_stride;
    unsigned* index_errors;
};

}  // namespace lec
