// Per-row helpers shared by the row-transform kernels (lec_rows.cu) and the fused update kernel (lec_update.cu).
#pragma once
#include <math.h>

#include "lec_common.cuh"

namespace lec {

constexpr int kMaxPeers = LEC_MAX_PEERS;

// Sum of the gradient replicas of one element.  Four independent partial sums keep eight loads in flight (the
// replicas sit n*ld floats apart in L2); the order is fixed, so every caller gets the same bits.
__device__ __forceinline__ float rsum(const float* g, int replicas, int64_t stride) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int r = 0;
#pragma unroll 2
    for (; r + 4 <= replicas; r += 4) {
        const float a = g[r * stride], b = g[(r + 1) * stride], c = g[(r + 2) * stride], d = g[(r + 3) * stride];
        s0 += a; s1 += b; s2 += c; s3 += d;
    }
    for (; r < replicas; ++r) s0 += g[r * stride];
    return (s0 + s1) + (s2 + s3);
}

__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// the same sum on one 16-byte chunk (same association per component as rsum)
__device__ __forceinline__ float4 rsum4(const float* g, int replicas, int64_t stride) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 s0 = z, s1 = z, s2 = z, s3 = z;
    int r = 0;
#pragma unroll 2
    for (; r + 4 <= replicas; r += 4) {
        const float4 a = *reinterpret_cast<const float4*>(g + r * stride);
        const float4 b = *reinterpret_cast<const float4*>(g + (r + 1) * stride);
        const float4 c = *reinterpret_cast<const float4*>(g + (r + 2) * stride);
        const float4 d = *reinterpret_cast<const float4*>(g + (r + 3) * stride);
        s0 = add4(s0, a); s1 = add4(s1, b); s2 = add4(s2, c); s3 = add4(s3, d);
    }
    for (; r < replicas; ++r) s0 = add4(s0, *reinterpret_cast<const float4*>(g + r * stride));
    return add4(add4(s0, s1), add4(s2, s3));
}

// shell projection (order_embeddings_h.py:217-228): out = (add + e) / div * mul
__device__ __forceinline__ void shell_factor(float r, float r_in, bool feat, float& mul, float& add, float& div) {
    mul = 1.f; add = 0.f; div = 1.f;
    if (r <= r_in) { mul = r_in; div = feat ? (1e-6f + r) : r; add = feat ? 1e-6f : 0.f; }
    if (r >= 1.0f) { mul = (float)(1.0 - 1e-5); div = r; add = 0.f; }
}

// r_in = 2K / (1 + sqrt(1 + 4K^2)) (order_embeddings_h.py:1089) and c0 = atanh(clamp(r_in)) (oe_h.py:106-110)
inline void hyp_constants(float K, float& r_in, float& c0) {
    const double k = (double)K;
    const double rin = 2.0 * k / (1.0 + sqrt(1.0 + 4.0 * k * k));
    double v = rin;
    if (v < -1 + 1e-5) v = -1 + 1e-5;
    if (v > 1 - 1e-5) v = 1 - 1e-5;
    r_in = (float)rin;
    c0 = (float)(0.5 * (log(1 + v) - log(1 - v)));
}

// Per-row aperture terms, batched.  row_aux<double> is an fp64 sqrt + division + asin (several hundred issue slots) and
// only ONE lane of a row's team has to run it, so inline it occupied a whole warp for 32 / TT rows at a time (cfg4,
// 82 K rows x 50: ~35 us of an 83 us launch).  Teams park |row|^2 in shared memory instead; every kThreads rows (or at the
// end) the block computes one row per THREAD.  Every thread of the block must call push()/flush() the same number of
// times; blockDim.x == kThreads.
struct AuxBatch {
    double A[kThreads];
    int64_t row[kThreads];
};
template <int TT>
__device__ __forceinline__ void aux_flush(AuxBatch& b, int& fill, int geom, float K, double* __restrict__ aux) {
    __syncthreads();
    if ((int)threadIdx.x < fill && b.row[threadIdx.x] >= 0) {
        const Aux<double> x = row_aux_fast(geom, b.A[threadIdx.x], K);
        double2* dst = reinterpret_cast<double2*>(aux + 4 * b.row[threadIdx.x]);
        dst[0] = make_double2(x.A, x.ria);
        dst[1] = make_double2(x.t0, x.t1);
    }
    __syncthreads();
    fill = 0;
}
template <int TT>
__device__ __forceinline__ void aux_push(AuxBatch& b, int& fill, double A, int64_t row, bool valid, int geom, float K,
                                         double* __restrict__ aux) {
    constexpr int kTeams = kThreads / TT;
    if (threadIdx.x % TT == 0) {
        b.A[fill + threadIdx.x / TT] = A;
        b.row[fill + threadIdx.x / TT] = valid ? row : -1;
    }
    fill += kTeams;
    if (fill + kTeams > kThreads) aux_flush<TT>(b, fill, geom, K, aux);
}

}  // namespace lec
