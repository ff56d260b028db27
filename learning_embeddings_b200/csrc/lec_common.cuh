// Shared device helpers for the entailment-cone kernels (sm_100a).
//
// Work decomposition used by every pair kernel: a TEAM of T lanes (T a power of two <= 32) owns one
// work item (a pair, or a positive with its 2N negatives).  A row of the transformed table is
// [ld] floats, ld % 4 == 0, and is held by the team as V float4 chunks per lane, chunk q = lane + T*j
// so that one load instruction of the team covers 16*T contiguous bytes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lec_b200.h"

namespace lec {

constexpr int kThreads = 256;
constexpr float kNormEps = 1e-12f;   // F.normalize eps (order_embeddings.py:197, :965)
constexpr float kClampEps = 1e-5f;   // acos/asin clamp (order_embeddings_h.py:1113-1114)

extern unsigned long long g_launches;  // host-side counter (lec_api.cu)

template <int V>
struct Vec {
    float4 c[V];
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// d += a (vector float4 reduction into global memory: one REDG.E.ADD.F32x4 on sm_100a)
__device__ __forceinline__ void red_add4(float* p, float4 v) { atomicAdd(reinterpret_cast<float4*>(p), v); }

template <int T, typename S>
__device__ __forceinline__ S team_sum(S v) {
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename S>
__device__ __forceinline__ S warp_sum(S v) {
    return team_sum<32, S>(v);
}

template <int T, int V>
__device__ __forceinline__ void load_row(Vec<V>& r, const float* __restrict__ rows, int64_t row, int ld, int lane_t) {
    const float* base = rows + row * (int64_t)ld;
    const int Q = ld >> 2;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int q = lane_t + T * j;
        r.c[j] = (q < Q) ? ldg4(base + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <typename S, int V>
__device__ __forceinline__ S dot_part(const Vec<V>& a, const Vec<V>& b) {
    S s = 0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        s += (S)a.c[j].x * (S)b.c[j].x;
        s += (S)a.c[j].y * (S)b.c[j].y;
        s += (S)a.c[j].z * (S)b.c[j].z;
        s += (S)a.c[j].w * (S)b.c[j].w;
    }
    return s;
}

// sum (a-b)^2; the subtraction is done in the accumulator type (fp32: exactly what torch.norm(x - y)
// sees in the reference's fp32 run; fp64: exact difference of the fp32 inputs)
template <typename S, int V>
__device__ __forceinline__ S dist2_part(const Vec<V>& a, const Vec<V>& b) {
    S s = 0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        S d;
        d = (S)a.c[j].x - (S)b.c[j].x; s += d * d;
        d = (S)a.c[j].y - (S)b.c[j].y; s += d * d;
        d = (S)a.c[j].z - (S)b.c[j].z; s += d * d;
        d = (S)a.c[j].w - (S)b.c[j].w; s += d * d;
    }
    return s;
}

// <x, y - x> (Euclidean cone numerator)
template <typename S, int V>
__device__ __forceinline__ S dot_diff_part(const Vec<V>& x, const Vec<V>& y) {
    S s = 0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        s += (S)x.c[j].x * ((S)y.c[j].x - (S)x.c[j].x);
        s += (S)x.c[j].y * ((S)y.c[j].y - (S)x.c[j].y);
        s += (S)x.c[j].z * ((S)y.c[j].z - (S)x.c[j].z);
        s += (S)x.c[j].w * ((S)y.c[j].w - (S)x.c[j].w);
    }
    return s;
}

// sum max(0, x - y)^2 (order-embedding energy)
template <typename S, int V>
__device__ __forceinline__ S relu_diff2_part(const Vec<V>& x, const Vec<V>& y) {
    S s = 0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        float d;
        d = fmaxf(x.c[j].x - y.c[j].x, 0.f); s += (S)d * (S)d;
        d = fmaxf(x.c[j].y - y.c[j].y, 0.f); s += (S)d * (S)d;
        d = fmaxf(x.c[j].z - y.c[j].z, 0.f); s += (S)d * (S)d;
        d = fmaxf(x.c[j].w - y.c[j].w, 0.f); s += (S)d * (S)d;
    }
    return s;
}

__device__ __forceinline__ float4 axpby4(float a, float4 x, float b, float4 y) {
    return make_float4(fmaf(a, x.x, b * y.x), fmaf(a, x.y, b * y.y), fmaf(a, x.z, b * y.z), fmaf(a, x.w, b * y.w));
}
__device__ __forceinline__ void fma4(float4& acc, float a, float4 x) {
    acc.x = fmaf(a, x.x, acc.x); acc.y = fmaf(a, x.y, acc.y); acc.z = fmaf(a, x.z, acc.z); acc.w = fmaf(a, x.w, acc.w);
}
__device__ __forceinline__ float4 relu_diff4(float4 x, float4 y) {
    return make_float4(fmaxf(x.x - y.x, 0.f), fmaxf(x.y - y.y, 0.f), fmaxf(x.z - y.z, 0.f), fmaxf(x.w - y.w, 0.f));
}

// clamp(v, min=0) with torch semantics (NaN propagates; fmaxf would swallow it)
template <typename S>
__device__ __forceinline__ S relu_nan(S v) { return v < (S)0 ? (S)0 : v; }

// ------------------------------------------------------------------------------------------------
// Per-pair scalar cores.  Input: team-reduced dot products.  Output: raw z (energy before the outer
// max(0,.)), and the four coefficients of dz/dx = zxx*x + zxy*y, dz/dy = zyx*x + zyy*y.
// ------------------------------------------------------------------------------------------------
struct PairGrad {
    float z;
    float zxx, zxy, zyx, zyy;
};

template <typename S> __device__ __forceinline__ S s_sqrt(S v);
// a*b - c*d with both products rounded before the subtraction (no FMA contraction), so that the
// numerator of the hyperbolic angle is exactly 0 when x == y, as in the reference (0/0 -> NaN)
__device__ __forceinline__ float diff_of_products(float a, float b, float c, float d) {
    return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d));
}
__device__ __forceinline__ double diff_of_products(double a, double b, double c, double d) {
    return __dsub_rn(__dmul_rn(a, b), __dmul_rn(c, d));
}
template <> __device__ __forceinline__ float s_sqrt<float>(float v) { return sqrtf(v); }
template <> __device__ __forceinline__ double s_sqrt<double>(double v) { return sqrt(v); }

// Euclidean cone, cos-space (order_embeddings.py:954-969).  A=<x,x>, DD=<d,d>, XD=<x,d>, d=y-x.
template <typename S, bool GRAD>
__device__ __forceinline__ void euc_core(S A, S DD, S XD, float K, PairGrad& o) {
    const S a = s_sqrt<S>(A), b = s_sqrt<S>(DD);
    const S an = a > (S)kNormEps ? a : (S)kNormEps;
    const S bn = b > (S)kNormEps ? b : (S)kNormEps;
    const S inv_ab = (S)1 / (an * bn);
    const S c = XD * inv_ab;
    const S K2 = (S)K * (S)K;
    const S root = s_sqrt<S>((S)1 - K2 / A);
    o.z = (float)(root - c);
    if (GRAD) {
        const S inv_a2 = (S)1 / (an * an), inv_b2 = (S)1 / (bn * bn);
        const S kt = K2 / (A * A * root);
        const S m = -inv_ab - c * inv_b2;  // coefficient of x in dz/dy (and of y in dz/dx)
        o.zyx = (float)m;
        o.zyy = (float)(c * inv_b2);
        o.zxy = (float)m;
        o.zxx = (float)((S)2 * inv_ab + c * inv_a2 + c * inv_b2 + kt);
    }
}

// Poincare cone (order_embeddings_h.py:1097-1120).  A=<x,x>, B=<y,y>, P=<x,y>, S2=|x-y|^2.
template <typename S, bool GRAD>
__device__ __forceinline__ void hyp_core(S A, S B, S P, S S2, float K, PairGrad& o) {
    const S one = (S)1;
    const S lo = (S)(-1.0 + 1e-5), hi = (S)(1.0 - 1e-5);
    const S a = s_sqrt<S>(A);
    const S w2 = one + A * B - (S)2 * P;
    const S den = a * s_sqrt<S>(S2) * s_sqrt<S>(w2);
    const S inv_den = one / den;
    const S g = diff_of_products(P, one + A, A, one + B) * inv_den;
    const S h = (S)K * (one - A) / a;
    const S gc = g < lo ? lo : (g > hi ? hi : g);  // NaN falls through unchanged, like torch.clamp
    const S hc = h < lo ? lo : (h > hi ? hi : h);
    const S sg = s_sqrt<S>((one - gc) * (one + gc));
    const S sh = s_sqrt<S>((one - hc) * (one + hc));
    float theta, psi;
    if (sizeof(S) == 8) {
        // fp64 core: angle from (cos, sin) both known to double accuracy
        theta = atan2f((float)sg, (float)gc);
        psi = atan2f((float)hc, (float)sh);
    } else {
        theta = acosf((float)gc);
        psi = asinf((float)hc);
    }
    o.z = theta - psi;
    if (GRAD) {
        const S th = (g >= lo && g <= hi) ? -one / sg : (S)0;  // d acos(clamp(g))
        const S ps = (h >= lo && h <= hi) ? one / sh : (S)0;   // d asin(clamp(h))
        const S inv_w2 = one / w2;
        const S g_p = (one + A) * inv_den + g * inv_w2;
        const S g_A = (P - one - B) * inv_den - g * ((S)0.5 / A + (S)0.5 * B * inv_w2);
        const S g_B = -A * inv_den - g * (S)0.5 * A * inv_w2;
        const S gs_over_s = -g / S2;  // (dg/ds)/s
        const S h_A = -(S)K * (one + A) / ((S)2 * a * A);
        o.zxx = (float)(th * ((S)2 * g_A + gs_over_s) - ps * h_A * (S)2);
        o.zxy = (float)(th * (g_p - gs_over_s));
        o.zyx = (float)(th * (g_p - gs_over_s));
        o.zyy = (float)(th * ((S)2 * g_B + gs_over_s));
    }
}

// hinge: returns d loss / d z and adds this pair's loss term
//   positive: w*E            -> w*[z>=0]
//   negative: w*max(0,a-E)   -> -w*[a-E>=0]*[z>=0]
__device__ __forceinline__ float hinge(float z, bool is_pos, float w, float alpha, float& E, double& loss) {
    E = relu_nan(z);
    const float act = (z >= 0.f) ? 1.f : 0.f;
    if (is_pos) {
        loss += (double)(w * E);
        return w * act;
    }
    const float m = alpha - E;
    loss += (double)(w * relu_nan(m));
    return (m >= 0.f) ? -w * act : 0.f;
}

template <int IDXB>
__device__ __forceinline__ int64_t load_idx(const void* p, int64_t i) {
    if (IDXB == 4) return (int64_t) __ldg(reinterpret_cast<const int32_t*>(p) + i);
    return (int64_t) __ldg(reinterpret_cast<const long long*>(p) + i);
}

// block-level double sum -> one atomic per block
__device__ __forceinline__ void block_add_double(double v, double* out) {
    __shared__ double warp_part[kThreads / 32];
    v = warp_sum<double>(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_part[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = (lane < (blockDim.x >> 5)) ? warp_part[lane] : 0.0;
        t = warp_sum<double>(t);
        if (lane == 0 && out != nullptr && t != 0.0) atomicAdd(out, t);
    }
}

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace lec
