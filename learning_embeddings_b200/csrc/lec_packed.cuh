// Packed fp32x2 arithmetic (one FFMA2 / FMUL2 / FADD2 issue slot does two fp32 operations) and the clamped
// acos used by the scoring kernels (lec_score_fast_impl.cuh, lec_score_mma.cu).
#pragma once
#include "lec_common.cuh"

namespace lec {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float rsqrt_approx(float v) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
// min / max that propagate NaN like torch.clamp
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float min_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// acos(|g|-part) coefficients, Abramowitz & Stegun 4.4.46: acos(x) = sqrt(1-x) * sum a_i x^i, |err| <= 2e-8 on [0,1]
#define LEC_ACOS_A0 1.5707963050f
#define LEC_ACOS_A1 -0.2145988016f
#define LEC_ACOS_A2 0.0889789874f
#define LEC_ACOS_A3 -0.0501743046f
#define LEC_ACOS_A4 0.0308918810f
#define LEC_ACOS_A5 -0.0170881256f
#define LEC_ACOS_A6 0.0066700901f
#define LEC_ACOS_A7 -0.0012624911f

// theta = acos(clamp(g, -1+1e-5, 1-1e-5)) for a packed pair; NaN propagates
__device__ __forceinline__ u64 acos_clamped2(u64 g2) {
    float g0, g1;
    unpack2(g2, g0, g1);
    // clamp(g, -1+eps, 1-eps) only matters through |g| (the sign is taken from g's sign bit below), so one
    // NaN-propagating min on |g| replaces the max/min pair
    const float hi = 1.f - kClampEps;
    const float a0 = min_nan(fabsf(g0), hi), a1 = min_nan(fabsf(g1), hi);
    const u64 ax = pack2(a0, a1);
    const u64 t = ffma2(ax, pack2(-1.f, -1.f), pack2(1.f, 1.f));  // 1 - |g| (exact for |g| >= 0.5)
    float t0, t1;
    unpack2(t, t0, t1);
    const u64 sq = pack2(sqrt_approx(t0), sqrt_approx(t1));
    u64 P = pack2(LEC_ACOS_A7, LEC_ACOS_A7);
    P = ffma2(P, ax, pack2(LEC_ACOS_A6, LEC_ACOS_A6));
    P = ffma2(P, ax, pack2(LEC_ACOS_A5, LEC_ACOS_A5));
    P = ffma2(P, ax, pack2(LEC_ACOS_A4, LEC_ACOS_A4));
    P = ffma2(P, ax, pack2(LEC_ACOS_A3, LEC_ACOS_A3));
    P = ffma2(P, ax, pack2(LEC_ACOS_A2, LEC_ACOS_A2));
    P = ffma2(P, ax, pack2(LEC_ACOS_A1, LEC_ACOS_A1));
    P = ffma2(P, ax, pack2(LEC_ACOS_A0, LEC_ACOS_A0));
    const u64 r = fmul2(sq, P);
    // g < 0: pi - r ; else r.   sg = copysign(1, g) (one LOP3 each); offset = pi/2 - sg * pi/2 is exactly 0 or
    // fp32(pi) and comes from the FMA pipe as a packed op instead of two more ALU ops per score
    const float sg0 = __int_as_float((__float_as_int(g0) & 0x80000000) | 0x3f800000);
    const float sg1 = __int_as_float((__float_as_int(g1) & 0x80000000) | 0x3f800000);
    const u64 sg = pack2(sg0, sg1);
    const u64 HP = pack2(1.57079637f, 1.57079637f);
    const u64 off = ffma2(sg, pack2(-1.57079637f, -1.57079637f), HP);
    return ffma2(sg, r, off);
}

// z = acos(clamp(g, -1+1e-5, 1-1e-5)) + c for a packed pair, with c = pi/2 - psi folded into one constant
// (hpc = {pi/2 - psi, pi/2 - psi'}): with r = acos(|g|) = sqrt(1-|g|) P(|g|) and h = pi/2 - r > 0,
//     acos(g) = pi/2 - sign(g) h      =>      z = (pi/2 - psi) + (h with the sign bit of -g)
// -- one FFMA2 for h, one LOP3 per score for the sign, one FADD2; the off / sign-multiply form of acos_clamped2
// above spends two more packed operations per pair.  |h - (pi/2 - r)| <= 6e-8, so z carries ~1.2e-7 absolute
// rounding (the reference's own fp32 acos is no better).  NaN propagates.
__device__ __forceinline__ u64 acos_clamped_plus2(u64 g2, u64 hpc) {
    float g0, g1;
    unpack2(g2, g0, g1);
    const float hi = 1.f - kClampEps;
    const float a0 = min_nan(fabsf(g0), hi), a1 = min_nan(fabsf(g1), hi);
    const u64 ax = pack2(a0, a1);
    const u64 t = ffma2(ax, pack2(-1.f, -1.f), pack2(1.f, 1.f));
    float t0, t1;
    unpack2(t, t0, t1);
    const u64 sq = pack2(sqrt_approx(t0), sqrt_approx(t1));
    u64 P = pack2(-(LEC_ACOS_A7), -(LEC_ACOS_A7));   // -P(|g|): the coefficients carry the minus sign of h = pi/2 - sqrt(1-|g|) P(|g|)
    P = ffma2(P, ax, pack2(-(LEC_ACOS_A6), -(LEC_ACOS_A6)));
    P = ffma2(P, ax, pack2(-(LEC_ACOS_A5), -(LEC_ACOS_A5)));
    P = ffma2(P, ax, pack2(-(LEC_ACOS_A4), -(LEC_ACOS_A4)));
    P = ffma2(P, ax, pack2(-(LEC_ACOS_A3), -(LEC_ACOS_A3)));
    P = ffma2(P, ax, pack2(-(LEC_ACOS_A2), -(LEC_ACOS_A2)));
    P = ffma2(P, ax, pack2(-(LEC_ACOS_A1), -(LEC_ACOS_A1)));
    P = ffma2(P, ax, pack2(-(LEC_ACOS_A0), -(LEC_ACOS_A0)));
    float h0, h1;
    unpack2(ffma2(sq, P, pack2(1.57079637f, 1.57079637f)), h0, h1);   // h = pi/2 - r >= 6e-8: sign bit clear (or NaN)
    // m = h carrying the sign of -g:  h | (~g & 0x80000000)  (one LOP3)
    const float m0 = __int_as_float(__float_as_int(h0) | (~__float_as_int(g0) & 0x80000000));
    const float m1 = __int_as_float(__float_as_int(h1) | (~__float_as_int(g1) & 0x80000000));
    return fadd2(hpc, pack2(m0, m1));
}

}  // namespace lec
