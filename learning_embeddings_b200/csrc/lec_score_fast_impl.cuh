// Fast all-pairs image x label scoring (fp32 core) with per-level top-k: the bandwidth/issue-bound tiled
// kernel of the scoring path (oe.py:1764-1779, oe_h.py:2018-2036 replaced).
//
// Work unit = (label segment, image group).  A segment is one level's label range [start, stop) (its
// top-k is private to the unit) or, when the full matrix is wanted, a gap between levels.  Units are
// ordered largest segment first so the small levels fill the tail of the last wave.
//
// Thread mapping: a thread owns RI images (rows in registers, packed two images per 64-bit register
// pair) and walks the segment's labels, which are staged tile by tile in shared memory with every value
// duplicated ({x_d, x_d}) so that one broadcast LDS.128 feeds two packed FFMA2 (fma.rn.f32x2: two fp32
// FMAs per issue slot; the B200 FMA pipe does 128 FMA/clk/SM either way, the packing frees issue slots
// for the MUFU/ALU/LSU work of the epilogue).  Everything that depends on the label only (|x|^2, 1+|x|^2,
// the half-aperture) is computed once per tile in fp64 and staged next to the row.
//
// Hyperbolic energy per pair, from p = <x,y>, A = |x|^2, B = |y|^2 (order_embeddings_h.py:1097-1120):
//   num = p(1+A) - A(1+B);  s2 = A + B - 2p;  w2 = 1 + AB - 2p;  g = num * rsqrt(A s2 w2)
//   E = max(0, acos(clamp(g)) - psi),  acos(x) = sqrt(1-|x|) P7(|x|), reflected for x < 0
// (9 packed FMA-pipe ops + 2 MUFU + an 8-term polynomial; checked against the reference's fp64 run in
// tests/test_gpu_parity.py -- its error is below the reference's own fp32 error on the golden vectors).
//
// Top-k: a thread cannot afford a sorted insertion per score (lanes diverge), so candidates that beat the
// current k-th best are appended, predicated and branch-free, to a small per-thread ring in shared
// memory; when any lane's ring is nearly full the whole warp merges its rings into the per-image sorted
// top-k lists (also in shared memory) and refreshes the thresholds.  NaN never passes `E < thr`.
//
// Deferred angle (hyperbolic, top-k only -- what the reference's caller consumes): acos is monotone, so
// E < thr  <=>  g > cos(thr + psi) = cos(thr) cos(psi) - sin(thr) sin(psi)  for thr + psi <= pi.  The main
// loop therefore tests g against that bound (2 packed FMAs, minus a 4e-6 slack so rounding can only let
// extra candidates through) and appends {g, label, -psi} to the ring; the acos polynomial runs only for
// ring entries inside merge(), with the same instruction sequence as the full-matrix path, so top-k
// values are bit-identical to the matrix entries.  thr > pi/2 or an unfilled list accepts everything.
#pragma once
#include <cstdio>
#include <cstdlib>

#include "lec_common.cuh"
#include "lec_packed.cuh"

namespace lec {

constexpr int kMaxSeg = 2 * LEC_MAX_LEVELS + 1;
constexpr int kRingDefault = 16;  // candidate ring entries per thread
constexpr int kTileLabels = 96;  // labels per shared-memory tile

struct FastArgs {
    const float* labels; const float* images; int64_t L, N; int D; float K;
    int k, n_levels;
    int n_seg;
    int seg_start[kMaxSeg], seg_stop[kMaxSeg], seg_level[kMaxSeg];
    float* scores; int64_t s_img, s_lab;  // element strides of the score matrix
    int32_t* topk_idx; float* topk_val;
    int64_t groups;  // image groups of NT*RI images
    int ring;        // candidate ring entries per thread (> RI)
};

// Shared-memory layout of one staged label: DQ float4 chunks of the row (zero padded), then one float4
// of constants.   hyp: {A, 1+A, A^2, -psi} {cos psi, sin psi, 0, 0}   euc: {A, t0, 0, 0}   oe: unused
// Label values enter the packed FMAs as scalar-broadcast operands (FFMA2 takes a .F32 operand that
// feeds both lanes), so nothing is duplicated in shared memory and an instruction reads two register
// pairs plus one scalar: the B200 register file sustains two pair reads per clock, three distinct pair
// operands cost three clocks (scripts/ubench_pipes.cu).
template <int V>
struct IntC {
    static constexpr int value = V;
};

template <int GEOM, int DH, int RI, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) score_fast_kernel(const FastArgs a) {
    static_assert(RI % 2 == 0, "images are packed in pairs");
    constexpr int DQ = (DH + 1) / 2;     // float4 chunks per row
    constexpr int LS = 4 * DQ + 8;       // floats per staged label: row chunks + two float4 of constants
    constexpr int NP = RI / 2;           // packed image pairs per thread
    extern __shared__ __align__(16) float smem[];
    float* lab = smem;                                      // [kTileLabels][LS]
    float2* top = reinterpret_cast<float2*>(smem + kTileLabels * LS);  // [RI*k][NT]   {E, idx}
    float2* ring = top + (size_t)RI * a.k * NT;             // [a.ring][NT]  {E or g, (label << 3) | image slot}
    float* ringp = reinterpret_cast<float*>(ring + (size_t)a.ring * NT);  // [a.ring][NT]  -psi of a deferred entry
    float* psi_max_s = ringp + (size_t)a.ring * NT;         // largest half-aperture of the staged tile

    const int tid = threadIdx.x;
    const int seg = (int)(blockIdx.x / a.groups);
    const int64_t group = blockIdx.x % a.groups;
    const int l_begin = a.seg_start[seg], l_end = a.seg_stop[seg], level = a.seg_level[seg];
    const bool want_topk = (level >= 0) && (a.topk_idx != nullptr);
    const int k = a.k;
    const int D = a.D;

    // ---- image rows -> registers (two images per packed pair), per-image constants
    const int64_t img0 = group * (int64_t)(NT * RI) + tid;  // image of slot r: img0 + r*NT
    u64 y[NP][4 * DQ];
    u64 B2[NP], C2[NP];  // hyp: B, -(1+B) ; euc: B, unused
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        const int64_t i0 = img0 + (2 * j) * NT, i1 = img0 + (2 * j + 1) * NT;
        const float* s0 = a.images + (i0 < a.N ? i0 : 0) * (int64_t)D;
        const float* s1 = a.images + (i1 < a.N ? i1 : 0) * (int64_t)D;
        float b0 = 0.f, b1 = 0.f;
#pragma unroll
        for (int d = 0; d < 4 * DQ; ++d) {
            const float v0 = (d < D) ? __ldg(s0 + d) : 0.f;
            const float v1 = (d < D) ? __ldg(s1 + d) : 0.f;
            b0 = fmaf(v0, v0, b0);
            b1 = fmaf(v1, v1, b1);
            y[j][d] = (GEOM == LEC_GEOM_OE) ? pack2(-v0, -v1) : pack2(v0, v1);
        }
        B2[j] = pack2(b0, b1);
        C2[j] = pack2(-1.f - b0, -1.f - b1);
    }

    // ---- top-k state: thresholds in registers, sorted lists and the candidate ring in shared memory
    float thr[RI];
#pragma unroll
    for (int r = 0; r < RI; ++r) thr[r] = INFINITY;
    const unsigned ring0 = (unsigned)__cvta_generic_to_shared(ring + tid);  // entry j of this thread: ring0 + j*NT*8
    unsigned rp = ring0;                                                     // next free entry
    const unsigned ring_trigger = ring0 + (unsigned)(a.ring - RI) * NT * 8;   // merge when rp > trigger
    if (want_topk) {
        for (int s = 0; s < RI * k; ++s) top[s * NT + tid] = make_float2(INFINITY, __int_as_float(-1));
    }
    const bool full = (group + 1) * (int64_t)(NT * RI) <= a.N;  // block-uniform: every image of the group exists

    auto merge = [&]() {
        // insert this thread's ring entries into its sorted per-image lists (stable: ties keep the lower label)
        const int cnt = (int)((rp - ring0) / (NT * 8));
        for (int j = 0; j < cnt; ++j) {
            const float2 e = ring[j * NT + tid];
            const int code = __float_as_int(e.y);
            const int r = code & 7;
            float2* t = top + (size_t)(r * k) * NT + tid;
            if (e.x < t[(k - 1) * NT].x) {
                int pos = k - 1;
                while (pos > 0) {
                    const float2 prev = t[(pos - 1) * NT];
                    if (!(prev.x > e.x)) break;
                    t[pos * NT] = prev;
                    --pos;
                }
                t[pos * NT] = make_float2(e.x, __int_as_float(code >> 3));
            }
        }
        rp = ring0;
#pragma unroll
        for (int r = 0; r < RI; ++r) thr[r] = top[(size_t)(r * k + k - 1) * NT + tid].x;
    };

    // <x, y> of one staged label against this thread's packed image pairs.  Few pairs per thread (large D)
    // means few independent FMA chains, so the sum is split over NA partial accumulators per pair.
    auto dots = [&](const float4* lp, u64 (&acc)[NP]) {
        constexpr int NA = NP >= 4 ? 1 : (NP == 2 ? 2 : 4);
        u64 part[NP][NA];
#pragma unroll
        for (int j = 0; j < NP; ++j)
#pragma unroll
            for (int c = 0; c < NA; ++c) part[j][c] = 0ull;
#pragma unroll
        for (int q = 0; q < DQ; ++q) {
            const float4 v = lp[q];
            const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (4 * q + c >= 2 * DH) continue;
                const u64 xb = pack2(xv[c], xv[c]);
#pragma unroll
                for (int j = 0; j < NP; ++j) part[j][c % NA] = ffma2(y[j][4 * q + c], xb, part[j][c % NA]);
            }
        }
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            u64 t = part[j][0];
            if (NA == 2) t = fadd2(t, part[j][1]);
            if (NA == 4) t = fadd2(fadd2(t, part[j][1]), fadd2(part[j][2 % NA], part[j][3 % NA]));
            acc[j] = t;
        }
    };
    // cos of the cone angle for a packed pair: g = (p(1+A) - A(1+B)) * rsqrt(A (A+B-2p) (1+AB-2p))
    auto hyp_g = [&](u64 P, int j, const float4& cst) -> u64 {
        const u64 A2 = pack2(cst.x, cst.x), A12 = pack2(cst.y, cst.y), ASQ = pack2(cst.z, cst.z);
        const u64 M2 = pack2(-2.f, -2.f), ONE2 = pack2(1.f, 1.f);
        const u64 q = fmul2(P, M2);
        const u64 num = ffma2(P, A12, fmul2(A2, C2[j]));          // p(1+A) - A(1+B)
        const u64 w2 = fadd2(q, ffma2(A2, B2[j], ONE2));          // 1 + AB - 2p
        const u64 as2 = ffma2(A2, q, ffma2(A2, B2[j], ASQ));      // A (A + B - 2p)
        const u64 d2 = fmul2(as2, w2);
        float d0, d1;
        unpack2(d2, d0, d1);
        return fmul2(num, pack2(rsqrt_approx(d0), rsqrt_approx(d1)));
    };

    // ---- deferred-angle filter state (hyperbolic, top-k only): accept g >= cT cos(psi) - sT sin(psi) + off
    u64 cT2[NP], nsT2[NP], off2[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) { cT2[j] = 0ull; nsT2[j] = 0ull; off2[j] = pack2(-INFINITY, -INFINITY); }
    const unsigned ringp0 = (unsigned)__cvta_generic_to_shared(ringp + tid);
    // psi_max = largest half-aperture of the staged tile: the bound is only valid while thr + psi <= pi
    auto filter_terms = [&](float t, float psi_max, float& c, float& ns, float& off) {
        if (t <= 0.f) { c = 0.f; ns = 0.f; off = INFINITY; }                         // k zeros already: nothing can beat them
        else if (!(t + psi_max <= 3.1415f)) { c = 0.f; ns = 0.f; off = -INFINITY; }  // list not full / thr + psi may pass pi
        else { float sn, cs; __sincosf(t, &sn, &cs); c = cs; ns = -sn; off = -4e-6f; }  // MUFU sin/cos: |err| < 2e-6 on [0, pi]
    };
    auto refresh_filter = [&]() {
        const float psi_max = *psi_max_s;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            float c0, c1, n0, n1, o0, o1;
            filter_terms(top[(size_t)((2 * j) * k + k - 1) * NT + tid].x, psi_max, c0, n0, o0);
            filter_terms(top[(size_t)((2 * j + 1) * k + k - 1) * NT + tid].x, psi_max, c1, n1, o1);
            cT2[j] = pack2(c0, c1); nsT2[j] = pack2(n0, n1); off2[j] = pack2(o0, o1);
        }
    };
    auto merge_deferred = [&]() {
        const int cnt = (int)((rp - ring0) / (NT * 8));
        for (int j = 0; j < cnt; ++j) {
            const float2 e = ring[j * NT + tid];
            const float np = ringp[j * NT + tid];
            const int code = __float_as_int(e.y);
            const int r = code & 7;
            float z0, z1;
            unpack2(fadd2(acos_clamped2(pack2(e.x, e.x)), pack2(np, np)), z0, z1);
            const float Ev = max_nan(z0, 0.f);
            float2* t = top + (size_t)(r * k) * NT + tid;
            if (Ev < t[(k - 1) * NT].x) {
                int pos = k - 1;
                while (pos > 0) {
                    const float2 prev = t[(pos - 1) * NT];
                    if (!(prev.x > Ev)) break;
                    t[pos * NT] = prev;
                    --pos;
                }
                t[pos * NT] = make_float2(Ev, __int_as_float(code >> 3));
            }
        }
        rp = ring0;
        refresh_filter();
    };
    auto tile_loop_deferred = [&](int l0, int tl) {
        int lcode = l0 << 3;
#pragma unroll 2
        for (int t = 0; t < tl; ++t) {
            const float4* lp = reinterpret_cast<const float4*>(lab + t * LS);
            u64 acc[NP];
            dots(lp, acc);
            const float4 cst = lp[DQ];
            const float4 cs2 = lp[DQ + 1];
            const u64 CP = pack2(cs2.x, cs2.x), SP = pack2(cs2.y, cs2.y);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const u64 g = hyp_g(acc[j], j, cst);
                const u64 cb = ffma2(cT2[j], CP, ffma2(nsT2[j], SP, off2[j]));
                float g0, g1, b0, b1;
                unpack2(g, g0, g1);
                unpack2(cb, b0, b1);
                if (g0 >= b0) {
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rp), "f"(g0), "r"(lcode + 2 * j) : "memory");
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(ringp0 + ((rp - ring0) >> 1)), "f"(cst.w) : "memory");
                    rp += NT * 8;
                }
                if (g1 >= b1) {
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rp), "f"(g1), "r"(lcode + 2 * j + 1) : "memory");
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(ringp0 + ((rp - ring0) >> 1)), "f"(cst.w) : "memory");
                    rp += NT * 8;
                }
            }
            lcode += 8;
            if (__any_sync(0xffffffffu, rp > ring_trigger)) merge_deferred();
        }
    };

    // energies of one staged label against this thread's RI images
    auto energies = [&](const float* lrow, float (&E)[RI]) {
        const float4* lp = reinterpret_cast<const float4*>(lrow);
        u64 acc[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) acc[j] = 0ull;  // {0.f, 0.f}
        if (GEOM == LEC_GEOM_OE) {
#pragma unroll
            for (int q = 0; q < DQ; ++q) {
                const float4 v = lp[q];
                const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (4 * q + c >= 2 * DH) continue;
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        float d0, d1;
                        unpack2(fadd2(pack2(xv[c], xv[c]), y[j][4 * q + c]), d0, d1);  // x_d - y_d for the two images
                        const u64 dd = pack2(fmaxf(d0, 0.f), fmaxf(d1, 0.f));
                        acc[j] = ffma2(dd, dd, acc[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NP; ++j) unpack2(acc[j], E[2 * j], E[2 * j + 1]);
            return;
        }
        dots(lp, acc);
        const float4 cst = lp[DQ];
        if (GEOM == LEC_GEOM_HYP) {
            // cst = {A, 1+A, A^2, -psi};  q = -2p;  w2 = q + (AB + 1);  A*s2 = A q + (AB + A^2)
            const u64 npsi = pack2(cst.w, cst.w);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const u64 g = hyp_g(acc[j], j, cst);
                const u64 z = fadd2(acos_clamped2(g), npsi);
                float z0, z1;
                unpack2(z, z0, z1);
                E[2 * j] = max_nan(z0, 0.f);
                E[2 * j + 1] = max_nan(z1, 0.f);
            }
        } else {
            const float Af = cst.x, t0 = cst.y;
            const u64 A2 = pack2(Af, Af);
            const u64 M2 = pack2(-2.f, -2.f);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const u64 P = acc[j];
                const u64 s2 = ffma2(M2, P, fadd2(A2, B2[j]));            // |y - x|^2 = A + B - 2p
                float s0, s1, p0, p1;
                unpack2(s2, s0, s1);
                unpack2(P, p0, p1);
                // <x, y-x> / (max(|x|,eps) max(|y-x|,eps)); a vanishing |y-x| gives 0 like F.normalize
                const float q0 = Af * fmaxf(s0, 0.f), q1 = Af * fmaxf(s1, 0.f);
                const float c_0 = q0 > 0.f ? (p0 - Af) * rsqrt_approx(q0) : 0.f;
                const float c_1 = q1 > 0.f ? (p1 - Af) * rsqrt_approx(q1) : 0.f;
                E[2 * j] = max_nan(t0 - c_0, 0.f);
                E[2 * j + 1] = max_nan(t0 - c_1, 0.f);
            }
        }
    };

    // STORE: 0 no matrix, 1 label-major and every image of the group exists, 2 generic strides / partial group
    auto tile_loop = [&](int l0, int tl, auto store_c, auto topk_c) {
        constexpr int STORE = decltype(store_c)::value;
        constexpr bool TOPK = decltype(topk_c)::value != 0;
        float* dst = nullptr;
        const int64_t s_lab = a.s_lab;
        if (STORE == 1) dst = a.scores + (int64_t)l0 * s_lab + img0;
        int lcode = l0 << 3;
#pragma unroll 2
        for (int t = 0; t < tl; ++t) {
            float E[RI];
            energies(lab + t * LS, E);
            if (STORE == 1) {
#pragma unroll
                for (int r = 0; r < RI; ++r) dst[r * NT] = E[r];
                dst += s_lab;
            } else if (STORE == 2) {
                float* row = a.scores + (int64_t)(l0 + t) * s_lab;
#pragma unroll
                for (int r = 0; r < RI; ++r) {
                    const int64_t i = img0 + r * NT;
                    if (i < a.N) row[i * a.s_img] = E[r];
                }
            }
            if (TOPK) {
#pragma unroll
                for (int r = 0; r < RI; ++r) {
                    if (E[r] < thr[r]) {
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rp), "f"(E[r]), "r"(lcode + r) : "memory");
                        rp += NT * 8;
                    }
                }
                lcode += 8;
                if (__any_sync(0xffffffffu, rp > ring_trigger)) merge();
            }
        }
    };

    const int store = (a.scores == nullptr) ? 0 : ((a.s_img == 1 && full) ? 1 : 2);
    for (int l0 = l_begin; l0 < l_end; l0 += kTileLabels) {
        const int tl = min(kTileLabels, l_end - l0);
        __syncthreads();
        // ---- stage the tile: rows zero-padded to whole float4 chunks
        for (int i = tid; i < tl * 4 * DQ; i += NT) {
            const int t = i / (4 * DQ), d = i - t * (4 * DQ);
            lab[t * LS + d] = (d < D) ? __ldg(a.labels + (int64_t)(l0 + t) * D + d) : 0.f;
        }
        // ---- per-label constants in fp64 (same terms as lec_rows_fwd's aux)
        if (GEOM != LEC_GEOM_OE) {
            for (int t = tid; t < tl; t += NT) {
                const float* src = a.labels + (int64_t)(l0 + t) * D;
                double A = 0.0;
                for (int d = 0; d < D; ++d) { const double v = (double)__ldg(src + d); A += v * v; }
                const Aux<double> x = row_aux<double>(GEOM, A, a.K);
                float4 c, c2 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (GEOM == LEC_GEOM_HYP) {
                    c = make_float4((float)A, (float)(1.0 + A), (float)(A * A), (float)(-x.t0));
                    const double sp = sin(x.t0);  // = the clamped asin argument
                    c2 = make_float4((float)sqrt(fmax(0.0, 1.0 - sp * sp)), (float)sp, 0.f, 0.f);
                } else {
                    c = make_float4((float)A, (float)x.t0, 0.f, 0.f);
                }
                *reinterpret_cast<float4*>(lab + t * LS + 4 * DQ) = c;
                *reinterpret_cast<float4*>(lab + t * LS + 4 * DQ + 4) = c2;
            }
        }
        __syncthreads();
        if (GEOM == LEC_GEOM_HYP && want_topk && store == 0) {
            if (tid < 32) {
                float mx = -INFINITY;
                for (int t = tid; t < tl; t += 32) mx = fmaxf(mx, -lab[t * LS + 4 * DQ + 3]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                if (tid == 0) *psi_max_s = mx;
            }
            __syncthreads();
            refresh_filter();
        }
        if (want_topk) {
            if (store == 0 && GEOM == LEC_GEOM_HYP) tile_loop_deferred(l0, tl);
            else if (store == 0) tile_loop(l0, tl, IntC<0>(), IntC<1>());
            else if (store == 1) tile_loop(l0, tl, IntC<1>(), IntC<1>());
            else tile_loop(l0, tl, IntC<2>(), IntC<1>());
        } else {
            if (store == 1) tile_loop(l0, tl, IntC<1>(), IntC<0>());
            else if (store == 2) tile_loop(l0, tl, IntC<2>(), IntC<0>());
        }
    }

    if (want_topk) {
        if (store == 0 && GEOM == LEC_GEOM_HYP) merge_deferred();
        else merge();
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            const int64_t i = img0 + r * NT;
            if (i >= a.N) continue;
            const int64_t o = (i * a.n_levels + level) * k;
            for (int j = 0; j < k; ++j) {
                const float2 e = top[(size_t)(r * k + j) * NT + tid];
                a.topk_idx[o + j] = __float_as_int(e.y);
                if (a.topk_val) a.topk_val[o + j] = e.x;
            }
        }
    }
}

// experiment knob: LEC_SCORE_CFG="<images per thread>,<threads>,<min blocks>,<ring entries>" picks another instantiation
static void tuning(int& ri, int& nt, int& minb, int& ring) {
    static int cfg[4] = {0, 0, 0, 0};
    static bool read = false;
    if (!read) {
        read = true;
        const char* e = getenv("LEC_SCORE_CFG");
        if (e) sscanf(e, "%d,%d,%d,%d", &cfg[0], &cfg[1], &cfg[2], &cfg[3]);
    }
    if (cfg[0]) ri = cfg[0];
    if (cfg[1]) nt = cfg[1];
    if (cfg[2]) minb = cfg[2];
    if (cfg[3]) ring = cfg[3];
}

template <int GEOM, int DH, int RI, int NT, int MINB>
static int fast_launch_cfg(FastArgs& a, cudaStream_t st) {
    constexpr int LS = 4 * ((DH + 1) / 2) + 8;
    const size_t smem = (size_t)kTileLabels * LS * sizeof(float) + ((size_t)RI * a.k + a.ring) * NT * sizeof(float2) +
                        (size_t)a.ring * NT * sizeof(float) + 16;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(score_fast_kernel<GEOM, DH, RI, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    a.groups = (a.N + (int64_t)NT * RI - 1) / ((int64_t)NT * RI);
    const int64_t blocks = a.groups * a.n_seg;
    if (blocks <= 0) return 0;
    if (blocks > 0x7fffffffLL) return LEC_E_SIZE;
    score_fast_kernel<GEOM, DH, RI, NT, MINB><<<(unsigned)blocks, NT, smem, st>>>(a);
    ++g_launches;
    return (int)cudaGetLastError();
}

template <int GEOM, int DH, int RI_DEFAULT>
static int fast_launch_dh(FastArgs& a, cudaStream_t st) {
    int ri = RI_DEFAULT, nt = 128, minb = 1, ring = kRingDefault;
    tuning(ri, nt, minb, ring);
    a.ring = ring < 2 * ri ? 2 * ri : ring;
#ifdef LEC_SCORE_TUNING_VARIANTS
    if (ri == 4 && nt == 256) return fast_launch_cfg<GEOM, DH, 4, 256, 1>(a, st);
    if (ri == 4 && nt == 128 && minb == 4) return fast_launch_cfg<GEOM, DH, 4, 128, 4>(a, st);
    if (ri == 2 && nt == 256 && minb == 4) return fast_launch_cfg<GEOM, DH, 2, 256, 4>(a, st);
    if (ri == 2 && nt == 128) return fast_launch_cfg<GEOM, DH, 2, 128, 1>(a, st);
    if (ri == 6 && nt == 128) return fast_launch_cfg<GEOM, DH, 6, 128, 1>(a, st);
    if (ri == 8 && nt == 128) return fast_launch_cfg<GEOM, DH, 8, 128, 1>(a, st);
#endif
    return fast_launch_cfg<GEOM, DH, RI_DEFAULT, 128, 1>(a, st);
}

template <int GEOM>
static int fast_launch_geom(FastArgs& a, cudaStream_t st) {
    const int dh = (a.D + 1) / 2;
    if (dh <= 1) return fast_launch_dh<GEOM, 1, 4>(a, st);
    if (dh <= 2) return fast_launch_dh<GEOM, 2, 4>(a, st);
    if (dh <= 5) return fast_launch_dh<GEOM, 5, 4>(a, st);
    if (dh <= 8) return fast_launch_dh<GEOM, 8, 4>(a, st);
    if (dh <= 16) return fast_launch_dh<GEOM, 16, 2>(a, st);
    if (dh <= 25) return fast_launch_dh<GEOM, 25, 2>(a, st);
    if (dh <= 32) return fast_launch_dh<GEOM, 32, 2>(a, st);
    return LEC_E_DIM;
}

}  // namespace lec
