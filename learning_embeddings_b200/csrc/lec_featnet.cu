// FeatNet.fc1 of the joint image+label trainers (oe.py:97,113; oe_h.py:127,143) on the step's GATHERED feature rows,
// forward and weight gradient, each as ONE pass over the gathered rows.
//
// The reference looks every image's 2048-d feature vector up in a Python dict (get_img_features, oe.py:680-707), builds a
// [B, 2048] tensor and runs nn.Linear(2048, D); autograd then re-reads the same rows for the weight gradient.  D is 10,
// so both GEMMs are thin and bound by the bytes of X: here the gather is fused into them (X is read straight from the
// device-resident feature matrix through the row index, never materialised), the weight lives in shared memory
// (forward) or the partial dW in registers (backward), and X streams through once per pass:
//   forward   Y[i, :] = W X[sel[i], :] + b                 one warp per 4 rows, W [D, F] in shared memory
//   wgrad     dW[d, f] += sum_i gY[i, d] X[sel[i], f]      one block per row range, a thread owns 4 columns x D outputs
//             db[d]    += sum_i gY[i, d]                   in registers; flushed with vector reductions into one of R
//                                                          replicas of the flat [D*F + D] gradient (lec_update_rows sums them)
// Algorithmic bytes: 4 F per gathered row per pass (+ 4 D out / in): cfg2 = 16 384 rows x 8 KB = 134 MB per pass.
#include "lec_featnet.cuh"
#include "lec_packed.cuh"

namespace lec {

// feature rows are read exactly once per pass: keep them out of L1
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// row of the feature matrix selected by entry i (NULL sel = identity); an id outside the pool reads nothing and is counted
__device__ __forceinline__ const float* feat_row(const FeatArgs& a, int64_t i, bool count) {
    int64_t ix = i;
    if (a.sel) ix = a.sel_bytes == 4 ? (int64_t)__ldg(reinterpret_cast<const int32_t*>(a.sel) + i)
                                     : (int64_t)__ldg(reinterpret_cast<const long long*>(a.sel) + i);
    if ((uint64_t)ix >= (uint64_t)a.n_pool) {
        if (count && a.index_errors) atomicAdd(a.index_errors, 1u);
        return nullptr;
    }
    return a.features + ix * (int64_t)a.F;
}

// One butterfly step that also halves the data: the lane whose bit `o` is clear keeps (and completes) the first half of
// the values, its partner the second half -- N/2 shuffles instead of N.
template <int N>
__device__ __forceinline__ void halve(const float (&in)[N], float (&out)[N / 2], bool upper, int o) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float send = upper ? in[i] : in[i + N / 2];
        const float keep = upper ? in[i + N / 2] : in[i];
        out[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
}

// NP output pairs (D <= 2 NP); a warp owns RW = 4 rows at a time.  The weight sits in shared memory with the two outputs
// of a pair interleaved, so one 128-bit load yields two packed operands and every packed FMA (fma.rn.f32x2) advances
// two outputs of one row: 16 NP packed FMAs + 2 NP shared loads per 4 rows x 128 columns.
template <int NP>
__global__ void __launch_bounds__(kThreads) featnet_fwd_kernel(const FeatArgs a) {
    constexpr int RW = 4;
    extern __shared__ __align__(16) float s_w[];   // float4 [(2 p + h) * C + c]: h = 0 columns 4c, 4c+1; h = 1 columns 4c+2, 4c+3
    pdl_launch_dependents();
    pdl_wait();
    const int C = a.F >> 2;                          // float4 chunks per row
    float4* s_w4 = reinterpret_cast<float4*>(s_w);
    for (int i = threadIdx.x; i < NP * C; i += kThreads) {
        const int p = i / C, c = i - p * C;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 w0 = 2 * p < a.D ? __ldg(reinterpret_cast<const float4*>(a.weight + (int64_t)(2 * p) * a.F) + c) : z;
        const float4 w1 = 2 * p + 1 < a.D ? __ldg(reinterpret_cast<const float4*>(a.weight + (int64_t)(2 * p + 1) * a.F) + c) : z;
        s_w4[(2 * p) * C + c] = make_float4(w0.x, w1.x, w0.y, w1.y);
        s_w4[(2 * p + 1) * C + c] = make_float4(w0.z, w1.z, w0.w, w1.w);
    }
    __syncthreads();
    const ulonglong2* s_wp = reinterpret_cast<const ulonglong2*>(s_w);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t total = (int64_t)gridDim.x * (kThreads / 32);
    const int64_t n_items = (a.m + RW - 1) / RW;
    for (int64_t item = (int64_t)blockIdx.x * (kThreads / 32) + warp; item < n_items; item += total) {
        const int64_t row0 = item * RW;
        const float* xr[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) xr[r] = row0 + r < a.m ? feat_row(a, row0 + r, lane == 0) : nullptr;
        u64 acc[RW][NP];
#pragma unroll
        for (int r = 0; r < RW; ++r)
#pragma unroll
            for (int p = 0; p < NP; ++p) acc[r][p] = 0ull;
        float4 xn[RW];
        int c = lane;
#pragma unroll
        for (int r = 0; r < RW; ++r) xn[r] = (c < C && xr[r]) ? ld_stream4(xr[r] + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (; c < C; c += 32) {
            u64 xd[RW][4];
#pragma unroll
            for (int r = 0; r < RW; ++r) {
                xd[r][0] = pack2(xn[r].x, xn[r].x); xd[r][1] = pack2(xn[r].y, xn[r].y);
                xd[r][2] = pack2(xn[r].z, xn[r].z); xd[r][3] = pack2(xn[r].w, xn[r].w);
            }
            const int cn = c + 32;
#pragma unroll
            for (int r = 0; r < RW; ++r) xn[r] = (cn < C && xr[r]) ? ld_stream4(xr[r] + 4 * cn) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const ulonglong2 wa = s_wp[(2 * p) * C + c], wb = s_wp[(2 * p + 1) * C + c];
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    acc[r][p] = ffma2(xd[r][0], wa.x, acc[r][p]);
                    acc[r][p] = ffma2(xd[r][1], wa.y, acc[r][p]);
                    acc[r][p] = ffma2(xd[r][2], wb.x, acc[r][p]);
                    acc[r][p] = ffma2(xd[r][3], wb.y, acc[r][p]);
                }
            }
        }
        // 8 NP partial sums per lane -> after three halving steps lane group (b4 b3 b2) holds outputs [b2 NP, (b2+1) NP) of
        // row 2 b4 + b3; two plain butterfly steps finish the sum over the group's four lanes
        float v0[RW * 2 * NP];
#pragma unroll
        for (int r = 0; r < RW; ++r)
#pragma unroll
            for (int p = 0; p < NP; ++p) unpack2(acc[r][p], v0[r * 2 * NP + 2 * p], v0[r * 2 * NP + 2 * p + 1]);
        float v1[RW * NP], v2[RW * NP / 2], v3[NP];
        halve<RW * 2 * NP>(v0, v1, (lane & 16) != 0, 16);
        halve<RW * NP>(v1, v2, (lane & 8) != 0, 8);
        halve<2 * NP>(v2, v3, (lane & 4) != 0, 4);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            v3[i] += __shfl_xor_sync(0xffffffffu, v3[i], 2);
            v3[i] += __shfl_xor_sync(0xffffffffu, v3[i], 1);
        }
        const int r = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1), d0 = ((lane >> 2) & 1) * NP;
        if ((lane & 3) == 0 && row0 + r < a.m) {
#pragma unroll
            for (int i = 0; i < NP; ++i)
                if (d0 + i < a.D) a.Y[(row0 + r) * a.D + d0 + i] = v3[i] + (a.bias ? __ldg(a.bias + d0 + i) : 0.f);
        }
    }
}

// NT threads per block, a thread owns CPT float4 column chunks (chunk = tid + NT*j) x NP output pairs; packed FMAs as above
template <int NP, int NT, int CPT, int RU>
__global__ void __launch_bounds__(NT, 1) featnet_wgrad_kernel(const FeatArgs a) {
    constexpr int RB = 32;                            // rows staged per round
    constexpr int DP = 2 * NP;
    __shared__ __align__(16) float s_g[RB][DP];
    __shared__ const float* s_x[RB];
    pdl_launch_dependents();
    pdl_wait();
    const int C = a.F >> 2;
    const int tid = threadIdx.x;
    // contiguous row range of this block
    const int64_t base_n = a.m / gridDim.x, rem = a.m % gridDim.x;
    const int64_t r_lo = blockIdx.x * base_n + (blockIdx.x < rem ? blockIdx.x : rem);
    const int64_t r_hi = r_lo + base_n + (blockIdx.x < rem ? 1 : 0);
    u64 acc[CPT][4][NP];
#pragma unroll
    for (int j = 0; j < CPT; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int p = 0; p < NP; ++p) acc[j][k][p] = 0ull;
    float bias_acc = 0.f;                             // thread d < D: sum_i gY[i, d]
    for (int64_t base = r_lo; base < r_hi; base += RB) {
        __syncthreads();
        const int nb = (int)((r_hi - base) < RB ? (r_hi - base) : RB);
        for (int i = tid; i < RB * DP; i += NT) {
            const int r = i / DP, d = i - r * DP;
            s_g[r][d] = (r < nb && d < a.D) ? __ldg(a.gY + (base + r) * a.D + d) : 0.f;
        }
        if (tid < RB) s_x[tid] = tid < nb ? feat_row(a, base + tid, true) : nullptr;
        __syncthreads();
        if (tid < a.D)
            for (int r = 0; r < nb; ++r) bias_acc += s_g[r][tid];
        for (int r0 = 0; r0 < nb; r0 += RU) {
            float4 x[RU][CPT];
#pragma unroll
            for (int rr = 0; rr < RU; ++rr) {
                const float* xp = r0 + rr < nb ? s_x[r0 + rr] : nullptr;
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const int c = tid + NT * j;
                    x[rr][j] = (xp && c < C) ? ld_stream4(xp + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int rr = 0; rr < RU; ++rr) {
                if (r0 + rr < nb) {
                    const u64* gp = reinterpret_cast<const u64*>(&s_g[r0 + rr][0]);
                    u64 g2[NP];
#pragma unroll
                    for (int p = 0; p < NP; ++p) g2[p] = gp[p];
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        const u64 x0 = pack2(x[rr][j].x, x[rr][j].x), x1 = pack2(x[rr][j].y, x[rr][j].y);
                        const u64 x2 = pack2(x[rr][j].z, x[rr][j].z), x3 = pack2(x[rr][j].w, x[rr][j].w);
#pragma unroll
                        for (int p = 0; p < NP; ++p) {
                            acc[j][0][p] = ffma2(x0, g2[p], acc[j][0][p]);
                            acc[j][1][p] = ffma2(x1, g2[p], acc[j][1][p]);
                            acc[j][2][p] = ffma2(x2, g2[p], acc[j][2][p]);
                            acc[j][3][p] = ffma2(x3, g2[p], acc[j][3][p]);
                        }
                    }
                }
            }
        }
    }
    if (r_hi <= r_lo) return;
    float* out = a.grad + (int64_t)(blockIdx.x % a.replicas) * a.grad_stride;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int c = tid + NT * j;
        if (c < C) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                float e0, o0, e1, o1, e2, o2, e3, o3;
                unpack2(acc[j][0][p], e0, o0); unpack2(acc[j][1][p], e1, o1);
                unpack2(acc[j][2][p], e2, o2); unpack2(acc[j][3][p], e3, o3);
                if (2 * p < a.D) red_add4(out + (int64_t)(2 * p) * a.F + 4 * c, make_float4(e0, e1, e2, e3));
                if (2 * p + 1 < a.D) red_add4(out + (int64_t)(2 * p + 1) * a.F + 4 * c, make_float4(o0, o1, o2, o3));
            }
        }
    }
    if (tid < a.D) atomicAdd(out + (int64_t)a.D * a.F + tid, bias_acc);
}

static int np_of(int D) { return (D + 1) / 2; }

bool featnet_supported(int F, int D) {
    return D >= 1 && D <= 16 && F >= 4 && (F & 3) == 0 && F <= 4096 && (size_t)2 * np_of(D) * F * 4 <= 200 * 1024;
}

template <int NP>
static int featnet_fwd_go(const FeatArgs& a, cudaStream_t st) {
    constexpr int RW = 4;
    const size_t smem = (size_t)2 * NP * a.F * 4;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(featnet_fwd_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done = true;
    }
    int per_sm = smem > 0 ? (int)((220 * 1024) / (smem + 1024)) : 8;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    const int64_t items = (a.m + RW - 1) / RW;
    int64_t need = (items + kThreads / 32 - 1) / (kThreads / 32);
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (need < 1) need = 1;
    const cudaError_t e = launch_step_kernel(featnet_fwd_kernel<NP>, (int)(need < cap ? need : cap), kThreads, st, a, smem);
    ++g_launches;
    return (int)(e != cudaSuccess ? e : cudaGetLastError());
}

template <int NP, int NT, int CPT, int RU>
static int featnet_wgrad_go2(const FeatArgs& a, cudaStream_t st) {
    int64_t grid = sm_count();
    const int64_t min_rows = 8;                       // a block's flush costs D*F/4 vector reductions: give it rows to amortise
    if (grid * min_rows > a.m) grid = (a.m + min_rows - 1) / min_rows;
    if (grid < 1) grid = 1;
    const cudaError_t e = launch_step_kernel(featnet_wgrad_kernel<NP, NT, CPT, RU>, (int)grid, NT, st, a);
    ++g_launches;
    return (int)(e != cudaSuccess ? e : cudaGetLastError());
}

template <int NP>
static int featnet_wgrad_go(const FeatArgs& a, cudaStream_t st) {
    const int C = a.F >> 2;
    if (C <= 256) return featnet_wgrad_go2<NP, 256, 1, 8>(a, st);
    if (C <= 512) return featnet_wgrad_go2<NP, 512, 1, 8>(a, st);
    return featnet_wgrad_go2<NP, 512, 2, 4>(a, st);
}

#define LEC_FEAT_DISPATCH(FN, A, ST)                  \
    switch (np_of((A).D)) {                           \
        case 1: return FN<1>(A, ST);                  \
        case 2: return FN<2>(A, ST);                  \
        case 3: return FN<3>(A, ST);                  \
        case 4: return FN<4>(A, ST);                  \
        case 5: return FN<5>(A, ST);                  \
        case 6: return FN<6>(A, ST);                  \
        case 7: return FN<7>(A, ST);                  \
        default: return FN<8>(A, ST);                 \
    }

int featnet_fwd_launch(const FeatArgs& a, cudaStream_t st) {
    if (a.m == 0) return 0;
    LEC_FEAT_DISPATCH(featnet_fwd_go, a, st)
}

int featnet_wgrad_launch(const FeatArgs& a, cudaStream_t st) {
    if (a.m == 0) return 0;
    LEC_FEAT_DISPATCH(featnet_wgrad_go, a, st)
}

}  // namespace lec
