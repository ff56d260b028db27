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

template <int DP, int RW>
__global__ void __launch_bounds__(kThreads) featnet_fwd_kernel(const FeatArgs a) {
    extern __shared__ __align__(16) float s_w[];   // [DP][F], rows D..DP-1 zero
    pdl_launch_dependents();
    pdl_wait();
    const int C = a.F >> 2;                          // float4 chunks per row
    float4* s_w4 = reinterpret_cast<float4*>(s_w);
    for (int i = threadIdx.x; i < DP * C; i += kThreads) {
        const int d = i / C, c = i - d * C;
        s_w4[i] = d < a.D ? __ldg(reinterpret_cast<const float4*>(a.weight + (int64_t)d * a.F) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t total = (int64_t)gridDim.x * (kThreads / 32);
    const int64_t n_items = (a.m + RW - 1) / RW;
    for (int64_t item = (int64_t)blockIdx.x * (kThreads / 32) + warp; item < n_items; item += total) {
        const int64_t row0 = item * RW;
        const float* xr[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) xr[r] = row0 + r < a.m ? feat_row(a, row0 + r, lane == 0) : nullptr;
        float acc[RW][DP];
#pragma unroll
        for (int r = 0; r < RW; ++r)
#pragma unroll
            for (int d = 0; d < DP; ++d) acc[r][d] = 0.f;
        float4 xn[RW];
        int c = lane;
#pragma unroll
        for (int r = 0; r < RW; ++r) xn[r] = (c < C && xr[r]) ? ld_stream4(xr[r] + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (; c < C; c += 32) {
            float4 x[RW];
#pragma unroll
            for (int r = 0; r < RW; ++r) x[r] = xn[r];
            const int cn = c + 32;
#pragma unroll
            for (int r = 0; r < RW; ++r) xn[r] = (cn < C && xr[r]) ? ld_stream4(xr[r] + 4 * cn) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                const float4 w = s_w4[d * C + c];
#pragma unroll
                for (int r = 0; r < RW; ++r)
                    acc[r][d] = fmaf(x[r].x, w.x, fmaf(x[r].y, w.y, fmaf(x[r].z, w.z, fmaf(x[r].w, w.w, acc[r][d]))));
            }
        }
#pragma unroll
        for (int r = 0; r < RW; ++r)
#pragma unroll
            for (int d = 0; d < DP; ++d) acc[r][d] = warp_sum<float>(acc[r][d]);
        // every lane holds every sum; lane d stores output d (one 4*D-byte row per store instruction)
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            if (row0 + r < a.m) {
                float out = 0.f;
#pragma unroll
                for (int d = 0; d < DP; ++d) out = lane == d ? acc[r][d] : out;
                if (lane < a.D) a.Y[(row0 + r) * a.D + lane] = out + (a.bias ? __ldg(a.bias + lane) : 0.f);
            }
        }
    }
}

// NT threads per block, a thread owns CPT float4 column chunks (chunk = tid + NT*j) x DP outputs
template <int DP, int NT, int CPT, int RU>
__global__ void __launch_bounds__(NT, 1) featnet_wgrad_kernel(const FeatArgs a) {
    constexpr int RB = 32;                            // rows staged per round
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
    float acc[CPT][4][DP];
#pragma unroll
    for (int j = 0; j < CPT; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int d = 0; d < DP; ++d) acc[j][k][d] = 0.f;
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
#pragma unroll
                    for (int d4 = 0; d4 < DP / 4; ++d4) {
                        const float4 g = *reinterpret_cast<const float4*>(&s_g[r0 + rr][4 * d4]);
                        const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
#pragma unroll
                            for (int j = 0; j < CPT; ++j) {
                                acc[j][0][4 * d4 + k] = fmaf(x[rr][j].x, gv[k], acc[j][0][4 * d4 + k]);
                                acc[j][1][4 * d4 + k] = fmaf(x[rr][j].y, gv[k], acc[j][1][4 * d4 + k]);
                                acc[j][2][4 * d4 + k] = fmaf(x[rr][j].z, gv[k], acc[j][2][4 * d4 + k]);
                                acc[j][3][4 * d4 + k] = fmaf(x[rr][j].w, gv[k], acc[j][3][4 * d4 + k]);
                            }
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
            for (int d = 0; d < DP; ++d)
                if (d < a.D) red_add4(out + (int64_t)d * a.F + 4 * c, make_float4(acc[j][0][d], acc[j][1][d], acc[j][2][d], acc[j][3][d]));
        }
    }
    if (tid < a.D) atomicAdd(out + (int64_t)a.D * a.F + tid, bias_acc);
}

static int dp_of(int D) { return D <= 4 ? 4 : (D <= 8 ? 8 : (D <= 12 ? 12 : 16)); }

bool featnet_supported(int F, int D) {
    return D >= 1 && D <= 16 && F >= 4 && (F & 3) == 0 && F <= 4096 && (size_t)dp_of(D) * F * 4 <= 200 * 1024;
}

template <int DP>
static int featnet_fwd_go(const FeatArgs& a, cudaStream_t st) {
    constexpr int RW = 4;
    const size_t smem = (size_t)DP * a.F * 4;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(featnet_fwd_kernel<DP, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done = true;
    }
    int per_sm = smem > 0 ? (int)((220 * 1024) / (smem + 1024)) : 8;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    const int64_t items = (a.m + RW - 1) / RW;
    int64_t need = (items + kThreads / 32 - 1) / (kThreads / 32);
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (need < 1) need = 1;
    const cudaError_t e = launch_step_kernel(featnet_fwd_kernel<DP, RW>, (int)(need < cap ? need : cap), kThreads, st, a, smem);
    ++g_launches;
    return (int)(e != cudaSuccess ? e : cudaGetLastError());
}

template <int DP, int NT, int CPT, int RU>
static int featnet_wgrad_go2(const FeatArgs& a, cudaStream_t st) {
    int64_t grid = sm_count();
    const int64_t min_rows = 8;                       // a block's flush costs D*F/4 vector reductions: give it rows to amortise
    if (grid * min_rows > a.m) grid = (a.m + min_rows - 1) / min_rows;
    if (grid < 1) grid = 1;
    const cudaError_t e = launch_step_kernel(featnet_wgrad_kernel<DP, NT, CPT, RU>, (int)grid, NT, st, a);
    ++g_launches;
    return (int)(e != cudaSuccess ? e : cudaGetLastError());
}

template <int DP>
static int featnet_wgrad_go(const FeatArgs& a, cudaStream_t st) {
    const int C = a.F >> 2;
    if (C <= 256) return featnet_wgrad_go2<DP, 256, 1, 8>(a, st);
    if (C <= 512) return featnet_wgrad_go2<DP, 512, 1, 8>(a, st);
    return featnet_wgrad_go2<DP, 512, 2, 4>(a, st);
}

int featnet_fwd_launch(const FeatArgs& a, cudaStream_t st) {
    if (a.m == 0) return 0;
    switch (dp_of(a.D)) {
        case 4: return featnet_fwd_go<4>(a, st);
        case 8: return featnet_fwd_go<8>(a, st);
        case 12: return featnet_fwd_go<12>(a, st);
        default: return featnet_fwd_go<16>(a, st);
    }
}

int featnet_wgrad_launch(const FeatArgs& a, cudaStream_t st) {
    if (a.m == 0) return 0;
    switch (dp_of(a.D)) {
        case 4: return featnet_wgrad_go<4>(a, st);
        case 8: return featnet_wgrad_go<8>(a, st);
        case 12: return featnet_wgrad_go<12>(a, st);
        default: return featnet_wgrad_go<16>(a, st);
    }
}

}  // namespace lec
