// Per-row kernels: Embedder / FeatNet row transforms (forward + VJP), the per-row aperture terms
// ("aux") consumed by the pair kernels, and the fused Riemannian-SGD table update.
// One team of TT lanes per table row; lanes stride over the D columns so a team's accesses are
// contiguous.  These are full-table elementwise passes: 8*D bytes per row for the transforms,
// 12*D bytes per row for the update (read w, read g, write w).
#include "lec_rowops.cuh"

namespace lec {

struct RowsArgs {
    const float* in; int64_t n; int D; int mode; int geom; float K; float r_in; float c0;
    float* out; int ld; double* aux; float* zero_out; int replicas; int64_t replica_stride; double* zero_scalar;
    const float* grad_rows; float* grad_in; int accumulate;
};

template <int TT>
__device__ __forceinline__ float tsum(float v) { return team_sum<TT, float>(v); }
template <int TT>
__device__ __forceinline__ double tsumd(double v) { return team_sum<TT, double>(v); }

template <int TT>
__global__ void __launch_bounds__(kThreads) rows_fwd_kernel(const RowsArgs a) {
    __shared__ AuxBatch s_aux;
    int aux_fill = 0;
    const int lane = threadIdx.x % TT;
    const int64_t n_teams = (int64_t)gridDim.x * (kThreads / TT);
    const int64_t team = (int64_t)blockIdx.x * (kThreads / TT) + threadIdx.x / TT;
    const int64_t iters = (a.n + n_teams - 1) / n_teams;
    const int D = a.D;
    if (a.zero_scalar && blockIdx.x == 0 && threadIdx.x == 0) *a.zero_scalar = 0.0;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t row = team + it * n_teams;
        const bool valid = row < a.n;
        const float* e = a.in + (valid ? row : 0) * (int64_t)D;
        const bool hyp = a.mode >= LEC_ROWS_HYP_SHELL;
        float ss = 0.f;
        for (int d = lane; d < D; d += TT) {
            float v = __ldg(e + d);
            if (hyp) v += 1e-15f;
            ss = fmaf(v, v, ss);
        }
        ss = tsum<TT>(ss);
        const float r = sqrtf(ss);
        float scale = 1.f;   // first-stage multiplier
        if (a.mode == LEC_ROWS_EUC_SOFTCLIP) {
            scale = (r + a.K);  // applied to e / max(r, eps)
        } else if (a.mode == LEC_ROWS_HYP_TANH || a.mode == LEC_ROWS_HYP_TANH_FEAT) {
            scale = tanhf(fminf(fmaxf(a.c0 + r, -15.f), 15.f));
        }
        const float rn = fmaxf(r, kNormEps);
        // second stage (hyperbolic only): projection on the norm of the first-stage output
        float mul = 1.f, add = 0.f, div = 1.f;
        if (hyp) {
            float r2;
            if (a.mode == LEC_ROWS_HYP_SHELL) {
                r2 = r;
            } else {
                float s2 = 0.f;
                for (int d = lane; d < D; d += TT) {
                    const float v = scale * ((__ldg(e + d) + 1e-15f) / rn);
                    s2 = fmaf(v, v, s2);
                }
                r2 = sqrtf(tsum<TT>(s2));
            }
            shell_factor(r2, a.r_in, a.mode == LEC_ROWS_HYP_TANH_FEAT, mul, add, div);
        }
        double A = 0.0;  // |row|^2 of the values as stored
        float* o = a.out + (valid ? row : 0) * (int64_t)a.ld;
        float* zo = a.zero_out ? a.zero_out + (valid ? row : 0) * (int64_t)a.ld : nullptr;
        for (int d = lane; d < a.ld; d += TT) {
            float v = 0.f;
            if (d < D) {
                v = __ldg(e + d);
                if (hyp) v += 1e-15f;
                if (a.mode == LEC_ROWS_EUC_SOFTCLIP) v = (v / rn) * scale;
                else if (a.mode == LEC_ROWS_HYP_TANH || a.mode == LEC_ROWS_HYP_TANH_FEAT) v = scale * (v / rn);
                if (hyp && (mul != 1.f || div != 1.f)) v = ((add + v) / div) * mul;
            }
            A += (double)v * (double)v;
            if (valid) {
                o[d] = v;
                if (zo)
                    for (int rr = 0; rr < a.replicas; ++rr) zo[rr * a.replica_stride + d] = 0.f;
            }
        }
        if (a.aux) {
            A = tsumd<TT>(A);
            aux_push<TT>(s_aux, aux_fill, A, row, valid, a.geom, a.K, a.aux);
        }
    }
    if (a.aux) aux_flush<TT>(s_aux, aux_fill, a.geom, a.K, a.aux);
}

template <int TT>
__global__ void __launch_bounds__(kThreads) rows_bwd_kernel(const RowsArgs a) {
    const int lane = threadIdx.x % TT;
    const int64_t n_teams = (int64_t)gridDim.x * (kThreads / TT);
    const int64_t team = (int64_t)blockIdx.x * (kThreads / TT) + threadIdx.x / TT;
    const int64_t iters = (a.n + n_teams - 1) / n_teams;
    const int D = a.D;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t row = team + it * n_teams;
        const bool valid = row < a.n;
        const int64_t rc = valid ? row : 0;
        const float* e = a.in + rc * (int64_t)D;
        const float* g = a.grad_rows + rc * (int64_t)a.ld;
        float c_g = 1.f, c_e = 0.f;  // grad_in = c_g * g + c_e * e'
        if (a.mode == LEC_ROWS_EUC_SOFTCLIP) {
            float ss = 0.f, eg = 0.f;
            for (int d = lane; d < D; d += TT) {
                const float ev = __ldg(e + d), gv = rsum(g + d, a.replicas, a.replica_stride);
                ss = fmaf(ev, ev, ss);
                eg = fmaf(ev, gv, eg);
            }
            ss = tsum<TT>(ss); eg = tsum<TT>(eg);
            const float r = sqrtf(ss);
            c_g = 1.f + a.K / r;
            c_e = -a.K * eg / (r * ss);
        } else if (a.mode == LEC_ROWS_HYP_TANH || a.mode == LEC_ROWS_HYP_TANH_FEAT) {
            float ss = 0.f, eg = 0.f;
            for (int d = lane; d < D; d += TT) {
                const float ev = __ldg(e + d) + 1e-15f, gv = rsum(g + d, a.replicas, a.replica_stride);
                ss = fmaf(ev, ev, ss);
                eg = fmaf(ev, gv, eg);
            }
            ss = tsum<TT>(ss); eg = tsum<TT>(eg);
            const float r = sqrtf(ss);
            const float arg = a.c0 + r;
            const float t = tanhf(fminf(fmaxf(arg, -15.f), 15.f));
            const float tp = (arg >= -15.f && arg <= 15.f) ? (1.f - t * t) : 0.f;
            // grad = tp*(eh.g)*eh + (t/r)*(g - eh*(eh.g)),  eh = e'/r
            c_g = t / r;
            c_e = (tp - t / r) * eg / ss;
        }
        if (valid) {
            float* o = a.grad_in + row * (int64_t)D;
            const bool hyp = a.mode >= LEC_ROWS_HYP_SHELL;
            for (int d = lane; d < D; d += TT) {
                float ev = __ldg(e + d);
                if (hyp) ev += 1e-15f;
                float v = fmaf(c_e, ev, c_g * rsum(g + d, a.replicas, a.replica_stride));
                if (a.accumulate) v += o[d];
                o[d] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Riemannian SGD (order_embeddings_h.py:764-775) on a DENSE gradient (lec_rsgd_update: the drop-in for a trainer that
// called loss.backward() and holds weight.grad).  The training engine uses the fused kernel of lec_update.cu instead.
// A team keeps its row (w and the rescaled gradient) in registers: E elements per lane, element d = lane + TT*j.
// Per-element arithmetic is fp32 like the reference's; the row-wide sums and the Moebius coefficients built from them
// are carried in fp64 so the update stays accurate when |w| is close to 1 (the denominators there cancel heavily).
// ------------------------------------------------------------------------------------------------
struct RsgdArgs {
    float* table; const float* grad; int64_t n; int D; int ld_g; float lr; float r_in; int lambda_mode;
    float* grad_out; int replicas; int64_t replica_stride;
};

template <int TT, int E>
__global__ void __launch_bounds__(kThreads) rsgd_kernel(const RsgdArgs a) {
    const int lane = threadIdx.x % TT;
    const int64_t n_teams = (int64_t)gridDim.x * (kThreads / TT);
    const int64_t team = (int64_t)blockIdx.x * (kThreads / TT) + threadIdx.x / TT;
    const int64_t iters = (a.n + n_teams - 1) / n_teams;
    const int D = a.D;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t row = team + it * n_teams;
        const bool valid = row < a.n;
        const int64_t rc = valid ? row : 0;
        float* w = a.table + rc * (int64_t)D;
        const float* g = a.grad + rc * (int64_t)a.ld_g;
        float wv[E], gv[E];
        double uu = 0.0;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int d = lane + TT * j;
            wv[j] = (d < D) ? w[d] : 0.f;
            gv[j] = (d < D) ? rsum(g + d, a.replicas, a.replica_stride) : 0.f;
            uu += (double)wv[j] * (double)wv[j];
        }
        uu = tsumd<TT>(uu);
        // conformal factor (order_embeddings_h.py:662-666; SURVEY F4: the norm, not its square)
        const float wn = (float)sqrt(uu);
        const float lam = 2.f / (1.f - (a.lambda_mode == 1 ? (float)uu : wn));
        const float inv = 1.f / lam;
        const float gs = inv * inv;
        // v = -lr * g' + 1e-15
        double vv = 0.0;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            gv[j] *= gs;                                   // the Riemannian gradient, as left in weight.grad
            const float v = -a.lr * gv[j] + 1e-15f;
            if (lane + TT * j < D) vv += (double)v * (double)v;
        }
        vv = tsumd<TT>(vv);
        const float vn = (float)sqrt(vv);
        const float th = tanhf(fminf(fmaxf(lam * vn / 2.f, -15.f), 15.f));
        // t = th * v / |v| + 1e-6 ; Moebius sums
        float tv[E];
        double uv = 0.0, tt = 0.0;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const float v = -a.lr * gv[j] + 1e-15f;
            tv[j] = th * v / vn + 1e-6f;
            if (lane + TT * j < D) {
                uv += (double)wv[j] * (double)tv[j];
                tt += (double)tv[j] * (double)tv[j];
            }
        }
        uv = 2.0 * tsumd<TT>(uv);
        tt = tsumd<TT>(tt);
        const double den = 1.0 + uv + tt * uu;
        const float cw = (float)((1.0 + uv + tt) / den);
        const float ct = (float)((1.0 - uu) / den);
        double rr = 0.0;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            tv[j] = cw * wv[j] + ct * tv[j];
            if (lane + TT * j < D) rr += (double)tv[j] * (double)tv[j];
        }
        const float rn = (float)sqrt(tsumd<TT>(rr));
        float mul, add, div;
        shell_factor(rn, a.r_in, false, mul, add, div);
        float* go = (valid && a.grad_out) ? a.grad_out + row * (int64_t)D : nullptr;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int d = lane + TT * j;
            float res = tv[j];
            if (mul != 1.f || div != 1.f) res = __fmul_rn(res / div, mul);
            if (valid && d < D) {
                w[d] = res;
                if (go) go[d] = gv[j];
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) reduce_replicas_kernel(const float* in, int replicas, int64_t count,
                                                                   float* out) {
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < count; i += stride)
        out[i] = rsum(in + i, replicas, count);
}

int reduce_replicas_launch(const float* in, int replicas, int64_t count, float* out, cudaStream_t st) {
    if (count == 0) return 0;
    int64_t need = (count + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    reduce_replicas_kernel<<<(int)(need < cap ? need : cap), kThreads, 0, st>>>(in, replicas, count, out);
    ++g_launches;
    return (int)cudaGetLastError();
}

static int team_width(int D) {
    int t = 1;
    while (t < D && t < 32) t <<= 1;
    return t;
}

static int grid_rows(int64_t n, int tt) {
    const int tpb = kThreads / tt;
    int64_t need = (n + tpb - 1) / tpb;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

#define LEC_ROWS_DISPATCH(KERNEL, ARGS, N, D, ST)                                            \
    do {                                                                                     \
        switch (team_width(D)) {                                                             \
            case 1: KERNEL<1><<<grid_rows(N, 1), kThreads, 0, ST>>>(ARGS); break;             \
            case 2: KERNEL<2><<<grid_rows(N, 2), kThreads, 0, ST>>>(ARGS); break;             \
            case 4: KERNEL<4><<<grid_rows(N, 4), kThreads, 0, ST>>>(ARGS); break;             \
            case 8: KERNEL<8><<<grid_rows(N, 8), kThreads, 0, ST>>>(ARGS); break;             \
            case 16: KERNEL<16><<<grid_rows(N, 16), kThreads, 0, ST>>>(ARGS); break;          \
            default: KERNEL<32><<<grid_rows(N, 32), kThreads, 0, ST>>>(ARGS); break;          \
        }                                                                                    \
        ++g_launches;                                                                        \
    } while (0)

int rows_fwd_launch(const float* in, int64_t n, int D, int mode, int geom, float K, float* out, int ld, double* aux,
                    float* zero_out, int zero_replicas, int64_t zero_stride, double* zero_scalar, cudaStream_t st) {
    RowsArgs a{};
    a.zero_scalar = zero_scalar;
    a.in = in; a.n = n; a.D = D; a.mode = mode; a.geom = geom; a.K = K; a.out = out; a.ld = ld; a.aux = aux;
    a.zero_out = zero_out; a.replicas = zero_replicas; a.replica_stride = zero_stride > 0 ? zero_stride : n * (int64_t)ld;
    hyp_constants(K, a.r_in, a.c0);
    if (n == 0) return 0;
    LEC_ROWS_DISPATCH(rows_fwd_kernel, a, n, D, st);
    return (int)cudaGetLastError();
}

int rows_bwd_launch(const float* in, const float* grad_rows, int replicas, int64_t grad_stride, int64_t n, int D, int ld,
                    int mode, float K, float* grad_in, int accumulate, cudaStream_t st) {
    RowsArgs a{};
    a.replicas = replicas; a.replica_stride = grad_stride > 0 ? grad_stride : n * (int64_t)ld;
    a.in = in; a.n = n; a.D = D; a.mode = mode; a.K = K; a.ld = ld; a.grad_rows = grad_rows; a.grad_in = grad_in;
    a.accumulate = accumulate;
    hyp_constants(K, a.r_in, a.c0);
    if (n == 0) return 0;
    LEC_ROWS_DISPATCH(rows_bwd_kernel, a, n, D, st);
    return (int)cudaGetLastError();
}

template <int TT, int E>
static void rsgd_go(const RsgdArgs& a, cudaStream_t st) {
    rsgd_kernel<TT, E><<<grid_rows(a.n, TT), kThreads, 0, st>>>(a);
}

int rsgd_launch(float* table, const float* grad, int replicas, int64_t n, int D, int ld_g, float lr, float r_in,
                int lambda_mode, float* grad_out, cudaStream_t st) {
    RsgdArgs a{};
    a.table = table; a.grad = grad; a.n = n; a.D = D; a.ld_g = ld_g; a.lr = lr; a.r_in = r_in;
    a.lambda_mode = lambda_mode; a.grad_out = grad_out; a.replicas = replicas; a.replica_stride = n * (int64_t)ld_g;
    if (n == 0) return 0;
    // E elements per lane in registers; TT lanes per row
    if (D <= 1) rsgd_go<1, 1>(a, st);
    else if (D <= 2) rsgd_go<1, 2>(a, st);
    else if (D <= 4) rsgd_go<1, 4>(a, st);
    else if (D <= 8) rsgd_go<2, 4>(a, st);
    else if (D <= 16) rsgd_go<4, 4>(a, st);
    else if (D <= 32) rsgd_go<8, 4>(a, st);
    else if (D <= 64) rsgd_go<16, 4>(a, st);
    else if (D <= 128) rsgd_go<32, 4>(a, st);
    else if (D <= 256) rsgd_go<32, 8>(a, st);
    else if (D <= 512) rsgd_go<32, 16>(a, st);
    else rsgd_go<32, 32>(a, st);
    ++g_launches;
    return (int)cudaGetLastError();
}

}  // namespace lec
