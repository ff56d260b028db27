// Per-row kernels: Embedder / FeatNet row transforms (forward + VJP), the per-row aperture terms
// ("aux") consumed by the pair kernels, and the fused Riemannian-SGD table update.
// One team of TT lanes per table row; lanes stride over the D columns so a team's accesses are
// contiguous.  These are full-table elementwise passes: 8*D bytes per row for the transforms,
// 12*D bytes per row for the update (read w, read g, write w).
#include "lec_common.cuh"

namespace lec {

struct RowsArgs {
    const float* in; int64_t n; int D; int mode; int geom; float K; float r_in; float c0;
    float* out; int ld; double* aux; float* zero_out; int replicas; int64_t replica_stride; double* zero_scalar;
    const float* grad_rows; float* grad_in; int accumulate;
};

// sum of the gradient replicas at one element
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

template <int TT>
__device__ __forceinline__ float tsum(float v) { return team_sum<TT, float>(v); }
template <int TT>
__device__ __forceinline__ double tsumd(double v) { return team_sum<TT, double>(v); }

// shell projection (order_embeddings_h.py:217-228): out = (add + e) / div * mul
__device__ __forceinline__ void shell_factor(float r, float r_in, bool feat, float& mul, float& add, float& div) {
    mul = 1.f; add = 0.f; div = 1.f;
    if (r <= r_in) { mul = r_in; div = feat ? (1e-6f + r) : r; add = feat ? 1e-6f : 0.f; }
    if (r >= 1.0f) { mul = (float)(1.0 - 1e-5); div = r; add = 0.f; }
}

// Per-row aperture terms, batched.  row_aux<double> is an fp64 sqrt + division + asin (several hundred issue slots) and
// only ONE lane of a row's team has to run it, so inline it occupied a whole warp for 32 / TT rows at a time (cfg4,
// 82 K rows x 50: ~35 us of an 83 us launch).  Teams park |row|^2 in shared memory instead; every kThreads rows (or at the
// end) the block computes one row per THREAD.  Every thread of the block must call push()/flush() the same number of times.
struct AuxBatch {
    double A[kThreads];
    int64_t row[kThreads];
};
template <int TT>
__device__ __forceinline__ void aux_flush(AuxBatch& b, int& fill, int geom, float K, double* __restrict__ aux) {
    __syncthreads();
    if ((int)threadIdx.x < fill && b.row[threadIdx.x] >= 0) {
        const Aux<double> x = row_aux<double>(geom, b.A[threadIdx.x], K);
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
// Riemannian SGD (order_embeddings_h.py:764-775).  A team keeps its row (w and the rescaled gradient)
// in registers: E elements per lane, element d = lane + TT*j.  Per-element arithmetic is fp32 like the
// reference's; the row-wide sums and the Moebius coefficients built from them are carried in fp64 so
// the update stays accurate when |w| is close to 1 (the denominators there cancel heavily).
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
static void hyp_constants(float K, float& r_in, float& c0);

struct RsgdArgs {
    float* table; const float* grad; int64_t n; int D; int ld_g; float lr; float r_in; int lambda_mode;
    float* grad_out; int replicas; int64_t replica_stride;
    // peer-memory mode (lec_rsgd_update_p2p): the gradient is the sum over ranks of peer[p][row*D + d]
    int world; int rank; int slot; unsigned tag; int64_t slot_floats;
    const float* peer[kMaxPeers]; double* loss_out; int* error_out;
    int local_sources;   // peer mode: 1 = every rank PUSHED its gradient into my buffer (slot[2][world][slot_floats])
    // fused row transform of the updated table (lec_rsgd_update_rows): Embedder.forward of order_embeddings_h.py:205-228
    // for the next step, its per-row aperture terms, and the clearing of the gradient replicas
    float* rows_out; int ld_rows; double* aux_out; float K; float* zero_grad;
    float r_in_rows;     // inner radius as lec_rows_fwd derives it from K (the update's own r_in is the caller's)
    double* loss_acc; double* loss_step;   // *loss_step = *loss_acc; *loss_acc = 0   (single-GPU fused step)
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// layout of one rank's exchange buffer: [2 slots][slot_floats] fp32, then flags[2][world] u32
__device__ __forceinline__ const unsigned* p2p_flags(const float* buf, int64_t slot_floats) {
    return reinterpret_cast<const unsigned*>(buf + 2 * slot_floats);
}

// where rank p's partial gradient of the current slot lives: in p's own buffer (pull layout, slot[2][slot_floats]) or,
// after lec_p2p_push, in MY buffer (push layout, slot[2][world][slot_floats])
__device__ __forceinline__ const float* p2p_src(const RsgdArgs& a, int p) {
    return a.local_sources ? a.peer[a.rank] + ((int64_t)a.slot * a.world + p) * a.slot_floats
                           : a.peer[p] + (int64_t)a.slot * a.slot_floats;
}

template <int TT, int E, bool P2P>
__global__ void __launch_bounds__(kThreads) rsgd_kernel(const RsgdArgs a) {
    __shared__ AuxBatch s_aux;
    int aux_fill = 0;
    pdl_launch_dependents();
    pdl_wait();   // the pair kernel's (or the push kernel's) stores and reductions are complete
    if (P2P) {
        // wait until every rank has published its partial gradient of this step into slot a.slot
        __shared__ int s_ok;
        if (threadIdx.x == 0) {
            const unsigned* flags = p2p_flags(a.peer[a.rank], a.slot_floats * (a.local_sources ? a.world : 1)) + a.slot * a.world;
            int ok = 1;
            const long long t0 = clock64();
            for (int p = 0; p < a.world; ++p) {
                while (ld_acquire_sys(flags + p) < a.tag) {
                    if (clock64() - t0 > 8000000000LL) { ok = 0; break; }   // ~4 s: a peer died; do not hang
                }
            }
            if (!ok && a.error_out) atomicExch(a.error_out, 1);
            s_ok = ok;
            if (blockIdx.x == 0 && a.loss_out) {
                double l = 0.0;
                for (int p = 0; p < a.world; ++p)
                    l += __ldcv(reinterpret_cast<const double*>(p2p_src(a, p) + a.slot_floats - 2));
                *a.loss_out = l;
            }
        }
        __syncthreads();
        if (!s_ok) return;
    }
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
            if (P2P) {
                float acc = 0.f;
                if (d < D)
                    for (int p = 0; p < a.world; ++p)   // fixed rank order: every rank computes the identical sum
                        acc += __ldcv(p2p_src(a, p) + rc * (int64_t)D + d);
                gv[j] = acc;
            } else {
                gv[j] = (d < D) ? rsum(g + d, a.replicas, a.replica_stride) : 0.f;
            }
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
            // __fmul_rn / __fadd_rn below: the stored table value is rounded to fp32 BEFORE the row transform adds its
            // 1e-15, exactly as when a separate lec_rows_fwd reloads it (no FMA contraction across the two stages)
            if (mul != 1.f || div != 1.f) res = __fmul_rn(res / div, mul);
            tv[j] = (d < D) ? res : 0.f;     // the updated row stays in registers for the fused transform
            if (valid && d < D) {
                w[d] = res;
                if (go) go[d] = gv[j];
            }
        }
        if (a.rows_out) {
            // ---- rows_fwd_kernel, LEC_ROWS_HYP_SHELL, on the row just written (same operation order, so the fused and
            //      the separate launch produce the same bits) ----
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < E; ++j) {
                if (lane + TT * j < D) { const float v = __fadd_rn(tv[j], 1e-15f); ss = fmaf(v, v, ss); }
            }
            ss = tsum<TT>(ss);
            float m2, a2, d2;
            shell_factor(sqrtf(ss), a.r_in_rows, false, m2, a2, d2);
            double A = 0.0;
            float* o = a.rows_out + rc * (int64_t)a.ld_rows;
            float* zo = a.zero_grad ? a.zero_grad + rc * (int64_t)a.ld_rows : nullptr;
#pragma unroll
            for (int j = 0; j < E; ++j) {
                const int d = lane + TT * j;
                if (d < a.ld_rows) {
                    float v = 0.f;
                    if (d < D) {
                        v = __fadd_rn(tv[j], 1e-15f);
                        if (m2 != 1.f || d2 != 1.f) v = __fmul_rn((a2 + v) / d2, m2);
                    }
                    A += (double)v * (double)v;
                    if (valid) {
                        o[d] = v;
                        if (zo)
                            for (int rr = 0; rr < a.replicas; ++rr) zo[rr * a.replica_stride + d] = 0.f;
                    }
                }
            }
            for (int d = lane + TT * E; d < a.ld_rows; d += TT) {   // pad columns beyond the register tile (D <= 2)
                if (valid) {
                    o[d] = 0.f;
                    if (zo)
                        for (int rr = 0; rr < a.replicas; ++rr) zo[rr * a.replica_stride + d] = 0.f;
                }
            }
            A = tsumd<TT>(A);
            if (a.aux_out) aux_push<TT>(s_aux, aux_fill, A, row, valid, LEC_GEOM_HYP, a.K, a.aux_out);
        }
    }
    if (a.rows_out && a.aux_out) aux_flush<TT>(s_aux, aux_fill, LEC_GEOM_HYP, a.K, a.aux_out);
    if (a.loss_acc && blockIdx.x == 0 && threadIdx.x == 0) {
        if (a.loss_step) *a.loss_step = *a.loss_acc;
        *a.loss_acc = 0.0;
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

static void hyp_constants(float K, float& r_in, float& c0) {
    const double k = (double)K;
    const double rin = 2.0 * k / (1.0 + sqrt(1.0 + 4.0 * k * k));  // order_embeddings_h.py:1089
    double v = rin;                                                 // oe_h.py:106-110 (arctanh with clamp)
    if (v < -1 + 1e-5) v = -1 + 1e-5;
    if (v > 1 - 1e-5) v = 1 - 1e-5;
    r_in = (float)rin;
    c0 = (float)(0.5 * (log(1 + v) - log(1 - v)));
}

int rows_fwd_launch(const float* in, int64_t n, int D, int mode, int geom, float K, float* out, int ld, double* aux,
                    float* zero_out, int zero_replicas, double* zero_scalar, cudaStream_t st) {
    RowsArgs a{};
    a.zero_scalar = zero_scalar;
    a.in = in; a.n = n; a.D = D; a.mode = mode; a.geom = geom; a.K = K; a.out = out; a.ld = ld; a.aux = aux;
    a.zero_out = zero_out; a.replicas = zero_replicas; a.replica_stride = n * (int64_t)ld;
    hyp_constants(K, a.r_in, a.c0);
    if (n == 0) return 0;
    LEC_ROWS_DISPATCH(rows_fwd_kernel, a, n, D, st);
    return (int)cudaGetLastError();
}

int rows_bwd_launch(const float* in, const float* grad_rows, int replicas, int64_t n, int D, int ld, int mode, float K,
                    float* grad_in, int accumulate, cudaStream_t st) {
    RowsArgs a{};
    a.replicas = replicas; a.replica_stride = n * (int64_t)ld;
    a.in = in; a.n = n; a.D = D; a.mode = mode; a.K = K; a.ld = ld; a.grad_rows = grad_rows; a.grad_in = grad_in;
    a.accumulate = accumulate;
    hyp_constants(K, a.r_in, a.c0);
    if (n == 0) return 0;
    LEC_ROWS_DISPATCH(rows_bwd_kernel, a, n, D, st);
    return (int)cudaGetLastError();
}

template <int TT, int E>
static void rsgd_go(const RsgdArgs& a, cudaStream_t st) {
    if (a.world > 0) launch_step_kernel(rsgd_kernel<TT, E, true>, grid_rows(a.n, TT), kThreads, st, a);
    else launch_step_kernel(rsgd_kernel<TT, E, false>, grid_rows(a.n, TT), kThreads, st, a);
}

static int rsgd_dispatch(const RsgdArgs& a, cudaStream_t st);

int rsgd_launch(float* table, const float* grad, int replicas, int64_t n, int D, int ld_g, float lr, float r_in,
                int lambda_mode, float* grad_out, cudaStream_t st) {
    RsgdArgs a{};
    a.table = table; a.grad = grad; a.n = n; a.D = D; a.ld_g = ld_g; a.lr = lr; a.r_in = r_in;
    a.lambda_mode = lambda_mode; a.grad_out = grad_out; a.replicas = replicas; a.replica_stride = n * (int64_t)ld_g;
    a.world = 0;
    return rsgd_dispatch(a, st);
}

// ---- peer-memory exchange (NVLink / NVSwitch P2P) ------------------------------------------------
// publish: stores this rank's loss next to its partial gradient (already written into the slot by
// lec_rows_bwd), then raises flag[slot][rank] = tag in EVERY rank's buffer with release semantics.
__global__ void p2p_publish_kernel(const double* loss_local, RsgdArgs a) {
    float* mine = const_cast<float*>(a.peer[a.rank]) + a.slot * a.slot_floats;
    if (threadIdx.x == 0 && loss_local)
        *reinterpret_cast<double*>(mine + a.slot_floats - 2) = *loss_local;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < a.world) {
        unsigned* flags = const_cast<unsigned*>(p2p_flags(a.peer[threadIdx.x], a.slot_floats)) + a.slot * a.world;
        st_release_sys(flags + a.rank, a.tag);
    }
}

int p2p_publish_launch(const double* loss_local, void* const* peer_bufs, int64_t slot_floats, int world, int rank,
                       int slot, unsigned tag, cudaStream_t st) {
    RsgdArgs a{};
    a.world = world; a.rank = rank; a.slot = slot; a.tag = tag; a.slot_floats = slot_floats;
    for (int p = 0; p < world; ++p) a.peer[p] = static_cast<const float*>(peer_bufs[p]);
    p2p_publish_kernel<<<1, 32, 0, st>>>(loss_local, a);
    ++g_launches;
    return (int)cudaGetLastError();
}

// ---- push exchange: every rank writes its partial gradient into EVERY rank's buffer --------------------------------
// The replica sum of each element (straight-through rows: that IS d loss / d table) is stored into slot[slot][rank] of
// all `world` buffers -- remote stores are fire-and-forget, so the NVLink latency is paid once, behind the kernel, not
// per dependent load -- and the replicas are cleared on the way.  The last block to finish (device counter) stores
// the loss and release-stores the flags; the update kernel then reads local memory only.
struct PushArgs {
    float* grad_rows; int replicas; int64_t replica_stride; int64_t n; int D; int ld;
    double* loss_acc; double* loss_step;
    int world, rank, slot; unsigned tag; int64_t slot_floats; float* peer[kMaxPeers];
    unsigned* counter;
};

__global__ void __launch_bounds__(kThreads) p2p_push_kernel(const PushArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t total = a.n * (int64_t)a.ld;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    const int64_t dst0 = ((int64_t)a.slot * a.world + a.rank) * a.slot_floats;
    // the step's loss is complete when this kernel starts (the pair kernel has finished): it travels with the first
    // block's data, under the same fence, so the last block has nothing left to store but the flags
    if (blockIdx.x == 0 && (int)threadIdx.x < a.world)
        *reinterpret_cast<double*>(a.peer[threadIdx.x] + dst0 + a.slot_floats - 2) = *a.loss_acc;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
        const int64_t row = i / a.ld;
        const int d = (int)(i - row * a.ld);
        float* g = a.grad_rows + i;
        if (d < a.D) {
            const float v = rsum(g, a.replicas, a.replica_stride);
            for (int p = 0; p < a.world; ++p) a.peer[p][dst0 + row * a.D + d] = v;
        }
        for (int r = 0; r < a.replicas; ++r) g[r * a.replica_stride] = 0.f;
    }
    __threadfence_system();     // this thread's remote stores are performed before its block is counted
    __syncthreads();
    __shared__ unsigned s_last;
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(a.counter, 1u);
        s_last = (prev == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();            // all blocks' counts (and the stores fenced before them) are visible here
    if (threadIdx.x == 0) {
        *a.counter = 0;
        if (a.loss_step) *a.loss_step = *a.loss_acc;
        *a.loss_acc = 0.0;
    }
    if ((int)threadIdx.x < a.world) {
        const int p = threadIdx.x;
        unsigned* flags = reinterpret_cast<unsigned*>(a.peer[p] + 2 * (int64_t)a.world * a.slot_floats) + a.slot * a.world;
        st_release_sys(flags + a.rank, a.tag);
    }
}

int p2p_push_launch(float* grad_rows, int replicas, int64_t n, int D, int ld, double* loss_acc, double* loss_step,
                    void* const* peer_bufs, int64_t slot_floats, int world, int rank, int slot, unsigned tag,
                    unsigned* counter, cudaStream_t st) {
    PushArgs a{};
    a.grad_rows = grad_rows; a.replicas = replicas; a.replica_stride = n * (int64_t)ld; a.n = n; a.D = D; a.ld = ld;
    a.loss_acc = loss_acc; a.loss_step = loss_step;
    a.world = world; a.rank = rank; a.slot = slot; a.tag = tag; a.slot_floats = slot_floats; a.counter = counter;
    for (int p = 0; p < world; ++p) a.peer[p] = static_cast<float*>(peer_bufs[p]);
    const int64_t total = n * (int64_t)ld;
    int64_t need = (total + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (need < 1) need = 1;
    launch_step_kernel(p2p_push_kernel, (int)(need < cap ? need : cap), kThreads, st, a);
    ++g_launches;
    return (int)cudaGetLastError();
}

int rsgd_rows_launch(float* table, float* grad_rows, int replicas, int64_t n, int D, int ld, float lr, float r_in,
                     int lambda_mode, float K, float* rows_out, double* aux_out, double* loss_acc, double* loss_step,
                     float* grad_out, cudaStream_t st) {
    RsgdArgs a{};
    a.table = table; a.grad = grad_rows; a.n = n; a.D = D; a.ld_g = ld; a.lr = lr; a.r_in = r_in;
    a.lambda_mode = lambda_mode; a.grad_out = grad_out; a.replicas = replicas; a.replica_stride = n * (int64_t)ld;
    a.world = 0;
    a.rows_out = rows_out; a.ld_rows = ld; a.aux_out = aux_out; a.K = K; a.zero_grad = grad_rows;
    a.loss_acc = loss_acc; a.loss_step = loss_step;
    float c0_unused;
    hyp_constants(K, a.r_in_rows, c0_unused);
    return rsgd_dispatch(a, st);
}

int rsgd_p2p_launch(float* table, void* const* peer_bufs, int64_t slot_floats, int world, int rank, int slot,
                    unsigned tag, int64_t n, int D, float lr, float r_in, int lambda_mode, double* loss_out,
                    int* error_out, int local_sources, float K, float* rows_out, int ld_rows, double* aux_out,
                    cudaStream_t st) {
    RsgdArgs a{};
    a.local_sources = local_sources; a.K = K; a.rows_out = rows_out; a.ld_rows = ld_rows; a.aux_out = aux_out;
    float c0_unused;
    hyp_constants(K, a.r_in_rows, c0_unused);
    a.table = table; a.grad = nullptr; a.n = n; a.D = D; a.ld_g = D; a.lr = lr; a.r_in = r_in;
    a.lambda_mode = lambda_mode; a.grad_out = nullptr; a.replicas = 1; a.replica_stride = 0;
    a.world = world; a.rank = rank; a.slot = slot; a.tag = tag; a.slot_floats = slot_floats;
    a.loss_out = loss_out; a.error_out = error_out;
    for (int p = 0; p < world; ++p) a.peer[p] = static_cast<const float*>(peer_bufs[p]);
    return rsgd_dispatch(a, st);
}

static int rsgd_dispatch(const RsgdArgs& a, cudaStream_t st) {
    const int64_t n = a.n;
    const int D = a.D;
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
