// All-pairs image x label scoring with per-level top-k (replaces the per-image CPU loop of
// JointEmbeddings.calculate_classification_metrics, oe.py:1764-1779 / oe_h.py:2018-2036).
//
// Layout: one thread per image, the image row lives in registers; label rows and their per-label
// scalars (|x|^2 and the half-aperture term, which depend on the label only) are staged in shared
// memory tile by tile and read as warp-wide broadcasts.  The only HBM traffic is the image rows in
// (4*D bytes per image) and the top-k out (8*n_levels*k bytes per image), or the full [N, L] matrix
// when the caller asks for it.
#include "lec_common.cuh"

namespace lec {

struct ScoreArgs {
    const float* labels; int64_t L; const float* images; int64_t N; int D; float K;
    int n_levels; int k;
    int level_start[LEC_MAX_LEVELS]; int level_stop[LEC_MAX_LEVELS];
    float* scores; int64_t s_img, s_lab; int32_t* topk_idx; float* topk_val;
    int tile_labels;  // labels per shared-memory tile
};

enum { SC_EUC = 0, SC_HYP32 = 1, SC_HYP64 = 2, SC_OE = 3 };

// per-label scalars (fp64, once per label per block): s0 = |x|^2 ; s1 = sqrt(1-K^2/|x|^2) (euc) or
// asin(clamp(K(1-|x|^2)/|x|)) (hyp) -- the same per-row terms lec_rows_fwd stores as "aux"
template <int GEOMC>
__device__ __forceinline__ void label_scalars(double A, float K, double& s0, double& s1) {
    const int geom = (GEOMC == SC_EUC) ? LEC_GEOM_EUC : (GEOMC == SC_OE ? LEC_GEOM_OE : LEC_GEOM_HYP);
    const Aux<double> x = row_aux<double>(geom, A, K);
    s0 = x.A;
    s1 = x.t0;
}

template <int GEOMC, int DQ>
__global__ void __launch_bounds__(kThreads) score_kernel(const ScoreArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int DP = 4 * DQ;
    float* lab = smem;                                                                  // [tile][DP]
    double* ls0 = reinterpret_cast<double*>(smem + (size_t)a.tile_labels * DP);         // [tile]
    double* ls1 = ls0 + a.tile_labels;                                                  // [tile]

    const int64_t img = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool valid = img < a.N;
    float y[DP];
    float B = 0.f;
    {
        const float* src = a.images + (valid ? img : 0) * (int64_t)a.D;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            y[d] = (d < a.D) ? __ldg(src + d) : 0.f;
            B = fmaf(y[d], y[d], B);
        }
    }
    float tv[LEC_MAX_TOPK];
    int ti[LEC_MAX_TOPK];
#pragma unroll
    for (int j = 0; j < LEC_MAX_TOPK; ++j) { tv[j] = INFINITY; ti[j] = -1; }
    float thr = INFINITY;  // current k-th best of the level being scanned
    int level = 0;  // current level (levels are ascending, disjoint label ranges)

    for (int64_t l0 = 0; l0 < a.L; l0 += a.tile_labels) {
        const int tl = (int)min((int64_t)a.tile_labels, a.L - l0);
        __syncthreads();
        for (int i = threadIdx.x; i < tl * DP; i += kThreads) {
            const int r = i / DP, d = i - r * DP;
            lab[i] = (d < a.D) ? __ldg(a.labels + (l0 + r) * (int64_t)a.D + d) : 0.f;
        }
        __syncthreads();
        for (int r = threadIdx.x; r < tl; r += kThreads) {
            double A = 0.0;
            for (int d = 0; d < DP; ++d) A += (double)lab[r * DP + d] * (double)lab[r * DP + d];
            label_scalars<GEOMC>(A, a.K, ls0[r], ls1[r]);
        }
        __syncthreads();
        for (int r = 0; r < tl; ++r) {
            const int l = (int)(l0 + r);
            // level bookkeeping (uniform across the block)
            while (level < a.n_levels && l >= a.level_stop[level]) {
                if (valid && a.topk_idx) {
                    for (int j = 0; j < LEC_MAX_TOPK; ++j)
                        if (j < a.k) {
                            const int64_t o = (img * a.n_levels + level) * a.k + j;
                            a.topk_idx[o] = ti[j];
                            if (a.topk_val) a.topk_val[o] = tv[j];
                        }
                }
#pragma unroll
                for (int j = 0; j < LEC_MAX_TOPK; ++j) { tv[j] = INFINITY; ti[j] = -1; }
                thr = INFINITY;
                ++level;
            }
            const float4* xr = reinterpret_cast<const float4*>(lab + r * DP);
            float E;
            if (GEOMC == SC_OE) {
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < DQ; ++q) {
                    const float4 x = xr[q];
                    float d;
                    d = fmaxf(x.x - y[4 * q + 0], 0.f); s = fmaf(d, d, s);
                    d = fmaxf(x.y - y[4 * q + 1], 0.f); s = fmaf(d, d, s);
                    d = fmaxf(x.z - y[4 * q + 2], 0.f); s = fmaf(d, d, s);
                    d = fmaxf(x.w - y[4 * q + 3], 0.f); s = fmaf(d, d, s);
                }
                E = s;
            } else if (GEOMC == SC_EUC) {
                float dd = 0.f, xd = 0.f;
#pragma unroll
                for (int q = 0; q < DQ; ++q) {
                    const float4 x = xr[q];
                    float d;
                    d = y[4 * q + 0] - x.x; dd = fmaf(d, d, dd); xd = fmaf(x.x, d, xd);
                    d = y[4 * q + 1] - x.y; dd = fmaf(d, d, dd); xd = fmaf(x.y, d, xd);
                    d = y[4 * q + 2] - x.z; dd = fmaf(d, d, dd); xd = fmaf(x.z, d, xd);
                    d = y[4 * q + 3] - x.w; dd = fmaf(d, d, dd); xd = fmaf(x.w, d, xd);
                }
                const float A = (float)ls0[r];
                const float an = fmaxf(sqrtf(A), kNormEps), bn = fmaxf(sqrtf(dd), kNormEps);
                E = relu_nan((float)ls1[r] - xd / (an * bn));
            } else {
                float p = 0.f, s2 = 0.f;
#pragma unroll
                for (int q = 0; q < DQ; ++q) {
                    const float4 x = xr[q];
                    float d;
                    p = fmaf(x.x, y[4 * q + 0], p); d = x.x - y[4 * q + 0]; s2 = fmaf(d, d, s2);
                    p = fmaf(x.y, y[4 * q + 1], p); d = x.y - y[4 * q + 1]; s2 = fmaf(d, d, s2);
                    p = fmaf(x.z, y[4 * q + 2], p); d = x.z - y[4 * q + 2]; s2 = fmaf(d, d, s2);
                    p = fmaf(x.w, y[4 * q + 3], p); d = x.w - y[4 * q + 3]; s2 = fmaf(d, d, s2);
                }
                const float A = (float)ls0[r];
                float theta;
                if (GEOMC == SC_HYP64) {
                    const double Ad = ls0[r], Bd = B, pd = p;
                    const double w2 = 1.0 + Ad * Bd - 2.0 * pd;
                    const double g = (s2 == 0.f) ? (double)NAN : diff_of_products(pd, 1.0 + Ad, Ad, 1.0 + Bd) / (sqrt(Ad) * sqrt((double)s2) * sqrt(w2));
                    const double gc = g < -1.0 + 1e-5 ? -1.0 + 1e-5 : (g > 1.0 - 1e-5 ? 1.0 - 1e-5 : g);
                    theta = atan2f((float)sqrt((1.0 - gc) * (1.0 + gc)), (float)gc);
                } else {
                    const float w2 = 1.f + A * B - 2.f * p;
                    const float g = (s2 == 0.f) ? NAN : diff_of_products(p, 1.f + A, A, 1.f + B) / (sqrtf(A) * sqrtf(s2) * sqrtf(w2));
                    const float gc = g < -1.f + kClampEps ? -1.f + kClampEps : (g > 1.f - kClampEps ? 1.f - kClampEps : g);
                    theta = acosf(gc);
                }
                E = relu_nan(theta - (float)ls1[r]);
            }
            if (valid && a.scores) a.scores[img * a.s_img + (int64_t)l * a.s_lab] = E;
            if (level < a.n_levels && l >= a.level_start[level]) {
                // sorted insertion, ascending; NaN never enters (comparison false), ties keep the lower label
                if (E < thr) {
                    float cv = E; int ci = l;
#pragma unroll
                    for (int j = 0; j < LEC_MAX_TOPK; ++j) {
                        if (j < a.k && cv < tv[j]) {
                            const float ov = tv[j]; const int oi = ti[j];
                            tv[j] = cv; ti[j] = ci; cv = ov; ci = oi;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < LEC_MAX_TOPK; ++j)
                        if (j == a.k - 1) thr = tv[j];
                }
            }
        }
    }
    // flush the levels that end at (or after) L
    while (level < a.n_levels) {
        if (valid && a.topk_idx) {
            for (int j = 0; j < LEC_MAX_TOPK; ++j)
                if (j < a.k) {
                    const int64_t o = (img * a.n_levels + level) * a.k + j;
                    a.topk_idx[o] = ti[j];
                    if (a.topk_val) a.topk_val[o] = tv[j];
                }
        }
#pragma unroll
        for (int j = 0; j < LEC_MAX_TOPK; ++j) { tv[j] = INFINITY; ti[j] = -1; }
        ++level;
    }
}

template <int GEOMC, int DQ>
static int score_launch_dq(ScoreArgs& a, cudaStream_t st) {
    constexpr int DP = 4 * DQ;
    const size_t per_label = (size_t)(DP + 4) * sizeof(float);  // row + two doubles
    const size_t budget = 200 * 1024;
    int64_t tile = (int64_t)(budget / per_label);
    if (tile > a.L) tile = a.L;
    if (tile < 1) tile = 1;
    a.tile_labels = (int)tile;
    const size_t smem = (size_t)tile * per_label;
    cudaError_t e = cudaFuncSetAttribute(score_kernel<GEOMC, DQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int grid = (int)((a.N + kThreads - 1) / kThreads);
    score_kernel<GEOMC, DQ><<<grid, kThreads, smem, st>>>(a);
    ++g_launches;
    return (int)cudaGetLastError();
}

template <int GEOMC>
static int score_launch_geom(ScoreArgs& a, cudaStream_t st) {
    const int dq = (a.D + 3) / 4;
    if (dq <= 1) return score_launch_dq<GEOMC, 1>(a, st);
    if (dq <= 2) return score_launch_dq<GEOMC, 2>(a, st);
    if (dq <= 3) return score_launch_dq<GEOMC, 3>(a, st);
    if (dq <= 4) return score_launch_dq<GEOMC, 4>(a, st);
    if (dq <= 8) return score_launch_dq<GEOMC, 8>(a, st);
    if (dq <= 13) return score_launch_dq<GEOMC, 13>(a, st);
    if (dq <= 16) return score_launch_dq<GEOMC, 16>(a, st);
    if (dq <= 32) return score_launch_dq<GEOMC, 32>(a, st);
    return LEC_E_DIM;  // scoring keeps the image row in registers: D <= 128
}

bool score_fast_supported(int geom, int precision, int D, int64_t L);
int score_fast_launch(int geom, const float* labels, int64_t L, const float* images, int64_t N, int D, float K,
                      const int32_t* level_start, const int32_t* level_stop, int n_levels, int k, float* scores,
                      int64_t s_img, int64_t s_lab, int32_t* topk_idx, float* topk_val, cudaStream_t st);

// s_img / s_lab: element strides of the score matrix (row-major [N, L]: L, 1; label-major [L, N]: 1, N)
int score_launch(int geom, int precision, const float* labels, int64_t L, const float* images, int64_t N, int D,
                 float K, const int32_t* level_start, const int32_t* level_stop, int n_levels, int k, float* scores,
                 int64_t s_img, int64_t s_lab, int32_t* topk_idx, float* topk_val, cudaStream_t st) {
    if (score_fast_supported(geom, precision, D, L))
        return score_fast_launch(geom, labels, L, images, N, D, K, level_start, level_stop, n_levels, k, scores, s_img,
                                 s_lab, topk_idx, topk_val, st);
    ScoreArgs a{};
    a.labels = labels; a.L = L; a.images = images; a.N = N; a.D = D; a.K = K; a.n_levels = n_levels; a.k = k;
    for (int i = 0; i < n_levels; ++i) { a.level_start[i] = level_start[i]; a.level_stop[i] = level_stop[i]; }
    a.scores = scores; a.s_img = s_img; a.s_lab = s_lab; a.topk_idx = topk_idx; a.topk_val = topk_val;
    if (N == 0 || L == 0) return 0;
    switch (geom) {
        case LEC_GEOM_EUC: return score_launch_geom<SC_EUC>(a, st);
        case LEC_GEOM_HYP:
            return precision == LEC_PREC_F64CORE ? score_launch_geom<SC_HYP64>(a, st) : score_launch_geom<SC_HYP32>(a, st);
        case LEC_GEOM_OE: return score_launch_geom<SC_OE>(a, st);
    }
    return LEC_E_ENUM;
}

}  // namespace lec
