// Negative-edge sampler of the cone path, off the Python critical path.
//
// Replaces sample_negative_edge (order_embeddings.py:989-1008; joint variant oe.py:755-808) and the B x N x 2
// Python loop that calls it (order_embeddings.py:1070-1091, oe.py:846-863).  The reference materialises
// np.where(negative_G[u, :] == 1) -- a dense n x n bool row -- for every draw and hands it to random.choice.
// negative_G is always "ones - transitive closure - diagonal" (order_embeddings.py:417-423, oe.py:465-474), so
// the candidate list of a node is [0, n) minus a short sorted EXCLUDED list (the node and its closure
// descendants for a row, the node and its closure ancestors for a column).  The k-th candidate in ascending
// order is then found by a binary search over the excluded list, with no adjacency at all.
//
// Two modes:
//   * exact  (host, this file's lec_sample_negatives): consumes CPython's Mersenne-Twister stream exactly as
//     random.choice does (Random._randbelow_with_getrandbits: k = n.bit_length(), r = genrand_uint32() >> (32-k),
//     redraw while r >= n), so indices are BIT-EXACT with the reference and the caller's `random` state can be
//     loaded before and stored back after the call.
//   * fast   (device, lec_sample_negatives_philox): the same candidate sets and the same uniform law, drawn from
//     a counter-based generator (one Philox4x32-10 block per draw, keyed by (seed, step), counter = draw id);
//     reproducible, order-independent, not the reference's stream.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/lec_b200.h"

namespace lec {
extern std::atomic<unsigned long long> g_launches;  // lec_api.cu
}

// ------------------------------------------------------------------------------------------------
// MT19937 exactly as CPython's _randommodule.c drives it
// ------------------------------------------------------------------------------------------------
static inline void mt_init_genrand(lec_mt19937* s, uint32_t seed) {
    uint32_t* mt = s->mt;
    mt[0] = seed;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    s->index = 624;
}

static inline void mt_twist(lec_mt19937* s) {
    uint32_t* mt = s->mt;
    for (int k = 0; k < 624; ++k) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s->index = 0;
}

static inline uint32_t mt_uint32(lec_mt19937* s) {
    if (s->index >= 624) mt_twist(s);
    uint32_t y = s->mt[s->index++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// Random._randbelow_with_getrandbits for 0 < n < 2^32
static inline uint32_t mt_randbelow(lec_mt19937* s, uint32_t n) {
    int k = 32 - __builtin_clz(n);
    uint32_t r = mt_uint32(s) >> (32 - k);
    while (r >= n) r = mt_uint32(s) >> (32 - k);
    return r;
}

extern "C" int lec_mt_seed(lec_mt19937* s, const uint32_t* key, int key_words) {
    if (!s || !key) return LEC_E_NULL;
    if (key_words < 1) return LEC_E_SIZE;
    mt_init_genrand(s, 19650218u);
    uint32_t* mt = s->mt;
    int i = 1, j = 0;
    int k = key_words > 624 ? key_words : 624;
    for (; k; --k) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        ++i; ++j;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
        if (j >= key_words) j = 0;
    }
    for (k = 623; k; --k) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        ++i;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000u;
    s->index = 624;
    return 0;
}

extern "C" uint32_t lec_mt_uint32(lec_mt19937* s) { return mt_uint32(s); }

extern "C" int64_t lec_mt_randbelow(lec_mt19937* s, uint32_t n) {
    if (!s) return LEC_E_NULL;
    if (n == 0) return LEC_E_EMPTY;
    return (int64_t)mt_randbelow(s, n);
}

// ------------------------------------------------------------------------------------------------
// candidate sets
// ------------------------------------------------------------------------------------------------
namespace {

template <typename T>
__host__ __device__ inline int64_t lower_bound_i32(const T* a, int64_t lo, int64_t hi, int64_t key) {
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// The window [lo, hi) of node ids a draw may come from (per-level filtering, order_embeddings.py:1001-1002;
// joint variant with the extra "image level", oe.py:786-804).
struct Window { int64_t lo, hi; };

__host__ __device__ inline Window draw_window(const lec_sampler_graph* g, int p, int64_t fixed_endpoint) {
    Window w{0, g->n_nodes};
    if (!g->pick_per_level) return w;
    int level = p % g->level_mod;
    if (level < g->n_levels) {
        w.lo = g->level_start[level];
        w.hi = g->level_stop[level];
    } else if (g->n_labels > 0) {  // joint graphs: the slot after the label levels is the image level
        if (fixed_endpoint >= g->n_labels) w.hi = g->n_labels;  // fixed endpoint is an image -> label candidates
        else w.lo = g->n_labels;                                 // fixed endpoint is a label -> image candidates
    }
    return w;
}

// candidates = [w.lo, w.hi) minus excl[a..b); returns their count and the slice of the excluded list
__host__ __device__ inline int64_t candidate_count(const int32_t* excl, int64_t e0, int64_t e1, Window w, int64_t* a_out) {
    int64_t a = lower_bound_i32(excl, e0, e1, w.lo);
    int64_t b = lower_bound_i32(excl, a, e1, w.hi);
    *a_out = a;
    int64_t cnt = (w.hi - w.lo) - (b - a);
    // remember b in the sign-free way: callers recompute m = (w.hi - w.lo) - cnt
    return cnt;
}

// r-th (0-based) candidate in ascending order: x = lo + r + j with j = #excluded below x, i.e. the smallest j
// with j == m or excl[a + j] - lo - j > r
__host__ __device__ inline int64_t kth_candidate(const int32_t* excl, int64_t a, int64_t m, int64_t lo, int64_t r) {
    int64_t jl = 0, jh = m;
    while (jl < jh) {
        int64_t mid = (jl + jh) >> 1;
        if ((int64_t)excl[a + mid] - lo - mid > r) jh = mid; else jl = mid + 1;
    }
    return lo + r + jl;
}

inline int graph_ok(const lec_sampler_graph* g) {
    if (!g || !g->row_excl_ptr || !g->row_excl || !g->col_excl_ptr || !g->col_excl) return LEC_E_NULL;
    if (g->n_nodes < 1 || g->n_nodes > 0x7fffffffLL) return LEC_E_SIZE;
    if (g->pick_per_level) {
        if (g->n_levels < 1 || g->n_levels > LEC_MAX_LEVELS || g->level_mod < 1) return LEC_E_ENUM;
    }
    return 0;
}

}  // namespace

extern "C" int lec_sample_negatives(lec_mt19937* rng, const lec_sampler_graph* g, const int64_t* u, const int64_t* v,
                                    int64_t B, int N, int64_t* neg_to, int64_t* neg_from) {
    int rc = graph_ok(g);
    if (rc) return rc;
    if (!rng || !u || !v || !neg_to || !neg_from) return LEC_E_NULL;
    if (B < 0 || N < 0) return LEC_E_SIZE;
    for (int64_t i = 0; i < B; ++i) {
        int64_t ui = u[i], vi = v[i];
        if (ui < 0 || ui >= g->n_nodes || vi < 0 || vi >= g->n_nodes) return LEC_E_INDEX;
        int64_t r0 = g->row_excl_ptr[ui], r1 = g->row_excl_ptr[ui + 1];
        int64_t c0 = g->col_excl_ptr[vi], c1 = g->col_excl_ptr[vi + 1];
        for (int p = 0; p < N; ++p) {
            // corrupt the child: a candidate of row u (order_embeddings.py:1074)
            Window w = draw_window(g, p, ui);
            int64_t a, cnt = candidate_count(g->row_excl, r0, r1, w, &a);
            if (cnt <= 0) return LEC_E_EMPTY;  // random.choice([]) -> IndexError
            int64_t r = mt_randbelow(rng, (uint32_t)cnt);
            neg_to[i * N + p] = kth_candidate(g->row_excl, a, (w.hi - w.lo) - cnt, w.lo, r);
            // corrupt the parent: a candidate of column v (order_embeddings.py:1083)
            w = draw_window(g, p, vi);
            cnt = candidate_count(g->col_excl, c0, c1, w, &a);
            if (cnt <= 0) return LEC_E_EMPTY;
            r = mt_randbelow(rng, (uint32_t)cnt);
            neg_from[i * N + p] = kth_candidate(g->col_excl, a, (w.hi - w.lo) - cnt, w.lo, r);
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// fast mode: Philox4x32-10, one block per (i, p) slot (two draws)
// ------------------------------------------------------------------------------------------------
namespace {

__host__ __device__ inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// 64 random bits -> uniform integer in [0, n) by multiply-high (bias < n / 2^64, i.e. none that a test can see)
__host__ __device__ inline int64_t mulhi_below(uint64_t x, uint64_t n) {
#ifdef __CUDA_ARCH__
    return (int64_t)__umul64hi(x, n);
#else
    return (int64_t)(((unsigned __int128)x * n) >> 64);
#endif
}

// Draw `draw` of stream `stream`: one Philox4x32-10 block (counter = draw / 2, stream) serves the two draws of an (i, p)
// slot -- the row draw (even id) takes its low 64 bits, the column draw (odd id) its high 64 bits.
__host__ __device__ inline void philox_pair(uint64_t seed, uint64_t stream, uint64_t pair_id, uint64_t& x_even, uint64_t& x_odd) {
    uint32_t c[4] = {(uint32_t)pair_id, (uint32_t)(pair_id >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    x_even = ((uint64_t)c[1] << 32) | c[0];
    x_odd = ((uint64_t)c[3] << 32) | c[2];
}

__host__ __device__ inline int64_t philox_below(uint64_t seed, uint64_t stream, uint64_t draw, uint64_t n) {
    uint64_t xe, xo;
    philox_pair(seed, stream, draw >> 1, xe, xo);
    return mulhi_below((draw & 1) ? xo : xe, n);
}

struct PhiloxArgs {
    lec_sampler_graph g;  // device pointers
    const void* u; const void* v;
    void* neg_to; void* neg_from;
    int64_t B; int N; int idx_bytes;
    uint64_t seed, stream;
    int* status;
};

// One side of an (i, p) slot, 32-bit arithmetic throughout (node ids and list lengths are < 2^31): the excluded ids of
// `node` inside the draw window, then the r-th candidate.  Without per-level windows the whole list counts and the two
// window searches disappear.
struct SideDraw { const int32_t* ex; int m; int lo; int cnt; };

__device__ __forceinline__ SideDraw side_setup(const lec_sampler_graph& g, const int64_t* __restrict__ ptr,
                                               const int32_t* __restrict__ excl, int p, int64_t node) {
    const int64_t e0 = __ldg(ptr + node), e1 = __ldg(ptr + node + 1);
    SideDraw d;
    if (!g.pick_per_level) {
        d.ex = excl + e0; d.m = (int)(e1 - e0); d.lo = 0; d.cnt = (int)g.n_nodes - d.m;
        return d;
    }
    const Window w = draw_window(&g, p, node);
    const int64_t a = lower_bound_i32(excl, e0, e1, w.lo);
    const int64_t b = lower_bound_i32(excl, a, e1, w.hi);
    d.ex = excl + a; d.m = (int)(b - a); d.lo = (int)w.lo; d.cnt = (int)(w.hi - w.lo) - d.m;
    return d;
}

template <typename I>
__global__ void __launch_bounds__(256) sample_philox_kernel(PhiloxArgs a) {
    const int64_t total = a.B * a.N;     // (i, p) slots; a thread makes the slot's row draw and its column draw
    for (int64_t ip = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ip < total; ip += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = ip / a.N;
        const int p = (int)(ip - i * a.N);
        const int64_t nu = (int64_t)((const I*)a.u)[i], nv = (int64_t)((const I*)a.v)[i];
        if (nu < 0 || nu >= a.g.n_nodes || nv < 0 || nv >= a.g.n_nodes) { atomicExch(a.status, LEC_E_INDEX); continue; }
        const SideDraw du = side_setup(a.g, a.g.row_excl_ptr, a.g.row_excl, p, nu);
        const SideDraw dv = side_setup(a.g, a.g.col_excl_ptr, a.g.col_excl, p, nv);
        if (du.cnt <= 0 || dv.cnt <= 0) { atomicExch(a.status, LEC_E_EMPTY); continue; }
        uint64_t xe, xo;
        philox_pair(a.seed, a.stream, (uint64_t)ip, xe, xo);   // draw ids 2 ip (row) and 2 ip + 1 (column)
        const int ru = (int)mulhi_below(xe, (uint64_t)du.cnt), rv = (int)mulhi_below(xo, (uint64_t)dv.cnt);
        // r-th candidate = lo + r + j, j = the smallest index with j == m or ex[j] - lo - j > r.  The two searches run in
        // lock step (independent loads in flight together), branch-free halving.
        int bu = 0, lu = du.m, bv = 0, lv = dv.m;
        while ((lu | lv) > 0) {
            const int hu = lu >> 1, hv = lv >> 1;
            const int mu = bu + hu, mv = bv + hv;
            const int eu = lu > 0 ? __ldg(du.ex + mu) : 0, ev = lv > 0 ? __ldg(dv.ex + mv) : 0;
            const bool right_u = lu > 0 && !(eu - du.lo - mu > ru);   // keep searching above mid
            const bool right_v = lv > 0 && !(ev - dv.lo - mv > rv);
            bu = right_u ? mu + 1 : bu;  lu = right_u ? lu - hu - 1 : hu;
            bv = right_v ? mv + 1 : bv;  lv = right_v ? lv - hv - 1 : hv;
        }
        ((I*)a.neg_to)[ip] = (I)(du.lo + ru + bu);
        ((I*)a.neg_from)[ip] = (I)(dv.lo + rv + bv);
    }
}

}  // namespace

extern "C" int64_t lec_philox_below(uint64_t seed, uint64_t stream, uint64_t draw, uint64_t n) {
    if (n == 0) return LEC_E_EMPTY;
    return philox_below(seed, stream, draw, n);
}

extern "C" int lec_sample_negatives_philox(const lec_sampler_graph* g_dev, const void* u, const void* v, int idx_bytes,
                                           int64_t B, int N, uint64_t seed, uint64_t stream_id, void* neg_to,
                                           void* neg_from, int* status, void* stream) {
    if (!g_dev || !g_dev->row_excl_ptr || !g_dev->row_excl || !g_dev->col_excl_ptr || !g_dev->col_excl) return LEC_E_NULL;
    if (!u || !v || !neg_to || !neg_from || !status) return LEC_E_NULL;
    if (g_dev->n_nodes < 1 || g_dev->n_nodes > 0x7fffffffLL || B < 0 || N < 0) return LEC_E_SIZE;
    if (g_dev->pick_per_level && (g_dev->n_levels < 1 || g_dev->n_levels > LEC_MAX_LEVELS || g_dev->level_mod < 1))
        return LEC_E_ENUM;
    if (idx_bytes == 2 && g_dev->n_nodes > 65536) return LEC_E_ENUM;
    const int64_t total = B * N;   // one thread per (i, p) slot: its row draw and its column draw
    if (total == 0) return 0;
    PhiloxArgs a{*g_dev, u, v, neg_to, neg_from, B, N, idx_bytes, seed, stream_id, status};
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    cudaStream_t st = (cudaStream_t)stream;
    switch (idx_bytes) {
        case 2: sample_philox_kernel<uint16_t><<<(int)blocks, 256, 0, st>>>(a); break;
        case 4: sample_philox_kernel<int32_t><<<(int)blocks, 256, 0, st>>>(a); break;
        case 8: sample_philox_kernel<int64_t><<<(int)blocks, 256, 0, st>>>(a); break;
        default: return LEC_E_ENUM;
    }
    ++lec::g_launches;
    return (int)cudaGetLastError();
}
