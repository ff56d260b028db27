// Issue-rate microbenchmarks for the B200 SM (sm_100a): which instruction mixes reach 1 warp-instruction
// per clock per SM sub-partition.  Results decide the layout of the scoring kernel (packed FFMA2 or not,
// how expensive MUFU / FMNMX / predicated STS are next to the FMA pipe).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_pipes scripts/ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float lo(unsigned long long v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a + b;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// 0: FFMA x8 chains
__global__ void k_ffma(float* out, float s) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    float b = s, c = s * 0.5f;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// 1: FFMA2 x8 chains (16 fp32 FMAs per iteration)
__global__ void k_ffma2(float* out, float s) {
    unsigned long long a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pk(threadIdx.x * 0.001f + i, i);
    unsigned long long b = pk(s, s), c = pk(s * 0.5f, s);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], b, c);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += lo(a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// 2: FMNMX x8 chains
__global__ void k_fmnmx(float* out, float s) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    float b = s, c = -s;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fminf(fmaxf(a[i], c), b), c += 0.f;
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// 3: MUFU.RSQ x8 chains
__global__ void k_mufu(float* out, float s) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i + 1.f;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + s;
}
// 4: FFMA + FMNMX interleaved 1:1 (8 + 8 per iteration)
__global__ void k_mix_fma_alu(float* out, float s) {
    float a[8], m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; m[i] = a[i] + 1.f; }
    float b = s, c = s * 0.5f;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], b, c); m[i] = fminf(m[i], b); b += 0.f; }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i] + m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// 5: FFMA2 + FMNMX interleaved 1:1
__global__ void k_mix_fma2_alu(float* out, float s) {
    unsigned long long a[8];
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = pk(threadIdx.x * 0.001f + i, i); m[i] = i + 1.f + threadIdx.x; }
    unsigned long long b = pk(s, s), c = pk(s * 0.5f, s);
    float bb = s;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] = fma2(a[i], b, c); m[i] = fminf(m[i], bb); bb += 0.f; }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += lo(a[i]) + m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// 6: FFMA2 x6 + MUFU x1 + FMNMX x2 per group (the scoring mix)
__global__ void k_mix_score(float* out, float s) {
    unsigned long long a[6];
    float m[2], q[2];
#pragma unroll
    for (int i = 0; i < 6; ++i) a[i] = pk(threadIdx.x * 0.001f + i, i);
    m[0] = threadIdx.x; m[1] = 2.f; q[0] = 1.f + threadIdx.x; q[1] = 3.f;
    unsigned long long b = pk(s, s), c = pk(s * 0.5f, s);
    float bb = s;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
            for (int i = 0; i < 6; ++i) a[i] = fma2(a[i], b, c);
            asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(q[rep]));
            m[0] = fminf(m[0], bb); m[1] = fmaxf(m[1], bb); bb += 0.f;
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) r += lo(a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + m[0] + m[1] + q[0] + q[1];
}
// 7: broadcast LDS.128 feeding 2 FFMA2 each (label row read pattern)
__global__ void k_lds_ffma2(float* out, float s) {
    __shared__ __align__(16) float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = s + i * 1e-6f;
    __syncthreads();
    unsigned long long a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pk(threadIdx.x * 0.001f + i, i);
    unsigned long long c = pk(s * 0.5f, s);
    for (int it = 0; it < ITER; ++it) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(sm) + (it & 63) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const ulonglong2 v = p[i];
            a[2 * i] = fma2(a[2 * i], v.x, c);
            a[2 * i + 1] = fma2(a[2 * i + 1], v.y, c);
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += lo(a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// 8: FFMA2 x8 + (FSETP + predicated STS.64 + predicated IADD) x2 : the candidate append next to the FMA stream
__global__ void k_append(float* out, float s, float thr) {
    extern __shared__ float2 buf[];  // [16][blockDim]
    unsigned long long a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pk(threadIdx.x * 0.001f + i, i);
    unsigned long long b = pk(s, s), c = pk(s * 0.5f, s);
    int cnt = 0;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], b, c);
        float e0, e1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(e0), "=f"(e1) : "l"(a[0]));
        if (e0 < thr) { buf[(cnt & 15) * blockDim.x + threadIdx.x] = make_float2(e0, (float)it); ++cnt; }
        if (e1 < thr) { buf[(cnt & 15) * blockDim.x + threadIdx.x] = make_float2(e1, (float)it); ++cnt; }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += lo(a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + cnt + buf[threadIdx.x].x;
}
// 9: FMUL2/FADD2 alternating
__global__ void k_muladd2(float* out, float s) {
    unsigned long long a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pk(threadIdx.x * 0.001f + i, i);
    unsigned long long b = pk(s, s);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = (i & 1) ? mul2(a[i], b) : add2(a[i], b);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += lo(a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// 10: DFMA x8 chains
__global__ void k_dfma(float* out, float s) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001 + i;
    double b = s, c = s * 0.5;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)r;
}

// 11: FFMA2 with three distinct register-pair operands per instruction (no operand reuse possible)
__global__ void k_ffma2_3op(float* out, float s) {
    unsigned long long a[8], y[8], x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = pk(threadIdx.x * 0.001f + i, i); y[i] = pk(s + i + threadIdx.x, s - i * threadIdx.x); x[i] = pk(s * i + threadIdx.x * 0.5f, s + 2 * i - threadIdx.x); }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma2(y[i], x[i], a[i]);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += lo(a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// 12: FFMA2, one operand shared by consecutive instructions (acc[i] = y[i] * x0 + acc[i])
__global__ void k_ffma2_reuse(float* out, float s) {
    unsigned long long a[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = pk(threadIdx.x * 0.001f + i, i); y[i] = pk(s + i + threadIdx.x, s - i * threadIdx.x); }
    unsigned long long x0 = pk(s + threadIdx.x, s * 0.5f - threadIdx.x);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma2(y[i], x0, a[i]);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += lo(a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// 13: scalar FFMA with three distinct register operands
__global__ void k_ffma_3op(float* out, float s) {
    float a[8], y[8], x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; y[i] = s + i + threadIdx.x; x[i] = s * i - threadIdx.x; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(y[i]), "f"(x[i]));
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// 14: scalar FFMA, one operand shared
__global__ void k_ffma_reuse(float* out, float s) {
    float a[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; y[i] = s + i * threadIdx.x; }
    float x0 = s - threadIdx.x;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(y[i]), "f"(x0));
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// 15: FMNMX.NAN clamp pairs
__global__ void k_fmnmx_nan(float* out, float s) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("max.NaN.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(-s));
            asm volatile("min.NaN.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s));
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
static void run(const char* name, F launch, double warp_instr_per_thread_iter, double flops_per_thread_iter, int threads, int blocks_per_sm) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int grid = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(grid, threads);
    launch(grid, threads);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        launch(grid, threads);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    const double warps = (double)grid * threads / 32.0;
    const double winstr = warps * ITER * warp_instr_per_thread_iter;
    const double cycles = best * 1e-3 * clk_khz * 1e3;
    printf("%-28s thr=%4d bps=%d  %8.3f ms  %6.3f warp-instr/clk/SMSP (at %d MHz nominal)  %7.2f TFLOP/s  %s\n", name, threads,
           blocks_per_sm, best, winstr / cycles / sms / 4.0, clk_khz / 1000, (double)grid * threads * ITER * flops_per_thread_iter / (best * 1e-3) / 1e12,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    cudaFuncSetAttribute(k_append, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 8);
    for (int threads : {256}) {
        for (int bps : {4}) {
            if (threads * bps > 2048) continue;
            run("FFMA x8", [&](int g, int t) { k_ffma<<<g, t>>>(out, 1.0001f); }, 8, 16, threads, bps);
            run("FFMA2 x8", [&](int g, int t) { k_ffma2<<<g, t>>>(out, 1.0001f); }, 8, 32, threads, bps);
            run("FMNMX x16", [&](int g, int t) { k_fmnmx<<<g, t>>>(out, 1.0001f); }, 16, 0, threads, bps);
            run("MUFU.RSQ x8", [&](int g, int t) { k_mufu<<<g, t>>>(out, 1.0001f); }, 8, 0, threads, bps);
            run("FFMA+FMNMX 8+8", [&](int g, int t) { k_mix_fma_alu<<<g, t>>>(out, 1.0001f); }, 16, 16, threads, bps);
            run("FFMA2+FMNMX 8+8", [&](int g, int t) { k_mix_fma2_alu<<<g, t>>>(out, 1.0001f); }, 16, 32, threads, bps);
            run("score mix 12F2+2MUFU+4MNMX", [&](int g, int t) { k_mix_score<<<g, t>>>(out, 1.0001f); }, 18, 48, threads, bps);
            run("LDS.128 x4 + FFMA2 x8", [&](int g, int t) { k_lds_ffma2<<<g, t>>>(out, 1.0001f); }, 12, 32, threads, bps);
            run("FFMA2 x8 + append x2", [&](int g, int t) { k_append<<<g, t, 16 * t * 8>>>(out, 1.0001f, 0.5f); }, 14, 32, threads, bps);
            run("FMUL2/FADD2 x8", [&](int g, int t) { k_muladd2<<<g, t>>>(out, 1.0001f); }, 8, 16, threads, bps);
            run("FFMA2 3 distinct operands", [&](int g, int t) { k_ffma2_3op<<<g, t>>>(out, 1.0001f); }, 8, 32, threads, bps);
            run("FFMA2 one operand reused", [&](int g, int t) { k_ffma2_reuse<<<g, t>>>(out, 1.0001f); }, 8, 32, threads, bps);
            run("FFMA 3 distinct operands", [&](int g, int t) { k_ffma_3op<<<g, t>>>(out, 1.0001f); }, 8, 16, threads, bps);
            run("FFMA one operand reused", [&](int g, int t) { k_ffma_reuse<<<g, t>>>(out, 1.0001f); }, 8, 16, threads, bps);
            run("FMNMX.NAN x16", [&](int g, int t) { k_fmnmx_nan<<<g, t>>>(out, 1.0001f); }, 16, 0, threads, bps);
            run("DFMA x8", [&](int g, int t) { k_dfma<<<g, t>>>(out, 1.0001f); }, 8, 16, threads, bps);
        }
    }
    return 0;
}
