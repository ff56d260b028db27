// Store-pattern ceiling of the label-major score matrix [L, N]: how fast can 148 SMs write it with the tile shapes of
// the scoring kernels, without any arithmetic?   nvcc -arch=sm_100a -O3 -o scripts/bin/ubench_store_pattern scripts/ubench_store_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_fill(float4* p, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
// tile = IMG images x all L labels per CTA; warp stores 32 consecutive images of one label row (128 B)
template <int IMG, int HINT>
__global__ void k_label_major(float* s, int L, size_t N) {
    const int warps = blockDim.x / 32, w = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int img_warps = IMG / 32;             // warps side by side along the image axis
    const int lab_split = warps / img_warps;    // warps stacked along the label axis
    const size_t img = (size_t)blockIdx.x * IMG + (w % img_warps) * 32 + lane;
    if (img >= N) return;
    for (int l = w / img_warps; l < L; l += lab_split) {
        float* q = s + (size_t)l * N + img;
        const float v = (float)l;
        if (HINT == 0) *q = v;
        else if (HINT == 1) asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(q), "f"(v) : "memory");
        else asm volatile("st.global.wt.f32 [%0], %1;" ::"l"(q), "f"(v) : "memory");
    }
}
// same tile, but each thread stores a float4 = 4 consecutive images (a warp covers 128 images = 512 B of a row)
__global__ void k_label_major_v4(float* s, int L, size_t N) {
    const int warps = blockDim.x / 32, w = threadIdx.x / 32, lane = threadIdx.x % 32;
    const size_t img = (size_t)blockIdx.x * 128 + lane * 4;
    if (img + 3 >= N) return;
    for (int l = w; l < L; l += warps) *reinterpret_cast<float4*>(s + (size_t)l * N + img) = make_float4(l, l, l, l);
}

template <typename F>
static void run(const char* name, size_t bytes, F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e9f;
    for (int i = 0; i < 5; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    printf("%-44s %8.3f ms  %7.0f GB/s written  (%s)\n", name, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int L = 723; const size_t N = 1000064;   // multiple of 128
    const size_t n = (size_t)L * N, bytes = n * 4;
    float* s; cudaMalloc(&s, bytes);
    run("fill float4 grid-stride", bytes, [&] { k_fill<<<148 * 8, 256>>>((float4*)s, n / 4); });
    run("label-major 128 img/CTA, 256 thr", bytes, [&] { k_label_major<128, 0><<<(unsigned)(N / 128), 256>>>(s, L, N); });
    run("label-major 128 img/CTA, 256 thr, .cs", bytes, [&] { k_label_major<128, 1><<<(unsigned)(N / 128), 256>>>(s, L, N); });
    run("label-major 128 img/CTA, 256 thr, .wt", bytes, [&] { k_label_major<128, 2><<<(unsigned)(N / 128), 256>>>(s, L, N); });
    run("label-major 256 img/CTA, 256 thr", bytes, [&] { k_label_major<256, 0><<<(unsigned)(N / 256), 256>>>(s, L, N); });
    run("label-major 128 img/CTA, 512 thr", bytes, [&] { k_label_major<128, 0><<<(unsigned)(N / 128), 512>>>(s, L, N); });
    run("label-major 128 img/CTA float4/thread", bytes, [&] { k_label_major_v4<<<(unsigned)(N / 128), 256>>>(s, L, N); });
    cudaMemset(s, 0, bytes);
    run("cudaMemset", bytes, [&] { cudaMemsetAsync(s, 0, bytes); });
    return 0;
}
