// Microbenchmark (not product code): throughput of the memory primitives a Gaussian-major splat would use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_gather red_gather.cu
// Pattern: "row lanes" -- lane i owns a random (x0, y) in a 1600x1200 fp32 image and touches CH aligned
// float4 chunks of that row (what a (Gaussian,row) work item does with a ~9 px span).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hash(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ void red_v4(float *a, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void red_s(float *a, float x) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a), "f"(x) : "memory");
}

// items are grouped 12 consecutive rows per "Gaussian" at a random position
template <int MODE, int CH>
__global__ void k(float *img, int W, int H, long long n_items, float *sink) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    unsigned g = (unsigned)(i / 12), r = (unsigned)(i % 12);
    unsigned h = hash(g * 2654435761u + 12345u);
    int x0 = (h % (W - 32)) & ~3, y = (hash(h) % (H - 12)) + r;
    float *p = img + (long long)y * W + x0;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        if (MODE == 0) red_v4(p + 4 * c, 1.f, 2.f, 3.f, 4.f);
        if (MODE == 1) { red_s(p + 4 * c, 1.f); red_s(p + 4 * c + 1, 1.f); red_s(p + 4 * c + 2, 1.f); red_s(p + 4 * c + 3, 1.f); }
        if (MODE == 2) { float4 v = __ldg(reinterpret_cast<const float4 *>(p + 4 * c)); acc += v.x + v.y + v.z + v.w; }
        if (MODE == 3) { acc += __ldg(p + 4 * c) + __ldg(p + 4 * c + 1) + __ldg(p + 4 * c + 2) + __ldg(p + 4 * c + 3); }
    }
    if (MODE >= 2 && acc == 123.456f) *sink = acc;
}

template <int MODE, int CH>
void run(const char *name, float *img, int W, int H, long long n_items, float *sink) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int thr = 256; int grid = (int)((n_items + thr - 1) / thr);
    for (int it = 0; it < 3; ++it) k<MODE, CH><<<grid, thr>>>(img, W, H, n_items, sink);
    cudaEventRecord(a);
    const int R = 10;
    for (int it = 0; it < R; ++it) k<MODE, CH><<<grid, thr>>>(img, W, H, n_items, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s items %lld chunks %d : %8.2f us/launch  (%.2f G lane-ops/s)\n", name, n_items, CH, ms * 1000 / R,
           n_items * CH / (ms / R * 1e-3) * 1e-9);
}

int main() {
    const int W = 1600, H = 1200;
    float *img, *sink; cudaMalloc(&img, sizeof(float) * W * H); cudaMalloc(&sink, 4);
    cudaMemset(img, 0, sizeof(float) * W * H);
    const long long n = 5354231;  // (Gaussian,row) items of the 500k init scene
    run<0, 3>("red.v4 x3/row", img, W, H, n, sink);
    run<0, 1>("red.v4 x1/row", img, W, H, n, sink);
    run<1, 3>("red.f32 x12/row", img, W, H, n, sink);
    run<2, 3>("ldg.128 x3/row", img, W, H, n, sink);
    run<3, 3>("ldg.32 x12/row", img, W, H, n, sink);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
