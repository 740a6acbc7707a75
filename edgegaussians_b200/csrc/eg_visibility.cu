// eg_visibility.cu -- batched visibility filter ("next"): for every Gaussian the
// fraction of views in which its mean projects inside the image AND onto an edge pixel.
//
// Replaces the per-view CPU loop of cull_gaussians_not_projecting,
// /root/reference/edgegaussians/models/edge_gs.py:578-601 (D2H of the means, one [N,4] x [4,3] matmul, round,
// mask gather and a [N,V] bool matrix per call), up to its threshold `mean < min_projecting_fraction`.
// Reference semantics kept: P = K @ viewmat[:3,:4] in fp32, pixel = round-half-even(P X / (P X)_z) with NO depth
// test (a point behind the camera may still land inside), mean over ALL views.
// mode 1 serves the post-processing filter filter_by_projection (edge_extraction/filtering.py:80-123): the same
// projection, but the MEAN over the views of the edge map's value (uint8 / 255) at the pixel instead of a mask hit rate.
// One thread per Gaussian, the V projection matrices staged in shared memory; 12 B read per Gaussian plus one
// byte gather per (Gaussian, view) from the L2-resident masks.
#include "eg_common.cuh"

namespace {

__global__ void __launch_bounds__(256) projecting_fraction_kernel(
    const int n, const float *__restrict__ means, const int n_views, const float *__restrict__ viewmats,
    const float *__restrict__ Ks, const int32_t *__restrict__ sizes, const unsigned char *__restrict__ masks,
    const long long *__restrict__ mask_offsets, const int mode, float *__restrict__ fraction) {
    extern __shared__ float sP[];  // [V][12]
    for (int e = threadIdx.x; e < 12 * n_views; e += blockDim.x) {
        const int v = e / 12, i = (e % 12) / 4, j = e % 4;
        const float *K = Ks + 9 * v, *vm = viewmats + 16 * v;
        sP[e] = __fadd_rn(__fadd_rn(__fmul_rn(K[3 * i], vm[j]), __fmul_rn(K[3 * i + 1], vm[4 + j])),
                          __fmul_rn(K[3 * i + 2], vm[8 + j]));
    }
    __syncthreads();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const float x = __ldg(means + 3 * g), y = __ldg(means + 3 * g + 1), z = __ldg(means + 3 * g + 2);
    int count = 0;  // mode 0: views whose mask is set at the projection; mode 1: sum of the u8 edge values there
    for (int v = 0; v < n_views; ++v) {
        const float *P = sP + 12 * v;
        float pr[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
            pr[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[4 * i], x), __fmul_rn(P[4 * i + 1], y)),
                                        __fmul_rn(P[4 * i + 2], z)), P[4 * i + 3]);
        const float u = rintf(__fdiv_rn(pr[0], pr[2])), w = rintf(__fdiv_rn(pr[1], pr[2]));
        const int W = __ldg(sizes + 2 * v), H = __ldg(sizes + 2 * v + 1);
        if (u >= 0.0f && u < (float)W && w >= 0.0f && w < (float)H) {  // NaN / inf fail
            const int m = __ldg(masks + __ldg(mask_offsets + v) + (long long)w * W + (long long)u);
            count += mode == 0 ? (m != 0) : m;
        }
    }
    // mode 1: the integer sum of the edge values (<= 255 * 1024 views: exact in fp32); the caller divides by 255 V
    fraction[g] = mode == 0 ? (float)count / (float)n_views : (float)count;
}

}  // namespace

extern "C" int eg_projecting_fraction(int n, const float *means, int n_views, const float *viewmats, const float *Ks,
                                      const int32_t *sizes, const uint8_t *masks, const int64_t *mask_offsets,
                                      int mode, float *fraction, void *stream) {
    if (n < 0 || n_views <= 0 || means == nullptr || viewmats == nullptr || Ks == nullptr || sizes == nullptr ||
        masks == nullptr || mask_offsets == nullptr || fraction == nullptr) {
        eg_set_error("eg_projecting_fraction: bad arguments");
        return 1;
    }
    if (mode != 0 && mode != 1) {
        eg_set_error("eg_projecting_fraction: mode must be 0 (mask hit fraction) or 1 (mean edge value)");
        return 1;
    }
    if (n == 0) return 0;
    const size_t smem = (size_t)n_views * 12 * sizeof(float);
    if (smem > 48 * 1024) {
        eg_set_error("eg_projecting_fraction: at most %d views per call", (int)(48 * 1024 / (12 * sizeof(float))));
        return 1;
    }
    projecting_fraction_kernel<<<(n + 255) / 256, 256, smem, (cudaStream_t)stream>>>(
        n, means, n_views, viewmats, Ks, sizes, masks, (const long long *)mask_offsets, mode, fraction);
    return eg_check_launch("eg_projecting_fraction");
}
