// eg_adam.cu -- fused Adam update of one parameter tensor ("next").
// Mirrors torch.optim.Adam (no weight decay, no amsgrad) as configured at
// /root/reference/edgegaussians/utils/train_utils.py:48-65 (4 independent Adams):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// with bc1 = 1 - b1^t, bc2 = 1 - b2^t supplied by the host.  One pass: 16 B read + 12 B written per
// element (+ optional zeroing of the gradient, replacing optimizer.zero_grad()).
#include "eg_common.cuh"

namespace {

__global__ void __launch_bounds__(256) adam_kernel(long long n, float *__restrict__ p, float *__restrict__ g,
                                                   float *__restrict__ m, float *__restrict__ v, float step_size,
                                                   float b1, float omb1, float b2, float omb2, float eps,
                                                   float sqrt_bc2, int zero_grad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float m0 = m[i];
    const float mi = m0 + omb1 * (gi - m0);            // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = b2 * v[i] + omb2 * gi * gi;       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * (mi / (sqrtf(vi) / sqrt_bc2 + eps));  // param.addcdiv_(exp_avg, denom, value=-step_size)
    if (zero_grad) g[i] = 0.0f;
}

}  // namespace

extern "C" int eg_adam_step(int64_t n, float *param, float *grad, float *exp_avg, float *exp_avg_sq, double lr,
                            double beta1, double beta2, double eps, double bias_correction1, double bias_correction2,
                            int zero_grad, void *stream) {
    if (n <= 0) return 0;
    if (!(bias_correction1 > 0.0) || !(bias_correction2 > 0.0)) {
        eg_set_error("eg_adam_step: bias corrections must be positive");
        return 1;
    }
    const int block = 256;
    const long long grid = (n + block - 1) / block;
    adam_kernel<<<(unsigned)grid, block, 0, (cudaStream_t)stream>>>(n, param, grad, exp_avg, exp_avg_sq,
                                                                   (float)(lr / bias_correction1), (float)beta1,
                                                                   (float)(1.0 - beta1), (float)beta2,
                                                                   (float)(1.0 - beta2), (float)eps,
                                                                   (float)sqrt(bias_correction2), zero_grad);
    return eg_check_launch("eg_adam_step");
}

// ---------------------------------------------------------------------------------------------------------
// eg_adam_multi -- ONE launch for the reference's four Adams (utils/train_utils.py:48-65; stepped back to back at
// train_gaussians.py:104-106 and, for means / scales / quats only, at :118-121 and :128-131).  The gradients are
// the segments of the fused step's flat buffer (eg_grad_layout); each segment has its own parameter tensor, moment
// tensors, learning rate and step count.  Hyper-parameters that change from step to step (lr, step count, enabled
// flag) are read from DEVICE memory, so a captured CUDA graph keeps working while the schedulers move the learning
// rates; the step counts are advanced by the kernel itself (by the last CTA to finish).
// ---------------------------------------------------------------------------------------------------------
namespace {

constexpr int AM_MAX_SEGS = 8;

struct AdamSegs {
    float *param[AM_MAX_SEGS], *m[AM_MAX_SEGS], *v[AM_MAX_SEGS];
    long long goff[AM_MAX_SEGS], count[AM_MAX_SEGS], first4[AM_MAX_SEGS + 1];  // first4: prefix of ceil(count / 4)
    int n;
};

// hyper [n_segs] on the device: lr (f64) | step (i64, completed steps) | enabled (i64)
struct AdamHyper {
    double lr;
    long long step, enabled;
};

__global__ void __launch_bounds__(256) adam_multi_kernel(const AdamSegs segs, float *__restrict__ grads,
                                                         AdamHyper *__restrict__ hyper, const double beta1,
                                                         const double beta2, const double eps, const int zero_grad,
                                                         unsigned int *__restrict__ ticket) {
    __shared__ float s_step[AM_MAX_SEGS], s_sqbc2[AM_MAX_SEGS];
    __shared__ int s_on[AM_MAX_SEGS];
    if ((int)threadIdx.x < segs.n) {
        const AdamHyper h = hyper[threadIdx.x];
        const double t = (double)(h.step + 1);
        // torch: step_size = lr / (1 - beta1^t), denom = sqrt(v) / sqrt(1 - beta2^t) + eps -- doubles, rounded to
        // fp32 only when they meet the tensors
        s_step[threadIdx.x] = (float)(h.lr / (1.0 - pow(beta1, t)));
        s_sqbc2[threadIdx.x] = (float)sqrt(1.0 - pow(beta2, t));
        s_on[threadIdx.x] = h.enabled != 0;
    }
    __syncthreads();
    const float b2 = (float)beta2, omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2), epsf = (float)eps;
    const long long total4 = segs.first4[segs.n];
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += (long long)gridDim.x * blockDim.x) {
        int s = 0;
#pragma unroll
        for (int k = 1; k < AM_MAX_SEGS; ++k)
            if (k < segs.n && q >= segs.first4[k]) s = k;
        if (!s_on[s]) continue;
        const long long e0 = 4 * (q - segs.first4[s]);
        const int cnt = (int)min(4ll, segs.count[s] - e0);
        float *p = segs.param[s] + e0, *m = segs.m[s] + e0, *v = segs.v[s] + e0, *g = grads + segs.goff[s] + e0;
        float gi[4], mi[4], vi[4], pi[4];
        // parameter / moment tensors are torch allocations (16-byte aligned bases), the gradient segments are
        // 16-byte aligned by construction: full quads move as 128-bit accesses
        const bool vec = cnt == 4 && ((((uintptr_t)p) | ((uintptr_t)m) | ((uintptr_t)v) | ((uintptr_t)g)) & 15) == 0;
        if (vec) {
            const float4 g4 = *reinterpret_cast<const float4 *>(g), m4 = *reinterpret_cast<const float4 *>(m),
                         v4 = *reinterpret_cast<const float4 *>(v), p4 = *reinterpret_cast<const float4 *>(p);
            gi[0] = g4.x; gi[1] = g4.y; gi[2] = g4.z; gi[3] = g4.w;
            mi[0] = m4.x; mi[1] = m4.y; mi[2] = m4.z; mi[3] = m4.w;
            vi[0] = v4.x; vi[1] = v4.y; vi[2] = v4.z; vi[3] = v4.w;
            pi[0] = p4.x; pi[1] = p4.y; pi[2] = p4.z; pi[3] = p4.w;
        } else {
            for (int k = 0; k < cnt; ++k) { gi[k] = g[k]; mi[k] = m[k]; vi[k] = v[k]; pi[k] = p[k]; }
        }
        const float step_size = s_step[s], sq = s_sqbc2[s];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < cnt) {
                mi[k] = mi[k] + omb1 * (gi[k] - mi[k]);
                vi[k] = b2 * vi[k] + omb2 * gi[k] * gi[k];
                pi[k] -= step_size * (mi[k] / (sqrtf(vi[k]) / sq + epsf));
            }
        }
        if (vec) {
            *reinterpret_cast<float4 *>(m) = make_float4(mi[0], mi[1], mi[2], mi[3]);
            *reinterpret_cast<float4 *>(v) = make_float4(vi[0], vi[1], vi[2], vi[3]);
            *reinterpret_cast<float4 *>(p) = make_float4(pi[0], pi[1], pi[2], pi[3]);
            if (zero_grad) *reinterpret_cast<float4 *>(g) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int k = 0; k < cnt; ++k) {
                m[k] = mi[k]; v[k] = vi[k]; p[k] = pi[k];
                if (zero_grad) g[k] = 0.0f;
            }
        }
    }
    // the last CTA to finish advances the step counts (every CTA has read them by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            for (int s = 0; s < segs.n; ++s)
                if (hyper[s].enabled != 0) hyper[s].step += 1;
            *ticket = 0u;
        }
    }
}

struct GatherArrays {
    const float *src[16];
    float *dst[16];
    int width[16];
    long long zero_from[16];
    int n;
};

// dst[a][r, :] = r < zero_from[a] ? src[a][idx[r], :] : 0   for every array a: one launch moves parameters, both
// Adam moments and the abs-grad statistic through a cull (idx = surviving rows) or a duplication (idx = old rows
// followed by the duplicated ones, whose moments start at zero)
__global__ void __launch_bounds__(256) gather_rows_kernel(const long long n_out, const int *__restrict__ idx,
                                                          const GatherArrays arr) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_out) return;
    const long long src_row = __ldg(idx + r);
    for (int a = 0; a < arr.n; ++a) {
        const int w = arr.width[a];
        float *d = arr.dst[a] + r * w;
        if (r >= arr.zero_from[a]) {
            for (int k = 0; k < w; ++k) d[k] = 0.0f;
        } else {
            const float *s = arr.src[a] + src_row * w;
            for (int k = 0; k < w; ++k) d[k] = __ldg(s + k);
        }
    }
}

}  // namespace

extern "C" int eg_adam_multi(int n_segs, const eg_adam_segment *segs, float *grads, void *hyper, double beta1,
                             double beta2, double eps, int zero_grad, uint32_t *ticket, void *stream) {
    if (n_segs < 1 || n_segs > AM_MAX_SEGS || segs == nullptr || grads == nullptr || hyper == nullptr || ticket == nullptr) {
        eg_set_error("eg_adam_multi: bad arguments (1..%d segments; grads, hyper and ticket are required)", AM_MAX_SEGS);
        return 1;
    }
    AdamSegs a;
    a.n = n_segs;
    a.first4[0] = 0;
    for (int s = 0; s < n_segs; ++s) {
        if (segs[s].param == nullptr || segs[s].exp_avg == nullptr || segs[s].exp_avg_sq == nullptr || segs[s].count < 0) {
            eg_set_error("eg_adam_multi: segment %d incomplete", s);
            return 1;
        }
        a.param[s] = segs[s].param;
        a.m[s] = segs[s].exp_avg;
        a.v[s] = segs[s].exp_avg_sq;
        a.goff[s] = segs[s].grad_offset;
        a.count[s] = segs[s].count;
        a.first4[s + 1] = a.first4[s] + (segs[s].count + 3) / 4;
    }
    const long long total4 = a.first4[n_segs];
    if (total4 == 0) return 0;
    long long grid = (total4 + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    adam_multi_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(a, grads, (AdamHyper *)hyper, beta1, beta2, eps,
                                                                       zero_grad, ticket);
    return eg_check_launch("eg_adam_multi");
}

extern "C" int eg_gather_rows(int64_t n_out, const int32_t *idx, int n_arrays, const eg_row_array *arrays, void *stream) {
    if (n_out < 0 || n_arrays < 0 || n_arrays > 16 || (n_arrays > 0 && arrays == nullptr) || (n_out > 0 && idx == nullptr)) {
        eg_set_error("eg_gather_rows: bad arguments (at most 16 arrays)");
        return 1;
    }
    if (n_out == 0 || n_arrays == 0) return 0;
    GatherArrays g;
    g.n = n_arrays;
    for (int a = 0; a < n_arrays; ++a) {
        if (arrays[a].dst == nullptr || arrays[a].width <= 0 || (arrays[a].src == nullptr && arrays[a].zero_from_row > 0)) {
            eg_set_error("eg_gather_rows: array %d incomplete", a);
            return 1;
        }
        g.src[a] = arrays[a].src;
        g.dst[a] = arrays[a].dst;
        g.width[a] = arrays[a].width;
        g.zero_from[a] = arrays[a].zero_from_row < 0 ? n_out : arrays[a].zero_from_row;
    }
    gather_rows_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_out, idx, g);
    return eg_check_launch("eg_gather_rows");
}
