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
