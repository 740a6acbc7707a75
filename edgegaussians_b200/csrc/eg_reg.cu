// eg_reg.cu -- a10 + a11: edge-direction and anisotropy regularisers, forward and backward fused.
//
// Semantics (reference-owned code, pinned by tests/golden/regularisers.npz):
//   compute_direction_loss  /root/reference/edgegaussians/models/edge_gs.py:346-373
//   quats_to_rotmats_tensor /root/reference/edgegaussians/utils/misc_utils.py:53-86
//   compute_ratio_loss      /root/reference/edgegaussians/models/edge_gs.py:375-380
// One thread per Gaussian; the gradient towards each neighbour's mean is scattered with atomics.
#include "eg_common.cuh"

namespace {

constexpr int REG_MAX_NN = 64;

__global__ void __launch_bounds__(128) reg_kernel(int n, const float *__restrict__ means,
                                                  const float *__restrict__ quats,
                                                  const float *__restrict__ log_scales,
                                                  const int32_t *__restrict__ nn, int cols, int k, int half,
                                                  float dir_w, float ratio_w, double *__restrict__ losses,
                                                  float *__restrict__ v_means, float *__restrict__ v_quats,
                                                  float *__restrict__ v_log_scales) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float dir_part = 0.0f, ratio_part = 0.0f;
    if (i < n) {
        const float ls0 = __ldg(log_scales + 3 * i), ls1 = __ldg(log_scales + 3 * i + 1), ls2 = __ldg(log_scales + 3 * i + 2);
        const float e[3] = {expf(ls0), expf(ls1), expf(ls2)};
        // ---- ratio loss: second largest / largest of exp(s) ----
        {
            int i0 = 0;
            if (e[1] > e[i0]) i0 = 1;
            if (e[2] > e[i0]) i0 = 2;
            int i1 = -1;
            for (int c = 0; c < 3; ++c)
                if (c != i0 && (i1 < 0 || e[c] > e[i1])) i1 = c;
            const float r = e[i1] / e[i0];
            ratio_part = r;
            if (v_log_scales != nullptr && ratio_w != 0.0f) {
                const float gsc = ratio_w * r / (float)n;
                v_log_scales[3 * i + i1] += gsc;
                v_log_scales[3 * i + i0] -= gsc;
            }
        }
        // ---- direction loss ----
        if (nn != nullptr && cols > 0) {
            int js = 0;  // torch.argmax: first maximal element
            if (fabsf(e[1]) > fabsf(e[js])) js = 1;
            if (fabsf(e[2]) > fabsf(e[js])) js = 2;
            const float4 q4 = __ldg(reinterpret_cast<const float4 *>(quats) + i);
            const float qn = fmaxf(sqrtf(q4.x * q4.x + q4.y * q4.y + q4.z * q4.z + q4.w * q4.w), 1e-12f);
            const float iqn = 1.0f / qn;
            const float w = q4.x * iqn, x = q4.y * iqn, y = q4.z * iqn, z = q4.w * iqn;
            float m[3];
            if (js == 0) { m[0] = 1.f - 2.f * (y * y + z * z); m[1] = 2.f * (x * y + w * z); m[2] = 2.f * (x * z - w * y); }
            else if (js == 1) { m[0] = 2.f * (x * y - w * z); m[1] = 1.f - 2.f * (x * x + z * z); m[2] = 2.f * (y * z + w * x); }
            else { m[0] = 2.f * (x * z + w * y); m[1] = 2.f * (y * z - w * x); m[2] = 1.f - 2.f * (x * x + y * y); }
            const float mu[3] = {__ldg(means + 3 * i), __ldg(means + 3 * i + 1), __ldg(means + 3 * i + 2)};
            float a[REG_MAX_NN];
            const int c_n = min(cols, REG_MAX_NN);
            for (int c = 0; c < c_n; ++c) {
                const int j = __ldg(nn + (long long)i * cols + c);
                const float vx = mu[0] - __ldg(means + 3 * j), vy = mu[1] - __ldg(means + 3 * j + 1), vz = mu[2] - __ldg(means + 3 * j + 2);
                const float inv = 1.0f / sqrtf(vx * vx + vy * vy + vz * vz);
                a[c] = (m[0] * vx + m[1] * vy + m[2] * vz) * inv;
            }
            const int denom = half ? k : c_n;
            float vmaj[3] = {0.f, 0.f, 0.f}, vself[3] = {0.f, 0.f, 0.f};
            for (int c = 0; c < c_n; ++c) {
                bool sel = true;
                if (half) {  // top-k of |a| (descending sort, first k)
                    int rank = 0;
                    const float ac = fabsf(a[c]);
                    for (int d = 0; d < c_n; ++d) {
                        const float ad = fabsf(a[d]);
                        rank += (ad > ac) || (ad == ac && d < c);
                    }
                    sel = rank < k;
                }
                if (!sel) continue;
                dir_part += fabsf(a[c]) / (float)denom;
                if (v_means == nullptr || dir_w == 0.0f) continue;
                const int j = __ldg(nn + (long long)i * cols + c);
                const float vx = mu[0] - __ldg(means + 3 * j), vy = mu[1] - __ldg(means + 3 * j + 1), vz = mu[2] - __ldg(means + 3 * j + 2);
                const float inv = 1.0f / sqrtf(vx * vx + vy * vy + vz * vz);
                const float d[3] = {vx * inv, vy * inv, vz * inv};
                const float sg = a[c] > 0.f ? 1.f : (a[c] < 0.f ? -1.f : 0.f);
                const float va = -sg * dir_w / ((float)denom * (float)n);
                float vv[3];
#pragma unroll
                for (int t = 0; t < 3; ++t) vmaj[t] += va * d[t];
                const float md = va * a[c];  // (v_d . d) with v_d = va * m
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    vv[t] = (va * m[t] - md * d[t]) * inv;
                    vself[t] += vv[t];
                    atomicAdd(v_means + 3 * j + t, -vv[t]);
                }
            }
            if (v_means != nullptr && dir_w != 0.0f) {
#pragma unroll
                for (int t = 0; t < 3; ++t) atomicAdd(v_means + 3 * i + t, vself[t]);
                // rotation-matrix column js receives vmaj; quaternion VJP (SURVEY.md A.6 form)
                float vR[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
                vR[0][js] = vmaj[0]; vR[1][js] = vmaj[1]; vR[2][js] = vmaj[2];
                float vq[4];
                vq[0] = 2.f * (x * (vR[2][1] - vR[1][2]) + y * (vR[0][2] - vR[2][0]) + z * (vR[1][0] - vR[0][1]));
                vq[1] = 2.f * (-2.f * x * (vR[1][1] + vR[2][2]) + y * (vR[0][1] + vR[1][0]) + z * (vR[0][2] + vR[2][0]) + w * (vR[2][1] - vR[1][2]));
                vq[2] = 2.f * (x * (vR[0][1] + vR[1][0]) - 2.f * y * (vR[0][0] + vR[2][2]) + z * (vR[1][2] + vR[2][1]) + w * (vR[0][2] - vR[2][0]));
                vq[3] = 2.f * (x * (vR[0][2] + vR[2][0]) + y * (vR[1][2] + vR[2][1]) - 2.f * z * (vR[0][0] + vR[1][1]) + w * (vR[1][0] - vR[0][1]));
                const float dq = vq[0] * w + vq[1] * x + vq[2] * y + vq[3] * z;
                v_quats[4 * i + 0] += (vq[0] - dq * w) * iqn;
                v_quats[4 * i + 1] += (vq[1] - dq * x) * iqn;
                v_quats[4 * i + 2] += (vq[2] - dq * y) * iqn;
                v_quats[4 * i + 3] += (vq[3] - dq * z) * iqn;
            }
        }
    }
    // block reduction of the two loss partial sums -> one fp64 atomic each
    __shared__ float red[2][4];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        dir_part += __shfl_xor_sync(0xffffffffu, dir_part, d);
        ratio_part += __shfl_xor_sync(0xffffffffu, ratio_part, d);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = dir_part; red[1][warp] = ratio_part; }
    __syncthreads();
    if (threadIdx.x == 0 && losses != nullptr) {
        atomicAdd(losses + 0, (double)(red[0][0] + red[0][1] + red[0][2] + red[0][3]));
        atomicAdd(losses + 1, (double)(red[1][0] + red[1][1] + red[1][2] + red[1][3]));
    }
}

}  // namespace

extern "C" int eg_reg_fwd_bwd(int n, const float *means, const float *quats, const float *log_scales,
                              const int32_t *nn_indices, int nn_cols, int k, int enforce_half, float dir_weight,
                              float ratio_weight, double *losses, float *v_means, float *v_quats,
                              float *v_log_scales, void *stream) {
    if (n <= 0) return 0;
    if (nn_cols > REG_MAX_NN) {
        eg_set_error("eg_reg_fwd_bwd: nn_cols %d > %d", nn_cols, REG_MAX_NN);
        return 1;
    }
    if (enforce_half && (k <= 0 || k > nn_cols)) {
        eg_set_error("eg_reg_fwd_bwd: enforce_half needs 0 < k <= nn_cols");
        return 1;
    }
    reg_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, means, quats, log_scales, nn_indices, nn_cols, k,
                                                                  enforce_half, dir_weight, ratio_weight, losses,
                                                                  v_means, v_quats, v_log_scales);
    return eg_check_launch("eg_reg_fwd_bwd");
}
