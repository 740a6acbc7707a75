// eg_project_bwd.cu -- K7: projection backward, fused with the reference's activation VJPs and
// update_absgrads.
//
// Semantics: SURVEY.md Appendix A.6 (gsplat==1.0.0 fully_fused_projection bwd, viewmat gradients
// off, behind /root/reference/edgegaussians/models/edge_gs.py:250-268), preceded by the VJP of
// gsplat rendering.py's  opacities * compensations  (antialiased) and followed by ExpBackward /
// SigmoidBackward of edge_gs.py:253-254 when cfg.raw_params = 1, and by
// absgrads += ||means2d.absgrad||_2  (edge_gs.py:612).
// One thread per Gaussian: reads 44 B parameters + 32 B record + 32 B 2D gradients, writes 44 B.
#include "eg_project_vjp.cuh"

namespace {

template <bool RAW>
__global__ void __launch_bounds__(128) project_bwd_kernel(
    const eg_config cfg, const float *__restrict__ means, const float *__restrict__ quats,
    const float *__restrict__ scales, const float *__restrict__ opacities, const float *__restrict__ viewmat,
    const float *__restrict__ Kmat, const float4 *__restrict__ rec, const int2 *__restrict__ gint,
    float4 *__restrict__ grad2d, int zero_grad2d, const float *__restrict__ v_depths, float *__restrict__ v_means,
    float *__restrict__ v_quats, float *__restrict__ v_scales, float *__restrict__ v_opacities,
    float *__restrict__ absgrad_accum, const eg_push_target push) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= cfg.n) return;
    float vm[3] = {0.f, 0.f, 0.f}, vs[3] = {0.f, 0.f, 0.f}, vq[4] = {0.f, 0.f, 0.f, 0.f}, vo = 0.f;
    // issue every load up front (streaming kernel: bytes in flight per thread are what hides DRAM latency)
    const int2 gi = __ldg(gint + g);
    const float4 r1 = __ldg(rec + 2 * g + 1);
    const float4 g0 = grad2d[2 * g], g1 = grad2d[2 * g + 1];
    const float4 q4 = __ldg(reinterpret_cast<const float4 *>(quats) + g);
    const float mx = __ldg(means + 3 * g), my = __ldg(means + 3 * g + 1), mz = __ldg(means + 3 * g + 2);
    float s[3];
    s[0] = __ldg(scales + 3 * g); s[1] = __ldg(scales + 3 * g + 1); s[2] = __ldg(scales + 3 * g + 2);
    float o = __ldg(opacities + g);
    if (zero_grad2d) {  // leave the accumulator clean for the next iteration's atomics
        grad2d[2 * g] = make_float4(0.f, 0.f, 0.f, 0.f);
        grad2d[2 * g + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (gi.x > 0) {
        const EgCam cam = eg_load_cam(viewmat, Kmat);
        if (absgrad_accum != nullptr) absgrad_accum[g] += sqrtf(g0.z * g0.z + g0.w * g0.w);
        eg_project_vjp<RAW>(cfg, cam, r1, g0, g1, mx, my, mz, q4, s, o, v_depths != nullptr ? __ldg(v_depths + g) : 0.0f,
                            vm, vs, vq, vo);
    }
    const EgGradOut out = eg_grad_out(push, g, v_means, v_scales, v_quats, v_opacities);  // local tensors, or the owner's slot
    out.means[3 * g] = vm[0]; out.means[3 * g + 1] = vm[1]; out.means[3 * g + 2] = vm[2];
    out.scales[3 * g] = vs[0]; out.scales[3 * g + 1] = vs[1]; out.scales[3 * g + 2] = vs[2];
    eg_store_quat_grad(out.quats, g, vq);
    out.opac[g] = vo;
}

}  // namespace

static int project_bwd_launch(const eg_config *cfg, const float *means, const float *quats, const float *scales,
                              const float *opacities, const float *viewmat, const float *K, const float *rec,
                              const int32_t *gint, float *grad2d, int zero_grad2d, const float *v_depths,
                              float *v_means, float *v_quats, float *v_scales, float *v_opacities,
                              float *absgrad_accum, const eg_push_target &push, void *stream) {
    if (cfg->n <= 0) return 0;
    const int block = 128, grid = (cfg->n + block - 1) / block;
    cudaStream_t s = (cudaStream_t)stream;
    if (cfg->raw_params)
        project_bwd_kernel<true><<<grid, block, 0, s>>>(*cfg, means, quats, scales, opacities, viewmat, K,
                                                        (const float4 *)rec, (const int2 *)gint, (float4 *)grad2d,
                                                        zero_grad2d, v_depths, v_means, v_quats, v_scales, v_opacities,
                                                        absgrad_accum, push);
    else
        project_bwd_kernel<false><<<grid, block, 0, s>>>(*cfg, means, quats, scales, opacities, viewmat, K,
                                                         (const float4 *)rec, (const int2 *)gint, (float4 *)grad2d,
                                                         zero_grad2d, v_depths, v_means, v_quats, v_scales, v_opacities,
                                                         absgrad_accum, push);
    return eg_check_launch("eg_project_bwd");
}

extern "C" int eg_project_bwd(const eg_config *cfg, const float *means, const float *quats, const float *scales,
                              const float *opacities, const float *viewmat, const float *K, const float *rec,
                              const int32_t *gint, float *grad2d, int zero_grad2d, const float *v_depths,
                              float *v_means, float *v_quats, float *v_scales, float *v_opacities,
                              float *absgrad_accum, void *stream) {
    if (cfg == nullptr) {
        eg_set_error("eg_project_bwd: null config");
        return 1;
    }
    eg_push_target none = {};
    return project_bwd_launch(cfg, means, quats, scales, opacities, viewmat, K, rec, gint, grad2d, zero_grad2d, v_depths,
                              v_means, v_quats, v_scales, v_opacities, absgrad_accum, none, stream);
}

// the same kernel with its gradient stores redirected into the owners' staging slots (eg_push_target)
extern "C" int eg_project_bwd_push(const eg_config *cfg, const float *means, const float *quats, const float *scales,
                                   const float *opacities, const float *viewmat, const float *K, const float *rec,
                                   const int32_t *gint, float *grad2d, int zero_grad2d, const eg_push_target *push,
                                   float *absgrad_accum, void *stream) {
    if (cfg == nullptr || !eg_push_target_ok("eg_project_bwd_push", push, cfg->n)) return 1;
    if (push->world < 2) {
        eg_set_error("eg_project_bwd_push: needs world >= 2 (a single rank calls eg_project_bwd)");
        return 1;
    }
    return project_bwd_launch(cfg, means, quats, scales, opacities, viewmat, K, rec, gint, grad2d, zero_grad2d, nullptr,
                              nullptr, nullptr, nullptr, nullptr, absgrad_accum, *push, stream);
}
