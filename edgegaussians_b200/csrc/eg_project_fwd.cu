// eg_project_fwd.cu -- K1 + K2(pass 1): per-Gaussian camera projection, 3D covariance -> 2D conic,
// blur compensation, radius / frustum culling, tile rectangle and per-tile intersection counts.
//
// Semantics: SURVEY.md Appendix A.1 / A.2 (gsplat==1.0.0 fully_fused_projection fwd + isect_tiles
// pass 1, reached from /root/reference/edgegaussians/models/edge_gs.py:250-268), with the
// reference's activations (edge_gs.py:253-254) optionally fused.
//
// Integer outputs (radius, tile rectangle) must be bit-exact against the oracle, so every fp32
// operation that feeds them is an explicitly rounded intrinsic (__fmul_rn, __fadd_rn, ...): the
// compiler cannot contract them into FMAs.  This is the "canonical evaluation order" of DESIGN.md.
// The kernel is HBM-bound (44 B in, 40 B out per Gaussian), the extra instructions are free.
#include "eg_common.cuh"

namespace {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }
// (a*b + c*d) + e*f, one rounding per operation, left to right
__device__ __forceinline__ float dot3(float a, float b, float c, float d, float e, float f) {
    return add(add(mul(a, b), mul(c, d)), mul(e, f));
}

#ifndef EG_PF_MINBLOCKS
#define EG_PF_MINBLOCKS 5   // resident CTAs per SM the register allocation is bounded for
#endif

template <bool RAW>
__global__ void __launch_bounds__(256, EG_PF_MINBLOCKS) project_fwd_kernel(
    const eg_config cfg, const float *__restrict__ means, const float *__restrict__ quats,
    const float *__restrict__ scales, const float *__restrict__ opacities, const float *__restrict__ colors,
    const float *__restrict__ viewmat, const float *__restrict__ Kmat, float4 *__restrict__ rec,
    int2 *__restrict__ gint, int32_t *__restrict__ tile_counts, unsigned long long *__restrict__ keys,
    int32_t *__restrict__ status, int tw, int th) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= cfg.n) return;
    const EgCam cam = eg_load_cam(viewmat, Kmat);
    const float *R = cam.R;

    if (colors != nullptr) {
        const float c0 = __ldg(colors + 3 * g), c1 = __ldg(colors + 3 * g + 1), c2 = __ldg(colors + 3 * g + 2);
        if (c0 != 1.0f || c1 != 1.0f || c2 != 1.0f) status[EG_ST_BADCOLOR] = 1;
    }

    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = make_float4(0.f, 0.f, 0.f, 0.f);
    int radius_i = 0, ntiles = 0;
    uint32_t x0 = 0, y0 = 0, x1 = 0, y1 = 0;

    do {
        const float mx = __ldg(means + 3 * g), my = __ldg(means + 3 * g + 1), mz = __ldg(means + 3 * g + 2);
        const float x = add(dot3(R[0], mx, R[1], my, R[2], mz), cam.t[0]);
        const float y = add(dot3(R[3], mx, R[4], my, R[5], mz), cam.t[1]);
        const float z = add(dot3(R[6], mx, R[7], my, R[8], mz), cam.t[2]);
        if (z < cfg.near_plane || z > cfg.far_plane) break;

        const float4 q4 = __ldg(reinterpret_cast<const float4 *>(quats) + g);
        float qw = q4.x, qx = q4.y, qy = q4.z, qz = q4.w;
        const float n2 = add(add(add(mul(qx, qx), mul(qy, qy)), mul(qz, qz)), mul(qw, qw));
        const float inv_n = dvd(1.0f, __fsqrt_rn(n2));
        qw = mul(qw, inv_n); qx = mul(qx, inv_n); qy = mul(qy, inv_n); qz = mul(qz, inv_n);
        const float x2 = mul(qx, qx), y2 = mul(qy, qy), z2 = mul(qz, qz);
        const float xy = mul(qx, qy), xz = mul(qx, qz), yz = mul(qy, qz);
        const float wx = mul(qw, qx), wy = mul(qw, qy), wz = mul(qw, qz);
        float Rq[3][3];
        Rq[0][0] = sub(1.0f, mul(2.0f, add(y2, z2))); Rq[0][1] = mul(2.0f, sub(xy, wz)); Rq[0][2] = mul(2.0f, add(xz, wy));
        Rq[1][0] = mul(2.0f, add(xy, wz)); Rq[1][1] = sub(1.0f, mul(2.0f, add(x2, z2))); Rq[1][2] = mul(2.0f, sub(yz, wx));
        Rq[2][0] = mul(2.0f, sub(xz, wy)); Rq[2][1] = mul(2.0f, add(yz, wx)); Rq[2][2] = sub(1.0f, mul(2.0f, add(x2, y2)));

        float s[3];
        s[0] = __ldg(scales + 3 * g); s[1] = __ldg(scales + 3 * g + 1); s[2] = __ldg(scales + 3 * g + 2);
        if (RAW) { s[0] = expf(s[0]); s[1] = expf(s[1]); s[2] = expf(s[2]); }
        float M[3][3], S[3][3], Tm[3][3], Sc[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) M[i][j] = mul(Rq[i][j], s[j]);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i; j < 3; ++j) {
                S[i][j] = dot3(M[i][0], M[j][0], M[i][1], M[j][1], M[i][2], M[j][2]);
                S[j][i] = S[i][j];
            }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Tm[i][j] = dot3(R[3 * i], S[0][j], R[3 * i + 1], S[1][j], R[3 * i + 2], S[2][j]);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i; j < 3; ++j) {
                Sc[i][j] = dot3(Tm[i][0], R[3 * j], Tm[i][1], R[3 * j + 1], Tm[i][2], R[3 * j + 2]);
                Sc[j][i] = Sc[i][j];
            }
        const float tan_fovx = dvd(mul(0.5f, (float)cfg.width), cam.fx);
        const float tan_fovy = dvd(mul(0.5f, (float)cfg.height), cam.fy);
        const float lim_x = mul(1.3f, tan_fovx), lim_y = mul(1.3f, tan_fovy);
        const float rz = dvd(1.0f, z);
        const float rz2 = mul(rz, rz);
        const float tx = mul(z, fminf(lim_x, fmaxf(-lim_x, mul(x, rz))));
        const float ty = mul(z, fminf(lim_y, fmaxf(-lim_y, mul(y, rz))));
        const float J00 = mul(cam.fx, rz), J11 = mul(cam.fy, rz);
        const float J02 = -mul(mul(cam.fx, tx), rz2);
        const float J12 = -mul(mul(cam.fy, ty), rz2);
        const float a00 = add(mul(J00, Sc[0][0]), mul(J02, Sc[2][0]));
        const float a01 = add(mul(J00, Sc[0][1]), mul(J02, Sc[2][1]));
        const float a02 = add(mul(J00, Sc[0][2]), mul(J02, Sc[2][2]));
        const float a11 = add(mul(J11, Sc[1][1]), mul(J12, Sc[2][1]));
        const float a12 = add(mul(J11, Sc[1][2]), mul(J12, Sc[2][2]));
        const float c00_0 = add(mul(a00, J00), mul(a02, J02));
        const float c01 = add(mul(a01, J11), mul(a02, J12));
        const float c11_0 = add(mul(a11, J11), mul(a12, J12));
        const float m2x = add(mul(mul(cam.fx, x), rz), cam.cx);
        const float m2y = add(mul(mul(cam.fy, y), rz), cam.cy);
        const float det0 = sub(mul(c00_0, c11_0), mul(c01, c01));
        const float c00 = add(c00_0, cfg.eps2d), c11 = add(c11_0, cfg.eps2d);
        const float det = sub(mul(c00, c11), mul(c01, c01));
        if (!(det > 0.0f)) break;
        const float comp = __fsqrt_rn(fmaxf(0.0f, dvd(det0, det)));
        const float inv_det = dvd(1.0f, det);
        const float cA = mul(c11, inv_det);
        const float cB = mul(-c01, inv_det);
        const float cC = mul(c00, inv_det);
        const float b = mul(0.5f, add(c00, c11));
        const float v1 = add(b, __fsqrt_rn(fmaxf(0.01f, sub(mul(b, b), det))));
        const float radius = ceilf(mul(3.0f, __fsqrt_rn(v1)));
        if (radius <= cfg.radius_clip) break;
        if (add(m2x, radius) <= 0.0f || sub(m2x, radius) >= (float)cfg.width || add(m2y, radius) <= 0.0f ||
            sub(m2y, radius) >= (float)cfg.height)
            break;
        radius_i = (int)radius;
        float o = __ldg(opacities + g);
        if (RAW) o = dvd(1.0f, add(1.0f, expf(-o)));
        const float o_eff = cfg.antialiased ? mul(o, comp) : o;
        r0 = make_float4(m2x, m2y, o_eff, z);
        r1 = make_float4(cA, cB, cC, comp);
        eg_tile_rect(m2x, m2y, radius_i, tw, th, x0, y0, x1, y1);
        ntiles = (int)((y1 - y0) * (x1 - x0));
    } while (0);

    rec[2 * g] = r0;
    rec[2 * g + 1] = r1;
    gint[g] = make_int2(radius_i, ntiles);
    // status[EG_ST_NISECT] = sum of tiles_per_gauss (gsplat's n_isects), whatever is emitted below: one atomic per warp
    {
        const unsigned am = __activemask();
        const int nt = __reduce_add_sync(am, ntiles);
        if ((int)(threadIdx.x & 31) == __ffs(am) - 1 && nt > 0) atomicAdd(status + EG_ST_NISECT, nt);
    }
    if (cfg.flags & EG_FLAG_NO_EMIT) return;  // Gaussian-major forward: no tile lists at all
    if (ntiles <= 0) return;
    // K2 emission: append (depth_bits << 32 | id) to the bucket of every tile of the rectangle -- with
    // EG_FLAG_CULL_TILES only of the tiles the alpha >= 1/255 footprint can reach (fused step: the others would
    // fail the alpha test at every pixel; elongated Gaussians touch a fraction of their bounding square)
    const unsigned long long key = ((unsigned long long)__float_as_uint(r0.w) << 32) | (unsigned int)g;
    const uint32_t w = x1 - x0;
    const bool cull = (cfg.flags & EG_FLAG_CULL_TILES) != 0;
    const bool count_only = (cfg.flags & EG_FLAG_COMPACT_KEYS) != 0;  // eg_bin emits after the scan
    float hu = 1e30f, hv = 1e30f, tau = 0.0f;
    if (cull && !eg_extent(r0.z, r1.x, r1.y, r1.z, hu, hv, tau)) return;  // opacity' < 1/255: never composited
    if (!cull && !count_only && ntiles <= 12) {
        // common case: issue all the (independent) atomics first so that they overlap, then the stores
        int pos[12];
#pragma unroll
        for (int q = 0; q < 12; ++q)
            if (q < ntiles) {
                const uint32_t i = y0 + (uint32_t)q / w, j = x0 + (uint32_t)q % w;
                pos[q] = atomicAdd(tile_counts + (size_t)(i * tw + j) * EG_CNT_STRIDE, 1);
            }
#pragma unroll
        for (int q = 0; q < 12; ++q)
            if (q < ntiles) {
                const uint32_t i = y0 + (uint32_t)q / w, j = x0 + (uint32_t)q % w;
                if (pos[q] < cfg.tile_capacity) keys[(size_t)(i * tw + j) * (size_t)cfg.tile_capacity + pos[q]] = key;
            }
        return;
    }
    // The slot a key goes to is the RETURN value of the tile counter's atomic: keep three atomics in flight and store
    // the key of the oldest one (a store right behind its own atomic would expose every atomic's full latency in turn).
    int p1 = -1, p2 = -1, p3 = -1;
    size_t t1 = 0, t2 = 0, t3 = 0;
    for (uint32_t i = y0; i < y1; ++i) {
        int j0 = (int)x0, j1 = (int)x1 - 1;
        if (cull && !eg_tile_row_cols(r0.x, r0.y, r1.x, r1.y, r1.z, tau, hu, hv, (int)i, (int)x0, (int)x1, j0, j1)) continue;
        for (int j = j0; j <= j1; ++j) {
            const size_t t = (size_t)(i * tw + j);
            const int pos = atomicAdd(tile_counts + t * EG_CNT_STRIDE, 1);
            if (count_only) continue;
            if (p3 >= 0 && p3 < cfg.tile_capacity) keys[t3 * (size_t)cfg.tile_capacity + p3] = key;
            p3 = p2; t3 = t2;
            p2 = p1; t2 = t1;
            p1 = pos; t1 = t;
        }
    }
    if (p3 >= 0 && p3 < cfg.tile_capacity) keys[t3 * (size_t)cfg.tile_capacity + p3] = key;
    if (p2 >= 0 && p2 < cfg.tile_capacity) keys[t2 * (size_t)cfg.tile_capacity + p2] = key;
    if (p1 >= 0 && p1 < cfg.tile_capacity) keys[t1 * (size_t)cfg.tile_capacity + p1] = key;
}

}  // namespace

extern "C" int eg_project_fwd(const eg_config *cfg, const float *means, const float *quats, const float *scales,
                              const float *opacities, const float *colors, const float *viewmat, const float *K,
                              float *rec, int32_t *gint, int32_t *tile_counts, uint64_t *keys, int32_t *status,
                              void *stream) {
    if (cfg == nullptr || cfg->tile_size != EG_TILE) {
        eg_set_error("eg_project_fwd: tile_size must be %d", EG_TILE);
        return 1;
    }
    if (cfg->n <= 0) return 0;
    int tw, th;
    eg_tile_grid(cfg->width, cfg->height, cfg->tile_size, &tw, &th);
    const int block = 256, grid = (cfg->n + block - 1) / block;
    cudaStream_t s = (cudaStream_t)stream;
    if (cfg->raw_params)
        project_fwd_kernel<true><<<grid, block, 0, s>>>(*cfg, means, quats, scales, opacities, colors, viewmat, K,
                                                        (float4 *)rec, (int2 *)gint, tile_counts,
                                                        (unsigned long long *)keys, status, tw, th);
    else
        project_fwd_kernel<false><<<grid, block, 0, s>>>(*cfg, means, quats, scales, opacities, colors, viewmat, K,
                                                         (float4 *)rec, (int2 *)gint, tile_counts,
                                                         (unsigned long long *)keys, status, tw, th);
    return eg_check_launch("eg_project_fwd");
}
