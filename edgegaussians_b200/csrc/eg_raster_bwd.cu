// eg_raster_bwd.cu -- K6: compositing backward with abs-grad, one CTA per 16x16 tile.
//
// Semantics: SURVEY.md Appendix A.5 (gsplat==1.0.0 rasterize_to_pixels bwd behind
// /root/reference/edgegaussians/models/edge_gs.py:250-268, absgrad=True edge_gs.py:266).
//
// gsplat walks every pixel's blend list backwards, recovering T by division and carrying a running
// colour buffer, then warp-reduces 11 floats per (pixel-warp, Gaussian) and issues up to 88 atomics
// per (tile, Gaussian).  The reference only ever splats colors == 1 (edge_gs.py:247), for which
//     render_ch(p) = sum_i alpha_i T_i = 1 - prod_i (1 - alpha_i) = alpha(p)
// and therefore   d out(p) / d alpha_k = T_final(p) / (1 - alpha_k)   for every composited k:
// gsplat's  (color*T_k - buffer_k*ra_k)  is this quantity computed with cancellation.  The backward
// is thus a plain sum over (pixel, Gaussian) pairs with no ordering dependence, and is organised the
// other way round: one THREAD per (tile, Gaussian), looping over the pixels of the tile that the
// Gaussian's alpha >= 1/255 footprint can reach.  Per-Gaussian gradients accumulate in registers --
// no warp reductions -- and leave as two 128-bit vector reductions (red.global.add.v4.f32) per
// (tile, Gaussian).  Footprints larger than BIG_AREA pixels are deferred to a cooperative pass in
// which a whole warp shares one Gaussian.  The per-pixel state (seed * T_final, last contributor)
// lives in shared memory.
#include "eg_common.cuh"

namespace {

constexpr int RB_THREADS = 256;
constexpr int BIG_AREA = 48;    // footprints (pixels inside the tile) above this go to the warp pass
constexpr int BIG_QUEUE = 1024; // per-CTA queue of deferred Gaussians (overflow handled inline)

struct PairAcc {
    float gx, gy, ax, ay, ca, cb, cc, go;
};

// contribution of pixel (w = seed*T_final, valid) to Gaussian (mx,my,o | A,B,C)
__device__ __forceinline__ void pair_grad(float w, float mx, float my, float o, float A, float B, float C,
                                          float px, float py, PairAcc &acc) {
    const float dx = mx - px, dy = my - py;
    const float sigma = eg_sigma(A, B, C, dx, dy);
    const float vis = eg_vis(sigma);
    const float ov = __fmul_rn(o, vis);
    const float al = fminf(EG_ALPHA_MAX, ov);
    if (sigma < 0.0f || al < EG_ALPHA_MIN) return;
    const float ra = __fdividef(1.0f, 1.0f - al);
    const float v_alpha = w * ra;
    if (ov <= EG_ALPHA_MAX) {
        const float v_sigma = -ov * v_alpha;
        const float gx = v_sigma * (A * dx + B * dy);
        const float gy = v_sigma * (B * dx + C * dy);
        acc.gx += gx;
        acc.gy += gy;
        acc.ax += fabsf(gx);
        acc.ay += fabsf(gy);
        acc.ca += 0.5f * v_sigma * dx * dx;
        acc.cb += v_sigma * dx * dy;
        acc.cc += 0.5f * v_sigma * dy * dy;
        acc.go += vis * v_alpha;
    }
}

__global__ void __launch_bounds__(RB_THREADS) raster_bwd_kernel(
    const eg_config cfg, int tw, const float4 *__restrict__ rec, const int32_t *__restrict__ tile_offsets,
    const int32_t *__restrict__ flatten_ids, const int32_t *__restrict__ last_ids, const float *__restrict__ alpha,
    const float *__restrict__ v_render, int vr_ch, const float *__restrict__ v_alpha,
    const float *__restrict__ wpix, float seed_scale, float *__restrict__ grad2d,
    const int32_t *__restrict__ status) {
    __shared__ float s_w[EG_TILE * EG_TILE];    // seed * T_final per pixel (0 outside the image)
    __shared__ int s_last[EG_TILE * EG_TILE];   // last contributor, relative to the segment start
    __shared__ int s_big[BIG_QUEUE];
    __shared__ int s_nbig;

    if (status[EG_ST_OVERFLOW]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tile_y = tile / tw, tile_x = tile - tile_y * tw;
    const int start = tile_offsets[tile];
    const int L = tile_offsets[tile + 1] - start;
    if (L <= 0) return;
    const int X0 = tile_x * EG_TILE, Y0 = tile_y * EG_TILE;

    {
        const int lx = tid & 15, ly = tid >> 4;
        const int pxi = X0 + lx, pyi = Y0 + ly;
        float w = 0.0f;
        int last = -1;
        if (pxi < cfg.width && pyi < cfg.height) {
            const long long pix = (long long)pyi * cfg.width + pxi;
            if (wpix != nullptr) {
                w = seed_scale * __ldg(wpix + pix);
            } else {
                float gsum = 0.0f;
                if (v_render != nullptr)
                    for (int c = 0; c < vr_ch; ++c) gsum += __ldg(v_render + pix * vr_ch + c);
                if (v_alpha != nullptr) gsum += __ldg(v_alpha + pix);
                w = gsum * (1.0f - __ldg(alpha + pix));
            }
            last = __ldg(last_ids + pix) - start;
        }
        s_w[tid] = w;
        s_last[tid] = last;
        if (tid == 0) s_nbig = 0;
    }
    __syncthreads();

    const int xmax = min(EG_TILE, cfg.width - X0) - 1, ymax = min(EG_TILE, cfg.height - Y0) - 1;

    // ---- pass 1: one thread per (tile, Gaussian) ----
    for (int k = tid; k < L; k += RB_THREADS) {
        const int gid = __ldg(flatten_ids + start + k);
        const float4 r0 = __ldg(rec + 2 * gid), r1 = __ldg(rec + 2 * gid + 1);
        float hx, hy, tau;
        if (!eg_extent(r0.z, r1.x, r1.y, r1.z, hx, hy, tau)) continue;
        // pixel j (centre j + 0.5) is reachable iff  mx - hx <= j + 0.5 <= mx + hx
        const float fx0 = r0.x - (float)X0, fy0 = r0.y - (float)Y0;
        const int xlo = max(0, (int)ceilf(fminf(fx0 - hx - 0.5f, 64.0f)));
        const int xhi = min(xmax, (int)floorf(fmaxf(fx0 + hx - 0.5f, -64.0f)));
        const int ylo = max(0, (int)ceilf(fminf(fy0 - hy - 0.5f, 64.0f)));
        const int yhi = min(ymax, (int)floorf(fmaxf(fy0 + hy - 0.5f, -64.0f)));
        if (xlo > xhi || ylo > yhi) continue;
        const int area = (xhi - xlo + 1) * (yhi - ylo + 1);
        if (area > BIG_AREA) {
            const int slot = atomicAdd(&s_nbig, 1);
            if (slot < BIG_QUEUE) {
                s_big[slot] = k;
                continue;
            }
        }
        PairAcc acc = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int y = ylo; y <= yhi; ++y) {
            const float py = (float)(Y0 + y) + 0.5f;
            for (int x = xlo; x <= xhi; ++x) {
                const int p = y * EG_TILE + x;
                const float w = s_w[p];
                if (w == 0.0f || k > s_last[p]) continue;
                pair_grad(w, r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, (float)(X0 + x) + 0.5f, py, acc);
            }
        }
        if (acc.ax != 0.0f || acc.ay != 0.0f || acc.go != 0.0f || acc.ca != 0.0f || acc.cc != 0.0f) {
            float *dst = grad2d + 8ll * gid;
            eg_red_add_v4(dst, acc.gx, acc.gy, acc.ax, acc.ay);
            eg_red_add_v4(dst + 4, acc.ca, acc.cb, acc.cc, acc.go);
        }
    }
    __syncthreads();

    // ---- pass 2: large footprints, one warp per Gaussian, lanes = pixels of the tile ----
    const int nbig = min(s_nbig, BIG_QUEUE);
    for (int q = warp; q < nbig; q += RB_THREADS / 32) {
        const int k = s_big[q];
        const int gid = __ldg(flatten_ids + start + k);
        const float4 r0 = __ldg(rec + 2 * gid), r1 = __ldg(rec + 2 * gid + 1);
        float hx, hy, tau;
        eg_extent(r0.z, r1.x, r1.y, r1.z, hx, hy, tau);
        const float fx0 = r0.x - (float)X0, fy0 = r0.y - (float)Y0;
        const int xlo = max(0, (int)ceilf(fminf(fx0 - hx - 0.5f, 64.0f)));
        const int xhi = min(xmax, (int)floorf(fmaxf(fx0 + hx - 0.5f, -64.0f)));
        const int ylo = max(0, (int)ceilf(fminf(fy0 - hy - 0.5f, 64.0f)));
        const int yhi = min(ymax, (int)floorf(fmaxf(fy0 + hy - 0.5f, -64.0f)));
        const int wbox = xhi - xlo + 1, area = wbox * (yhi - ylo + 1);
        PairAcc acc = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = lane; i < area; i += 32) {
            const int y = ylo + i / wbox, x = xlo + i % wbox;
            const int p = y * EG_TILE + x;
            const float w = s_w[p];
            if (w == 0.0f || k > s_last[p]) continue;
            pair_grad(w, r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, (float)(X0 + x) + 0.5f, (float)(Y0 + y) + 0.5f, acc);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            acc.gx += __shfl_xor_sync(0xffffffffu, acc.gx, d);
            acc.gy += __shfl_xor_sync(0xffffffffu, acc.gy, d);
            acc.ax += __shfl_xor_sync(0xffffffffu, acc.ax, d);
            acc.ay += __shfl_xor_sync(0xffffffffu, acc.ay, d);
            acc.ca += __shfl_xor_sync(0xffffffffu, acc.ca, d);
            acc.cb += __shfl_xor_sync(0xffffffffu, acc.cb, d);
            acc.cc += __shfl_xor_sync(0xffffffffu, acc.cc, d);
            acc.go += __shfl_xor_sync(0xffffffffu, acc.go, d);
        }
        if (lane == 0 && (acc.ax != 0.0f || acc.ay != 0.0f || acc.go != 0.0f || acc.ca != 0.0f || acc.cc != 0.0f)) {
            float *dst = grad2d + 8ll * gid;
            eg_red_add_v4(dst, acc.gx, acc.gy, acc.ax, acc.ay);
            eg_red_add_v4(dst + 4, acc.ca, acc.cb, acc.cc, acc.go);
        }
    }
}

}  // namespace

extern "C" int eg_raster_bwd(const eg_config *cfg, const float *rec, const int32_t *tile_offsets,
                             const int32_t *flatten_ids, const int32_t *last_ids, const float *alpha,
                             const float *v_render, int v_render_channels, const float *v_alpha, const float *wpix,
                             float seed_scale, float *grad2d, const int32_t *status, void *stream) {
    if (cfg == nullptr || cfg->tile_size != EG_TILE) {
        eg_set_error("eg_raster_bwd: tile_size must be %d", EG_TILE);
        return 1;
    }
    if (wpix == nullptr && alpha == nullptr) {
        eg_set_error("eg_raster_bwd: need either wpix or alpha (+ v_render / v_alpha)");
        return 1;
    }
    int tw, th;
    eg_tile_grid(cfg->width, cfg->height, cfg->tile_size, &tw, &th);
    raster_bwd_kernel<<<tw * th, RB_THREADS, 0, (cudaStream_t)stream>>>(
        *cfg, tw, (const float4 *)rec, tile_offsets, flatten_ids, last_ids, alpha, v_render, v_render_channels,
        v_alpha, wpix, seed_scale, grad2d, status);
    return eg_check_launch("eg_raster_bwd");
}
