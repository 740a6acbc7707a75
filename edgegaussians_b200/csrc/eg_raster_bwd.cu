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
// is thus a plain sum over (pixel, Gaussian) pairs with no ordering dependence, and is organised
// Gaussian-major with a BALANCED pair distribution:
//   1. batches of 256 Gaussians of the tile are staged in shared memory together with the pixel
//      rectangle (inside the tile) their alpha >= 1/255 footprint can reach;
//   2. a block-wide prefix sum over the rectangle areas linearises all (Gaussian, pixel) pairs of the
//      batch; every thread takes an equal contiguous slice of that pair list and walks it,
//      accumulating the 8 per-Gaussian gradient values in registers (no warp reductions);
//   3. whenever the walk leaves a Gaussian its partial sums leave as two 128-bit vector reductions
//      (red.global.add.v4.f32) -- about two flushes per thread.
// The per-pixel state (seed * T_final, last contributor) lives in shared memory.
#include "eg_common.cuh"

namespace {

constexpr int RB_THREADS = 256;
constexpr int PAIR_CAP = 32;  // per-thread buffer of candidate pairs between the cheap and the heavy phase

struct PairAcc {
    float gx, gy, ax, ay, ca, cb, cc, go;
};

__device__ __forceinline__ void acc_zero(PairAcc &a) { a.gx = a.gy = a.ax = a.ay = a.ca = a.cb = a.cc = a.go = 0.0f; }

__device__ __forceinline__ void acc_flush(const PairAcc &a, float *__restrict__ grad2d, int gid) {
    if (a.go != 0.0f || a.ax != 0.0f || a.ay != 0.0f || a.ca != 0.0f || a.cc != 0.0f) {
        float *dst = grad2d + 8ll * gid;
        eg_red_add_v4(dst, a.gx, a.gy, a.ax, a.ay);
        eg_red_add_v4(dst + 4, a.ca, a.cb, a.cc, a.go);
    }
}

__global__ void __launch_bounds__(RB_THREADS) raster_bwd_kernel(
    const eg_config cfg, int tw, const float4 *__restrict__ rec, const int32_t *__restrict__ tile_offsets,
    const int32_t *__restrict__ flatten_ids, const int32_t *__restrict__ last_ids, const float *__restrict__ alpha,
    const float *__restrict__ v_render, int vr_ch, const float *__restrict__ v_alpha,
    const float *__restrict__ wpix, float seed_scale, float *__restrict__ grad2d,
    const int32_t *__restrict__ status) {
    __shared__ float2 s_pix[EG_TILE * EG_TILE];        // (seed * T_final, last contributor rel. to segment start)
    __shared__ __align__(16) float4 sA[RB_THREADS];    // mean2d.x, mean2d.y, opacity, packed pixel rectangle
    __shared__ __align__(16) float4 sB[RB_THREADS];    // conic a, b, c, gaussian id
    __shared__ __align__(16) float4 sC[RB_THREADS];    // 2*A*tau, det(conic), 1/A (0 = no row span), -
    __shared__ unsigned short s_buf[PAIR_CAP * RB_THREADS];  // candidate pairs: pixel index | local Gaussian << 8
    __shared__ int s_off[RB_THREADS + 1];              // exclusive prefix of the rectangle areas
    __shared__ int s_wsum[RB_THREADS / 32];

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
        s_pix[tid] = make_float2(w, __int_as_float(last));
    }
    const int xmax = min(EG_TILE, cfg.width - X0) - 1, ymax = min(EG_TILE, cfg.height - Y0) - 1;

    const float X0f = (float)X0 + 0.5f, Y0f = (float)Y0 + 0.5f;  // centre of pixel (0,0) of the tile

    for (int b0 = 0; b0 < L; b0 += RB_THREADS) {
        __syncthreads();  // s_pix visible / previous batch fully consumed
        // ---- 1. stage one Gaussian per thread, with its reachable pixel rectangle ----
        const int k = b0 + tid;
        int area = 0;
        if (k < L) {
            const int gid = __ldg(flatten_ids + start + k);
            const float4 r0 = __ldg(rec + 2 * gid), r1 = __ldg(rec + 2 * gid + 1);
            float hx, hy, tau;
            int rect = 0;
            float two_tau_a = 0.0f, det = 0.0f, inv_a = 0.0f;  // inv_a == 0: no per-row span (degenerate conic)
            if (eg_extent(r0.z, r1.x, r1.y, r1.z, hx, hy, tau)) {
                // pixel j (centre j + 0.5) is reachable iff  mx - hx <= j + 0.5 <= mx + hx
                const float fx0 = r0.x - (float)X0, fy0 = r0.y - (float)Y0;
                const int xlo = max(0, (int)ceilf(fminf(fx0 - hx - 0.5f, 64.0f)));
                const int xhi = min(xmax, (int)floorf(fmaxf(fx0 + hx - 0.5f, -64.0f)));
                const int ylo = max(0, (int)ceilf(fminf(fy0 - hy - 0.5f, 64.0f)));
                const int yhi = min(ymax, (int)floorf(fmaxf(fy0 + hy - 0.5f, -64.0f)));
                if (xlo <= xhi && ylo <= yhi) {
                    area = (xhi - xlo + 1) * (yhi - ylo + 1);
                    rect = xlo | (xhi << 4) | (ylo << 8) | (yhi << 12);
                }
                if (hx < 1e29f) {
                    det = r1.x * r1.z - r1.y * r1.y;
                    two_tau_a = 2.0f * tau * r1.x;
                    inv_a = 1.0f / r1.x;
                }
            }
            sA[tid] = make_float4(r0.x, r0.y, r0.z, __int_as_float(rect));
            sB[tid] = make_float4(r1.x, r1.y, r1.z, __int_as_float(gid));
            sC[tid] = make_float4(two_tau_a, det, inv_a, 0.0f);
        }
        // ---- 2. block-wide exclusive prefix sum of the areas ----
        int incl = area;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < RB_THREADS / 32; ++w) {
            const int v = s_wsum[w];
            if (w < warp) wbase += v;
            total += v;
        }
        s_off[tid] = wbase + incl - area;
        if (tid == 0) s_off[RB_THREADS] = total;
        __syncthreads();
        if (total == 0) continue;

        // ---- 3. every thread takes an equal slice of the (Gaussian, rectangle pixel) pair list ----
        const int chunk = (total + RB_THREADS - 1) / RB_THREADS;
        int p = tid * chunk;
        const int p_end = min(total, p + chunk);
        // binary search: largest g with s_off[g] <= p
        int lo = 0, hi = RB_THREADS;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_off[mid] <= p) lo = mid; else hi = mid;
        }
        int g = lo;
        // walk state of the cheap phase (valid while p < seg_end)
        int seg_end = p, x = 0, y = 0, xlo = 0, xhi = 0, xs = 0, xe = -1, kk = 0;
        float two_tau_a = 0.f, det = 0.f, inv_a = 0.f, mx = 0.f, my = 0.f, cb = 0.f;
        // accumulation state of the heavy phase
        int cur_g = -1;
        PairAcc acc;
        acc_zero(acc);

        for (;;) {
            // ---- 3a. cheap phase: collect up to PAIR_CAP pairs that can contribute ----
            int cnt = 0;
            while (cnt < PAIR_CAP && p < p_end) {
                if (p >= seg_end) {  // enter the next Gaussian of the slice
                    while (s_off[g + 1] <= p) ++g;
                    const float4 a = sA[g], sp = sC[g];
                    const int rect = __float_as_int(a.w);
                    xlo = rect & 15; xhi = (rect >> 4) & 15;
                    const int ylo = (rect >> 8) & 15, wbox = xhi - xlo + 1;
                    const int local = p - s_off[g], row0 = local / wbox;
                    y = ylo + row0; x = xlo + local - row0 * wbox;
                    seg_end = min(p_end, s_off[g + 1]);
                    kk = b0 + g;
                    mx = a.x; my = a.y; cb = sB[g].y;
                    two_tau_a = sp.x; det = sp.y; inv_a = sp.z;
                    xe = -2;  // forces the row span computation below
                }
                if (xe == -2) {  // first pixel of a row: sigma <= tau  <=>  |dx - c| <= hw
                    xs = xlo; xe = xhi;
                    if (inv_a != 0.0f) {
                        const float dy = my - (Y0f + (float)y);
                        const float D = fmaf(-det * dy, dy, two_tau_a);
                        if (D < 0.0f) {
                            xe = -1;
                        } else {
                            const float hw = sqrtf(D) * inv_a * 1.0001f + 2e-3f;
                            const float cpx = mx + cb * dy * inv_a - X0f;
                            xs = max(xs, (int)ceilf(fminf(cpx - hw, 64.0f)));
                            xe = min(xe, (int)floorf(fmaxf(cpx + hw, -64.0f)));
                        }
                    }
                }
                if (x >= xs && x <= xe) {
                    const int idx = y * EG_TILE + x;
                    const float2 pw = s_pix[idx];
                    if (pw.x != 0.0f && kk <= __float_as_int(pw.y)) {
                        s_buf[cnt * RB_THREADS + tid] = (unsigned short)(idx | (g << 8));
                        ++cnt;
                    }
                }
                ++p;
                if (++x > xhi) { x = xlo; ++y; xe = -2; }
            }
            // ---- 3b. heavy phase: all lanes of the warp evaluate their collected pairs together ----
            for (int i = 0; i < cnt; ++i) {
                const int e = s_buf[i * RB_THREADS + tid];
                const int eg = e >> 8;
                if (eg != cur_g) {
                    if (cur_g >= 0) acc_flush(acc, grad2d, __float_as_int(sB[cur_g].w));
                    acc_zero(acc);
                    cur_g = eg;
                }
                const float4 a = sA[eg], cn = sB[eg];
                const float w = s_pix[e & 255].x;
                const float dx = a.x - (X0f + (float)(e & 15)), dy = a.y - (Y0f + (float)((e >> 4) & 15));
                const float sigma = eg_sigma(cn.x, cn.y, cn.z, dx, dy);
                const float vis = eg_vis(sigma);
                const float ov = __fmul_rn(a.z, vis);
                if (sigma >= 0.0f && ov >= EG_ALPHA_MIN && ov <= EG_ALPHA_MAX) {
                    const float ra = __fdividef(1.0f, 1.0f - ov);
                    const float v_al = w * ra;
                    const float v_sigma = -ov * v_al;
                    const float gx = v_sigma * fmaf(cn.x, dx, cn.y * dy);
                    const float gy = v_sigma * fmaf(cn.y, dx, cn.z * dy);
                    const float hs = 0.5f * v_sigma;
                    acc.gx += gx;
                    acc.gy += gy;
                    acc.ax += fabsf(gx);
                    acc.ay += fabsf(gy);
                    acc.ca = fmaf(hs * dx, dx, acc.ca);
                    acc.cb = fmaf(v_sigma * dx, dy, acc.cb);
                    acc.cc = fmaf(hs * dy, dy, acc.cc);
                    acc.go = fmaf(vis, v_al, acc.go);
                }
            }
            if (__all_sync(0xffffffffu, p >= p_end)) break;
        }
        if (cur_g >= 0) acc_flush(acc, grad2d, __float_as_int(sB[cur_g].w));
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
