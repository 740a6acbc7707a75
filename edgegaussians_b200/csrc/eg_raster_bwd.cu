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
// Gaussian-major, in three divergence-free stages per batch of 256 Gaussians of the tile:
//   A. one thread per Gaussian: record + the pixel rectangle (inside the tile) that its
//      alpha >= 1/255 footprint can reach; block scan of the rectangle heights -> a list of
//      (Gaussian, pixel row) work items;
//   B. one thread per work item: the exact pixel span of that row inside the footprint ellipse
//      (sigma <= ln(255*opacity), solved analytically) intersected with the row's bitmask of pixels
//      whose backward seed is non-zero -> a 16-bit candidate mask per item;
//      a block scan of the popcounts linearises all candidate (Gaussian, pixel) pairs of the batch;
//   C. every thread takes an EQUAL contiguous slice of that candidate list: each step is one pair
//      that almost surely contributes, every lane of every warp does the same work.  The 8
//      per-Gaussian gradient values accumulate in registers (no warp reductions) and leave as two
//      128-bit vector reductions (red.global.add.v4.f32) when the slice moves on to the next
//      Gaussian (about 2 flushes per thread).
// The per-pixel state (seed * T_final, last contributor) lives in shared memory.
#include "eg_common.cuh"

namespace {

constexpr int RB_THREADS = 256;
constexpr int MAX_ITEMS = RB_THREADS * EG_TILE;  // (Gaussian, row) items of one batch

struct PairAcc {
    float gx, gy, ax, ay, ca, cb, cc, gs;  // gs = sum of v_sigma = -opacity * sum(vis * v_alpha)
};

__device__ __forceinline__ void acc_zero(PairAcc &a) { a.gx = a.gy = a.ax = a.ay = a.ca = a.cb = a.cc = a.gs = 0.0f; }

// v_opacity = sum(vis * v_alpha) = -gs / opacity
__device__ __forceinline__ void acc_flush(const PairAcc &a, float *__restrict__ grad2d, int gid, float opac) {
    // unconditional: candidates are pre-filtered, a segment without any contribution is rare (adds zeros)
    float *dst = grad2d + 8ll * gid;
    eg_red_add_v4(dst, a.gx, a.gy, a.ax, a.ay);
    eg_red_add_v4(dst + 4, a.ca, a.cb, a.cc, __fdividef(-a.gs, opac));
}

__global__ void __launch_bounds__(RB_THREADS) raster_bwd_kernel(
    const eg_config cfg, int tw, const float4 *__restrict__ rec, const int32_t *__restrict__ tile_offsets,
    const int32_t *__restrict__ flatten_ids, const int32_t *__restrict__ last_ids, const float *__restrict__ alpha,
    const float *__restrict__ v_render, int vr_ch, const float *__restrict__ v_alpha,
    const float *__restrict__ wpix, float seed_scale, float *__restrict__ grad2d,
    const int32_t *__restrict__ status) {
    __shared__ float2 s_pix[EG_TILE * EG_TILE];        // (seed * T_final, last contributor rel. to segment start)
    __shared__ unsigned s_rowmask[EG_TILE];            // pixels of the row with a non-zero seed
    __shared__ int s_rowlast[EG_TILE];                 // max last contributor over those pixels
    __shared__ __align__(16) float4 sA[RB_THREADS];    // mean2d.x, mean2d.y, opacity, packed pixel rectangle
    __shared__ __align__(16) float4 sB[RB_THREADS];    // conic a, b, c, gaussian id
    __shared__ __align__(16) float4 sC[RB_THREADS];    // 2*A*tau, det(conic), 1/A (0 = no row span), -
    __shared__ int s_wsum[RB_THREADS / 32];
    __shared__ unsigned s_item[MAX_ITEMS];             // candidate mask | row << 16 | Gaussian << 20
    __shared__ int s_ioff[MAX_ITEMS];                  // exclusive prefix of the items' candidate counts

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
        const unsigned bal = __ballot_sync(0xffffffffu, w != 0.0f);
        int ml = (w != 0.0f) ? last : -1;
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) ml = max(ml, __shfl_xor_sync(0xffffffffu, ml, d));
        if ((lane & 15) == 0) {
            s_rowmask[ly] = (lane == 0) ? (bal & 0xffffu) : (bal >> 16);
            s_rowlast[ly] = ml;
        }
    }
    const int xmax = min(EG_TILE, cfg.width - X0) - 1, ymax = min(EG_TILE, cfg.height - Y0) - 1;
    const float X0f = (float)X0 + 0.5f, Y0f = (float)Y0 + 0.5f;  // centre of pixel (0,0) of the tile

    for (int b0 = 0; b0 < L; b0 += RB_THREADS) {
        __syncthreads();  // pixel state visible / previous batch fully consumed
        // ---- A. one Gaussian per thread: record, reachable rectangle, row items ----
        const int k = b0 + tid;
        int nrows = 0;
        if (k < L) {
            const int gid = __ldg(flatten_ids + start + k);
            const float4 r0 = __ldg(rec + 2 * gid), r1 = __ldg(rec + 2 * gid + 1);
            float hx, hy, tau;
            int rect = 0;
            float two_tau_a = 0.0f, det = 0.0f, inv_a = 0.0f;
            if (eg_extent(r0.z, r1.x, r1.y, r1.z, hx, hy, tau)) {
                // pixel j (centre j + 0.5) is reachable iff  mx - hx <= j + 0.5 <= mx + hx
                const float fx0 = r0.x - (float)X0, fy0 = r0.y - (float)Y0;
                const int xlo = max(0, (int)ceilf(fminf(fx0 - hx - 0.5f, 64.0f)));
                const int xhi = min(xmax, (int)floorf(fmaxf(fx0 + hx - 0.5f, -64.0f)));
                const int ylo = max(0, (int)ceilf(fminf(fy0 - hy - 0.5f, 64.0f)));
                const int yhi = min(ymax, (int)floorf(fmaxf(fy0 + hy - 0.5f, -64.0f)));
                if (xlo <= xhi && ylo <= yhi) {
                    nrows = yhi - ylo + 1;
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
        int incl = nrows;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int wbase = 0, n_items = 0;
#pragma unroll
        for (int w = 0; w < RB_THREADS / 32; ++w) {
            const int v = s_wsum[w];
            if (w < warp) wbase += v;
            n_items += v;
        }
        if (n_items == 0) continue;  // uniform over the CTA
        const int excl = wbase + incl - nrows;
        {
            const int ylo = (__float_as_int(sA[tid].w) >> 8) & 15;
            for (int r = 0; r < nrows; ++r) s_item[excl + r] = ((unsigned)(ylo + r) << 16) | ((unsigned)tid << 20);
        }
        __syncthreads();

        // ---- B. (Gaussian, row) items: candidate pixel mask of the row; equal runs of items per thread ----
        const int per = (n_items + RB_THREADS - 1) / RB_THREADS;
        const int i0 = tid * per, i1 = min(n_items, i0 + per);
        int cnt = 0;
        for (int i = i0; i < i1; ++i) {
            const unsigned item = s_item[i];
            const int g = (int)(item >> 20), y = (int)((item >> 16) & 15u);
            const float4 a = sA[g], sp = sC[g];
            const int rect = __float_as_int(a.w);
            int xs = rect & 15, xe = (rect >> 4) & 15;
            unsigned mask = 0;
            if (b0 + g <= s_rowlast[y]) {
                if (sp.z != 0.0f) {
                    // sigma(dx, dy) <= tau  <=>  |dx - c| <= hw,  c = -B dy / A,  hw = sqrt(2 A tau - det dy^2) / A
                    const float dy = a.y - (Y0f + (float)y);
                    const float D = fmaf(-sp.y * dy, dy, sp.x);
                    if (D < 0.0f) {
                        xe = -1;
                    } else {
                        const float hw = sqrtf(D) * sp.z * 1.0001f + 2e-3f;
                        const float cpx = a.x + sB[g].y * dy * sp.z - X0f;  // in-tile pixel coordinate of the span centre
                        xs = max(xs, (int)ceilf(fminf(cpx - hw, 64.0f)));
                        xe = min(xe, (int)floorf(fmaxf(cpx + hw, -64.0f)));
                    }
                }
                if (xs <= xe) mask = ((2u << xe) - 1u) & ~((1u << xs) - 1u) & s_rowmask[y];
            }
            s_item[i] = item | mask;
            s_ioff[i] = cnt;  // thread-local exclusive candidate offset, rebased below
            cnt += __popc(mask);
        }
        // block scan of the per-thread candidate counts
        int cincl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, cincl, d);
            if (lane >= d) cincl += v;
        }
        __syncthreads();  // everyone is done reading s_wsum of stage A
        if (lane == 31) s_wsum[warp] = cincl;
        __syncthreads();
        int cbase = 0, n_cand = 0;
#pragma unroll
        for (int w = 0; w < RB_THREADS / 32; ++w) {
            const int v = s_wsum[w];
            if (w < warp) cbase += v;
            n_cand += v;
        }
        const int coff = cbase + cincl - cnt;
        for (int i = i0; i < i1; ++i) s_ioff[i] += coff;
        __syncthreads();
        if (n_cand == 0) continue;  // uniform over the CTA

        // ---- C. equal slices of the (implicit) candidate list; one (Gaussian, pixel) pair per step ----
        const int per_c = (n_cand + RB_THREADS - 1) / RB_THREADS;
        const int c0 = tid * per_c;
        int remaining = min(n_cand, c0 + per_c) - c0;
        if (remaining <= 0) continue;
        // largest item i with s_ioff[i] <= c0 that still has candidates beyond c0
        int lo = 0, hi = n_items;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_ioff[mid] <= c0) lo = mid; else hi = mid;
        }
        int i = lo;
        unsigned item = s_item[i];
        unsigned mask = item & 0xffffu;
        for (int q = c0 - s_ioff[i]; q > 0; --q) mask &= mask - 1;  // candidates of this item owned by the previous slice
        int g = (int)(item >> 20);
        float4 a = sA[g], cn = sB[g];
        EgFold f = eg_fold(cn.x, cn.y, cn.z, a.z);
        int y = (int)((item >> 16) & 15u);
        float dy = a.y - (Y0f + (float)y);
        int rowbase = y * EG_TILE, kk = b0 + g;
        PairAcc acc;
        acc_zero(acc);
        while (remaining > 0) {
            if (mask == 0) {  // next non-empty item (exists because remaining > 0)
                do {
                    item = s_item[++i];
                    mask = item & 0xffffu;
                } while (mask == 0);
                const int gn = (int)(item >> 20);
                if (gn != g) {
                    acc_flush(acc, grad2d, __float_as_int(cn.w), a.z);
                    acc_zero(acc);
                    g = gn;
                    a = sA[g];
                    cn = sB[g];
                    f = eg_fold(cn.x, cn.y, cn.z, a.z);
                    kk = b0 + g;
                }
                y = (int)((item >> 16) & 15u);
                dy = a.y - (Y0f + (float)y);
                rowbase = y * EG_TILE;
            }
            const int x = __ffs(mask) - 1;
            mask &= mask - 1;
            --remaining;
            const float2 pw = s_pix[rowbase + x];
            const float dx = a.x - (X0f + (float)x);
            const float pw2 = eg_pow2arg(f.fa, f.fb, f.fc, f.lo, dx, dy);
            const float ov = eg_ex2(pw2);  // opacity * exp(-sigma), exactly as the forward computed it
            if (kk <= __float_as_int(pw.y) && pw2 <= f.lo && ov >= EG_ALPHA_MIN && ov <= EG_ALPHA_MAX) {
                const float ra = __fdividef(1.0f, 1.0f - ov);
                const float v_sigma = -ov * pw.x * ra;
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
                acc.gs += v_sigma;
            }
        }
        acc_flush(acc, grad2d, __float_as_int(cn.w), a.z);
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
