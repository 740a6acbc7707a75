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
// is thus a plain sum over the (pixel, Gaussian) pairs the forward composited, with no ordering
// dependence.  The forward kernel recorded those pairs as a 256-bit contribution mask per tile
// intersection (eg_raster_fwd, `cmask`), so this kernel
//   A. stages 256 Gaussians of the tile per batch (record + mask), block-scans the mask popcounts;
//   B. gives every thread an EQUAL contiguous slice of the batch's pair list and walks its set bits:
//      each step is one pair that did contribute (no footprint, alpha or stop tests to redo).  The
//      8 per-Gaussian gradient values accumulate in registers (no warp reductions) and leave as two
//      128-bit vector reductions (red.global.add.v4.f32) when the slice moves on to the next Gaussian
//      (about 2 flushes per thread).
// The per-pixel seed (seed * T_final) lives in shared memory.
#include "eg_common.cuh"

namespace {

constexpr int RB_THREADS = 256;

struct PairAcc {
    float gx, gy, ax, ay, ca, cb, cc, gs;  // gs = sum of v_sigma = -opacity * sum(vis * v_alpha)
};

__device__ __forceinline__ void acc_zero(PairAcc &a) { a.gx = a.gy = a.ax = a.ay = a.ca = a.cb = a.cc = a.gs = 0.0f; }

// v_opacity = sum(vis * v_alpha) = -gs / opacity
__device__ __forceinline__ void acc_flush(const PairAcc &a, float *__restrict__ grad2d, int gid, float opac) {
    float *dst = grad2d + 8ll * gid;
    eg_red_add_v4(dst, a.gx, a.gy, a.ax, a.ay);
    eg_red_add_v4(dst + 4, a.ca, a.cb, a.cc, -a.gs * eg_rcp(opac));
}

__global__ void __launch_bounds__(RB_THREADS) raster_bwd_kernel(
    const eg_config cfg, int tw, const float4 *__restrict__ rec, const int32_t *__restrict__ tile_offsets,
    const int32_t *__restrict__ flatten_ids, const uint4 *__restrict__ cmask, const int32_t *__restrict__ tile_done,
    const float *__restrict__ alpha, const float *__restrict__ v_render, int vr_ch, const float *__restrict__ v_alpha,
    const float *__restrict__ wpix, float seed_scale, float *__restrict__ grad2d,
    const int32_t *__restrict__ status) {
    __shared__ __align__(16) float4 s_px[EG_TILE * EG_TILE];  // per pixel: seed * T_final (0 outside the image),
                                                              // pixel centre x, y, -
    __shared__ __align__(16) float4 sA[RB_THREADS];    // mean2d.x, mean2d.y, opacity, gaussian id
    __shared__ __align__(16) float4 sB[RB_THREADS];    // conic a, b, c, -
    __shared__ __align__(16) float4 sC[RB_THREADS];    // folded conic fa, fb, fc, log2(opacity)  (eg_fold)
    __shared__ __align__(16) uint32_t s_cm[RB_THREADS * 8];  // contribution masks of the batch
    __shared__ int s_off[RB_THREADS + 1];              // exclusive prefix of the mask popcounts
    __shared__ unsigned char s_nz[RB_THREADS];         // which of the 8 mask words are non-zero
    __shared__ __align__(16) int s_wsum[RB_THREADS / 32];  // 16-byte aligned: read back as two 128-bit loads (unaligned, ptxas widens them over the tail of s_nz: a racecheck false positive)

    if (status[EG_ST_OVERFLOW]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tile_y = tile / tw, tile_x = tile - tile_y * tw;
    const int start = tile_offsets[tile];
    // EG_FLAG_FRONT_SORT: the forward stopped after tile_done[tile] entries (everything behind is invisible and its
    // flatten_ids / cmask are undefined)
    const int L = tile_done != nullptr ? min(tile_done[tile], tile_offsets[tile + 1] - start) : tile_offsets[tile + 1] - start;
    if (L <= 0) return;
    const int X0 = tile_x * EG_TILE, Y0 = tile_y * EG_TILE;

    {
        const int lx = tid & 15, ly = tid >> 4;
        const int pxi = X0 + lx, pyi = Y0 + ly;
        float w = 0.0f;
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
        }
        s_px[tid] = make_float4(w, (float)pxi + 0.5f, (float)pyi + 0.5f, 0.0f);
    }

    for (int b0 = 0; b0 < L; b0 += RB_THREADS) {
        __syncthreads();  // pixel seeds visible / previous batch fully consumed
        // ---- A. one Gaussian per thread: record + contribution mask; block scan of the pair counts ----
        // Gaussians that composited nowhere in the tile are dropped here: the staged arrays are DENSE in the
        // Gaussians that have pairs (one packed scan gives both the pair offset and the dense rank).
        const int k = b0 + tid;
        int cnt = 0, gid = 0;
        uint4 c0 = make_uint4(0u, 0u, 0u, 0u), c1 = c0;
        if (k < L) {
            gid = __ldg(flatten_ids + start + k);
            c0 = __ldg(cmask + 2 * (size_t)(start + k));
            c1 = __ldg(cmask + 2 * (size_t)(start + k) + 1);
            cnt = __popc(c0.x) + __popc(c0.y) + __popc(c0.z) + __popc(c0.w) + __popc(c1.x) + __popc(c1.y) +
                  __popc(c1.z) + __popc(c1.w);
        }
        const int packed = cnt | ((cnt > 0) << 20);  // pairs (<= 2^16 per batch) | non-empty flag
        int incl = packed;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
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
        const int n_pairs = total & 0xfffff, n_live = total >> 20;
        const int excl = wbase + incl - packed;
        if (cnt > 0) {
            const int d = excl >> 20;  // dense index of this Gaussian
            const float4 r0 = __ldg(rec + 2 * gid), r1 = __ldg(rec + 2 * gid + 1);
            sA[d] = make_float4(r0.x, r0.y, r0.z, __int_as_float(gid));
            sB[d] = r1;
            const EgFold f0 = eg_fold(r1.x, r1.y, r1.z, r0.z);
            sC[d] = make_float4(f0.fa, f0.fb, f0.fc, f0.lo);
            reinterpret_cast<uint4 *>(s_cm)[2 * d] = c0;
            reinterpret_cast<uint4 *>(s_cm)[2 * d + 1] = c1;
            s_nz[d] = (unsigned char)((c0.x != 0) | ((c0.y != 0) << 1) | ((c0.z != 0) << 2) | ((c0.w != 0) << 3) |
                                      ((c1.x != 0) << 4) | ((c1.y != 0) << 5) | ((c1.z != 0) << 6) | ((c1.w != 0) << 7));
            s_off[d] = excl & 0xfffff;
        }
        if (tid == 0) s_off[n_live] = n_pairs;
        __syncthreads();
        if (n_pairs == 0) continue;  // uniform over the CTA

        // ---- B. equal slices of the pair list; one composited (Gaussian, pixel) pair per step ----
        const int per = (n_pairs + RB_THREADS - 1) / RB_THREADS;
        const int c0i = tid * per;
        int remaining = min(n_pairs, c0i + per) - c0i;
        if (remaining <= 0) continue;
        // Gaussian holding pair c0i: the g with s_off[g] <= c0i < s_off[g+1]
        int lo = 0, hi = n_live;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_off[mid] <= c0i) lo = mid; else hi = mid;
        }
        int g = lo;
        unsigned nzleft = s_nz[g];
        unsigned mask = 0;
        int wv = 0;
        {
            int skip = c0i - s_off[g];  // pairs of this Gaussian owned by the previous slice
            for (;;) {
                wv = __ffs(nzleft) - 1;
                nzleft &= nzleft - 1;
                mask = s_cm[g * 8 + wv];
                const int pc = __popc(mask);
                if (skip < pc) break;
                skip -= pc;
            }
            for (; skip > 0; --skip) mask &= mask - 1;
        }
        float4 a = sA[g], cn = sB[g], f = sC[g];
        int ib = ((wv >> 1) << 6) | ((wv & 1) << 3);  // tile-local index of the word's first pixel
        PairAcc acc;
        acc_zero(acc);
        while (remaining > 0) {
            if (mask == 0) {  // next non-empty mask word (exists because remaining > 0)
                if (nzleft == 0) {  // ... of the next Gaussian that has pairs
                    acc_flush(acc, grad2d, __float_as_int(a.w), a.z);
                    acc_zero(acc);
                    ++g;  // dense: the next staged Gaussian has pairs
                    nzleft = s_nz[g];
                    a = sA[g];
                    cn = sB[g];
                    f = sC[g];
                }
                wv = __ffs(nzleft) - 1;
                nzleft &= nzleft - 1;
                mask = s_cm[g * 8 + wv];
                ib = ((wv >> 1) << 6) | ((wv & 1) << 3);
            }
            const int l = __ffs(mask) - 1;
            mask &= mask - 1;
            --remaining;
            const float4 pxl = s_px[ib + l + (l & 24)];  // word bit l -> pixel (l & 7, l >> 3) of the 8x4 block
            const float w = pxl.x;
            const float dx = a.x - pxl.y, dy = a.y - pxl.z;
            const float pw2 = eg_pow2arg(f.x, f.y, f.z, f.w, dx, dy);
            const float ov = eg_ex2(pw2);  // opacity * exp(-sigma), exactly as the forward computed it
            if (ov <= EG_ALPHA_MAX) {      // gsplat: no gradient through a clamped alpha
                const float ra = eg_rcp(1.0f - ov);
                const float v_sigma = -ov * w * ra;
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
        acc_flush(acc, grad2d, __float_as_int(a.w), a.z);
    }
}

}  // namespace

extern "C" int eg_raster_bwd(const eg_config *cfg, const float *rec, const int32_t *tile_offsets,
                             const int32_t *flatten_ids, const uint32_t *cmask, const int32_t *tile_done, const float *alpha,
                             const float *v_render, int v_render_channels, const float *v_alpha, const float *wpix,
                             float seed_scale, float *grad2d, const int32_t *status, void *stream) {
    if (cfg == nullptr || cfg->tile_size != EG_TILE) {
        eg_set_error("eg_raster_bwd: tile_size must be %d", EG_TILE);
        return 1;
    }
    if (cmask == nullptr) {
        eg_set_error("eg_raster_bwd: the contribution masks written by eg_raster_fwd are required");
        return 1;
    }
    if (wpix == nullptr && alpha == nullptr) {
        eg_set_error("eg_raster_bwd: need either wpix or alpha (+ v_render / v_alpha)");
        return 1;
    }
    int tw, th;
    eg_tile_grid(cfg->width, cfg->height, cfg->tile_size, &tw, &th);
    raster_bwd_kernel<<<tw * th, RB_THREADS, 0, (cudaStream_t)stream>>>(
        *cfg, tw, (const float4 *)rec, tile_offsets, flatten_ids, (const uint4 *)cmask, tile_done, alpha, v_render,
        v_render_channels, v_alpha, wpix, seed_scale, grad2d, status);
    return eg_check_launch("eg_raster_bwd");
}
