// eg_splat_fwd.cu -- Gaussian-major forward of the fused training step (no tile lists, no sort):
//
//   eg_splat_fwd     every warp owns 32 Gaussians, walks the pixel rows of their alpha >= 1/255 footprints
//                    (eg_splat.cuh) and adds  log2(1 - alpha)  of every pair gsplat would composite into a
//                    per-pixel accumulator with 128-bit vector reductions (red.global.add.v4.f32, one per
//                    aligned 4-pixel chunk; the image is L2-resident).
//   eg_splat_resolve one CTA per 16x16 tile: T = 2^(sum) = prod (1 - alpha), render = alpha = 1 - T, the
//                    reference's clamp -> channel 0 -> mean |render - gt| (edge_gs.py:279,290-296;
//                    train_gaussians.py:84-94) and the per-pixel backward seed; re-zeroes the accumulator.
//   eg_emit_flagged  exactness fallback, see below.
//
// Why this is gsplat's result (colors == 1 as at edge_gs.py:247): a pixel's render
// is  sum_i alpha_i T_i = 1 - prod_i (1 - alpha_i)  over the Gaussians it composites, and it composites every
// Gaussian of its tile list that passes the alpha test UNLESS the running transmittance crosses the stop
// threshold (T (1 - alpha) <= 1e-4).  Prefix products only decrease, so if the order-free product over ALL
// passing Gaussians stays above the threshold no prefix in any order crossed it, gsplat never stopped and
// the order-free product IS its result (up to fp32 rounding of the product, ~1e-6 relative).  Tiles in which
// some pixel's product comes within 0.1 % of the threshold are flagged (tile_stop) and redone exactly:
// eg_emit_flagged appends the (depth, id) keys of the Gaussians that touch flagged tiles to per-tile buckets
// and eg_raster_fwd sorts and composites just those tiles front to back with the stop rule.  With the
// reference's translucent Gaussians (opacity 0.08 at init, configs/*.json:35) no tile is flagged and both
// fallback kernels return immediately.
#include <cstdlib>

#include "eg_splat.cuh"

namespace {

constexpr int SF_WARPS = 4;

__device__ __forceinline__ float eg_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void red_add_f32(float *addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// log2(1 - alpha) of one pair, 0 when gsplat skips it (sigma < 0 or alpha < 1/255)
template <bool CLIP>
__device__ __forceinline__ float pair_fwd(const EgSplatG &G, const float b1, const float c0, const float px,
                                          const bool in_span) {
    const float dx = G.mx - px;
    const float p = eg_pow2row(G.fa, b1, c0, dx);
    const float ov = eg_ex2(p);
    const float l = eg_lg2(1.0f - fminf(EG_ALPHA_MAX, ov));
    if (!CLIP) return eg_select_valid(l, ov, p, G.lo);
    return eg_pair_valid(ov, p, G.lo, in_span) ? l : 0.0f;
}

// log2(1 - alpha) of two adjacent pixels; the exponent runs on the packed fp32x2 pipe (element-wise bit-identical
// to eg_pow2row), MUFU and the validity select stay scalar.  ALIGNED rows only (no per-pixel clipping).
__device__ __forceinline__ void pair2_fwd(const EgSplatG &G, const eg_f2 mx2, const eg_f2 fa2, const eg_f2 b12,
                                          const eg_f2 c02, const eg_f2 npx, float &l0, float &l1) {
    const eg_f2 dx = f2_add(mx2, npx);
    float p0, p1;
    f2_unpack(f2_fma(f2_fma(fa2, dx, b12), dx, c02), p0, p1);
    const float ov0 = eg_ex2(p0), ov1 = eg_ex2(p1);
    l0 = eg_select_valid(eg_lg2(1.0f - fminf(EG_ALPHA_MAX, ov0)), ov0, p0, G.lo);
    l1 = eg_select_valid(eg_lg2(1.0f - fminf(EG_ALPHA_MAX, ov1)), ov1, p1, G.lo);
}

template <bool ALIGNED>
__device__ __forceinline__ void walk_row_fwd(const EgSplatG &G, const int y, const int W, float *__restrict__ logT) {
    const float dy = G.my - ((float)y + 0.5f);
    const float b1 = eg_pow2row_b1(G.fb, dy), c0 = eg_pow2row_c0(G.fc, G.lo, dy);
    int xa, xb;
    if (!eg_row_span(G, b1, c0, xa, xb)) return;
    int x = xa & ~3;
    const int xend = xb & ~3;
    float *ptr = logT + ((size_t)y * (size_t)W + (size_t)x);
    if (ALIGNED) {
        // W % 4 == 0 and the tile rectangle's columns are multiples of 16, so an aligned chunk that overlaps the span
        // lies entirely inside the rectangle -- no per-pixel clipping needed
        const eg_f2 mx2 = f2_dup(G.mx), fa2 = f2_dup(G.fa), b12 = f2_dup(b1), c02 = f2_dup(c0);
        float nb = -((float)x + 0.5f);  // negated pixel centre, exact; carried from chunk to chunk (no int->float per chunk)
        for (; x <= xend; x += 4, ptr += 4, nb -= 4.0f) {
            float v0, v1, v2, v3;
            pair2_fwd(G, mx2, fa2, b12, c02, f2_pack(nb, nb - 1.0f), v0, v1);
            pair2_fwd(G, mx2, fa2, b12, c02, f2_pack(nb - 2.0f, nb - 3.0f), v2, v3);
            if ((__float_as_uint(v0) | __float_as_uint(v1) | __float_as_uint(v2) | __float_as_uint(v3)) << 1)
                eg_red_add_v4(ptr, v0, v1, v2, v3);
        }
        return;
    }
    for (; x <= xend; x += 4, ptr += 4) {
        const float px = (float)x + 0.5f;  // pixel centres x + 0.5 .. x + 3.5 are exact in fp32
        const float v0 = pair_fwd<true>(G, b1, c0, px, x >= xa && x <= xb);
        const float v1 = pair_fwd<true>(G, b1, c0, px + 1.0f, x + 1 >= xa && x + 1 <= xb);
        const float v2 = pair_fwd<true>(G, b1, c0, px + 2.0f, x + 2 >= xa && x + 2 <= xb);
        const float v3 = pair_fwd<true>(G, b1, c0, px + 3.0f, x + 3 >= xa && x + 3 <= xb);
        if (v0 != 0.0f) red_add_f32(ptr, v0);
        if (v1 != 0.0f) red_add_f32(ptr + 1, v1);
        if (v2 != 0.0f) red_add_f32(ptr + 2, v2);
        if (v3 != 0.0f) red_add_f32(ptr + 3, v3);
    }
}

template <bool ALIGNED>
__global__ void __launch_bounds__(SF_WARPS * 32) splat_fwd_kernel(const eg_config cfg, const int tw, const int th,
                                                                  const float4 *__restrict__ rec,
                                                                  const int2 *__restrict__ gint,
                                                                  float *__restrict__ logT,
                                                                  const int32_t *__restrict__ status) {
    __shared__ EgSplatG s_g[SF_WARPS][32];
    if (status[EG_ST_OVERFLOW]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = (blockIdx.x * SF_WARPS + warp) * 32 + lane;
    EgSplatG G;
    int nrows = 0;
    if (g < cfg.n) {
        const int2 gi = __ldg(gint + g);
        nrows = eg_splat_setup(cfg, tw, th, g, __ldg(rec + 2 * g), __ldg(rec + 2 * g + 1), gi.x, G);
    }
    int incl = nrows;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    G.start = incl - nrows;
    const unsigned ne = __ballot_sync(0xffffffffu, nrows > 0);
    if (nrows > 0) s_g[warp][__popc(ne & ((1u << lane) - 1u))] = G;  // compacted: see EgOwnerIter
    const int R = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    EgOwnerIter it;
    for (int base = 0; base < R; base += 32) {
        const int item = base + lane;
        const int k = it.owner(base, lane, incl, nrows > 0);
        if (item < R) {
            const EgSplatG Go = s_g[warp][k];
            const int r0 = EG_ROWS_PER_ITEM * (item - Go.start);
#pragma unroll
            for (int r = 0; r < EG_ROWS_PER_ITEM; ++r)
                if (r0 + r < Go.nrows) walk_row_fwd<ALIGNED>(Go, Go.ylo + r0 + r, cfg.width, logT);
        }
    }
}

// Chunk-balanced variant (W % 4 == 0): lane = one aligned 4-pixel chunk.
// The row-per-lane walk above leaves lanes idle while the longest row of a 32-row batch finishes (69 % of the slots
// are used) and makes every lane of a reduction instruction touch a different 128-byte line.  Here a batch of 32 rows
// first derives its spans (lane = row), then the chunks of those rows form a dense list (same owner lookup as for the
// rows) that the warp consumes 32 chunks at a time: every lane evaluates exactly one chunk, and the chunks of a row
// sit in adjacent lanes, so their vector reductions share cache lines.
struct __align__(16) EgRowDesc {
    float mx, fa, b1, c0;   // eg_pow2row constants of the row
    float lo, nb0;          // log2(opacity'), negated centre of the row's first visited pixel
    int cstart, pad;        // index of the row's first chunk in the batch's chunk list
    float *ptr;             // &logT[y][4 * first chunk]
    long long pad2;
};

__global__ void __launch_bounds__(SF_WARPS * 32) splat_fwd_chunks_kernel(const eg_config cfg, const int tw, const int th,
                                                                         const float4 *__restrict__ rec,
                                                                         const int2 *__restrict__ gint,
                                                                         float *__restrict__ logT,
                                                                         const int32_t *__restrict__ status) {
    __shared__ EgSplatG s_g[SF_WARPS][32];
    __shared__ EgRowDesc s_row[SF_WARPS][32];
    if (status[EG_ST_OVERFLOW]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int g = (blockIdx.x * SF_WARPS + warp) * 32 + lane;
    EgSplatG G;
    G.nrows = 0;
    if (g < cfg.n) {
        const int2 gi = __ldg(gint + g);
        eg_splat_setup(cfg, tw, th, g, __ldg(rec + 2 * g), __ldg(rec + 2 * g + 1), gi.x, G);
    }
    const int nrows = G.nrows;
    int incl = nrows;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    G.start = incl - nrows;
    const unsigned ne = __ballot_sync(0xffffffffu, nrows > 0);
    if (nrows > 0) s_g[warp][__popc(ne & lt)] = G;  // compacted: see EgOwnerIter
    const int R = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    EgOwnerIter rows;
    for (int base = 0; base < R; base += 32) {
        // ---- lane = row: span -> chunk count, row descriptor ----
        const int item = base + lane;
        const int k = rows.owner(base, lane, incl, nrows > 0);
        int nch = 0;
        EgRowDesc D;
        if (item < R) {
            const EgSplatG Go = s_g[warp][k];
            const int y = Go.ylo + (item - Go.start);
            const float dy = Go.my - ((float)y + 0.5f);
            const float b1 = eg_pow2row_b1(Go.fb, dy), c0 = eg_pow2row_c0(Go.fc, Go.lo, dy);
            int xa, xb;
            if (eg_row_span(Go, b1, c0, xa, xb)) {
                const int c_first = xa >> 2;
                nch = (xb >> 2) - c_first + 1;
                D.mx = Go.mx; D.fa = Go.fa; D.b1 = b1; D.c0 = c0; D.lo = Go.lo;
                D.nb0 = -((float)(4 * c_first) + 0.5f);
                D.ptr = logT + ((size_t)y * (size_t)cfg.width + (size_t)(4 * c_first));
            }
        }
        int cincl = nch;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, cincl, d);
            if (lane >= d) cincl += t;
        }
        const int C = __shfl_sync(0xffffffffu, cincl, 31);
        const unsigned nz = __ballot_sync(0xffffffffu, nch > 0);
        __syncwarp();  // the previous batch's descriptors are dead
        if (nch > 0) {
            D.cstart = cincl - nch;
            s_row[warp][__popc(nz & lt)] = D;
        }
        __syncwarp();
        // ---- lane = chunk ----
        EgOwnerIter chunks;
        for (int cb = 0; cb < C; cb += 32) {
            const int ci = cb + lane;
            const int r = chunks.owner(cb, lane, cincl, nch > 0);
            if (ci < C) {
                const EgRowDesc Dr = s_row[warp][r];
                const int j = ci - Dr.cstart;
                const float nb = Dr.nb0 - (float)(4 * j);  // exact
                float l0, l1, l2, l3;
                {
                    const eg_f2 mx2 = f2_dup(Dr.mx), fa2 = f2_dup(Dr.fa), b12 = f2_dup(Dr.b1), c02 = f2_dup(Dr.c0);
                    float p0, p1, p2, p3;
                    const eg_f2 dxa = f2_add(mx2, f2_pack(nb, nb - 1.0f)), dxb = f2_add(mx2, f2_pack(nb - 2.0f, nb - 3.0f));
                    f2_unpack(f2_fma(f2_fma(fa2, dxa, b12), dxa, c02), p0, p1);
                    f2_unpack(f2_fma(f2_fma(fa2, dxb, b12), dxb, c02), p2, p3);
                    const float o0 = eg_ex2(p0), o1 = eg_ex2(p1), o2 = eg_ex2(p2), o3 = eg_ex2(p3);
                    l0 = eg_select_valid(eg_lg2(1.0f - fminf(EG_ALPHA_MAX, o0)), o0, p0, Dr.lo);
                    l1 = eg_select_valid(eg_lg2(1.0f - fminf(EG_ALPHA_MAX, o1)), o1, p1, Dr.lo);
                    l2 = eg_select_valid(eg_lg2(1.0f - fminf(EG_ALPHA_MAX, o2)), o2, p2, Dr.lo);
                    l3 = eg_select_valid(eg_lg2(1.0f - fminf(EG_ALPHA_MAX, o3)), o3, p3, Dr.lo);
                }
                if ((__float_as_uint(l0) | __float_as_uint(l1) | __float_as_uint(l2) | __float_as_uint(l3)) << 1)
                    eg_red_add_v4(Dr.ptr + 4 * j, l0, l1, l2, l3);
            }
        }
    }
}

// Persistent: a few CTAs per SM stride over the tiles and issue ONE loss atomic each at the end (one atomic per
// tile would serialise 7 500 fp64 additions on a single L2 address -- measured 60 us).
template <int GT_KIND>
__global__ void __launch_bounds__(256) splat_resolve_kernel(const eg_config cfg, const int tw, const int n_tiles,
                                                            float *__restrict__ logT, const void *__restrict__ gt,
                                                            double *__restrict__ loss_sum, float *__restrict__ wpix,
                                                            float *__restrict__ render0, float *__restrict__ alpha_out,
                                                            int32_t *__restrict__ tile_stop,
                                                            int32_t *__restrict__ stop_list,
                                                            const float *__restrict__ loss_params,
                                                            const unsigned char *__restrict__ sel_mask,
                                                            int32_t *__restrict__ status) {
    __shared__ float s_red[8];
    if (status[EG_ST_OVERFLOW]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float absd = 0.0f;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tile_y = tile / tw, tile_x = tile - tile_y * tw;
        // a warp covers 2 rows x 16 pixels: each row segment is one 64-byte run
        const int pxi = tile_x * EG_TILE + (tid & 15), pyi = tile_y * EG_TILE + (tid >> 4);
        const bool inside = pxi < cfg.width && pyi < cfg.height;
        const long long pix = (long long)pyi * cfg.width + pxi;
        float T = 1.0f;
        if (inside) {
            T = eg_ex2(logT[pix]);
            logT[pix] = 0.0f;  // clean for the next iteration's reductions
        }
        // some pixel of the tile may have hit gsplat's stop rule: the tile is redone exactly by eg_raster_fwd
        if (__syncthreads_or(inside && !(T > EG_T_MIN * 1.001f))) {
            if (tid == 0) {
                tile_stop[tile] = 1;
                stop_list[atomicAdd(status + EG_ST_STOPPED, 1)] = tile;
            }
            continue;
        }
        if (inside) {
            const float out = 1.0f - T;
            if (alpha_out) alpha_out[pix] = out;
            if (render0) render0[pix] = out;
            if (GT_KIND != EG_GT_NONE) {
                float gv;
                if (GT_KIND == EG_GT_F32) gv = __ldg(reinterpret_cast<const float *>(gt) + pix);
                else gv = __fdiv_rn((float)__ldg(reinterpret_cast<const unsigned char *>(gt) + pix), 255.0f);
                const float d = fminf(fmaxf(out, 0.0f), 1.0f) - gv;
                const float coef = eg_loss_coef(loss_params, sel_mask, gv, pix);
                absd += coef * fabsf(d);
                const float sgn = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
                const float pass = (out >= 0.0f && out <= 1.0f) ? 1.0f : 0.0f;
                if (wpix) wpix[pix] = sgn * pass * T * coef;
            }
        }
    }
    if (GT_KIND != EG_GT_NONE && loss_sum != nullptr) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) absd += __shfl_xor_sync(0xffffffffu, absd, d);
        if (lane == 0) s_red[warp] = absd;
        __syncthreads();
        if (tid == 0) {
            double tsum = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) tsum += (double)s_red[w];
            if (tsum != 0.0) atomicAdd(loss_sum, tsum);
        }
    }
}

// Vectorised variant for W % 4 == 0 (the common case): a CTA handles 4 horizontally adjacent tiles per iteration
// (64 x 16 pixels = 256 threads x 4 pixels), every access is a 128-bit one on a 256-byte contiguous row run, and
// both loads of an iteration (accumulator, edge map) are issued before the first barrier.
template <int GT_KIND>
__global__ void __launch_bounds__(256) splat_resolve4_kernel(const eg_config cfg, const int tw, const int th,
                                                             float *__restrict__ logT, const void *__restrict__ gt,
                                                             double *__restrict__ loss_sum, float *__restrict__ wpix,
                                                             float *__restrict__ render0, float *__restrict__ alpha_out,
                                                             int32_t *__restrict__ tile_stop,
                                                             int32_t *__restrict__ stop_list,
                                                             const float *__restrict__ loss_params,
                                                             const unsigned char *__restrict__ sel_mask,
                                                             int32_t *__restrict__ status) {
    __shared__ float s_red[8];
    __shared__ int s_flag[4];
    if (status[EG_ST_OVERFLOW]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid >> 4, c4 = tid & 15, k = c4 >> 2;
    const int ngx = (tw + 3) >> 2, n_groups = ngx * th;
    float absd = 0.0f;
    // coefficients of the fused loss (eg_loss_coef): 1 everywhere unless loss_params is given
    const bool weighted = loss_params != nullptr;
    float w_edge = 1.0f, w_bg = 1.0f, w_sel = 0.0f, w_thr = 0.0f;
    if (weighted) {
        w_edge = __ldg(loss_params); w_bg = __ldg(loss_params + 1); w_sel = __ldg(loss_params + 2); w_thr = __ldg(loss_params + 3);
    }
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int gy = grp / ngx, gx = grp - gy * ngx;
        const int tile_x = gx * 4 + k;
        const int x = gx * 64 + c4 * 4, y = gy * EG_TILE + row;
        const bool inside = x < cfg.width && y < cfg.height;  // W % 4 == 0: the 4 pixels are in or out together
        const long long pix = (long long)y * cfg.width + x;
        if (tid < 4) s_flag[tid] = 0;
        float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = l4;
        if (inside) {
            l4 = *reinterpret_cast<const float4 *>(logT + pix);
            if (GT_KIND == EG_GT_F32) {
                g4 = __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(gt) + pix));
            } else if (GT_KIND == EG_GT_U8) {
                const uchar4 u = __ldg(reinterpret_cast<const uchar4 *>(reinterpret_cast<const unsigned char *>(gt) + pix));
                g4 = make_float4(__fdiv_rn((float)u.x, 255.0f), __fdiv_rn((float)u.y, 255.0f),
                                 __fdiv_rn((float)u.z, 255.0f), __fdiv_rn((float)u.w, 255.0f));
            }
        }
        const float T[4] = {eg_ex2(l4.x), eg_ex2(l4.y), eg_ex2(l4.z), eg_ex2(l4.w)};
        const float thr = EG_T_MIN * 1.001f;
        const bool cand = inside && !(T[0] > thr && T[1] > thr && T[2] > thr && T[3] > thr);
        __syncthreads();
        if (cand) s_flag[k] = 1;  // some pixel of tile k may have hit gsplat's stop rule
        __syncthreads();
        const bool flagged = s_flag[k] != 0;
        if (inside) *reinterpret_cast<float4 *>(logT + pix) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (flagged) {  // the tile is redone exactly by eg_raster_fwd
            if (row == 0 && (c4 & 3) == 0 && tile_x < tw) {
                const int tile = gy * tw + tile_x;
                tile_stop[tile] = 1;
                stop_list[atomicAdd(status + EG_ST_STOPPED, 1)] = tile;
            }
        } else if (inside) {
            float o4[4], w4[4];
            const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
            uchar4 sel = make_uchar4(0, 0, 0, 0);
            if (weighted && sel_mask != nullptr) sel = __ldg(reinterpret_cast<const uchar4 *>(sel_mask + pix));
            const unsigned char sl[4] = {sel.x, sel.y, sel.z, sel.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                o4[i] = 1.0f - T[i];
                const float d = fminf(fmaxf(o4[i], 0.0f), 1.0f) - gv[i];
                float coef = 1.0f;
                if (weighted) coef = (gv[i] >= w_thr ? w_edge : w_bg) + (sl[i] ? w_sel : 0.0f);
                absd += coef * fabsf(d);
                const float sgn = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
                const float pass = (o4[i] >= 0.0f && o4[i] <= 1.0f) ? 1.0f : 0.0f;
                w4[i] = sgn * pass * T[i] * coef;
            }
            const float4 ov = make_float4(o4[0], o4[1], o4[2], o4[3]);
            if (alpha_out) *reinterpret_cast<float4 *>(alpha_out + pix) = ov;
            if (render0) *reinterpret_cast<float4 *>(render0 + pix) = ov;
            if (GT_KIND != EG_GT_NONE && wpix)
                *reinterpret_cast<float4 *>(wpix + pix) = make_float4(w4[0], w4[1], w4[2], w4[3]);
        }
        __syncthreads();  // s_flag is reset at the top of the next iteration
    }
    if (GT_KIND != EG_GT_NONE && loss_sum != nullptr) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) absd += __shfl_xor_sync(0xffffffffu, absd, d);
        if (lane == 0) s_red[warp] = absd;
        __syncthreads();
        if (tid == 0) {
            double tsum = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) tsum += (double)s_red[w];
            if (tsum != 0.0) atomicAdd(loss_sum, tsum);
        }
    }
}

// keys of the Gaussians that touch FLAGGED tiles -> per-tile buckets (only runs when a tile was flagged)
__global__ void __launch_bounds__(256) emit_flagged_kernel(const eg_config cfg, const int tw, const int th,
                                                           const float4 *__restrict__ rec,
                                                           const int2 *__restrict__ gint,
                                                           const int32_t *__restrict__ tile_stop,
                                                           int32_t *__restrict__ tile_cnt,
                                                           unsigned long long *__restrict__ keys,
                                                           int32_t *__restrict__ status) {
    if (status[EG_ST_STOPPED] == 0 || status[EG_ST_OVERFLOW]) return;  // the usual case: a small grid, gone at once
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < cfg.n; g += gridDim.x * blockDim.x) {
        const int2 gi = __ldg(gint + g);
        if (gi.x <= 0 || gi.y <= 0) continue;
        const float4 r0 = __ldg(rec + 2 * g);
        uint32_t x0, y0, x1, y1;
        eg_tile_rect(r0.x, r0.y, gi.x, tw, th, x0, y0, x1, y1);
        const unsigned long long key = ((unsigned long long)__float_as_uint(r0.w) << 32) | (unsigned int)g;
        const bool cull = (cfg.flags & EG_FLAG_CULL_TILES) != 0;  // only the tiles the footprint can reach
        float hu = 1e30f, hv = 1e30f, tau = 0.0f;
        float4 r1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cull) {
            r1 = __ldg(rec + 2 * g + 1);
            if (!eg_extent(r0.z, r1.x, r1.y, r1.z, hu, hv, tau)) continue;
        }
        for (uint32_t i = y0; i < y1; ++i) {
            int j0 = (int)x0, j1 = (int)x1 - 1;
            if (cull && !eg_tile_row_cols(r0.x, r0.y, r1.x, r1.y, r1.z, tau, hu, hv, (int)i, (int)x0, (int)x1, j0, j1)) continue;
            for (int j = j0; j <= j1; ++j) {
                const size_t t = (size_t)i * tw + j;
                if (__ldg(tile_stop + t) == 0) continue;
                const int pos = atomicAdd(tile_cnt + t, 1);
                if (pos < cfg.tile_capacity) keys[t * (size_t)cfg.tile_capacity + pos] = key;
                else {  // bucket too small: the caller re-runs with tile_capacity >= status[EG_ST_MAXTILE]
                    status[EG_ST_OVERFLOW] = 1;
                    atomicMax(status + EG_ST_MAXTILE, pos + 1);
                }
            }
        }
    }
}

}  // namespace

extern "C" int eg_splat_fwd(const eg_config *cfg, const float *rec, const int32_t *gint, float *logT,
                            const int32_t *status, void *stream) {
    if (cfg == nullptr || cfg->tile_size != EG_TILE) {
        eg_set_error("eg_splat_fwd: tile_size must be %d", EG_TILE);
        return 1;
    }
    if (rec == nullptr || gint == nullptr || logT == nullptr || status == nullptr) {
        eg_set_error("eg_splat_fwd: rec, gint, logT and status are required");
        return 1;
    }
    if (cfg->n <= 0) return 0;
    if (cfg->width >= 65536 || cfg->height >= 65536) {
        eg_set_error("eg_splat_fwd: image sides must be < 65536");
        return 1;
    }
    int tw, th;
    eg_tile_grid(cfg->width, cfg->height, cfg->tile_size, &tw, &th);
    const int block = SF_WARPS * 32, grid = (cfg->n + block - 1) / block;
    cudaStream_t s = (cudaStream_t)stream;
    // EG_FWD_ROWS=1 selects the row-per-lane walk on aligned images too (A/B measurements)
    static const bool rows_only = getenv("EG_FWD_ROWS") != nullptr && getenv("EG_FWD_ROWS")[0] == '1';
    if ((cfg->width % 4 == 0) && (((uintptr_t)logT & 15) == 0) && !rows_only)
        splat_fwd_chunks_kernel<<<grid, block, 0, s>>>(*cfg, tw, th, (const float4 *)rec, (const int2 *)gint, logT, status);
    else if ((cfg->width % 4 == 0) && (((uintptr_t)logT & 15) == 0))
        splat_fwd_kernel<true><<<grid, block, 0, s>>>(*cfg, tw, th, (const float4 *)rec, (const int2 *)gint, logT, status);
    else
        splat_fwd_kernel<false><<<grid, block, 0, s>>>(*cfg, tw, th, (const float4 *)rec, (const int2 *)gint, logT, status);
    return eg_check_launch("eg_splat_fwd");
}

extern "C" int eg_splat_resolve(const eg_config *cfg, float *logT, const void *gt, int gt_kind, double *loss_sum,
                                float *wpix, float *render0, float *alpha, int32_t *tile_stop, int32_t *stop_list,
                                const float *loss_params, const uint8_t *sel_mask, int32_t *status, void *stream) {
    if (cfg == nullptr || cfg->tile_size != EG_TILE) {
        eg_set_error("eg_splat_resolve: tile_size must be %d", EG_TILE);
        return 1;
    }
    if (logT == nullptr || tile_stop == nullptr || stop_list == nullptr || status == nullptr) {
        eg_set_error("eg_splat_resolve: logT, tile_stop, stop_list and status are required");
        return 1;
    }
    if (gt == nullptr) gt_kind = EG_GT_NONE;
    if ((gt_kind == EG_GT_NONE || loss_params == nullptr) && sel_mask != nullptr) {
        eg_set_error("eg_splat_resolve: sel_mask needs gt and loss_params");
        return 1;
    }
    if (gt_kind == EG_GT_NONE) loss_params = nullptr;
    int tw, th;
    eg_tile_grid(cfg->width, cfg->height, cfg->tile_size, &tw, &th);
    cudaStream_t s = (cudaStream_t)stream;
    const int n_tiles = tw * th;
    auto al16 = [](const void *p) { return ((uintptr_t)p & 15) == 0; };
    const bool vec = cfg->width % 4 == 0 && al16(logT) && al16(wpix) && al16(render0) && al16(alpha) &&
                     (gt_kind == EG_GT_U8 ? ((uintptr_t)gt & 3) == 0 : al16(gt)) && ((uintptr_t)sel_mask & 3) == 0;
    const int n_units = vec ? ((tw + 3) / 4) * th : n_tiles;
    const int grid = n_units < 148 * 8 ? n_units : 148 * 8;
#define EG_RS_LAUNCH(KIND)                                                                                          \
    do {                                                                                                            \
        if (vec)                                                                                                    \
            splat_resolve4_kernel<KIND><<<grid, 256, 0, s>>>(*cfg, tw, th, logT, gt, loss_sum, wpix, render0,       \
                                                             alpha, tile_stop, stop_list, loss_params, sel_mask,    \
                                                             status);                                               \
        else                                                                                                        \
            splat_resolve_kernel<KIND><<<grid, 256, 0, s>>>(*cfg, tw, n_tiles, logT, gt, loss_sum, wpix, render0,   \
                                                            alpha, tile_stop, stop_list, loss_params, sel_mask,     \
                                                            status);                                                \
    } while (0)
    switch (gt_kind) {
        case EG_GT_NONE: EG_RS_LAUNCH(EG_GT_NONE); break;
        case EG_GT_F32: EG_RS_LAUNCH(EG_GT_F32); break;
        case EG_GT_U8: EG_RS_LAUNCH(EG_GT_U8); break;
        default: eg_set_error("eg_splat_resolve: bad gt_kind %d", gt_kind); return 1;
    }
#undef EG_RS_LAUNCH
    return eg_check_launch("eg_splat_resolve");
}

extern "C" int eg_emit_flagged(const eg_config *cfg, const float *rec, const int32_t *gint, const int32_t *tile_stop,
                               int32_t *tile_cnt, uint64_t *keys, int32_t *status, void *stream) {
    if (cfg == nullptr || cfg->tile_size != EG_TILE) {
        eg_set_error("eg_emit_flagged: tile_size must be %d", EG_TILE);
        return 1;
    }
    if (cfg->n <= 0) return 0;
    int tw, th;
    eg_tile_grid(cfg->width, cfg->height, cfg->tile_size, &tw, &th);
    const int blocks = (cfg->n + 255) / 256, grid = blocks < 148 * 4 ? blocks : 148 * 4;
    emit_flagged_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        *cfg, tw, th, (const float4 *)rec, (const int2 *)gint, tile_stop, tile_cnt, (unsigned long long *)keys, status);
    return eg_check_launch("eg_emit_flagged");
}
