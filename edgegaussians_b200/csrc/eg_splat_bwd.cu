// eg_splat_bwd.cu -- K6 + K7 in one kernel: Gaussian-major compositing backward (with abs-grad) fused with the
// projection backward, the activation VJPs and update_absgrads.  No tile lists, no atomics, no 2D-gradient
// round trip through HBM.
//
// Semantics: SURVEY.md Appendix A.5 + A.6 (gsplat==1.0.0 rasterize_to_pixels bwd + fully_fused_projection bwd
// behind /root/reference/edgegaussians/models/edge_gs.py:250-268, absgrad=True edge_gs.py:266, update_absgrads
// edge_gs.py:603-613).
//
// With colors == 1 (edge_gs.py:247) every render channel is 1 - prod_i (1 - alpha_i), so
//     d out(p) / d alpha_k = T_final(p) / (1 - alpha_k)      for every k that pixel p composited,
// a plain sum over composited (pixel, Gaussian) pairs with no ordering dependence (see eg_raster_bwd.cu).
// Which pairs were composited is decidable from the Gaussian's side (eg_splat.cuh): tile rectangle, sigma >= 0,
// alpha >= 1/255 -- and, for the pixels that hit gsplat's transmittance stop, "sort key of g <= key of the last
// Gaussian the pixel composited", which the forward kernels leave in two per-pixel planes (last_depth,
// last_gid; last_depth = 0xffffffff where the pixel never stopped).  So:
//   phase 1  lane = Gaussian : record -> tile rectangle, folded conic, row range (eg_splat_setup);
//   phase 2  lane = (Gaussian, EG_ROWS_PER_ITEM rows) : walk each row's span in aligned 4-pixel chunks (one LDG.128
//            of the per-pixel seed  w_p = seed * T_final(p)  per chunk, the next chunk's load in flight), two
//            pixels per instruction on the packed fp32x2 pipe; accumulate the row's moments S0 = sum v_sigma,
//            S1 = sum v_sigma dx, S2 = sum v_sigma dx^2 and the two abs-sums in registers, turn them into the 8
//            per-Gaussian 2D gradients, segmented-reduce over the lanes of the same Gaussian (shuffles) and
//            add into the warp's private shared-memory accumulators;
//   phase 3  lane = Gaussian : projection VJP + exp/sigmoid VJP + abs-grad norm, gradients written with plain
//            stores (each Gaussian has exactly one owner).
#include <cstdlib>

#include "eg_project_vjp.cuh"
#include "eg_splat.cuh"

namespace {

constexpr int SB_WARPS = 4;
#ifndef EG_SB_MINBLOCKS
#define EG_SB_MINBLOCKS 6   // resident CTAs per SM the register allocation is bounded for
#endif

template <bool ALIGNED>
__device__ __forceinline__ float4 ld_f4(const float *__restrict__ row, int c, int xlim) {
    if (ALIGNED) return __ldg(reinterpret_cast<const float4 *>(row) + c);
    float4 v;
    const int x = 4 * c;
    v.x = (x < xlim) ? __ldg(row + x) : 0.0f;
    v.y = (x + 1 < xlim) ? __ldg(row + x + 1) : 0.0f;
    v.z = (x + 2 < xlim) ? __ldg(row + x + 2) : 0.0f;
    v.w = (x + 3 < xlim) ? __ldg(row + x + 3) : 0.0f;
    return v;
}
template <bool ALIGNED>
__device__ __forceinline__ uint4 ld_u4(const unsigned *__restrict__ row, int c, int xlim) {
    if (ALIGNED) return __ldg(reinterpret_cast<const uint4 *>(row) + c);
    uint4 v;
    const int x = 4 * c;
    v.x = (x < xlim) ? __ldg(row + x) : 0u;
    v.y = (x + 1 < xlim) ? __ldg(row + x + 1) : 0u;
    v.z = (x + 2 < xlim) ? __ldg(row + x + 2) : 0u;
    v.w = (x + 3 < xlim) ? __ldg(row + x + 3) : 0u;
    return v;
}

struct RowAcc {
    float S0, S1, S2, ax, ay;
};

template <bool HAS_LAST, bool CLIP>
__device__ __forceinline__ void pair_bwd(const EgSplatG &G, const float b1, const float c0, const float Bdy,
                                         const float Cdy, const float px, const bool in_span, const float w,
                                         const unsigned dl, const int *__restrict__ gid_px, RowAcc &r) {
    const float dx = G.mx - px;
    const float p = eg_pow2row(G.fa, b1, c0, dx);  // log2(opacity * exp(-sigma)), bit-identical to the forward's
    const float ov = eg_ex2(p);
    // composited (sigma >= 0, alpha >= 1/255) and alpha not clamped (gsplat: no gradient through the clamp)
    const float ra = eg_rcp(1.0f - ov);
    float vs = (ov * w) * ra;  // -v_sigma = alpha * T_final * seed / (1 - alpha); the sign is applied per row
    if (!HAS_LAST && !CLIP) {
        vs = eg_select_valid_grad(vs, ov, p, G.lo);
    } else {
        bool valid = eg_pair_valid_grad(ov, p, G.lo, in_span);
        if (HAS_LAST) {
            if (valid && G.depth_bits >= dl)  // rare: at or behind the pixel's last composited Gaussian
                valid = G.depth_bits == dl && G.gid <= __ldg(gid_px);
        }
        vs = valid ? vs : 0.0f;
    }
    const float vd = vs * dx;
    r.S0 += vs;
    r.S1 += vd;
    r.S2 = fmaf(vd, dx, r.S2);
    r.ax += fabsf(vs * fmaf(G.A, dx, Bdy));
    r.ay += fabsf(vs * fmaf(G.B, dx, Cdy));
}

// Two pixels per step on the packed fp32x2 pipe (ALIGNED rows only: no per-pixel clipping, see eg_splat_fwd.cu).
// Same element-wise operations and roundings as pair_bwd; only MUFU and the validity select stay scalar.
struct RowConst2 {
    eg_f2 mx, fa, b1, c0, A, B, Bdy, Cdy;
};
// Loop-carried sums stay SCALAR: ptxas keeps scalar accumulators in place but copies 64-bit asm results around
// (14 MOVs per chunk on the ALU pipe, which is the busiest pipe of this kernel), and FADD takes |x| for free.
struct RowAcc2 {
    float S0a, S0b, S1a, S1b, S2a, S2b, ax0, ax1, ay0, ay1;
};

template <bool HAS_LAST>
__device__ __forceinline__ void pair2_bwd(const EgSplatG &G, const RowConst2 &k, const eg_f2 npx, const float w0,
                                          const float w1, const unsigned dl0, const unsigned dl1,
                                          const int *__restrict__ gid_px, RowAcc2 &r) {
    const eg_f2 dx = f2_add(k.mx, npx);                       // mean2d.x - pixel centre
    const eg_f2 p = f2_fma(f2_fma(k.fa, dx, k.b1), dx, k.c0);  // eg_pow2row
    float p0, p1;
    f2_unpack(p, p0, p1);
    const float ov0 = eg_ex2(p0), ov1 = eg_ex2(p1);
    const eg_f2 ov = f2_pack(ov0, ov1);
    float om0, om1;
    f2_unpack(f2_fma(ov, f2_dup(-1.0f), f2_dup(1.0f)), om0, om1);  // 1 - alpha
    const eg_f2 vsu = f2_mul(f2_mul(ov, f2_pack(w0, w1)), f2_pack(eg_rcp(om0), eg_rcp(om1)));
    float v0, v1;
    f2_unpack(vsu, v0, v1);
    if (!HAS_LAST) {
        v0 = eg_select_valid_grad(v0, ov0, p0, G.lo);
        v1 = eg_select_valid_grad(v1, ov1, p1, G.lo);
    } else {
        bool a0 = eg_pair_valid_grad(ov0, p0, G.lo, true), a1 = eg_pair_valid_grad(ov1, p1, G.lo, true);
        if (a0 && G.depth_bits >= dl0) a0 = G.depth_bits == dl0 && G.gid <= __ldg(gid_px);
        if (a1 && G.depth_bits >= dl1) a1 = G.depth_bits == dl1 && G.gid <= __ldg(gid_px + 1);
        v0 = a0 ? v0 : 0.0f;
        v1 = a1 ? v1 : 0.0f;
    }
    const eg_f2 vs = f2_pack(v0, v1);
    float d0, d1, e0, e1;
    f2_unpack(f2_mul(vs, dx), d0, d1);
    f2_unpack(dx, e0, e1);
    r.S0a += v0;
    r.S0b += v1;
    r.S1a += d0;
    r.S1b += d1;
    r.S2a = fmaf(d0, e0, r.S2a);
    r.S2b = fmaf(d1, e1, r.S2b);
    float t0, t1, u0, u1;
    f2_unpack(f2_mul(vs, f2_fma(k.A, dx, k.Bdy)), t0, t1);
    f2_unpack(f2_mul(vs, f2_fma(k.B, dx, k.Cdy)), u0, u1);
    r.ax0 += fabsf(t0);
    r.ax1 += fabsf(t1);
    r.ay0 += fabsf(u0);
    r.ay1 += fabsf(u1);
}

template <bool HAS_LAST, bool ALIGNED>
__device__ __forceinline__ void walk_row_bwd(const EgSplatG &G, const int y, const int W, const int tw,
                                             const float *__restrict__ wpix, const unsigned *__restrict__ last_depth,
                                             const int *__restrict__ last_gid, const int *__restrict__ tile_stop,
                                             float (&v)[8]) {
    const float dy = G.my - ((float)y + 0.5f);
    const float b1 = eg_pow2row_b1(G.fb, dy), c0 = eg_pow2row_c0(G.fc, G.lo, dy);
    int xa, xb;
    RowAcc r = {0.f, 0.f, 0.f, 0.f, 0.f};
    const float Bdy = G.B * dy, Cdy = G.C * dy;
    if (eg_row_span(G, b1, c0, xa, xb)) {
        const size_t off = (size_t)y * (size_t)W;
        const float *wrow = wpix + off;
        const unsigned *drow = HAS_LAST ? last_depth + off : nullptr;
        const int *grow = HAS_LAST ? last_gid + off : nullptr;
        const int *srow = (HAS_LAST && tile_stop != nullptr) ? tile_stop + (y >> 4) * tw : nullptr;
        int c = xa >> 2;
        const int cend = xb >> 2;
        float4 w4 = ld_f4<ALIGNED>(wrow, c, W);
        if (ALIGNED) {
            RowConst2 k2;
            k2.mx = f2_dup(G.mx); k2.fa = f2_dup(G.fa); k2.b1 = f2_dup(b1); k2.c0 = f2_dup(c0);
            k2.A = f2_dup(G.A); k2.B = f2_dup(G.B); k2.Bdy = f2_dup(Bdy); k2.Cdy = f2_dup(Cdy);
            RowAcc2 r2;
            r2.S0a = r2.S0b = r2.S1a = r2.S1b = r2.S2a = r2.S2b = 0.0f;
            r2.ax0 = r2.ax1 = r2.ay0 = r2.ay1 = 0.0f;
            // negated pixel centres of the chunk: -(x + 0.5), -(x + 1.5) | -(x + 2.5), -(x + 3.5)   (exact in fp32)
            float nb = -((float)(4 * c) + 0.5f);
            for (; c <= cend; ++c) {
                const eg_f2 npa = f2_pack(nb, nb - 1.0f), npb = f2_pack(nb - 2.0f, nb - 3.0f);
                nb -= 4.0f;
                float4 wn = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < cend) wn = ld_f4<true>(wrow, c + 1, W);  // next chunk in flight while this one is evaluated
                const int x = 4 * c;
                uint4 d4 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                if (HAS_LAST) {  // the planes are only defined in tiles where some pixel stopped
                    if (srow == nullptr || __ldg(srow + (x >> 4)) != 0) d4 = ld_u4<true>(drow, c, W);
                }
                pair2_bwd<HAS_LAST>(G, k2, npa, w4.x, w4.y, d4.x, d4.y, grow + x, r2);
                pair2_bwd<HAS_LAST>(G, k2, npb, w4.z, w4.w, d4.z, d4.w, grow + x + 2, r2);
                w4 = wn;
            }
            r.S0 = r2.S0a + r2.S0b; r.S1 = r2.S1a + r2.S1b; r.S2 = r2.S2a + r2.S2b;
            r.ax = r2.ax0 + r2.ax1; r.ay = r2.ay0 + r2.ay1;
        } else
        for (; c <= cend; ++c) {
            float4 wn = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < cend) wn = ld_f4<ALIGNED>(wrow, c + 1, W);
            const int x = 4 * c;
            uint4 d4 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            if (HAS_LAST) {
                if (srow == nullptr || __ldg(srow + (x >> 4)) != 0) d4 = ld_u4<ALIGNED>(drow, c, W);
            }
            const float px = (float)x + 0.5f;
            pair_bwd<HAS_LAST, true>(G, b1, c0, Bdy, Cdy, px, x >= xa && x <= xb, w4.x, d4.x, grow + x, r);
            pair_bwd<HAS_LAST, true>(G, b1, c0, Bdy, Cdy, px + 1.0f, x + 1 >= xa && x + 1 <= xb, w4.y, d4.y, grow + x + 1, r);
            pair_bwd<HAS_LAST, true>(G, b1, c0, Bdy, Cdy, px + 2.0f, x + 2 >= xa && x + 2 <= xb, w4.z, d4.z, grow + x + 2, r);
            pair_bwd<HAS_LAST, true>(G, b1, c0, Bdy, Cdy, px + 3.0f, x + 3 >= xa && x + 3 <= xb, w4.w, d4.w, grow + x + 3, r);
            w4 = wn;
        }
    }
    // row moments (of -v_sigma) -> (v_mean2d.x, .y, absgrad.x, .y, v_conic.a, .b, .c, sum v_sigma), accumulated
    v[0] -= fmaf(G.A, r.S1, Bdy * r.S0);
    v[1] -= fmaf(G.B, r.S1, Cdy * r.S0);
    v[2] += r.ax;
    v[3] += r.ay;
    v[4] -= 0.5f * r.S2;
    v[5] -= dy * r.S1;
    v[6] -= 0.5f * dy * dy * r.S0;
    v[7] -= r.S0;
}

// phase 3 of both kernels (lane = Gaussian): the accumulated 2D gradients -> projection VJP + exp / sigmoid VJP +
// abs-grad norm, gradients written with plain stores (each Gaussian has exactly one owner)
template <bool RAW>
__device__ __forceinline__ void finish_gaussian(const eg_config &cfg, const int g, const bool has_pairs, const float *acc,
                                                const float seed_scale, const float opac_eff, const float4 r1,
                                                const float *__restrict__ means, const float *__restrict__ quats,
                                                const float *__restrict__ scales, const float *__restrict__ opacities,
                                                const float *__restrict__ viewmat, const float *__restrict__ Kmat,
                                                float4 *__restrict__ grad2d_out, float *__restrict__ v_means,
                                                float *__restrict__ v_quats, float *__restrict__ v_scales,
                                                float *__restrict__ v_opacities, float *__restrict__ absgrad_accum) {
    float vm[3] = {0.f, 0.f, 0.f}, vs[3] = {0.f, 0.f, 0.f}, vq[4] = {0.f, 0.f, 0.f, 0.f}, vo = 0.f;
    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
    if (has_pairs) {
        const float4 a0 = make_float4(acc[0], acc[1], acc[2], acc[3]);
        const float4 a1 = make_float4(acc[4], acc[5], acc[6], acc[7]);
        const float sc = seed_scale, asc = fabsf(seed_scale);
        g0 = make_float4(a0.x * sc, a0.y * sc, a0.z * asc, a0.w * asc);
        // v_opacity' = sum vis * v_alpha = -(sum v_sigma) / opacity'
        g1 = make_float4(a1.x * sc, a1.y * sc, a1.z * sc, -(a1.w * sc) * eg_rcp(opac_eff));
        const float4 q4 = __ldg(reinterpret_cast<const float4 *>(quats) + g);
        const float mx = __ldg(means + 3 * g), my = __ldg(means + 3 * g + 1), mz = __ldg(means + 3 * g + 2);
        float s[3];
        s[0] = __ldg(scales + 3 * g); s[1] = __ldg(scales + 3 * g + 1); s[2] = __ldg(scales + 3 * g + 2);
        const float o = __ldg(opacities + g);
        const EgCam cam = eg_load_cam(viewmat, Kmat);
        if (absgrad_accum != nullptr) absgrad_accum[g] += sqrtf(g0.z * g0.z + g0.w * g0.w);
        eg_project_vjp<RAW>(cfg, cam, r1, g0, g1, mx, my, mz, q4, s, o, 0.0f, vm, vs, vq, vo);
    }
    if (grad2d_out != nullptr) {
        grad2d_out[2 * g] = g0;
        grad2d_out[2 * g + 1] = g1;
    }
    v_means[3 * g] = vm[0]; v_means[3 * g + 1] = vm[1]; v_means[3 * g + 2] = vm[2];
    v_scales[3 * g] = vs[0]; v_scales[3 * g + 1] = vs[1]; v_scales[3 * g + 2] = vs[2];
    eg_store_quat_grad(v_quats, g, vq);
    v_opacities[g] = vo;
}

template <bool RAW, bool ALIGNED>
__global__ void __launch_bounds__(SB_WARPS * 32, EG_SB_MINBLOCKS) splat_bwd_kernel(
    const eg_config cfg, const int g_begin, const int g_end, const int tw, const int th, const float *__restrict__ means, const float *__restrict__ quats,
    const float *__restrict__ scales, const float *__restrict__ opacities, const float *__restrict__ viewmat,
    const float *__restrict__ Kmat, const float4 *__restrict__ rec, const int2 *__restrict__ gint,
    const float *__restrict__ wpix, const float seed_scale, const unsigned *__restrict__ last_depth,
    const int *__restrict__ last_gid, const int *__restrict__ tile_stop, const int32_t *__restrict__ status,
    float4 *__restrict__ grad2d_out,
    float *__restrict__ v_means, float *__restrict__ v_quats, float *__restrict__ v_scales,
    float *__restrict__ v_opacities, float *__restrict__ absgrad_accum, const int opts) {
    __shared__ EgSplatG s_g[SB_WARPS][32];                   // compacted over the Gaussians that have rows
    __shared__ __align__(16) float s_acc[SB_WARPS][32][8];   // same (compact) index
    __shared__ __align__(16) float s_slot[SB_WARPS][32][8];  // per lane: the contribution of its work item

    if (status[EG_ST_OVERFLOW]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = g_begin + (blockIdx.x * SB_WARPS + warp) * 32 + lane;
    const bool live = g < g_end;
    const bool use_last = last_depth != nullptr && last_gid != nullptr && status[EG_ST_STOPPED] != 0;

    // ---------------- phase 1: lane = Gaussian ----------------
    EgSplatG G;
    float4 r1 = make_float4(0.f, 0.f, 0.f, 0.f);
    float opac_eff = 0.0f;
    int nrows = 0, radius = 0;
    if (live) {
        const int2 gi = __ldg(gint + g);
        radius = gi.x;
        const float4 r0 = __ldg(rec + 2 * g);
        r1 = __ldg(rec + 2 * g + 1);
        opac_eff = r0.z;
        nrows = eg_splat_setup(cfg, tw, th, g, r0, r1, radius, G);  // work items (row groups)
    }
    int incl = nrows;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    G.start = incl - nrows;
    const unsigned ne = __ballot_sync(0xffffffffu, nrows > 0);
    const int kc = __popc(ne & ((1u << lane) - 1u));  // compact index of this lane's Gaussian (EgOwnerIter)
    if (nrows > 0) s_g[warp][kc] = G;
#pragma unroll
    for (int k = 0; k < 8; ++k) s_acc[warp][lane][k] = 0.0f;
    const int R = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    EgOwnerIter it;

    // ---------------- phase 2: lane = (Gaussian, row) ----------------
    for (int base = 0; base < R; base += 32) {
        const int item = base + lane;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int owner = it.owner(base, lane, incl, nrows > 0);
        if (item >= R) owner = 32;
        if (item < R) {
            const EgSplatG Go = s_g[warp][owner];
            // the EG_ROWS_PER_ITEM rows of an item: consecutive ones, or (opts & 1) rows half a footprint apart -- a
            // short row near the rim of the ellipse then shares a lane with a long one near its middle
            const int t = item - Go.start;
            const int n_items = (Go.nrows + EG_ROWS_PER_ITEM - 1) / EG_ROWS_PER_ITEM;
#pragma unroll
            for (int r = 0; r < EG_ROWS_PER_ITEM; ++r) {
                const int ry = (opts & 1) ? t + r * n_items : EG_ROWS_PER_ITEM * t + r;
                if (ry >= Go.nrows) break;
                const int y = Go.ylo + ry;
                if (use_last) walk_row_bwd<true, ALIGNED>(Go, y, cfg.width, tw, wpix, last_depth, last_gid, tile_stop, v);
                else walk_row_bwd<false, ALIGNED>(Go, y, cfg.width, tw, wpix, nullptr, nullptr, nullptr, v);
            }
        }
        if (opts & 2) {
            // the items of one Gaussian sit in adjacent lanes: every lane parks its 8 values in shared memory and the first
            // lane of each run adds up the others (a handful of LDS.128 instead of a 5-step segmented shuffle reduction)
            float4 *slot = reinterpret_cast<float4 *>(&s_slot[warp][lane][0]);
            slot[0] = make_float4(v[0], v[1], v[2], v[3]);
            slot[1] = make_float4(v[4], v[5], v[6], v[7]);
            const int prev = __shfl_up_sync(0xffffffffu, owner, 1);
            const bool head = item < R && (lane == 0 || prev != owner);
            const unsigned head_m = __ballot_sync(0xffffffffu, head);
            __syncwarp();
            if (head) {
                const unsigned above = lane == 31 ? 0u : (head_m & ~((2u << lane) - 1u));
                const int end = above ? __ffs(above) - 1 : min(32, R - base);
                float4 q0 = make_float4(v[0], v[1], v[2], v[3]), q1 = make_float4(v[4], v[5], v[6], v[7]);
                for (int t = lane + 1; t < end; ++t) {
                    const float4 b0 = *reinterpret_cast<const float4 *>(&s_slot[warp][t][0]);
                    const float4 b1 = *reinterpret_cast<const float4 *>(&s_slot[warp][t][4]);
                    q0.x += b0.x; q0.y += b0.y; q0.z += b0.z; q0.w += b0.w;
                    q1.x += b1.x; q1.y += b1.y; q1.z += b1.z; q1.w += b1.w;
                }
                float4 *dst = reinterpret_cast<float4 *>(&s_acc[warp][owner][0]);
                float4 a0 = dst[0], a1 = dst[1];
                a0.x += q0.x; a0.y += q0.y; a0.z += q0.z; a0.w += q0.w;
                a1.x += q1.x; a1.y += q1.y; a1.z += q1.z; a1.w += q1.w;
                dst[0] = a0;
                dst[1] = a1;
            }
            __syncwarp();
            continue;
        }
        // segmented sum over the (contiguous) lanes that share an owner
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o2 = __shfl_down_sync(0xffffffffu, owner, d);
            const bool add = (lane + d < 32) && (o2 == owner);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float t = __shfl_down_sync(0xffffffffu, v[k], d);
                v[k] += add ? t : 0.0f;
            }
        }
        const int prev = __shfl_up_sync(0xffffffffu, owner, 1);
        if (item < R && (lane == 0 || prev != owner)) {  // one head lane per Gaussian of this batch
            float4 *dst = reinterpret_cast<float4 *>(&s_acc[warp][owner][0]);
            float4 a0 = dst[0], a1 = dst[1];
            a0.x += v[0]; a0.y += v[1]; a0.z += v[2]; a0.w += v[3];
            a1.x += v[4]; a1.y += v[5]; a1.z += v[6]; a1.w += v[7];
            dst[0] = a0;
            dst[1] = a1;
        }
        __syncwarp();
    }

    // ---------------- phase 3: lane = Gaussian ----------------
    if (!live) return;
    finish_gaussian<RAW>(cfg, g, nrows > 0, nrows > 0 ? &s_acc[warp][kc][0] : nullptr, seed_scale, opac_eff, r1, means, quats, scales,
                         opacities, viewmat, Kmat, grad2d_out, v_means, v_quats, v_scales, v_opacities, absgrad_accum);
}

// =====================================================================================================================
// Lane = Gaussian variant (W % 4 == 0): every lane walks ALL rows of its own Gaussian.  No owner look-up, no shared
// memory, no reduction: the 2D gradients stay in registers from the first pair to the projection VJP.  Row r of the
// 32 footprints is walked in lock-step, so the warp is as busy as its Gaussians are alike (same row count, same row
// lengths): ideal after a Morton sort of equally sized Gaussians, poor for a mix of sizes.
// =====================================================================================================================
template <bool RAW>
__global__ void __launch_bounds__(SB_WARPS * 32, EG_SB_MINBLOCKS) splat_bwd_gauss_kernel(
    const eg_config cfg, const int g_begin, const int g_end, const int tw, const int th, const float *__restrict__ means,
    const float *__restrict__ quats, const float *__restrict__ scales, const float *__restrict__ opacities,
    const float *__restrict__ viewmat, const float *__restrict__ Kmat, const float4 *__restrict__ rec,
    const int2 *__restrict__ gint, const float *__restrict__ wpix, const float seed_scale,
    const unsigned *__restrict__ last_depth, const int *__restrict__ last_gid, const int *__restrict__ tile_stop,
    const int32_t *__restrict__ status, float4 *__restrict__ grad2d_out, float *__restrict__ v_means,
    float *__restrict__ v_quats, float *__restrict__ v_scales, float *__restrict__ v_opacities,
    float *__restrict__ absgrad_accum) {
    if (status[EG_ST_OVERFLOW]) return;
    const int g = g_begin + blockIdx.x * (SB_WARPS * 32) + threadIdx.x;
    if (g >= g_end) return;
    const bool use_last = last_depth != nullptr && last_gid != nullptr && status[EG_ST_STOPPED] != 0;
    EgSplatG G;
    const int2 gi = __ldg(gint + g);
    const float4 r0 = __ldg(rec + 2 * g), r1 = __ldg(rec + 2 * g + 1);
    eg_splat_setup(cfg, tw, th, g, r0, r1, gi.x, G);
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < G.nrows; ++r) {
        if (use_last) walk_row_bwd<true, true>(G, G.ylo + r, cfg.width, tw, wpix, last_depth, last_gid, tile_stop, v);
        else walk_row_bwd<false, true>(G, G.ylo + r, cfg.width, tw, wpix, nullptr, nullptr, nullptr, v);
    }
    finish_gaussian<RAW>(cfg, g, G.nrows > 0, v, seed_scale, r0.z, r1, means, quats, scales, opacities, viewmat, Kmat, grad2d_out,
                         v_means, v_quats, v_scales, v_opacities, absgrad_accum);
}

// =====================================================================================================================
// Balanced variant (W % 4 == 0): lane = PART, a run of at most `plen` consecutive 4-pixel chunks of ONE Gaussian.
//
// The row-per-lane walk above keeps a lane busy for as long as its rows are long, while the other 31 lanes wait for the
// longest one (about half of the evaluated pixel slots fall outside a footprint), and pays a 40-shuffle segmented
// reduction per 32 row pairs.  Here every Gaussian's footprint is first turned into a ROW TABLE (lane = row, all lanes
// converged: the conservative span of eg_row_span -> first chunk, chunk count, running chunk total), then cut into
// parts of equal length that the warp consumes 32 at a time: a lane walks its part with a flattened, branch-light loop
// (the row switch is a table look-up), accumulates moments relative to the Gaussian (no per-row fold), and the parts of
// one Gaussian meet in shared memory once per part instead of once per row pair.  Long rows (large, elongated
// Gaussians) are split across lanes by construction.  Gaussians too tall or too wide for the table go through a
// warp-cooperative path built from the row walker above.  Same pair arithmetic, same roundings, same skip decisions.
// =====================================================================================================================
constexpr int SP_WARPS = 4;
constexpr int SP_ROWS = 512;        // row-table entries per warp window
constexpr int SP_BIG_ROWS = 128;    // taller / wider Gaussians take the cooperative path (keeps a window's chunk
constexpr int SP_BIG_COLS = 2040;   // count below 2^16 per Gaussian and every Gaussian inside one window)
#ifndef EG_SP_MINBLOCKS
#define EG_SP_MINBLOCKS 6
#endif

struct __align__(8) SpRow {
    unsigned ycf;   // y | first chunk << 16
    unsigned wcum;  // chunks of the window in front of this row
};

struct SpSums {
    float S0, Sx, Sxx, Sy, Sxy, Syy, ax, ay;  // sums of -v_sigma * {1, dx, dx^2, dy, dx dy, dy^2}, |v_mean2d| terms
};

struct SpRowState {
    eg_f2 b1, c0, Bdy, Cdy;
    float dy, dy2, nb0;
    unsigned wcur, wnext;  // window chunk index of the row's first chunk / of the next row's
    unsigned off0;         // pixel index (y * W + x) of the first pixel of the row's first chunk
    int tsx;               // HAS_LAST only: tile row base (y >> 4) * tw, and x of the first chunk via off0
    int y;
};

// one aligned 4-pixel chunk: the arithmetic of pair2_bwd twice, chunk sums folded into the Gaussian-relative moments
template <bool HAS_LAST>
__device__ __forceinline__ void chunk_bwd(const EgSplatG &G, const eg_f2 mx2, const eg_f2 fa2, const eg_f2 A2, const eg_f2 B2,
                                          const SpRowState &rs, const float nb, const float4 w4, const uint4 d4,
                                          const int *__restrict__ gid_px, SpSums &a) {
    const eg_f2 dxa = f2_add(mx2, f2_pack(nb, nb - 1.0f)), dxb = f2_add(mx2, f2_pack(nb - 2.0f, nb - 3.0f));
    float p0, p1, p2, p3;
    f2_unpack(f2_fma(f2_fma(fa2, dxa, rs.b1), dxa, rs.c0), p0, p1);  // eg_pow2row, bit-identical to the forward's
    f2_unpack(f2_fma(f2_fma(fa2, dxb, rs.b1), dxb, rs.c0), p2, p3);
    const float o0 = eg_ex2(p0), o1 = eg_ex2(p1), o2 = eg_ex2(p2), o3 = eg_ex2(p3);
    const eg_f2 ova = f2_pack(o0, o1), ovb = f2_pack(o2, o3);
    float m0, m1, m2, m3;
    f2_unpack(f2_fma(ova, f2_dup(-1.0f), f2_dup(1.0f)), m0, m1);  // 1 - alpha
    f2_unpack(f2_fma(ovb, f2_dup(-1.0f), f2_dup(1.0f)), m2, m3);
    float v0, v1, v2, v3;
    f2_unpack(f2_mul(f2_mul(ova, f2_pack(w4.x, w4.y)), f2_pack(eg_rcp(m0), eg_rcp(m1))), v0, v1);
    f2_unpack(f2_mul(f2_mul(ovb, f2_pack(w4.z, w4.w)), f2_pack(eg_rcp(m2), eg_rcp(m3))), v2, v3);
    if (!HAS_LAST) {
        v0 = eg_select_valid_grad(v0, o0, p0, G.lo);
        v1 = eg_select_valid_grad(v1, o1, p1, G.lo);
        v2 = eg_select_valid_grad(v2, o2, p2, G.lo);
        v3 = eg_select_valid_grad(v3, o3, p3, G.lo);
    } else {
        bool q0 = eg_pair_valid_grad(o0, p0, G.lo, true), q1 = eg_pair_valid_grad(o1, p1, G.lo, true);
        bool q2 = eg_pair_valid_grad(o2, p2, G.lo, true), q3 = eg_pair_valid_grad(o3, p3, G.lo, true);
        // rare: at or behind the last Gaussian the pixel composited before it hit the transmittance stop
        if (q0 && G.depth_bits >= d4.x) q0 = G.depth_bits == d4.x && G.gid <= __ldg(gid_px);
        if (q1 && G.depth_bits >= d4.y) q1 = G.depth_bits == d4.y && G.gid <= __ldg(gid_px + 1);
        if (q2 && G.depth_bits >= d4.z) q2 = G.depth_bits == d4.z && G.gid <= __ldg(gid_px + 2);
        if (q3 && G.depth_bits >= d4.w) q3 = G.depth_bits == d4.w && G.gid <= __ldg(gid_px + 3);
        v0 = q0 ? v0 : 0.0f; v1 = q1 ? v1 : 0.0f; v2 = q2 ? v2 : 0.0f; v3 = q3 ? v3 : 0.0f;
    }
    const eg_f2 vsa = f2_pack(v0, v1), vsb = f2_pack(v2, v3);
    float d0, d1, d2, d3, e0, e1, e2, e3;
    f2_unpack(f2_mul(vsa, dxa), d0, d1);
    f2_unpack(f2_mul(vsb, dxb), d2, d3);
    f2_unpack(dxa, e0, e1);
    f2_unpack(dxb, e2, e3);
    const float cs0 = (v0 + v1) + (v2 + v3);
    const float cs1 = (d0 + d1) + (d2 + d3);
    const float cs2 = fmaf(d0, e0, fmaf(d1, e1, fmaf(d2, e2, d3 * e3)));
    float t0, t1, t2, t3, u0, u1, u2, u3;
    f2_unpack(f2_mul(vsa, f2_fma(A2, dxa, rs.Bdy)), t0, t1);
    f2_unpack(f2_mul(vsb, f2_fma(A2, dxb, rs.Bdy)), t2, t3);
    f2_unpack(f2_mul(vsa, f2_fma(B2, dxa, rs.Cdy)), u0, u1);
    f2_unpack(f2_mul(vsb, f2_fma(B2, dxb, rs.Cdy)), u2, u3);
    a.S0 += cs0;
    a.Sx += cs1;
    a.Sxx += cs2;
    a.Sy = fmaf(rs.dy, cs0, a.Sy);
    a.Sxy = fmaf(rs.dy, cs1, a.Sxy);
    a.Syy = fmaf(rs.dy2, cs0, a.Syy);
    a.ax += (fabsf(t0) + fabsf(t1)) + (fabsf(t2) + fabsf(t3));
    a.ay += (fabsf(u0) + fabsf(u1)) + (fabsf(u2) + fabsf(u3));
}

// row constants from table entry r; called every chunk step (branch-free row switch: r only moves when the lane has
// finished its row, and reloading the same entry is cheaper than a divergent branch that some lane takes at almost
// every step)
template <bool HAS_LAST>
__device__ __forceinline__ void sp_load_row(const EgSplatG &G, const SpRow *__restrict__ rows, const int r, const int W,
                                            const int tw, SpRowState &rs) {
    const SpRow e = rows[r];
    rs.wcur = e.wcum;
    rs.wnext = rows[r + 1].wcum;
    const unsigned y = e.ycf & 0xffffu, x0 = (e.ycf >> 16) << 2;
    rs.off0 = y * (unsigned)W + x0;
    const float dy = G.my - ((float)y + 0.5f);
    rs.dy = dy;
    rs.dy2 = dy * dy;
    rs.b1 = f2_dup(eg_pow2row_b1(G.fb, dy));
    rs.c0 = f2_dup(eg_pow2row_c0(G.fc, G.lo, dy));
    rs.Bdy = f2_dup(G.B * dy);
    rs.Cdy = f2_dup(G.C * dy);
    rs.nb0 = -((float)x0 + 0.5f);  // negated centre of the row's first visited pixel (exact)
    if (HAS_LAST) {
        rs.y = (int)y;
        rs.tsx = (int)(y >> 4) * tw;
    }
}

// chunks [c, c + n) of the window, all of Gaussian G whose (non-empty) rows are table entries [r_lo, r_hi) -> moments
template <bool HAS_LAST>
__device__ __forceinline__ void walk_part_bwd(const EgSplatG &G, const SpRow *__restrict__ rows, int r_lo, int r_hi,
                                              unsigned c, const int n, const int n_max, const int W, const int tw,
                                              const float *__restrict__ wpix, const unsigned *__restrict__ last_depth,
                                              const int *__restrict__ last_gid, const int *__restrict__ tile_stop,
                                              SpSums &a) {
    a.S0 = a.Sx = a.Sxx = a.Sy = a.Sxy = a.Syy = a.ax = a.ay = 0.0f;
    SpRowState rs;
    int r = r_lo;
    const eg_f2 mx2 = f2_dup(G.mx), fa2 = f2_dup(G.fa), A2 = f2_dup(G.A), B2 = f2_dup(G.B);
    if (n > 0) {
        // the row that holds chunk c: the last one whose running total is <= c
        while (r_hi - r > 1) {
            const int mid = (r + r_hi) >> 1;
            if (rows[mid].wcum <= c) r = mid; else r_hi = mid;
        }
        sp_load_row<HAS_LAST>(G, rows, r, W, tw, rs);
    }
    for (int i = 0; i < n_max; ++i) {
        if (i < n) {
            const unsigned j4 = (c - rs.wcur) << 2;
            const unsigned off = rs.off0 + j4;
            const float4 w4 = __ldg(reinterpret_cast<const float4 *>(wpix + off));
            uint4 d4 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            if (HAS_LAST) {  // the planes are only defined in tiles where some pixel stopped
                const int x = (int)(off - (unsigned)rs.y * (unsigned)W);
                if (tile_stop == nullptr || __ldg(tile_stop + rs.tsx + (x >> 4)) != 0)
                    d4 = __ldg(reinterpret_cast<const uint4 *>(last_depth + off));
            }
            chunk_bwd<HAS_LAST>(G, mx2, fa2, A2, B2, rs, rs.nb0 - (float)j4, w4, d4, HAS_LAST ? last_gid + off : nullptr, a);
            ++c;
            r += (c == rs.wnext) ? 1 : 0;   // no empty rows in the table: the next entry is the next row with chunks
            sp_load_row<HAS_LAST>(G, rows, r, W, tw, rs);  // (one entry past the part's last row is still inside the table)
        }
    }
}

template <bool RAW>
__global__ void __launch_bounds__(SP_WARPS * 32, EG_SP_MINBLOCKS) splat_bwd_parts_kernel(
    const eg_config cfg, const int g_begin, const int g_end, const int tw, const int th, const float *__restrict__ means,
    const float *__restrict__ quats, const float *__restrict__ scales, const float *__restrict__ opacities,
    const float *__restrict__ viewmat, const float *__restrict__ Kmat, const float4 *__restrict__ rec,
    const int2 *__restrict__ gint, const float *__restrict__ wpix, const float seed_scale,
    const unsigned *__restrict__ last_depth, const int *__restrict__ last_gid, const int *__restrict__ tile_stop,
    const int32_t *__restrict__ status, float4 *__restrict__ grad2d_out, float *__restrict__ v_means,
    float *__restrict__ v_quats, float *__restrict__ v_scales, float *__restrict__ v_opacities,
    float *__restrict__ absgrad_accum, const int plen_min) {
    __shared__ EgSplatG s_g[SP_WARPS][32];                  // compacted: table Gaussians first, then the big ones
    __shared__ int4 s_gp[SP_WARPS][32];                     // per table Gaussian: first chunk, end chunk, part length, first part
    __shared__ int s_ts[SP_WARPS][33];                      // per window Gaussian: its first entry in the row table
    __shared__ SpRow s_rows[SP_WARPS][SP_ROWS + 2];  // + sentinel (+ one entry a finished part may still look at)
    __shared__ __align__(16) float s_acc[SP_WARPS][32][8];  // per (compact) Gaussian: the 8 2D gradients
    __shared__ __align__(16) float s_slot[SP_WARPS][32][8]; // per lane: its part's contribution
    __shared__ unsigned char s_rank[SP_WARPS][32];

    if (status[EG_ST_OVERFLOW]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int g = g_begin + (blockIdx.x * SP_WARPS + warp) * 32 + lane;
    const bool live = g < g_end;
    const bool use_last = last_depth != nullptr && last_gid != nullptr && status[EG_ST_STOPPED] != 0;
    const int W = cfg.width;
    SpRow *rows = s_rows[warp];

    // ---------------- phase 1: lane = Gaussian ----------------
    EgSplatG G;
    G.nrows = 0;
    float4 r1 = make_float4(0.f, 0.f, 0.f, 0.f);
    float opac_eff = 0.0f;
    if (live) {
        const int2 gi = __ldg(gint + g);
        const float4 r0 = __ldg(rec + 2 * g);
        r1 = __ldg(rec + 2 * g + 1);
        opac_eff = r0.z;
        eg_splat_setup(cfg, tw, th, g, r0, r1, gi.x, G);
    }
    const bool any_rows = G.nrows > 0;
    const bool big = any_rows && (G.nrows > SP_BIG_ROWS || G.X1() - G.X0() > SP_BIG_COLS);
    const int nr = (any_rows && !big) ? G.nrows : 0;
    int rincl = nr;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, rincl, d);
        if (lane >= d) rincl += t;
    }
    G.start = rincl - nr;  // first row in the warp's dense row list
    const unsigned tab_m = __ballot_sync(0xffffffffu, nr > 0), big_m = __ballot_sync(0xffffffffu, big);
    const int K = __popc(tab_m);
    const int kc = nr > 0 ? __popc(tab_m & lt) : (big ? K + __popc(big_m & lt) : -1);  // compact index of MY Gaussian
    if (kc >= 0) s_g[warp][kc] = G;
#pragma unroll
    for (int k = 0; k < 8; ++k) s_acc[warp][lane][k] = 0.0f;
    __syncwarp();
    // lane = compact table Gaussian `lane` (< K): its rows in the dense list
    const int cst = lane < K ? s_g[warp][lane].start : 0;
    const int cend = lane < K ? cst + s_g[warp][lane].nrows : 0;

    // ---------------- phase 2: windows of table Gaussians whose rows fit the table ----------------
    for (int k0 = 0; k0 < K;) {
        const int rs0 = __shfl_sync(0xffffffffu, cst, k0);
        const unsigned fit_m = __ballot_sync(0xffffffffu, lane >= k0 && lane < K && cend - rs0 <= SP_ROWS);
        const int k1 = k0 + __popc(fit_m);  // ends are increasing: the fitting Gaussians are k0 .. k1-1 (at least one)
        const bool in_win = lane >= k0 && lane < k1;
        const int Rw = __shfl_sync(0xffffffffu, cend, k1 - 1) - rs0;
        // ---- row table: lane = row; rows whose span is empty are dropped ----
        int Rt = 0;
        unsigned Cw = 0;
        {
            EgOwnerIter rit;
            for (int base = 0; base < Rw; base += 32) {
                const int item = base + lane;
                const int kk = k0 + rit.owner(base, lane, cend - rs0, in_win);
                int nch = 0, first_of = -1;
                unsigned ycf = 0;
                if (item < Rw) {
                    const EgSplatG Go = s_g[warp][kk];
                    const int ry = item + rs0 - Go.start;
                    if (ry == 0) first_of = kk - k0;
                    const int y = Go.ylo + ry;
                    const float dy = Go.my - ((float)y + 0.5f);
                    const float b1 = eg_pow2row_b1(Go.fb, dy), c0 = eg_pow2row_c0(Go.fc, Go.lo, dy);
                    int xa, xb;
                    if (eg_row_span(Go, b1, c0, xa, xb)) {
                        const int cf = xa >> 2;
                        nch = (xb >> 2) - cf + 1;
                        ycf = (unsigned)y | ((unsigned)cf << 16);
                    }
                }
                const unsigned nz = __ballot_sync(0xffffffffu, nch > 0);
                const int pos = Rt + __popc(nz & lt);
                int cincl = nch;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, cincl, d);
                    if (lane >= d) cincl += t;
                }
                if (nch > 0) {
                    SpRow e;
                    e.ycf = ycf;
                    e.wcum = Cw + (unsigned)(cincl - nch);
                    rows[pos] = e;
                }
                if (first_of >= 0) s_ts[warp][first_of] = pos;  // first table entry at or after the Gaussian's first row
                Rt += __popc(nz);
                Cw += (unsigned)__shfl_sync(0xffffffffu, cincl, 31);
            }
            if (lane == 0) {
                SpRow e;
                e.ycf = 0;
                e.wcum = Cw;  // sentinel: the window's chunk total
                rows[Rt] = e;
                rows[Rt + 1] = e;
                s_ts[warp][k1 - k0] = Rt;
            }
        }
        __syncwarp();
        // ---- parts: lane = table Gaussian of the window ----
        const int plen_t = min(32, max(plen_min, ((int)Cw + 127) >> 7));
        int np = 0, ts = 0, te = 0;
        if (in_win) {
            ts = s_ts[warp][lane - k0];
            te = s_ts[warp][lane - k0 + 1];
            const int cb = (int)rows[ts].wcum, ce = (int)rows[te].wcum;
            const int C = ce - cb;
            np = (C + plen_t - 1) / plen_t;
            const int plen = np > 0 ? (C + np - 1) / np : 0;  // equal parts: no short tail part
            s_gp[warp][lane] = make_int4(cb, ce, plen, 0);
        }
        int pincl = np;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, pincl, d);
            if (lane >= d) pincl += t;
        }
        const int NP = __shfl_sync(0xffffffffu, pincl, 31);
        const unsigned has_m = __ballot_sync(0xffffffffu, np > 0);
        if (np > 0) {
            s_gp[warp][lane].w = pincl - np;
            s_rank[warp][__popc(has_m & lt)] = (unsigned char)lane;
            s_ts[warp][lane - k0] = ts | (te << 16);   // (the plain starts are consumed: keep the Gaussian's entry range)
        }
        __syncwarp();
        // ---- lane = part ----
        EgOwnerIter pit;
        for (int pb = 0; pb < NP; pb += 32) {
            const int pi = pb + lane;
            const int rk = pit.owner(pb, lane, pincl, np > 0);
            const bool valid = pi < NP;
            int own = -1, n = 0, r_lo = 0, r_hi = 0;
            unsigned c = 0;
            EgSplatG Go;
            if (valid) {
                own = (int)s_rank[warp][rk];
                Go = s_g[warp][own];
                const int4 gp = s_gp[warp][own];
                const int j = pi - gp.w;
                c = (unsigned)min(gp.y, gp.x + j * gp.z);
                n = min(gp.y, (int)c + gp.z) - (int)c;
                const int tr = s_ts[warp][own - k0];
                r_lo = tr & 0xffff;
                r_hi = tr >> 16;
            }
            const int n_max = __reduce_max_sync(0xffffffffu, n);
            SpSums a;
            if (use_last) walk_part_bwd<true>(Go, rows, r_lo, r_hi, c, n, n_max, W, tw, wpix, last_depth, last_gid, tile_stop, a);
            else walk_part_bwd<false>(Go, rows, r_lo, r_hi, c, n, n_max, W, tw, wpix, nullptr, nullptr, nullptr, a);
            // moments -> (v_mean2d.x, .y, absgrad.x, .y, v_conic.a, .b, .c, sum v_sigma); the sums hold -v_sigma
            float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
            if (n > 0) {
                q0 = make_float4(-fmaf(Go.A, a.Sx, Go.B * a.Sy), -fmaf(Go.B, a.Sx, Go.C * a.Sy), a.ax, a.ay);
                q1 = make_float4(-0.5f * a.Sxx, -a.Sxy, -0.5f * a.Syy, -a.S0);
            }
            float4 *slot = reinterpret_cast<float4 *>(&s_slot[warp][lane][0]);
            slot[0] = q0;
            slot[1] = q1;
            // the parts of one Gaussian sit in adjacent lanes: the first of them adds the others' slots to its own
            const int prev = __shfl_up_sync(0xffffffffu, own, 1);
            const bool head = valid && (lane == 0 || prev != own);
            const unsigned head_m = __ballot_sync(0xffffffffu, head);
            __syncwarp();
            if (head) {
                const unsigned above = lane == 31 ? 0u : (head_m & ~((2u << lane) - 1u));
                const int end = above ? __ffs(above) - 1 : min(32, NP - pb);
                for (int t = lane + 1; t < end; ++t) {
                    const float4 b0 = *reinterpret_cast<const float4 *>(&s_slot[warp][t][0]);
                    const float4 b1 = *reinterpret_cast<const float4 *>(&s_slot[warp][t][4]);
                    q0.x += b0.x; q0.y += b0.y; q0.z += b0.z; q0.w += b0.w;
                    q1.x += b1.x; q1.y += b1.y; q1.z += b1.z; q1.w += b1.w;
                }
                float4 *dst = reinterpret_cast<float4 *>(&s_acc[warp][own][0]);
                float4 a0 = dst[0], a1 = dst[1];
                a0.x += q0.x; a0.y += q0.y; a0.z += q0.z; a0.w += q0.w;
                a1.x += q1.x; a1.y += q1.y; a1.z += q1.z; a1.w += q1.w;
                dst[0] = a0;
                dst[1] = a1;
            }
            __syncwarp();
        }
        k0 = k1;
    }

    // ---------------- Gaussians too large for the table: the whole warp walks one at a time, lane = row ----------------
    for (unsigned bm = big_m; bm; bm &= bm - 1) {
        const int idx = K + __popc(big_m & ((1u << (__ffs(bm) - 1)) - 1u));
        const EgSplatG Gb = s_g[warp][idx];
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int rb = lane; rb < Gb.nrows; rb += 32) {
            if (use_last) walk_row_bwd<true, true>(Gb, Gb.ylo + rb, W, tw, wpix, last_depth, last_gid, tile_stop, v);
            else walk_row_bwd<false, true>(Gb, Gb.ylo + rb, W, tw, wpix, nullptr, nullptr, nullptr, v);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], d);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) s_acc[warp][idx][k] = v[k];
        }
    }
    __syncwarp();

    // ---------------- phase 3: lane = Gaussian ----------------
    if (!live) return;
    finish_gaussian<RAW>(cfg, g, kc >= 0, kc >= 0 ? &s_acc[warp][kc][0] : nullptr, seed_scale, opac_eff, r1, means, quats, scales,
                         opacities, viewmat, Kmat, grad2d_out, v_means, v_quats, v_scales, v_opacities, absgrad_accum);
}

// per-pixel backward seed of the gsplat-shaped autograd path:  w_p = (sum_ch v_render[p,ch] + v_alpha[p]) * (1 - alpha[p])
__global__ void __launch_bounds__(256) seed_kernel(long long P, const float *__restrict__ alpha,
                                                   const float *__restrict__ v_render, int ch,
                                                   const float *__restrict__ v_alpha, float *__restrict__ wpix) {
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= P) return;
    float gsum = 0.0f;
    if (v_render != nullptr)
        for (int c = 0; c < ch; ++c) gsum += __ldg(v_render + p * ch + c);
    if (v_alpha != nullptr) gsum += __ldg(v_alpha + p);
    wpix[p] = gsum * (1.0f - __ldg(alpha + p));
}

}  // namespace

extern "C" int eg_splat_bwd(const eg_config *cfg, const float *means, const float *quats, const float *scales,
                            const float *opacities, const float *viewmat, const float *K, const float *rec,
                            const int32_t *gint, const float *wpix, float seed_scale, const uint32_t *last_depth,
                            const int32_t *last_gid, const int32_t *tile_stop, const int32_t *status,
                            int g_begin, int g_end, float *grad2d_out, float *v_means,
                            float *v_quats, float *v_scales, float *v_opacities, float *absgrad_accum, void *stream) {
    if (cfg == nullptr || cfg->tile_size != EG_TILE) {
        eg_set_error("eg_splat_bwd: tile_size must be %d", EG_TILE);
        return 1;
    }
    if (wpix == nullptr || rec == nullptr || gint == nullptr || status == nullptr) {
        eg_set_error("eg_splat_bwd: rec, gint, wpix and status are required");
        return 1;
    }
    if ((last_depth == nullptr) != (last_gid == nullptr)) {
        eg_set_error("eg_splat_bwd: last_depth and last_gid go together");
        return 1;
    }
    if (g_end < 0 || g_end > cfg->n) g_end = cfg->n;
    if (g_begin < 0) g_begin = 0;
    if (g_begin >= g_end) return 0;
    if (cfg->width >= 65536 || cfg->height >= 65536) {
        eg_set_error("eg_splat_bwd: image sides must be < 65536");
        return 1;
    }
    int tw, th;
    eg_tile_grid(cfg->width, cfg->height, cfg->tile_size, &tw, &th);
    const int block = SB_WARPS * 32, grid = (g_end - g_begin + block - 1) / block;
    cudaStream_t s = (cudaStream_t)stream;
    const bool aligned = (cfg->width % 4 == 0) && (((uintptr_t)wpix & 15) == 0) &&
                         (last_depth == nullptr || ((uintptr_t)last_depth & 15) == 0);
#define EG_SB_LAUNCH(RAWP, AL)                                                                                       \
    splat_bwd_kernel<RAWP, AL><<<grid, block, 0, s>>>(*cfg, g_begin, g_end, tw, th, means, quats, scales, opacities, viewmat, K,    \
                                                      (const float4 *)rec, (const int2 *)gint, wpix, seed_scale,    \
                                                      last_depth, last_gid, tile_stop, status,                      \
                                                      (float4 *)grad2d_out, v_means,                                \
                                                      v_quats, v_scales, v_opacities, absgrad_accum, opts)
#define EG_SP_LAUNCH(RAWP)                                                                                           \
    splat_bwd_parts_kernel<RAWP><<<grid, block, 0, s>>>(*cfg, g_begin, g_end, tw, th, means, quats, scales, opacities, viewmat, K,  \
                                                        (const float4 *)rec, (const int2 *)gint, wpix, seed_scale,  \
                                                        last_depth, last_gid, tile_stop, status,                    \
                                                        (float4 *)grad2d_out, v_means,                              \
                                                        v_quats, v_scales, v_opacities, absgrad_accum, plen_min)
    // part length of small footprints (chunks); EG_SP_PLEN overrides for tuning
    static const int plen_min = getenv("EG_SP_PLEN") != nullptr ? max(1, min(32, atoi(getenv("EG_SP_PLEN")))) : 8;
    // rows kernel options (bit 0: rows of an item half a footprint apart, bit 1: shared-memory slot reduction)
    static const int opts = getenv("EG_BWD_OPTS") != nullptr ? atoi(getenv("EG_BWD_OPTS")) : 0;
    // EG_BWD_MODE selects the enumeration on aligned images (A/B measurements): rows (default) | parts | gauss
    static const char *mode_env = getenv("EG_BWD_MODE");
    static const int mode = mode_env == nullptr ? 0 : (mode_env[0] == 'p' ? 1 : (mode_env[0] == 'g' ? 2 : 0));
    if (aligned && mode == 2) {
        if (cfg->raw_params)
            splat_bwd_gauss_kernel<true><<<grid, block, 0, s>>>(*cfg, g_begin, g_end, tw, th, means, quats, scales, opacities, viewmat, K,
                (const float4 *)rec, (const int2 *)gint, wpix, seed_scale, last_depth, last_gid, tile_stop, status,
                (float4 *)grad2d_out, v_means, v_quats, v_scales, v_opacities, absgrad_accum);
        else
            splat_bwd_gauss_kernel<false><<<grid, block, 0, s>>>(*cfg, g_begin, g_end, tw, th, means, quats, scales, opacities, viewmat, K,
                (const float4 *)rec, (const int2 *)gint, wpix, seed_scale, last_depth, last_gid, tile_stop, status,
                (float4 *)grad2d_out, v_means, v_quats, v_scales, v_opacities, absgrad_accum);
    } else if (aligned && mode == 1) {
        static_assert(SP_WARPS == SB_WARPS, "both kernels own 128 Gaussians per CTA");
        if (cfg->raw_params) EG_SP_LAUNCH(true); else EG_SP_LAUNCH(false);
    } else if (cfg->raw_params) {
        if (aligned) EG_SB_LAUNCH(true, true); else EG_SB_LAUNCH(true, false);
    } else {
        if (aligned) EG_SB_LAUNCH(false, true); else EG_SB_LAUNCH(false, false);
    }
#undef EG_SP_LAUNCH
#undef EG_SB_LAUNCH
    return eg_check_launch("eg_splat_bwd");
}

extern "C" int eg_make_seed(int64_t n_pixels, const float *alpha, const float *v_render, int v_render_channels,
                            const float *v_alpha, float *wpix, void *stream) {
    if (alpha == nullptr || wpix == nullptr) {
        eg_set_error("eg_make_seed: alpha and wpix are required");
        return 1;
    }
    if (n_pixels <= 0) return 0;
    seed_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (long long)n_pixels, alpha, v_render, v_render_channels, v_alpha, wpix);
    return eg_check_launch("eg_make_seed");
}
