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

// Raw sums over every pair a lane visited, of vs = -v_sigma and its moments (dx = mean.x - pixel.x, dy likewise):
//   T0 = sum vs, T1 = sum vs dx, T2 = sum vs dx^2, U0 = sum vs dy, U1 = sum vs dx dy, W0 = sum vs dy^2, and the two abs-sums.
// They are linear in the pairs, so they are summed over rows (and lanes) as they are and only turned into the eight 2D
// gradients once per Gaussian (finish_gaussian).  The dx^2 moment and the abs-sums of aligned rows are carried in the
// pair loop's own registers from row to row (S2p, ax0 ..) and folded in at the end (gacc_store).
struct GaussAcc {
    float T0, T1, T2, U0, U1, W0, ax, ay;
    eg_f2 S2p;
    float ax0, ax1, ay0, ay1;
};
__device__ __forceinline__ void gacc_zero(GaussAcc &a) {
    a.T0 = a.T1 = a.T2 = a.U0 = a.U1 = a.W0 = a.ax = a.ay = 0.0f;
    a.S2p = f2_dup(0.0f);
    a.ax0 = a.ax1 = a.ay0 = a.ay1 = 0.0f;
}
__device__ __forceinline__ void gacc_store(const GaussAcc &a, float (&v)[8]) {
    float s0, s1;
    f2_unpack(a.S2p, s0, s1);
    v[0] = a.T0; v[1] = a.T1; v[2] = a.T2 + (s0 + s1); v[3] = a.U0;
    v[4] = a.U1; v[5] = a.W0; v[6] = a.ax + (a.ax0 + a.ax1); v[7] = a.ay + (a.ay0 + a.ay1);
}

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
// Loop-carried sums stay SCALAR (FADD takes |x| for free).  Measured equal (117.4 - 117.6 us, round 2): packed in-place
// accumulators (FADD2 / FFMA2, 11 fewer instructions per chunk) and a loop unrolled by two that never copies the
// prefetched seed quad -- the kernel is not bound by its instruction count alone.
struct RowAcc2 {
    float S0a, S0b, S1a, S1b, S2a, S2b;
    float ax0, ax1, ay0, ay1;
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
    const eg_f2 d = f2_mul(vs, dx);
    float d0, d1, e0, e1;
    f2_unpack(d, d0, d1);
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
                                             GaussAcc &acc) {
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
            r2.S0a = r2.S0b = r2.S1a = r2.S1b = 0.0f;
            f2_unpack(acc.S2p, r2.S2a, r2.S2b);
            r2.ax0 = acc.ax0; r2.ax1 = acc.ax1; r2.ay0 = acc.ay0; r2.ay1 = acc.ay1;
            // negated pixel centres of the chunk: -(x + 0.5), -(x + 1.5) | -(x + 2.5), -(x + 3.5)   (exact in fp32)
            const float nb = -((float)(4 * c) + 0.5f);
            eg_f2 npa = f2_pack(nb, nb - 1.0f), npb = f2_pack(nb - 2.0f, nb - 3.0f);
            const eg_f2 m4 = f2_dup(-4.0f);
            // one chunk: 4 pairs on the packed pipe; the seed of the NEXT chunk is in flight meanwhile, into the other of
            // two register quads (the loop is unrolled by two so that no quad is ever copied)
            auto chunk = [&](const float4 &w, const int cc) {
                const int x = 4 * cc;
                uint4 d4 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                if (HAS_LAST) {  // the planes are only defined in tiles where some pixel stopped
                    if (srow == nullptr || __ldg(srow + (x >> 4)) != 0) d4 = ld_u4<true>(drow, cc, W);
                }
                pair2_bwd<HAS_LAST>(G, k2, npa, w.x, w.y, d4.x, d4.y, grow + x, r2);
                pair2_bwd<HAS_LAST>(G, k2, npb, w.z, w.w, d4.z, d4.w, grow + x + 2, r2);
                npa = f2_add(npa, m4);
                npb = f2_add(npb, m4);
            };
            for (; c <= cend; ++c) {
                float4 wn = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < cend) wn = ld_f4<true>(wrow, c + 1, W);  // next chunk in flight while this one is evaluated
                chunk(w4, c);
                w4 = wn;
            }
            r.S0 = r2.S0a + r2.S0b; r.S1 = r2.S1a + r2.S1b;
            acc.S2p = f2_pack(r2.S2a, r2.S2b);
            acc.ax0 = r2.ax0; acc.ax1 = r2.ax1; acc.ay0 = r2.ay0; acc.ay1 = r2.ay1;
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
    // the row's moments join the Gaussian's (r.S2 / r.ax / r.ay are zero for aligned rows: carried in acc instead)
    const float d0 = dy * r.S0;
    acc.T0 += r.S0;
    acc.T1 += r.S1;
    acc.T2 += r.S2;
    acc.U0 += d0;
    acc.U1 = fmaf(dy, r.S1, acc.U1);
    acc.W0 = fmaf(dy, d0, acc.W0);
    acc.ax += r.ax;
    acc.ay += r.ay;
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
                                                float *__restrict__ v_opacities, float *__restrict__ absgrad_accum,
                                                float *stage = nullptr) {
    float vm[3] = {0.f, 0.f, 0.f}, vs[3] = {0.f, 0.f, 0.f}, vq[4] = {0.f, 0.f, 0.f, 0.f}, vo = 0.f;
    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
    if (has_pairs) {
        // raw moments (GaussAcc: T0, T1, T2, U0 | U1, W0, ax, ay; of vs = -v_sigma) -> 2D gradients:
        //   v_mean2d = -(A T1 + B U0, B T1 + C U0), v_conic = -(T2 / 2, U1, W0 / 2), sum v_sigma = -T0
        const float4 m0 = *reinterpret_cast<const float4 *>(acc), m1 = *reinterpret_cast<const float4 *>(acc + 4);
        const float A = r1.x, B = r1.y, C = r1.z;
        const float4 a0 = make_float4(-fmaf(A, m0.y, B * m0.w), -fmaf(B, m0.y, C * m0.w), m1.z, m1.w);
        const float4 a1 = make_float4(-0.5f * m0.z, -m1.x, -0.5f * m1.y, -m0.x);
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
    if (stage != nullptr) {
        // Whole warp, 32 consecutive Gaussians (push form: the stores cross NVLink): transpose the [32,3] means / scales
        // gradients through `stage` (192 floats of this warp's shared memory) and write them as 24 + 24 full 16-byte
        // vectors, 384 contiguous bytes each, instead of 6 strided 4-byte stores per lane.
        const int lane = threadIdx.x & 31;
        stage[3 * lane] = vm[0]; stage[3 * lane + 1] = vm[1]; stage[3 * lane + 2] = vm[2];
        stage[96 + 3 * lane] = vs[0]; stage[96 + 3 * lane + 1] = vs[1]; stage[96 + 3 * lane + 2] = vs[2];
        __syncwarp();
        if (lane < 24) {
            const int g0 = g - lane;   // the warp's first Gaussian (a multiple of 32: the vectors are 16-byte aligned)
            reinterpret_cast<float4 *>(v_means + 3 * g0)[lane] = reinterpret_cast<const float4 *>(stage)[lane];
            reinterpret_cast<float4 *>(v_scales + 3 * g0)[lane] = reinterpret_cast<const float4 *>(stage + 96)[lane];
        }
    } else {
        v_means[3 * g] = vm[0]; v_means[3 * g + 1] = vm[1]; v_means[3 * g + 2] = vm[2];
        v_scales[3 * g] = vs[0]; v_scales[3 * g + 1] = vs[1]; v_scales[3 * g + 2] = vs[2];
    }
    eg_store_quat_grad(v_quats, g, vq);
    v_opacities[g] = vo;
}

template <bool RAW, bool ALIGNED, bool PUSH>
__global__ void __launch_bounds__(SB_WARPS * 32, EG_SB_MINBLOCKS) splat_bwd_kernel(
    const eg_config cfg, const int g_begin, const int g_end, const int tw, const int th, const float *__restrict__ means, const float *__restrict__ quats,
    const float *__restrict__ scales, const float *__restrict__ opacities, const float *__restrict__ viewmat,
    const float *__restrict__ Kmat, const float4 *__restrict__ rec, const int2 *__restrict__ gint,
    const float *__restrict__ wpix, const float seed_scale, const unsigned *__restrict__ last_depth,
    const int *__restrict__ last_gid, const int *__restrict__ tile_stop, const int32_t *__restrict__ status,
    float4 *__restrict__ grad2d_out,
    float *__restrict__ v_means, float *__restrict__ v_quats, float *__restrict__ v_scales,
    float *__restrict__ v_opacities, float *__restrict__ absgrad_accum, const int opts, const eg_push_target push) {
    __shared__ EgSplatG s_g[SB_WARPS][32];                   // compacted over the Gaussians that have rows
    __shared__ __align__(16) float s_acc[SB_WARPS][32][8];   // same (compact) index

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
    int R = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    EgOwnerIter it;

    // ---------------- phase 2 (a): lane = Gaussian, when the warp's footprints are alike ----------------
    // Every lane walks ALL rows of its own Gaussian: no owner look-up, no reduction, the 2D gradients never leave the
    // registers.  Row r of the 32 footprints is walked in lock-step, so this only pays when the Gaussians of the warp
    // have about the same number of rows and the same row lengths -- equally sized Gaussians in Morton order
    // (EdgeGaussianSplatting.sort_gaussians_morton) -- which the warp decides from a cheap estimate of its work:
    // rows x chunks of the widest row, largest over mean <= 1.5.
    if (ALIGNED && (opts & 1) == 0) {
        int work = 0;
        if (nrows > 0) {
            const float hw = sqrtf(fmaxf(0.0f, (G.lo - EG_L2AMIN_CONS) * eg_rcp(fmaxf(-G.fa, 1e-12f))));  // half width of the widest row
            work = G.nrows * ((int)fminf(hw * 0.5f, 16000.0f) + 2);
        }
        const int wmax = __reduce_max_sync(0xffffffffu, work), wsum = __reduce_add_sync(0xffffffffu, min(work, 1 << 20));
        if (wmax <= (1 << 20) && 2ll * wmax * __popc(ne) <= 3ll * wsum) {
            GaussAcc ga;
            gacc_zero(ga);
            const int my_rows = nrows > 0 ? G.nrows : 0;  // (G is only defined for lanes that own a visible Gaussian)
            for (int r = 0; r < my_rows; ++r) {
                if (use_last) walk_row_bwd<true, ALIGNED>(G, G.ylo + r, cfg.width, tw, wpix, last_depth, last_gid, tile_stop, ga);
                else walk_row_bwd<false, ALIGNED>(G, G.ylo + r, cfg.width, tw, wpix, nullptr, nullptr, nullptr, ga);
            }
            if (nrows > 0) {
                float v[8];
                gacc_store(ga, v);
#pragma unroll
                for (int k = 0; k < 8; ++k) s_acc[warp][kc][k] = v[k];
            }
            R = 0;  // nothing left for the row-item walk
        }
    }
    __syncwarp();

    // ---------------- phase 2 (b): lane = (Gaussian, row) ----------------
    for (int base = 0; base < R; base += 32) {
        const int item = base + lane;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int owner = it.owner(base, lane, incl, nrows > 0);
        if (item >= R) owner = 32;
        if (item < R) {
            const EgSplatG Go = s_g[warp][owner];
            const int r0 = EG_ROWS_PER_ITEM * (item - Go.start);
            GaussAcc ga;
            gacc_zero(ga);
#pragma unroll
            for (int r = 0; r < EG_ROWS_PER_ITEM; ++r) {
                if (r0 + r >= Go.nrows) break;
                const int y = Go.ylo + r0 + r;
                if (use_last) walk_row_bwd<true, ALIGNED>(Go, y, cfg.width, tw, wpix, last_depth, last_gid, tile_stop, ga);
                else walk_row_bwd<false, ALIGNED>(Go, y, cfg.width, tw, wpix, nullptr, nullptr, nullptr, ga);
            }
            gacc_store(ga, v);
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
    // (PUSH) vectorised stores need the whole warp on 32 consecutive Gaussians of ONE owner (per is a multiple of 128)
    const bool vec_out = PUSH && __all_sync(0xffffffffu, live) && ((g - lane) & 31) == 0;
    if (!live) return;
    EgGradOut out;   // the caller's tensors, or (PUSH) this rank's slot in the staging area of the Gaussian's owner
    if (PUSH) out = eg_grad_out(push, g, v_means, v_scales, v_quats, v_opacities);
    else { out.means = v_means; out.scales = v_scales; out.quats = v_quats; out.opac = v_opacities; }
    finish_gaussian<RAW>(cfg, g, nrows > 0, nrows > 0 ? &s_acc[warp][kc][0] : nullptr, seed_scale, opac_eff, r1, means, quats, scales,
                         opacities, viewmat, Kmat, grad2d_out, out.means, out.quats, out.scales, out.opac, absgrad_accum,
                         vec_out ? reinterpret_cast<float *>(&s_g[warp][0]) : nullptr);  // s_g is dead after phase 2
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

static int splat_bwd_launch(const eg_config *cfg, const float *means, const float *quats, const float *scales,
                            const float *opacities, const float *viewmat, const float *K, const float *rec,
                            const int32_t *gint, const float *wpix, float seed_scale, const uint32_t *last_depth,
                            const int32_t *last_gid, const int32_t *tile_stop, const int32_t *status,
                            int g_begin, int g_end, float *grad2d_out, float *v_means,
                            float *v_quats, float *v_scales, float *v_opacities, float *absgrad_accum,
                            const eg_push_target &push, void *stream) {
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
#define EG_SB_LAUNCH(RAWP, AL, PU)                                                                                     \
    splat_bwd_kernel<RAWP, AL, PU><<<grid, block, 0, s>>>(*cfg, g_begin, g_end, tw, th, means, quats, scales, opacities, viewmat, K,    \
                                                      (const float4 *)rec, (const int2 *)gint, wpix, seed_scale,    \
                                                      last_depth, last_gid, tile_stop, status,                      \
                                                      (float4 *)grad2d_out, v_means,                                \
                                                      v_quats, v_scales, v_opacities, absgrad_accum, opts, push)
    // EG_BWD_OPTS=1 switches the lane = Gaussian path of uniform warps off (A/B measurements)
    static const int opts = getenv("EG_BWD_OPTS") != nullptr ? atoi(getenv("EG_BWD_OPTS")) : 0;
    if (push.world > 1) {
        if (cfg->raw_params) {
            if (aligned) EG_SB_LAUNCH(true, true, true); else EG_SB_LAUNCH(true, false, true);
        } else {
            if (aligned) EG_SB_LAUNCH(false, true, true); else EG_SB_LAUNCH(false, false, true);
        }
    } else if (cfg->raw_params) {
        if (aligned) EG_SB_LAUNCH(true, true, false); else EG_SB_LAUNCH(true, false, false);
    } else {
        if (aligned) EG_SB_LAUNCH(false, true, false); else EG_SB_LAUNCH(false, false, false);
    }
#undef EG_SB_LAUNCH
    return eg_check_launch("eg_splat_bwd");
}

extern "C" int eg_splat_bwd(const eg_config *cfg, const float *means, const float *quats, const float *scales,
                            const float *opacities, const float *viewmat, const float *K, const float *rec,
                            const int32_t *gint, const float *wpix, float seed_scale, const uint32_t *last_depth,
                            const int32_t *last_gid, const int32_t *tile_stop, const int32_t *status,
                            int g_begin, int g_end, float *grad2d_out, float *v_means,
                            float *v_quats, float *v_scales, float *v_opacities, float *absgrad_accum, void *stream) {
    eg_push_target none = {};
    return splat_bwd_launch(cfg, means, quats, scales, opacities, viewmat, K, rec, gint, wpix, seed_scale, last_depth, last_gid,
                            tile_stop, status, g_begin, g_end, grad2d_out, v_means, v_quats, v_scales, v_opacities,
                            absgrad_accum, none, stream);
}

// the same backward with its gradient stores redirected into the owners' staging slots (eg_push_target)
extern "C" int eg_splat_bwd_push(const eg_config *cfg, const float *means, const float *quats, const float *scales,
                                 const float *opacities, const float *viewmat, const float *K, const float *rec,
                                 const int32_t *gint, const float *wpix, float seed_scale, const uint32_t *last_depth,
                                 const int32_t *last_gid, const int32_t *tile_stop, const int32_t *status, int g_begin,
                                 int g_end, const eg_push_target *push, float *absgrad_accum, void *stream) {
    if (cfg == nullptr || !eg_push_target_ok("eg_splat_bwd_push", push, cfg->n)) return 1;
    if (push->world < 2) {
        eg_set_error("eg_splat_bwd_push: needs world >= 2 (a single rank calls eg_splat_bwd)");
        return 1;
    }
    return splat_bwd_launch(cfg, means, quats, scales, opacities, viewmat, K, rec, gint, wpix, seed_scale, last_depth, last_gid,
                            tile_stop, status, g_begin, g_end, nullptr, nullptr, nullptr, nullptr, nullptr, absgrad_accum,
                            *push, stream);
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
