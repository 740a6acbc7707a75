// eg_common.cuh -- shared declarations of the sm_100a kernels behind include/edgegs.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/edgegs.h"

#define EG_TILE 16
#define EG_ALPHA_MAX 0.999f
#define EG_ALPHA_MIN (1.0f / 255.0f)
#define EG_T_MIN 1e-4f
#define EG_LOG2E 1.4426950408889634f

void eg_set_error(const char *fmt, ...);
int eg_check_launch(const char *what);

// camera block shared by the projection kernels (read from device memory; no host sync)
struct EgCam {
    float R[9];
    float t[3];
    float fx, fy, cx, cy;
};

__device__ __forceinline__ EgCam eg_load_cam(const float *__restrict__ vm, const float *__restrict__ K) {
    EgCam c;
    c.R[0] = __ldg(vm + 0); c.R[1] = __ldg(vm + 1); c.R[2] = __ldg(vm + 2);  c.t[0] = __ldg(vm + 3);
    c.R[3] = __ldg(vm + 4); c.R[4] = __ldg(vm + 5); c.R[5] = __ldg(vm + 6);  c.t[1] = __ldg(vm + 7);
    c.R[6] = __ldg(vm + 8); c.R[7] = __ldg(vm + 9); c.R[8] = __ldg(vm + 10); c.t[2] = __ldg(vm + 11);
    c.fx = __ldg(K + 0); c.fy = __ldg(K + 4); c.cx = __ldg(K + 2); c.cy = __ldg(K + 5);
    return c;
}

// gsplat's (uint32_t)floor(x) on CUDA: cvt.rzi.u32.f32 saturates (negatives, NaN -> 0)
__device__ __forceinline__ uint32_t eg_sat_u32(float v) { return __float2uint_rz(v); }

// tile rectangle of a Gaussian (gsplat isect_tiles); canonical fp32 order, see DESIGN.md
__device__ __forceinline__ void eg_tile_rect(float m2x, float m2y, int radius, int tw, int th,
                                             uint32_t &x0, uint32_t &y0, uint32_t &x1, uint32_t &y1) {
    const float ts = (float)EG_TILE;
    const float tr = __fdiv_rn((float)radius, ts);
    const float txc = __fdiv_rn(m2x, ts), tyc = __fdiv_rn(m2y, ts);
    x0 = min(eg_sat_u32(floorf(__fsub_rn(txc, tr))), (uint32_t)tw);
    y0 = min(eg_sat_u32(floorf(__fsub_rn(tyc, tr))), (uint32_t)th);
    x1 = min(eg_sat_u32(ceilf(__fadd_rn(txc, tr))), (uint32_t)tw);
    y1 = min(eg_sat_u32(ceilf(__fadd_rn(tyc, tr))), (uint32_t)th);
}

__device__ __forceinline__ float eg_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float eg_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void eg_red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// v_quats[g] = vq: one 128-bit store when the tensor is 16-byte aligned (always the case inside the padded flat gradient
// buffer, eg_grad_layout), four scalar stores otherwise (a caller-supplied [N,4] tensor at an odd offset)
__device__ __forceinline__ void eg_store_quat_grad(float *__restrict__ v_quats, int g, const float (&vq)[4]) {
    if ((reinterpret_cast<uintptr_t>(v_quats) & 15) == 0) {
        reinterpret_cast<float4 *>(v_quats)[g] = make_float4(vq[0], vq[1], vq[2], vq[3]);
    } else {
        v_quats[4 * g] = vq[0]; v_quats[4 * g + 1] = vq[1]; v_quats[4 * g + 2] = vq[2]; v_quats[4 * g + 3] = vq[3];
    }
}

// Push form of the gradient exchange (eg_push_target, include/edgegs.h): where the four gradient outputs of Gaussian g
// go.  world <= 1: the caller's tensors.  Else slot `rank` of the staging area of g's owner o = g / per, laid out
// means [3 per] | scales [3 per] | quats [4 per] | opacities [per] and indexed by g - o * per -- returned as pointers
// that are still indexed by g itself.
struct EgGradOut {
    float *means, *scales, *quats, *opac;
};
__device__ __forceinline__ EgGradOut eg_grad_out(const eg_push_target &push, const int g, float *v_means, float *v_scales,
                                                 float *v_quats, float *v_opacities) {
    EgGradOut o;
    o.means = v_means; o.scales = v_scales; o.quats = v_quats; o.opac = v_opacities;
    if (push.world > 1) {
        const int owner = g / push.per;
        const long long per = push.per, g0 = (long long)owner * per;
        float *st = push.stage[0];   // select chain: a dynamically indexed kernel-parameter array would be copied to local memory
#pragma unroll
        for (int r = 1; r < 8; ++r) st = (owner == r) ? push.stage[r] : st;
        float *base = st + (long long)push.rank * 11 * per;
        o.means = base - 3 * g0;
        o.scales = base + 3 * per - 3 * g0;
        o.quats = base + 6 * per - 4 * g0;
        o.opac = base + 10 * per - g0;
    }
    return o;
}
int eg_push_target_ok(const char *what, const eg_push_target *push, int n);

// Packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2: two IEEE-rn fp32 operations per issued instruction,
// operands in even-aligned register pairs).  Element-wise results are bit-identical to the scalar .rn forms.
typedef unsigned long long eg_f2;
__device__ __forceinline__ eg_f2 f2_pack(float lo, float hi) {
    eg_f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ eg_f2 f2_dup(float v) { return f2_pack(v, v); }
__device__ __forceinline__ void f2_unpack(eg_f2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ eg_f2 f2_fma(eg_f2 a, eg_f2 b, eg_f2 c) {
    eg_f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ eg_f2 f2_add(eg_f2 a, eg_f2 b) {
    eg_f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ eg_f2 f2_mul(eg_f2 a, eg_f2 b) {
    eg_f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// sigma and opacity*exp(-sigma) with a FIXED operation order (explicit intrinsics, no compiler
// contraction) so that the forward and the backward kernels take identical skip decisions
// (sigma < 0, alpha < 1/255) for every (pixel, Gaussian) pair.
__device__ __forceinline__ float eg_sigma(float A, float B, float C, float dx, float dy) {
    const float q = __fmaf_rn(__fmul_rn(A, dx), dx, __fmul_rn(__fmul_rn(C, dy), dy));
    return __fmaf_rn(0.5f, q, __fmul_rn(__fmul_rn(B, dx), dy));
}
__device__ __forceinline__ float eg_vis(float sigma) { return eg_ex2(__fmul_rn(-sigma, EG_LOG2E)); }

// Folded form used by the raster kernels: with  fa = -0.5*log2(e)*A, fb = -log2(e)*B, fc = -0.5*log2(e)*C,
// lo = log2(opacity)  the exponent  p = lo - sigma*log2(e)  is six multiply-adds and
// opacity*exp(-sigma) = 2^p  (one MUFU.EX2, no extra multiplies);  sigma >= 0  <=>  p <= lo.
// Both raster kernels derive (fa, fb, fc, lo) from the same record with these exact expressions, so
// their skip decisions agree bit for bit.
struct EgFold {
    float fa, fb, fc, lo;
};
__device__ __forceinline__ EgFold eg_fold(float A, float B, float C, float o) {
    EgFold f;
    f.fa = __fmul_rn(A, -0.5f * EG_LOG2E);
    f.fb = __fmul_rn(B, -EG_LOG2E);
    f.fc = __fmul_rn(C, -0.5f * EG_LOG2E);
    f.lo = __log2f(o);
    return f;
}
// Horner form in dx: both roundings of the per-row terms  b1 = fb*dy  and  c0 = fc*dy*dy + lo  are explicit, so a
// kernel that walks a pixel ROW (dy fixed) hoists them and still reproduces this value bit for bit.
__device__ __forceinline__ float eg_pow2row_b1(float fb, float dy) { return __fmul_rn(fb, dy); }
__device__ __forceinline__ float eg_pow2row_c0(float fc, float lo, float dy) { return __fmaf_rn(__fmul_rn(fc, dy), dy, lo); }
__device__ __forceinline__ float eg_pow2row(float fa, float b1, float c0, float dx) {
    return __fmaf_rn(__fmaf_rn(fa, dx, b1), dx, c0);
}
__device__ __forceinline__ float eg_pow2arg(float fa, float fb, float fc, float lo, float dx, float dy) {
    return eg_pow2row(fa, eg_pow2row_b1(fb, dy), eg_pow2row_c0(fc, lo, dy), dx);
}

// Conservative half-extents (pixels) of the region where a Gaussian can reach alpha >= 1/255:
// { p : sigma(p) <= tau },  tau = ln(255 * opacity) (+ safety margin).  Returns false when the
// Gaussian can never contribute (opacity < 1/255).  Used ONLY to skip work whose result the
// alpha test would discard anyway, so it never changes results.
__device__ __forceinline__ bool eg_extent(float o, float A, float B, float C, float &hx, float &hy, float &tau) {
    if (!(o >= EG_ALPHA_MIN)) return false;
    tau = __logf(255.0f * o) * 1.001f + 0.02f;
    const float det = A * C - B * B;
    if (!(det > 0.0f) || !(A > 0.0f) || !(C > 0.0f)) {
        hx = hy = 1e30f;
        return true;
    }
    const float k = 2.0f * tau / det;
    hx = sqrtf(k * C) * 1.0001f + 1e-3f;
    hy = sqrtf(k * A) * 1.0001f + 1e-3f;
    return true;
}

// Conservative x-range (pixel-CENTRE coordinates) of the alpha >= 1/255 footprint of a Gaussian inside the horizontal
// strip of pixel-centre rows [ya, yb].  The footprint is the ellipse  sigma(u, v) = (A u^2 + C v^2) / 2 + B u v <= tau
// around the mean, tau = ln(255 o) (+ the safety margin of eg_extent); the strip cuts a slab out of it whose u-extent
// is attained at the slab's ends or at the ellipse's left / right extreme points.  hu, hv = eg_extent's half extents
// (a degenerate conic, hu >= 1e29, must be handled by the caller).  Only ever used to skip (tile / sub-tile,
// Gaussian) pairs in which the exact per-pixel alpha test of the raster kernels could never pass.
__device__ __forceinline__ bool eg_strip_xrange(float mx, float my, float A, float B, float C, float tau, float hu,
                                                float hv, float ya, float yb, float &xlo, float &xhi) {
    const float va = ya - my, vb = yb - my;
    if (vb < -hv || va > hv) return false;
    const float v1 = fmaxf(va, -hv), v2 = fminf(vb, hv);
    const float det = A * C - B * B, iA = 1.0f / A;
    const float s1 = sqrtf(fmaxf(0.0f, 2.0f * tau * A - det * v1 * v1)), s2 = sqrtf(fmaxf(0.0f, 2.0f * tau * A - det * v2 * v2));
    float umax = fmaxf((-B * v1 + s1) * iA, (-B * v2 + s2) * iA);
    float umin = fminf((-B * v1 - s1) * iA, (-B * v2 - s2) * iA);
    const float vr = -B * hu / C;  // v of the right-most point (u = +hu); the left-most one is at -vr
    if (vr >= v1 && vr <= v2) umax = hu;
    if (-vr >= v1 && -vr <= v2) umin = -hu;
    xhi = mx + (umax + fabsf(umax) * 1e-4f + 0.02f);
    xlo = mx + (umin - fabsf(umin) * 1e-4f - 0.02f);
    return true;
}

// Tile columns of tile row `ty` that the footprint can reach (EG_FLAG_CULL_TILES); (x0, x1) = the tile-column range
// of gsplat's rectangle.  Returns false when the row holds no tile; else [j0, j1] (inclusive).
__device__ __forceinline__ bool eg_tile_row_cols(float mx, float my, float A, float B, float C, float tau, float hu,
                                                 float hv, int ty, int x0, int x1, int &j0, int &j1) {
    j0 = x0;
    j1 = x1 - 1;
    if (!(hu < 1e29f)) return j1 >= j0;  // degenerate conic: eg_extent gave up, keep the whole row
    const float ya = (float)(ty * EG_TILE) + 0.5f;
    float xlo, xhi;
    if (!eg_strip_xrange(mx, my, A, B, C, tau, hu, hv, ya, ya + (float)(EG_TILE - 1), xlo, xhi)) return false;
    // pixel columns whose centre px + 0.5 lies in [xlo, xhi]
    const float pa = ceilf(xlo - 0.5f), pb = floorf(xhi - 0.5f);
    if (!(pb >= pa)) return false;
    const int ja = (int)fmaxf(pa, 0.0f) >> 4, jb = (int)fminf(fmaxf(pb, -1.0f), 1e9f) >> 4;
    j0 = max(j0, ja);
    j1 = min(j1, jb);
    return j1 >= j0 && pb >= 0.0f;
}

// Which of the eight 8x4-pixel sub-tiles of tile (X0, Y0) the footprint can reach: bit (2 r + c) = sub-tile row r
// (pixel rows Y0 + 4r .. Y0 + 4r + 3), column half c.  Same conservative strip test, one per sub-tile row.
__device__ __forceinline__ int eg_subtile_mask(float mx, float my, float A, float B, float C, float o, float X0, float Y0) {
    float hu, hv, tau;
    if (!eg_extent(o, A, B, C, hu, hv, tau)) return 0;
    if (!(hu < 1e29f)) return 0xff;
    int mask = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float xlo, xhi;
        const float ya = Y0 + 4.0f * r + 0.5f;
        if (!eg_strip_xrange(mx, my, A, B, C, tau, hu, hv, ya, ya + 3.0f, xlo, xhi)) continue;
        int cx = 0;
        if (xhi >= X0 + 0.5f && xlo <= X0 + 7.5f) cx |= 1;
        if (xhi >= X0 + 8.5f && xlo <= X0 + 15.5f) cx |= 2;
        mask |= cx << (2 * r);
    }
    return mask;
}

// Coefficient of |clamp(render) - gt| at one pixel in the fused projection loss (and of that pixel's backward seed):
// loss_params == NULL -> 1 ("whole", edge_gs.py:290-296); else loss_params = (w_edge, w_bg, w_sel, threshold) on the
// device: w_edge where gt >= threshold (the reference's edge mask, edge_gs.py:154-159), w_bg elsewhere, plus w_sel
// where sel_mask != 0 -- "weighted" (edge_gs.py:177-193,316-319) and "bg_edge_ratio" (edge_gs.py:298-314) are both
// of this form, see edge_gs.py::loss_spec.
__device__ __forceinline__ float eg_loss_coef(const float *__restrict__ loss_params,
                                              const unsigned char *__restrict__ sel_mask, float g, long long pix) {
    if (loss_params == nullptr) return 1.0f;
    float c = g >= __ldg(loss_params + 3) ? __ldg(loss_params) : __ldg(loss_params + 1);
    if (sel_mask != nullptr && __ldg(sel_mask + pix) != 0) c += __ldg(loss_params + 2);
    return c;
}
