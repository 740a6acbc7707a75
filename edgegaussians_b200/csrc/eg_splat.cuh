// eg_splat.cuh -- Gaussian-major walk of the (pixel, Gaussian) pairs of the splat, shared by eg_splat_fwd and
// eg_splat_bwd.
//
// gsplat composites Gaussian g at pixel p iff p lies in one of the 16x16 tiles of g's tile rectangle
// (isect_tiles), sigma >= 0, alpha = min(0.999, o*exp(-sigma)) >= 1/255, and p had not stopped before g
// (behind /root/reference/edgegaussians/models/edge_gs.py:250-268).  The first three
// conditions only involve g and p, so the pair set can be enumerated from the Gaussian side:
//   * a warp owns 32 consecutive Gaussians (lane = Gaussian) and derives, per Gaussian, the pixel rectangle of
//     its tile rectangle and the conservative row range of its alpha >= 1/255 ellipse;
//   * the rows of those 32 Gaussians form a dense list of (Gaussian, row) items; the warp walks it 32 items at
//     a time (lane = item), each lane solving the row's quadratic for its conservative pixel span and then
//     visiting the span in aligned 4-pixel chunks (one 128-bit access per chunk).
// Every visited pixel is still put through the EXACT test of the pixel-major kernels (eg_pow2arg's rounding,
// same thresholds), so the span arithmetic only has to be conservative, never exact.
#pragma once
#include "eg_common.cuh"

// log2(1/255) minus a margin that covers ex2.approx / lg2.approx error: only used to BOUND spans
#define EG_L2AMIN_CONS (-7.9965f)

// consecutive pixel rows of one Gaussian handled by one lane per work item: amortises the per-item overhead
// (owner lookup, staging fetch, segmented reduction) and evens out the per-lane work
#ifndef EG_ROWS_PER_ITEM
#define EG_ROWS_PER_ITEM 2
#endif

struct __align__(16) EgSplatG {
    float mx, my, fa, fb;  // mean2d, folded conic (eg_fold)
    float fc, lo, A, B;    // ..., log2(opacity'), conic a, b
    float C;               // conic c
    unsigned depth_bits;   // float bits of the depth (sort key high word)
    int X01;               // pixel columns [X0, X1) of the tile rectangle, clipped to the image: X0 | X1 << 16
    int nrows;             // pixel rows [ylo, ylo + nrows)
    int ylo, start, gid;   // first row, index of its first work item in the warp's list, Gaussian id
    float inv2a;           // 0.5 / fa (< 0 for a proper conic)
    __device__ __forceinline__ int X0() const { return X01 & 0xffff; }
    __device__ __forceinline__ int X1() const { return (int)((unsigned)X01 >> 16); }
};

// gsplat's per-pair test (rasterize_to_pixels): composited iff sigma >= 0 (p <= lo) and alpha >= 1/255
__device__ __forceinline__ bool eg_pair_valid(float ov, float p, float lo, bool in_span) {
    return in_span && p <= lo && ov >= EG_ALPHA_MIN;
}
// ... and alpha not clamped (ov <= 0.999: gsplat passes no gradient through the clamp).  ov = ex2(p) is +0, a
// positive float or NaN, so  1/255 <= ov <= 0.999  is one unsigned range check on its bit pattern.
__device__ __forceinline__ bool eg_pair_valid_grad(float ov, float p, float lo, bool in_span) {
    const unsigned lo_bits = 0x3b808081u /* 1/255 */, hi_bits = 0x3f7fbe77u /* 0.999f */;
    return in_span && p <= lo && (__float_as_uint(ov) - lo_bits) <= (hi_bits - lo_bits);
}

// v if the pair is composited with an unclamped alpha (eg_pair_valid_grad), else 0 -- as ONE predicate chain
// (integer range check, float compare AND-ed into the same predicate, select), which nvcc otherwise expands into
// a chain of selects
__device__ __forceinline__ float eg_select_valid_grad(float v, float ov, float p, float lo) {
    float r;
    asm("{\n\t.reg .pred q;\n\t.reg .u32 t;\n\t"
        "sub.u32 t, %1, 0x3b808081;\n\t"
        "setp.le.u32 q, t, 0x03ff3df6;\n\t"
        "setp.le.and.f32 q, %2, %3, q;\n\t"
        "selp.f32 %0, %4, 0f00000000, q;\n\t}"
        : "=f"(r)
        : "r"(__float_as_uint(ov)), "f"(p), "f"(lo), "f"(v));
    return r;
}
// same for the forward's test (eg_pair_valid): alpha >= 1/255  <=>  bits(ov) - bits(1/255) <= bits(+inf) - bits(1/255)
__device__ __forceinline__ float eg_select_valid(float v, float ov, float p, float lo) {
    float r;
    asm("{\n\t.reg .pred q;\n\t.reg .u32 t;\n\t"
        "sub.u32 t, %1, 0x3b808081;\n\t"
        "setp.le.u32 q, t, 0x43ff7f7f;\n\t"
        "setp.le.and.f32 q, %2, %3, q;\n\t"
        "selp.f32 %0, %4, 0f00000000, q;\n\t}"
        : "=f"(r)
        : "r"(__float_as_uint(ov)), "f"(p), "f"(lo), "f"(v));
    return r;
}

// lane = Gaussian.  Returns the number of work items (groups of EG_ROWS_PER_ITEM pixel rows) of the Gaussian, 0 = none.
__device__ __forceinline__ int eg_splat_setup(const eg_config &cfg, int tw, int th, int gid, const float4 r0,
                                              const float4 r1, int radius, EgSplatG &G) {
    G.gid = gid;
    G.start = 0;
    G.nrows = 0;
    if (radius <= 0) return 0;
    const EgFold f = eg_fold(r1.x, r1.y, r1.z, r0.z);
    if (!(f.lo >= EG_L2AMIN_CONS)) return 0;  // opacity' < 1/255: alpha >= 1/255 can never hold
    uint32_t tx0, ty0, tx1, ty1;
    eg_tile_rect(r0.x, r0.y, radius, tw, th, tx0, ty0, tx1, ty1);
    const int X0 = (int)tx0 * EG_TILE, X1 = min((int)tx1 * EG_TILE, cfg.width);
    const int Y0 = (int)ty0 * EG_TILE, Y1 = min((int)ty1 * EG_TILE, cfg.height);
    if (X1 <= X0 || Y1 <= Y0) return 0;
    G.mx = r0.x; G.my = r0.y; G.fa = f.fa; G.fb = f.fb; G.fc = f.fc; G.lo = f.lo;
    G.A = r1.x; G.B = r1.y; G.C = r1.z;
    G.depth_bits = __float_as_uint(r0.w);
    G.X01 = X0 | (X1 << 16);  // image sides < 65536 (checked by the host wrappers)
    G.inv2a = 0.5f / f.fa;
    // rows where max_x p(x, y) >= log2(1/255):  (fc - fb^2 / (4 fa)) dy^2 + lo >= L
    float ylo = (float)Y0, yhi = (float)(Y1 - 1);
    const float Kq = f.fc - f.fb * f.fb / (4.0f * f.fa);
    if (f.fa < 0.0f && Kq < 0.0f) {
        const float hy = sqrtf((f.lo - EG_L2AMIN_CONS) / (-Kq)) * 1.0001f + 0.02f;
        ylo = fmaxf(ylo, ceilf(r0.y - hy - 0.5f));
        yhi = fminf(yhi, floorf(r0.y + hy - 0.5f));
    }
    if (!(yhi >= ylo)) return 0;
    G.ylo = (int)ylo;
    G.nrows = (int)yhi - (int)ylo + 1;
    return (G.nrows + EG_ROWS_PER_ITEM - 1) / EG_ROWS_PER_ITEM;
}

// lane = (Gaussian, row).  Conservative pixel span [xa, xb] of the row; false = empty.
__device__ __forceinline__ bool eg_row_span(const EgSplatG &G, float b1, float c0, int &xa, int &xb) {
    float lo_x = (float)G.X0(), hi_x = (float)(G.X1() - 1);
    if (G.fa < 0.0f) {
        const float disc = b1 * b1 - 4.0f * G.fa * (c0 - EG_L2AMIN_CONS);
        if (disc < 0.0f) return false;
        const float inv = G.inv2a;  // 0.5 / fa < 0
        const float dxc = -b1 * inv;
        float sq;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(disc));  // the margins below dwarf its 2^-22 error
        const float w = -sq * inv * 1.0001f + 0.02f;
        const float cx = G.mx - dxc - 0.5f;  // pixel index (continuous) of the row's maximum
        lo_x = fmaxf(lo_x, ceilf(cx - w));   // fmaxf / fminf drop a NaN operand: a NaN span is the full row
        hi_x = fminf(hi_x, floorf(cx + w));
    }
    xa = (int)lo_x;
    xb = (int)hi_x;
    return xb >= xa;
}

// Row items of a warp's 32 Gaussians, without searching: the Gaussians that have rows are staged COMPACTED (index k
// = rank among the non-empty ones), so their list ends e_k are strictly increasing.  For the batch of items
// [base, base + 32) every Gaussian whose end falls in (base, base + 32] sets bit (end - base - 1) of a warp-wide
// OR; the owner of item base + j is then  k_base + popc(mask & ((1 << j) - 1))  with k_base the number of ends <= base,
// carried from batch to batch.  `end` / `nonempty` are the calling lane's values as a GAUSSIAN.
struct EgOwnerIter {
    int k_base;
    __device__ __forceinline__ EgOwnerIter() : k_base(0) {}
    __device__ __forceinline__ int owner(const int base, const int lane, const int end, const bool nonempty) {
        const unsigned t = (unsigned)(end - base - 1);
        const unsigned mask = __reduce_or_sync(0xffffffffu, (nonempty && t < 32u) ? (1u << t) : 0u);
        const int k = k_base + __popc(mask & ((1u << lane) - 1u));
        k_base += __popc(mask);
        return k;
    }
};
