/*
 * eg_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the splat operator that the reference calls at
 *   /root/reference/edgegaussians/models/edge_gs.py:250-268  (gsplat.rasterization, gsplat==1.0.0,
 *   requirements.txt:64).
 * gsplat itself is an un-vendored third-party dependency that is absent from /root/reference and
 * from this image, so what is restated here is its *published algorithm* as tabulated in
 * SURVEY.md section 8a rows a3..a7 / Appendix A (projection, tile binning + stable sort, front-to-back
 * compositing, the reverse-walk backward with abs-grad, the projection VJP).
 *
 * PARITY UNPINNED for rows a3..a7: the reference holds no test, golden vector or fixture at the
 * gsplat boundary and gsplat cannot be run here.  (Rows a8, a10, a11, a12 -- losses, regularisers,
 * KNN -- ARE pinned: see oracle/reference_ports.py and tests/golden/.)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  Nothing under edgegaussians_b200/ may.
 *
 * Floating point: the forward projection and the tile arithmetic are written as one IEEE fp32
 * operation per statement and must be compiled with -ffp-contract=off so that the integer outputs
 * (radii, tile rectangles, sort order) are reproducible bit for bit by any implementation that
 * follows the same evaluation order (DESIGN.md "canonical evaluation order").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define EGO_ALPHA_MAX 0.999f
#define EGO_ALPHA_MIN (1.0f / 255.0f)
#define EGO_T_MIN 1e-4f

int ego_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ego_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline float f_min(float a, float b) { return a < b ? a : b; }
static inline float f_max(float a, float b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------------- */
/* A.1 projection forward (gsplat 1.0.0 fully_fused_projection_fwd).    */
/* scales/opacities are the ACTIVATED values, as passed at edge_gs.py:253-254.                  */
/* viewmat: row-major 4x4 world->camera, K: row-major 3x3 (cameras.py:84-96).                   */
/* Outputs for culled Gaussians: radii = 0, everything else 0.                                  */
/* ------------------------------------------------------------------------------------------- */
void ego_project_fwd(int N, const float *means, const float *quats, const float *scales,
                     const float *viewmat, const float *K, int W, int H, float eps2d,
                     float near_plane, float far_plane, float radius_clip, int32_t *radii,
                     float *means2d, float *depths, float *conics, float *comps) {
    const float R00 = viewmat[0], R01 = viewmat[1], R02 = viewmat[2], t0 = viewmat[3];
    const float R10 = viewmat[4], R11 = viewmat[5], R12 = viewmat[6], t1 = viewmat[7];
    const float R20 = viewmat[8], R21 = viewmat[9], R22 = viewmat[10], t2 = viewmat[11];
    const float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const float R[3][3] = {{R00, R01, R02}, {R10, R11, R12}, {R20, R21, R22}};

#pragma omp parallel for schedule(static)
    for (int g = 0; g < N; ++g) {
        radii[g] = 0;
        means2d[2 * g] = means2d[2 * g + 1] = 0.f;
        depths[g] = 0.f;
        conics[3 * g] = conics[3 * g + 1] = conics[3 * g + 2] = 0.f;
        comps[g] = 0.f;

        const float mx = means[3 * g], my = means[3 * g + 1], mz = means[3 * g + 2];
        const float x = ((R00 * mx + R01 * my) + R02 * mz) + t0;
        const float y = ((R10 * mx + R11 * my) + R12 * mz) + t1;
        const float z = ((R20 * mx + R21 * my) + R22 * mz) + t2;
        if (z < near_plane || z > far_plane) continue;

        /* quaternion (w,x,y,z), normalised inside the op (edge_gs.py:229-230 never normalises) */
        float qw = quats[4 * g], qx = quats[4 * g + 1], qy = quats[4 * g + 2], qz = quats[4 * g + 3];
        const float n2 = ((qx * qx + qy * qy) + qz * qz) + qw * qw;
        const float inv_n = 1.0f / sqrtf(n2);
        qw = qw * inv_n; qx = qx * inv_n; qy = qy * inv_n; qz = qz * inv_n;
        const float x2 = qx * qx, y2 = qy * qy, z2 = qz * qz;
        const float xy = qx * qy, xz = qx * qz, yz = qy * qz;
        const float wx = qw * qx, wy = qw * qy, wz = qw * qz;
        float Rq[3][3];
        Rq[0][0] = 1.0f - 2.0f * (y2 + z2); Rq[0][1] = 2.0f * (xy - wz); Rq[0][2] = 2.0f * (xz + wy);
        Rq[1][0] = 2.0f * (xy + wz); Rq[1][1] = 1.0f - 2.0f * (x2 + z2); Rq[1][2] = 2.0f * (yz - wx);
        Rq[2][0] = 2.0f * (xz - wy); Rq[2][1] = 2.0f * (yz + wx); Rq[2][2] = 1.0f - 2.0f * (x2 + y2);
        const float s[3] = {scales[3 * g], scales[3 * g + 1], scales[3 * g + 2]};
        float M[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) M[i][j] = Rq[i][j] * s[j];
        /* Sigma = M M^T, upper triangle then mirrored */
        float S[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j) {
                S[i][j] = (M[i][0] * M[j][0] + M[i][1] * M[j][1]) + M[i][2] * M[j][2];
                S[j][i] = S[i][j];
            }
        /* Sigma_c = R Sigma R^T, upper triangle then mirrored */
        float Tm[3][3], Sc[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                Tm[i][j] = (R[i][0] * S[0][j] + R[i][1] * S[1][j]) + R[i][2] * S[2][j];
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j) {
                Sc[i][j] = (Tm[i][0] * R[j][0] + Tm[i][1] * R[j][1]) + Tm[i][2] * R[j][2];
                Sc[j][i] = Sc[i][j];
            }
        /* perspective projection with the symmetric 1.3*tan_fov clamp of gsplat 1.0.0 */
        const float tan_fovx = (0.5f * (float)W) / fx;
        const float tan_fovy = (0.5f * (float)H) / fy;
        const float lim_x = 1.3f * tan_fovx, lim_y = 1.3f * tan_fovy;
        const float rz = 1.0f / z;
        const float rz2 = rz * rz;
        const float tx = z * f_min(lim_x, f_max(-lim_x, x * rz));
        const float ty = z * f_min(lim_y, f_max(-lim_y, y * rz));
        const float J00 = fx * rz, J11 = fy * rz;
        const float J02 = -((fx * tx) * rz2);
        const float J12 = -((fy * ty) * rz2);
        const float a00 = J00 * Sc[0][0] + J02 * Sc[2][0];
        const float a01 = J00 * Sc[0][1] + J02 * Sc[2][1];
        const float a02 = J00 * Sc[0][2] + J02 * Sc[2][2];
        const float a11 = J11 * Sc[1][1] + J12 * Sc[2][1];
        const float a12 = J11 * Sc[1][2] + J12 * Sc[2][2];
        const float c00_0 = a00 * J00 + a02 * J02;
        const float c01 = a01 * J11 + a02 * J12;
        const float c11_0 = a11 * J11 + a12 * J12;
        const float m2x = (fx * x) * rz + cx;
        const float m2y = (fy * y) * rz + cy;
        /* blur + compensation */
        const float det0 = c00_0 * c11_0 - c01 * c01;
        const float c00 = c00_0 + eps2d, c11 = c11_0 + eps2d;
        const float det = c00 * c11 - c01 * c01;
        if (!(det > 0.0f)) continue; /* det <= 0 (or NaN) culls */
        const float comp = sqrtf(f_max(0.0f, det0 / det));
        const float inv_det = 1.0f / det;
        const float cA = c11 * inv_det;
        const float cB = -c01 * inv_det;
        const float cC = c00 * inv_det;
        const float b = 0.5f * (c00 + c11);
        const float v1 = b + sqrtf(f_max(0.01f, b * b - det));
        const float radius = ceilf(3.0f * sqrtf(v1));
        if (radius <= radius_clip) continue;
        if (m2x + radius <= 0.0f || m2x - radius >= (float)W || m2y + radius <= 0.0f ||
            m2y - radius >= (float)H)
            continue;
        radii[g] = (int32_t)radius;
        means2d[2 * g] = m2x; means2d[2 * g + 1] = m2y;
        depths[g] = z;
        conics[3 * g] = cA; conics[3 * g + 1] = cB; conics[3 * g + 2] = cC;
        comps[g] = comp;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* A.2 tile binning (gsplat 1.0.0 isect_tiles / isect_offset_encode).                           */
/* ------------------------------------------------------------------------------------------- */
static inline uint32_t sat_u32(float v) { /* CUDA cvt.rzi.u32.f32 semantics: saturating */
    if (!(v > 0.0f)) return 0u;           /* negatives and NaN -> 0 */
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}

static inline void tile_rect(float m2x, float m2y, int32_t radius, int tile_size, int tw, int th,
                             uint32_t *x0, uint32_t *y0, uint32_t *x1, uint32_t *y1) {
    const float ts = (float)tile_size;
    const float tr = (float)radius / ts;
    const float txc = m2x / ts, tyc = m2y / ts;
    uint32_t a;
    a = sat_u32(floorf(txc - tr)); *x0 = a < (uint32_t)tw ? a : (uint32_t)tw;
    a = sat_u32(floorf(tyc - tr)); *y0 = a < (uint32_t)th ? a : (uint32_t)th;
    a = sat_u32(ceilf(txc + tr));  *x1 = a < (uint32_t)tw ? a : (uint32_t)tw;
    a = sat_u32(ceilf(tyc + tr));  *y1 = a < (uint32_t)th ? a : (uint32_t)th;
}

/* pass 1: tiles touched per Gaussian. Returns n_isects. */
int64_t ego_isect_count(int N, const float *means2d, const int32_t *radii, int tile_size, int tw,
                        int th, int32_t *tiles_per_gauss) {
    int64_t total = 0;
    for (int g = 0; g < N; ++g) {
        if (radii[g] <= 0) { tiles_per_gauss[g] = 0; continue; }
        uint32_t x0, y0, x1, y1;
        tile_rect(means2d[2 * g], means2d[2 * g + 1], radii[g], tile_size, tw, th, &x0, &y0, &x1, &y1);
        /* uint32 arithmetic as in gsplat; x1 >= x0 may fail only for NaN-free inputs never */
        tiles_per_gauss[g] = (int32_t)((y1 - y0) * (x1 - x0));
        total += tiles_per_gauss[g];
    }
    return total;
}

int ego_tile_bits(int n_tiles) { /* floor(log2(n_tiles)) + 1 */
    int b = 0;
    while ((1 << (b + 1)) <= n_tiles) ++b;
    return b + 1;
}

/* stable LSD radix sort on the low `nbits` bits of 64-bit keys with a 32-bit payload */
static void radix_sort_pairs(int64_t n, int64_t *keys, int32_t *vals, int nbits) {
    int64_t *k2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    int32_t *v2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int64_t *src_k = keys, *dst_k = k2;
    int32_t *src_v = vals, *dst_v = v2;
    for (int shift = 0; shift < nbits; shift += 8) {
        int64_t count[257];
        memset(count, 0, sizeof(count));
        int bits = nbits - shift < 8 ? nbits - shift : 8;
        uint64_t mask = ((uint64_t)1 << bits) - 1;
        for (int64_t i = 0; i < n; ++i) count[(((uint64_t)src_k[i]) >> shift & mask) + 1]++;
        for (int d = 0; d < 256; ++d) count[d + 1] += count[d];
        for (int64_t i = 0; i < n; ++i) {
            int64_t p = count[((uint64_t)src_k[i]) >> shift & mask]++;
            dst_k[p] = src_k[i];
            dst_v[p] = src_v[i];
        }
        int64_t *tk = src_k; src_k = dst_k; dst_k = tk;
        int32_t *tv = src_v; src_v = dst_v; dst_v = tv;
    }
    if (src_k != keys) {
        memcpy(keys, src_k, sizeof(int64_t) * (size_t)n);
        memcpy(vals, src_v, sizeof(int32_t) * (size_t)n);
    }
    free(k2);
    free(v2);
}

/* pass 2: emit (key,val) in Gaussian order, stable sort, offsets. C (cameras) = 1. */
void ego_isect_emit_sort(int N, const float *means2d, const int32_t *radii, const float *depths,
                         int tile_size, int tw, int th, int64_t n_isects, int64_t *isect_ids,
                         int32_t *flatten_ids, int32_t *isect_offsets /* [th*tw] */) {
    const int n_tiles = tw * th;
    const int tile_bits = ego_tile_bits(n_tiles);
    const int cam_bits = 1; /* floor(log2(1)) + 1 */
    int64_t cur = 0;
    for (int g = 0; g < N; ++g) {
        if (radii[g] <= 0) continue;
        uint32_t x0, y0, x1, y1;
        tile_rect(means2d[2 * g], means2d[2 * g + 1], radii[g], tile_size, tw, th, &x0, &y0, &x1, &y1);
        int32_t dbits;
        memcpy(&dbits, &depths[g], 4);
        const int64_t depth_enc = (int64_t)dbits;
        for (uint32_t i = y0; i < y1; ++i)
            for (uint32_t j = x0; j < x1; ++j) {
                const int64_t tile_id = (int64_t)i * tw + j;
                isect_ids[cur] = (tile_id << 32) | depth_enc; /* cam 0 */
                flatten_ids[cur] = g;
                ++cur;
            }
    }
    (void)n_isects;
    radix_sort_pairs(cur, isect_ids, flatten_ids, 32 + tile_bits + cam_bits);
    /* offsets[t] = first sorted position whose tile id >= t */
    int64_t p = 0;
    for (int t = 0; t < n_tiles; ++t) {
        while (p < cur && (isect_ids[p] >> 32) < t) ++p;
        isect_offsets[t] = (int32_t)p;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* A.3 compositing forward (gsplat 1.0.0 rasterize_to_pixels_fwd, COLOR_DIM = 3, colors == 1).  */
/* render: [H,W,3], alpha: [H,W], last_ids: [H,W] (absolute sorted position).                   */
/* ------------------------------------------------------------------------------------------- */
void ego_raster_fwd(int W, int H, int tile_size, int tw, int th, int64_t n_isects,
                    const int32_t *isect_offsets, const int32_t *flatten_ids, const float *means2d,
                    const float *conics, const float *opac, const float *colors /* [N,3] or NULL=1 */,
                    float *render, float *alpha, int32_t *last_ids) {
    const int n_tiles = tw * th;
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < n_tiles; ++t) {
        const int ti = t / tw, tj = t % tw;
        const int64_t start = isect_offsets[t];
        const int64_t end = (t == n_tiles - 1) ? n_isects : isect_offsets[t + 1];
        for (int i = ti * tile_size; i < (ti + 1) * tile_size && i < H; ++i)
            for (int j = tj * tile_size; j < (tj + 1) * tile_size && j < W; ++j) {
                const float px = (float)j + 0.5f, py = (float)i + 0.5f;
                float T = 1.0f, out[3] = {0.f, 0.f, 0.f};
                int32_t cur_idx = 0;
                for (int64_t k = start; k < end; ++k) {
                    const int32_t g = flatten_ids[k];
                    const float dx = means2d[2 * g] - px, dy = means2d[2 * g + 1] - py;
                    const float A = conics[3 * g], B = conics[3 * g + 1], C = conics[3 * g + 2];
                    const float sigma = 0.5f * (A * dx * dx + C * dy * dy) + B * dx * dy;
                    const float a = f_min(EGO_ALPHA_MAX, opac[g] * expf(-sigma));
                    if (sigma < 0.f || a < EGO_ALPHA_MIN) continue;
                    const float nT = T * (1.0f - a);
                    if (nT <= EGO_T_MIN) break;
                    const float vis = a * T;
                    for (int c = 0; c < 3; ++c) out[c] += (colors ? colors[3 * g + c] : 1.0f) * vis;
                    cur_idx = (int32_t)k;
                    T = nT;
                }
                const int64_t pix = (int64_t)i * W + j;
                alpha[pix] = 1.0f - T;
                render[3 * pix] = out[0]; render[3 * pix + 1] = out[1]; render[3 * pix + 2] = out[2];
                last_ids[pix] = cur_idx;
            }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* A.5 compositing backward (gsplat 1.0.0 rasterize_to_pixels_bwd), gsplat's own arithmetic     */
/* form per pixel (fp32 T recovery by division, running buffer); the per-Gaussian sums over     */
/* pixels are accumulated in double (gsplat: fp32 warp sums + fp32 atomics, order undefined).   */
/* v_render: [H,W,3], v_alpha: [H,W] or NULL.                                                    */
/* Outputs (zeroed here): v_means2d [N,2], v_means2d_abs [N,2], v_conics [N,3], v_opac [N].      */
/* ------------------------------------------------------------------------------------------- */
void ego_raster_bwd(int N, int W, int H, int tile_size, int tw, int th, int64_t n_isects,
                    const int32_t *isect_offsets, const int32_t *flatten_ids, const float *means2d,
                    const float *conics, const float *opac, const float *colors,
                    const float *alpha, const int32_t *last_ids, const float *v_render,
                    const float *v_alpha, float *v_means2d, float *v_means2d_abs, float *v_conics,
                    float *v_opac) {
    const int n_tiles = tw * th;
    double *acc = (double *)calloc((size_t)N * 8, sizeof(double));
    int nthreads = ego_num_threads();
    /* per-thread private accumulators would cost N*8*nthreads doubles; tiles write to disjoint
       pixels but shared Gaussians, so serialise the accumulation with a critical per Gaussian
       chunk: simplest correct choice is atomics on doubles. */
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < n_tiles; ++t) {
        const int ti = t / tw, tj = t % tw;
        const int64_t start = isect_offsets[t];
        const int64_t end = (t == n_tiles - 1) ? n_isects : isect_offsets[t + 1];
        if (end <= start) continue;
        const int64_t L = end - start;
        double *loc = (double *)calloc((size_t)L * 8, sizeof(double));
        for (int i = ti * tile_size; i < (ti + 1) * tile_size && i < H; ++i)
            for (int j = tj * tile_size; j < (tj + 1) * tile_size && j < W; ++j) {
                const int64_t pix = (int64_t)i * W + j;
                const float px = (float)j + 0.5f, py = (float)i + 0.5f;
                const float T_final = 1.0f - alpha[pix];
                float T = T_final;
                float buf[3] = {0.f, 0.f, 0.f};
                const int32_t bin_final = last_ids[pix];
                const float vr[3] = {v_render[3 * pix], v_render[3 * pix + 1], v_render[3 * pix + 2]};
                const float va = v_alpha ? v_alpha[pix] : 0.0f;
                for (int64_t k = end - 1; k >= start; --k) {
                    if (k > bin_final) continue;
                    const int32_t g = flatten_ids[k];
                    const float dx = means2d[2 * g] - px, dy = means2d[2 * g + 1] - py;
                    const float A = conics[3 * g], B = conics[3 * g + 1], C = conics[3 * g + 2];
                    const float o = opac[g];
                    const float sigma = 0.5f * (A * dx * dx + C * dy * dy) + B * dx * dy;
                    const float vis = expf(-sigma);
                    const float a = f_min(EGO_ALPHA_MAX, o * vis);
                    if (sigma < 0.f || a < EGO_ALPHA_MIN) continue;
                    const float ra = 1.0f / (1.0f - a);
                    T *= ra;
                    const float fac = a * T;
                    float v_a = 0.f;
                    float col[3];
                    for (int c = 0; c < 3; ++c) {
                        col[c] = colors ? colors[3 * g + c] : 1.0f;
                        v_a += (col[c] * T - buf[c] * ra) * vr[c];
                    }
                    v_a += T_final * ra * va;
                    if (o * vis <= EGO_ALPHA_MAX) {
                        const float v_sigma = -o * vis * v_a;
                        double *d = loc + (k - start) * 8;
                        const float gx = v_sigma * (A * dx + B * dy);
                        const float gy = v_sigma * (B * dx + C * dy);
                        d[0] += gx; d[1] += gy;
                        d[2] += fabsf(gx); d[3] += fabsf(gy);
                        d[4] += 0.5f * v_sigma * dx * dx;
                        d[5] += v_sigma * dx * dy;
                        d[6] += 0.5f * v_sigma * dy * dy;
                        d[7] += vis * v_a;
                    }
                    for (int c = 0; c < 3; ++c) buf[c] += col[c] * fac;
                }
            }
        for (int64_t k = 0; k < L; ++k) {
            const int32_t g = flatten_ids[start + k];
            for (int c = 0; c < 8; ++c) {
                const double v = loc[k * 8 + c];
                if (v != 0.0) {
#pragma omp atomic
                    acc[(size_t)g * 8 + c] += v;
                }
            }
        }
        free(loc);
    }
    for (int g = 0; g < N; ++g) {
        const double *d = acc + (size_t)g * 8;
        v_means2d[2 * g] = (float)d[0]; v_means2d[2 * g + 1] = (float)d[1];
        v_means2d_abs[2 * g] = (float)d[2]; v_means2d_abs[2 * g + 1] = (float)d[3];
        v_conics[3 * g] = (float)d[4]; v_conics[3 * g + 1] = (float)d[5]; v_conics[3 * g + 2] = (float)d[6];
        v_opac[g] = (float)d[7];
    }
    free(acc);
}

/* ------------------------------------------------------------------------------------------- */
/* A.6 projection backward (gsplat 1.0.0 fully_fused_projection_bwd, viewmat grads off).         */
/* Inputs are fp32; arithmetic in double.  v_opac_eff is the gradient w.r.t. opacity' = o*comp   */
/* (antialiased) -- the split into v_o and v_comp (rendering.py: opacities*compensations) is     */
/* done here.  Outputs: v_means [N,3], v_quats [N,4], v_scales [N,3] (w.r.t. ACTIVATED scales),  */
/* v_opacities [N] (w.r.t. ACTIVATED opacity).                                                   */
/* ------------------------------------------------------------------------------------------- */
void ego_project_bwd(int N, const float *means, const float *quats, const float *scales,
                     const float *opacities, const float *viewmat, const float *K, int W, int H,
                     float eps2d, int antialiased, const int32_t *radii, const float *conics,
                     const float *comps, const float *v_means2d, const float *v_depths /* or NULL */,
                     const float *v_conics, const float *v_opac_eff, float *v_means, float *v_quats,
                     float *v_scales, float *v_opacities) {
    double R[3][3], tr[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[i][j] = viewmat[4 * i + j];
        tr[i] = viewmat[4 * i + 3];
    }
    const double fx = K[0], fy = K[4];
#pragma omp parallel for schedule(static)
    for (int g = 0; g < N; ++g) {
        for (int c = 0; c < 3; ++c) v_means[3 * g + c] = 0.f, v_scales[3 * g + c] = 0.f;
        for (int c = 0; c < 4; ++c) v_quats[4 * g + c] = 0.f;
        v_opacities[g] = 0.f;
        if (radii[g] <= 0) continue;
        const double o = opacities[g];
        double v_comp = 0.0;
        if (antialiased) {
            v_opacities[g] = (float)((double)v_opac_eff[g] * (double)comps[g]);
            v_comp = (double)v_opac_eff[g] * o;
        } else {
            v_opacities[g] = v_opac_eff[g];
        }
        const double A = conics[3 * g], B = conics[3 * g + 1], C = conics[3 * g + 2];
        const double vA = v_conics[3 * g], vB = v_conics[3 * g + 1], vC = v_conics[3 * g + 2];
        /* v_Sigma2 = -Cn V Cn */
        const double V[2][2] = {{vA, 0.5 * vB}, {0.5 * vB, vC}};
        const double Cn[2][2] = {{A, B}, {B, C}};
        double CV[2][2], vS2[2][2];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) CV[i][j] = Cn[i][0] * V[0][j] + Cn[i][1] * V[1][j];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) vS2[i][j] = -(CV[i][0] * Cn[0][j] + CV[i][1] * Cn[1][j]);
        if (antialiased) {
            const double comp = comps[g];
            const double detc = A * C - B * B;
            const double v_sqr = v_comp * 0.5 / (comp + 1e-6);
            const double om = 1.0 - comp * comp;
            vS2[0][0] += v_sqr * (om * A - (double)eps2d * detc);
            vS2[0][1] += v_sqr * (om * B);
            vS2[1][0] += v_sqr * (om * B);
            vS2[1][1] += v_sqr * (om * C - (double)eps2d * detc);
        }
        /* recompute forward intermediates */
        const double mx = means[3 * g], my = means[3 * g + 1], mz = means[3 * g + 2];
        const double x = R[0][0] * mx + R[0][1] * my + R[0][2] * mz + tr[0];
        const double y = R[1][0] * mx + R[1][1] * my + R[1][2] * mz + tr[1];
        const double z = R[2][0] * mx + R[2][1] * my + R[2][2] * mz + tr[2];
        double qw = quats[4 * g], qx = quats[4 * g + 1], qy = quats[4 * g + 2], qz = quats[4 * g + 3];
        const double qn = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
        qw /= qn; qx /= qn; qy /= qn; qz /= qn;
        double Rq[3][3];
        Rq[0][0] = 1 - 2 * (qy * qy + qz * qz); Rq[0][1] = 2 * (qx * qy - qw * qz); Rq[0][2] = 2 * (qx * qz + qw * qy);
        Rq[1][0] = 2 * (qx * qy + qw * qz); Rq[1][1] = 1 - 2 * (qx * qx + qz * qz); Rq[1][2] = 2 * (qy * qz - qw * qx);
        Rq[2][0] = 2 * (qx * qz - qw * qy); Rq[2][1] = 2 * (qy * qz + qw * qx); Rq[2][2] = 1 - 2 * (qx * qx + qy * qy);
        const double s[3] = {scales[3 * g], scales[3 * g + 1], scales[3 * g + 2]};
        double M[3][3], S[3][3], Tm[3][3], Sc[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) M[i][j] = Rq[i][j] * s[j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) S[i][j] = M[i][0] * M[j][0] + M[i][1] * M[j][1] + M[i][2] * M[j][2];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Tm[i][j] = R[i][0] * S[0][j] + R[i][1] * S[1][j] + R[i][2] * S[2][j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Sc[i][j] = Tm[i][0] * R[j][0] + Tm[i][1] * R[j][1] + Tm[i][2] * R[j][2];
        const double lim_x = 1.3 * (0.5 * W / fx), lim_y = 1.3 * (0.5 * H / fy);
        const double rz = 1.0 / z, rz2 = rz * rz, rz3 = rz2 * rz;
        const double xr = x * rz, yr = y * rz;
        const double tx = z * fmin(lim_x, fmax(-lim_x, xr));
        const double ty = z * fmin(lim_y, fmax(-lim_y, yr));
        const double J[2][3] = {{fx * rz, 0.0, -fx * tx * rz2}, {0.0, fy * rz, -fy * ty * rz2}};
        /* v_Sigma_c = J^T v_S2 J */
        double vSc[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double acc = 0;
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b) acc += J[a][i] * vS2[a][b] * J[b][j];
                vSc[i][j] = acc;
            }
        /* v_J = v_S2 J Sc^T + v_S2^T J Sc */
        double JSct[2][3], JSc[2][3], vJ[2][3];
        for (int a = 0; a < 2; ++a)
            for (int j = 0; j < 3; ++j) {
                double p = 0, q = 0;
                for (int k = 0; k < 3; ++k) { p += J[a][k] * Sc[j][k]; q += J[a][k] * Sc[k][j]; }
                JSct[a][j] = p; JSc[a][j] = q;
            }
        for (int a = 0; a < 2; ++a)
            for (int j = 0; j < 3; ++j)
                vJ[a][j] = vS2[a][0] * JSct[0][j] + vS2[a][1] * JSct[1][j] + vS2[0][a] * JSc[0][j] + vS2[1][a] * JSc[1][j];
        const double vmx = v_means2d[2 * g], vmy = v_means2d[2 * g + 1];
        double vp[3];
        vp[0] = fx * rz * vmx;
        vp[1] = fy * rz * vmy;
        vp[2] = -(fx * x * vmx + fy * y * vmy) * rz2;
        if (xr <= lim_x && xr >= -lim_x) vp[0] += -fx * rz2 * vJ[0][2];
        else vp[2] += -fx * rz3 * vJ[0][2] * tx;
        if (yr <= lim_y && yr >= -lim_y) vp[1] += -fy * rz2 * vJ[1][2];
        else vp[2] += -fy * rz3 * vJ[1][2] * ty;
        vp[2] += -fx * rz2 * vJ[0][0] - fy * rz2 * vJ[1][1] + 2.0 * fx * tx * rz3 * vJ[0][2] + 2.0 * fy * ty * rz3 * vJ[1][2];
        if (v_depths) vp[2] += v_depths[g];
        /* v_mu = R^T v_p ; v_Sigma = R^T v_Sc R */
        for (int c = 0; c < 3; ++c)
            v_means[3 * g + c] = (float)(R[0][c] * vp[0] + R[1][c] * vp[1] + R[2][c] * vp[2]);
        double RtV[3][3], vS[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) RtV[i][j] = R[0][i] * vSc[0][j] + R[1][i] * vSc[1][j] + R[2][i] * vSc[2][j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) vS[i][j] = RtV[i][0] * R[0][j] + RtV[i][1] * R[1][j] + RtV[i][2] * R[2][j];
        /* v_M = (v_S + v_S^T) M */
        double vM[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double acc = 0;
                for (int k = 0; k < 3; ++k) acc += (vS[i][k] + vS[k][i]) * M[k][j];
                vM[i][j] = acc;
            }
        double vRq[3][3];
        for (int j = 0; j < 3; ++j) {
            double acc = 0;
            for (int i = 0; i < 3; ++i) { acc += Rq[i][j] * vM[i][j]; vRq[i][j] = vM[i][j] * s[j]; }
            v_scales[3 * g + j] = (float)acc;
        }
        const double w_ = qw, x_ = qx, y_ = qy, z_ = qz;
        double vq[4];
        vq[0] = 2.0 * (x_ * (vRq[2][1] - vRq[1][2]) + y_ * (vRq[0][2] - vRq[2][0]) + z_ * (vRq[1][0] - vRq[0][1]));
        vq[1] = 2.0 * (-2.0 * x_ * (vRq[1][1] + vRq[2][2]) + y_ * (vRq[0][1] + vRq[1][0]) + z_ * (vRq[0][2] + vRq[2][0]) + w_ * (vRq[2][1] - vRq[1][2]));
        vq[2] = 2.0 * (x_ * (vRq[0][1] + vRq[1][0]) - 2.0 * y_ * (vRq[0][0] + vRq[2][2]) + z_ * (vRq[1][2] + vRq[2][1]) + w_ * (vRq[0][2] - vRq[2][0]));
        vq[3] = 2.0 * (x_ * (vRq[0][2] + vRq[2][0]) + y_ * (vRq[1][2] + vRq[2][1]) - 2.0 * z_ * (vRq[0][0] + vRq[1][1]) + w_ * (vRq[1][0] - vRq[0][1]));
        const double qh[4] = {w_, x_, y_, z_};
        const double dot = vq[0] * qh[0] + vq[1] * qh[1] + vq[2] * qh[2] + vq[3] * qh[3];
        for (int c = 0; c < 4; ++c) v_quats[4 * g + c] = (float)((vq[c] - dot * qh[c]) / qn);
    }
}
