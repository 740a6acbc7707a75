// eg_knn.cu -- exact k-nearest neighbours on the device for the direction regulariser
// (SURVEY.md section 8f rank 1).  Replaces k_nearest_sklearn (KD-tree on the CPU, 6.5 s per call at
// N = 500k) of /root/reference/edgegaussians/models/edge_gs.py:135-151, 326-344.
//
// Uniform grid hash: bounding box -> G^3 cells -> counting sort of the points by cell -> one thread per
// query walks cubic shells of cells outwards, keeping the K best (distance^2 in fp64, index) pairs in
// registers, and stops when the K-th best distance is covered by the shells already visited -- exact.
// Output: the neighbours of rank `skip` .. skip+kk-1 (rank 0 is the point itself): the reference asks
// sklearn for kk+2 neighbours and drops the first column twice, so skip = 2.
#include <cfloat>

#include "eg_common.cuh"

namespace {

struct KnnGrid {
    float min[3];
    float inv_cell;
    float cell;
    int G;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ unsigned f2ord(float f) {  // order-preserving float -> uint
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ws layout (ints): [0..2] min (ordered uint), [3..5] max (ordered uint), [6] G, [7] cell bits, [8..] cells
__global__ void knn_bbox_kernel(int n, const float *__restrict__ pts, unsigned *__restrict__ ws) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = pts[3 * i + c];
        if (!(v == v)) v = 0.0f;  // NaN scrub, as update_nearest_neighbors does (edge_gs.py:331-333)
        atomicMin(ws + c, f2ord(v));
        atomicMax(ws + 3 + c, f2ord(v));
    }
}

__device__ __forceinline__ KnnGrid knn_grid(const unsigned *ws, int n, int g_max) {
    KnnGrid g;
    float ext = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        g.min[c] = ord2f(ws[c]);
        ext = fmaxf(ext, ord2f(ws[3 + c]) - g.min[c]);
    }
    int G = (int)ceilf(cbrtf(0.5f * (float)n));
    G = clampi(G, 1, g_max);
    g.G = G;
    g.cell = fmaxf(ext / (float)G, 1e-12f) * 1.0001f;
    g.inv_cell = 1.0f / g.cell;
    return g;
}

__device__ __forceinline__ int cell_coord(float v, float mn, const KnnGrid &g) {
    if (!(v == v)) v = 0.0f;
    return clampi((int)floorf((v - mn) * g.inv_cell), 0, g.G - 1);
}

__global__ void knn_count_kernel(int n, const float *__restrict__ pts, const unsigned *__restrict__ ws, int g_max,
                                 int *__restrict__ cells) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const KnnGrid g = knn_grid(ws, n, g_max);
    const int cx = cell_coord(pts[3 * i], g.min[0], g), cy = cell_coord(pts[3 * i + 1], g.min[1], g),
              cz = cell_coord(pts[3 * i + 2], g.min[2], g);
    atomicAdd(cells + ((cz * g.G + cy) * g.G + cx), 1);
}

// single-CTA exclusive scan of the cell counts -> cell starts (in place) and a copy used as cursors
__global__ void __launch_bounds__(1024) knn_scan_kernel(int n, const unsigned *__restrict__ ws, int g_max,
                                                        int *__restrict__ cells, int *__restrict__ cursor) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const KnnGrid g = knn_grid(ws, n, g_max);
    const int C = g.G * g.G * g.G;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < C; base += 1024) {
        const int i = base + tid;
        const int v = i < C ? cells[i] : 0;
        int x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int w = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += y;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        const int incl = x + (wid > 0 ? wsum[wid - 1] : 0) + carry_s;
        if (i < C) {
            cells[i] = incl - v;
            cursor[i] = incl - v;
        }
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
    if (tid == 0) cells[C] = carry_s;
}

__global__ void knn_scatter_kernel(int n, const float *__restrict__ pts, const unsigned *__restrict__ ws, int g_max,
                                   int *__restrict__ cursor, float4 *__restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const KnnGrid g = knn_grid(ws, n, g_max);
    float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    if (!(x == x)) x = 0.f;
    if (!(y == y)) y = 0.f;
    if (!(z == z)) z = 0.f;
    const int cx = cell_coord(x, g.min[0], g), cy = cell_coord(y, g.min[1], g), cz = cell_coord(z, g.min[2], g);
    const int pos = atomicAdd(cursor + ((cz * g.G + cy) * g.G + cx), 1);
    sorted[pos] = make_float4(x, y, z, __int_as_float(i));
}

template <int K>
__global__ void __launch_bounds__(128) knn_query_kernel(int n, const float *__restrict__ pts,
                                                        const unsigned *__restrict__ ws, int g_max,
                                                        const int *__restrict__ cells,
                                                        const float4 *__restrict__ sorted, int kk, int skip,
                                                        int32_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const KnnGrid g = knn_grid(ws, n, g_max);
    float qx = pts[3 * i], qy = pts[3 * i + 1], qz = pts[3 * i + 2];
    if (!(qx == qx)) qx = 0.f;
    if (!(qy == qy)) qy = 0.f;
    if (!(qz == qz)) qz = 0.f;
    const int cx = cell_coord(qx, g.min[0], g), cy = cell_coord(qy, g.min[1], g), cz = cell_coord(qz, g.min[2], g);
    const int need = min(kk + skip, n);
    double bd[K];
    int bi[K];
#pragma unroll
    for (int t = 0; t < K; ++t) { bd[t] = DBL_MAX; bi[t] = 0x7fffffff; }
    // K-th best so far lives in slot need-1 (slots >= need stay at +inf and are never read)
    for (int r = 0; r < g.G + 1; ++r) {
        for (int dz = -r; dz <= r; ++dz) {
            const int z = cz + dz;
            if (z < 0 || z >= g.G) continue;
            for (int dy = -r; dy <= r; ++dy) {
                const int y = cy + dy;
                if (y < 0 || y >= g.G) continue;
                const bool face = (abs(dz) == r) || (abs(dy) == r);
                const int step = face ? 1 : max(2 * r, 1);  // interior rows of the shell: only the two end cells
                for (int dx = -r; dx <= r; dx += step) {
                    const int x = cx + dx;
                    if (x < 0 || x >= g.G) continue;
                    const int cell = (z * g.G + y) * g.G + x;
                    const int s = cells[cell], e = cells[cell + 1];
                    for (int p = s; p < e; ++p) {
                        const float4 c = __ldg(sorted + p);
                        const double ddx = (double)c.x - (double)qx, ddy = (double)c.y - (double)qy,
                                     ddz = (double)c.z - (double)qz;
                        double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                        int id = __float_as_int(c.w);
                        if (id == i) d2 = -1.0;  // the point itself is rank 0, as in the reference
                        // insert (d2, id) into the sorted list if it beats the current K-th best
                        if (d2 < bd[K - 1] || (d2 == bd[K - 1] && id < bi[K - 1])) {
#pragma unroll
                            for (int t = 0; t < K; ++t) {
                                const bool lt = d2 < bd[t] || (d2 == bd[t] && id < bi[t]);
                                if (lt) {
                                    const double td = bd[t]; const int ti = bi[t];
                                    bd[t] = d2; bi[t] = id;
                                    d2 = td; id = ti;
                                }
                            }
                        }
                    }
                }
            }
        }
        // every point closer than r * cell lies in the shells visited so far
        double kth = DBL_MAX;
#pragma unroll
        for (int t = 0; t < K; ++t)
            if (t == need - 1) kth = bd[t];
        const double reach = (double)r * (double)g.cell;
        if (kth <= reach * reach) break;
    }
#pragma unroll
    for (int t = 0; t < K; ++t) {
        const int col = t - skip;
        if (col >= 0 && col < kk) out[(long long)i * kk + col] = (t < need) ? bi[t] : i;
    }
}

}  // namespace

extern "C" size_t eg_knn_workspace_bytes(int n) {
    const int g_max = 160;
    long long G = (long long)ceil(cbrt(0.5 * (double)(n > 0 ? n : 1)));
    if (G < 1) G = 1;
    if (G > g_max) G = g_max;
    const long long cells = G * G * G + 1;
    // header (8 words) + cell starts + cursors + sorted points (float4)
    return (size_t)(8 + 2 * cells) * 4 + 16 + (size_t)(n > 0 ? n : 1) * 16;
}

extern "C" int eg_knn(int n, const float *points, int kk, int skip, int32_t *out, void *workspace,
                      size_t workspace_bytes, void *stream) {
    if (n <= 0) return 0;
    if (kk <= 0 || skip < 0 || kk + skip > 48) {
        eg_set_error("eg_knn: need 0 < kk and kk + skip <= 48 (got kk=%d skip=%d)", kk, skip);
        return 1;
    }
    if (workspace == nullptr || workspace_bytes < eg_knn_workspace_bytes(n)) {
        eg_set_error("eg_knn: workspace too small (%zu < %zu)", workspace_bytes, eg_knn_workspace_bytes(n));
        return 1;
    }
    const int g_max = 160;
    long long G = (long long)ceil(cbrt(0.5 * (double)n));
    if (G < 1) G = 1;
    if (G > g_max) G = g_max;
    const long long C = G * G * G + 1;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned *ws = (unsigned *)workspace;
    int *cells = (int *)(ws + 8);
    int *cursor = cells + C;
    size_t off = (size_t)(8 + 2 * C) * 4;
    off = (off + 15) / 16 * 16;
    float4 *sorted = (float4 *)((char *)workspace + off);
    // header: min = +inf (0xffffffff in ordered form), max = -inf (0); cells = 0
    cudaMemsetAsync(ws, 0xff, 3 * 4, s);
    cudaMemsetAsync(ws + 3, 0, (size_t)(5 + C) * 4, s);
    const int block = 256, grid = (n + block - 1) / block;
    knn_bbox_kernel<<<grid, block, 0, s>>>(n, points, ws);
    knn_count_kernel<<<grid, block, 0, s>>>(n, points, ws, g_max, cells);
    knn_scan_kernel<<<1, 1024, 0, s>>>(n, ws, g_max, cells, cursor);
    knn_scatter_kernel<<<grid, block, 0, s>>>(n, points, ws, g_max, cursor, sorted);
    const int qgrid = (n + 127) / 128;
    const int K = kk + skip;
    if (K <= 8) knn_query_kernel<8><<<qgrid, 128, 0, s>>>(n, points, ws, g_max, cells, sorted, kk, skip, out);
    else if (K <= 12) knn_query_kernel<12><<<qgrid, 128, 0, s>>>(n, points, ws, g_max, cells, sorted, kk, skip, out);
    else if (K <= 24) knn_query_kernel<24><<<qgrid, 128, 0, s>>>(n, points, ws, g_max, cells, sorted, kk, skip, out);
    else knn_query_kernel<48><<<qgrid, 128, 0, s>>>(n, points, ws, g_max, cells, sorted, kk, skip, out);
    return eg_check_launch("eg_knn");
}
