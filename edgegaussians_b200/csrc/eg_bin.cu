// eg_bin.cu -- K2/K4: tile binning bookkeeping without a host sync and without a global sort.
//
// gsplat 1.0.0 (behind /root/reference/edgegaussians/models/edge_gs.py:250-268) does
//   cumsum(tiles_per_gauss) -> D2H sync for n_isects -> emit (cam|tile|depth) keys in Gaussian order
//   -> 6-pass cub radix sort over ALL intersections -> isect_offset_encode.
// Here eg_project_fwd has already appended every (depth_bits<<32 | id) key to a fixed-capacity bucket
// of its tile (one atomic per intersection), so what is left is
//   scan_kernel : exclusive scan of the T tile counts -> tile_offsets (== gsplat isect_offsets),
//                 the number of emitted keys and the largest tile count, kept on the device (status words;
//                 gsplat's n_isects = sum of tiles_per_gauss is accumulated by eg_project_fwd);
// and the sort happens per tile on chip inside eg_raster_fwd.  The key is unique, so the result is
// deterministic and equals gsplat's stable sort order (tile, depth bits, Gaussian id).
#include "eg_common.cuh"

namespace {

constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(int32_t *__restrict__ counts,
                                                           int32_t *__restrict__ offsets, int T,
                                                           int32_t *__restrict__ status, long long capacity,
                                                           int tile_capacity, int compact) {
    __shared__ int32_t warp_sums[SCAN_THREADS / 32];
    __shared__ int32_t warp_max[SCAN_THREADS / 32];
    __shared__ int32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    int32_t vmax = 0;
    __syncthreads();
    constexpr int PER = 8;  // consecutive tiles per thread: 8 independent loads in flight, one block scan per 8192 tiles
    for (int base = 0; base < T; base += SCAN_THREADS * PER) {
        const int i0 = base + tid * PER;
        int32_t vals[PER];
        int32_t v = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            vals[q] = (i0 + q < T) ? counts[(size_t)(i0 + q) * EG_CNT_STRIDE] : 0;
            vmax = max(vmax, vals[q]);
            v += vals[q];
        }
        int32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int32_t w = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t y = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int32_t carry = carry_s;
        const int32_t incl = x + (wid > 0 ? warp_sums[wid - 1] : 0) + carry;
        int32_t run = incl - v;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            if (i0 + q < T) {
                offsets[i0 + q] = run;
                if (compact) counts[(size_t)(i0 + q) * EG_CNT_STRIDE + 1] = run;  // append cursor of the tile
            }
            run += vals[q];
        }
        __syncthreads();
        if (tid == SCAN_THREADS - 1) carry_s = incl;
        __syncthreads();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
    if (lane == 0) warp_max[wid] = vmax;
    __syncthreads();
    if (tid == 0) {
        int32_t m = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) m = max(m, warp_max[w]);
        const int32_t total = carry_s;
        offsets[T] = total;
        status[EG_ST_NKEYS] = total;  // emitted keys (== n_isects unless EG_FLAG_CULL_TILES dropped unreachable tiles)
        status[EG_ST_MAXTILE] = m;
        if ((long long)total > capacity || (!compact && m > tile_capacity)) status[EG_ST_OVERFLOW] = 1;
    }
}

// EG_FLAG_COMPACT_KEYS: second pass over the Gaussians, appending keys into the compact per-tile segments
__global__ void __launch_bounds__(256) emit_kernel(int n, const float4 *__restrict__ rec,
                                                   const int2 *__restrict__ gint, int32_t *__restrict__ counts,
                                                   unsigned long long *__restrict__ keys, long long capacity,
                                                   int tw, int th, int cull) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const int2 gi = __ldg(gint + g);
    if (gi.y <= 0) return;
    const float4 r0 = __ldg(rec + 2 * g);
    uint32_t x0, y0, x1, y1;
    eg_tile_rect(r0.x, r0.y, gi.x, tw, th, x0, y0, x1, y1);
    const unsigned long long key = ((unsigned long long)__float_as_uint(r0.w) << 32) | (unsigned int)g;
    float hu = 1e30f, hv = 1e30f, tau = 0.0f;
    float4 r1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cull) {  // the same tile set eg_project_fwd counted (EG_FLAG_CULL_TILES)
        r1 = __ldg(rec + 2 * g + 1);
        if (!eg_extent(r0.z, r1.x, r1.y, r1.z, hu, hv, tau)) return;
    }
    for (uint32_t i = y0; i < y1; ++i) {
        int j0 = (int)x0, j1 = (int)x1 - 1;
        if (cull && !eg_tile_row_cols(r0.x, r0.y, r1.x, r1.y, r1.z, tau, hu, hv, (int)i, (int)x0, (int)x1, j0, j1)) continue;
        for (int j = j0; j <= j1; ++j) {
            const long long pos = atomicAdd(counts + (size_t)(i * tw + j) * EG_CNT_STRIDE + 1, 1);
            if (pos < capacity) keys[pos] = key;
        }
    }
}

}  // namespace

extern "C" int eg_bin(const eg_config *cfg, int32_t *tile_counts, int32_t *tile_offsets, int32_t *status,
                      const float *rec, const int32_t *gint, uint64_t *keys, void *stream) {
    if (cfg == nullptr || cfg->tile_size != EG_TILE) {
        eg_set_error("eg_bin: tile_size must be %d", EG_TILE);
        return 1;
    }
    int tw, th;
    eg_tile_grid(cfg->width, cfg->height, cfg->tile_size, &tw, &th);
    const int compact = (cfg->flags & EG_FLAG_COMPACT_KEYS) ? 1 : 0;
    scan_kernel<<<1, SCAN_THREADS, 0, (cudaStream_t)stream>>>(tile_counts, tile_offsets, tw * th, status,
                                                             (long long)cfg->isect_capacity, cfg->tile_capacity,
                                                             compact);
    if (int e = eg_check_launch("eg_bin/scan")) return e;
    if (compact && cfg->n > 0) {
        if (rec == nullptr || gint == nullptr || keys == nullptr) {
            eg_set_error("eg_bin: EG_FLAG_COMPACT_KEYS needs rec, gint and keys");
            return 1;
        }
        emit_kernel<<<(cfg->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
            cfg->n, (const float4 *)rec, (const int2 *)gint, tile_counts, (unsigned long long *)keys,
            (long long)cfg->isect_capacity, tw, th, (cfg->flags & EG_FLAG_CULL_TILES) ? 1 : 0);
        return eg_check_launch("eg_bin/emit");
    }
    return 0;
}
