// eg_bin.cu -- K2 pass 2: tile binning without a host sync and without a global sort.
//
// gsplat 1.0.0 (behind /root/reference/edgegaussians/models/edge_gs.py:250-268) does
//   cumsum(tiles_per_gauss) -> D2H sync for n_isects -> emit (cam|tile|depth) keys in Gaussian order
//   -> 6-pass cub radix sort over ALL intersections -> isect_offset_encode.
// Here the per-tile counts already exist (eg_project_fwd), so:
//   scan_kernel : exclusive scan of the T tile counts -> tile_offsets (== gsplat isect_offsets) and
//                 n_isects, kept on the device (status[EG_ST_NISECT]);
//   emit_kernel : every Gaussian appends (depth_bits<<32 | id) to each tile segment it touches through
//                 a per-tile cursor (order inside a segment is arbitrary at this point);
// and the sort happens per tile in shared memory inside eg_raster_fwd.  The key is unique, so the
// result is deterministic and equals gsplat's stable sort order (tile, depth bits, Gaussian id).
// Algorithmic traffic: 12 B/Gaussian read (+ rect recompute) and 8 B/intersection written once.
#include "eg_common.cuh"

namespace {

constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(int32_t *__restrict__ counts,
                                                           int32_t *__restrict__ offsets, int T,
                                                           int32_t *__restrict__ status, long long capacity) {
    __shared__ int32_t warp_sums[SCAN_THREADS / 32];
    __shared__ int32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < T; base += SCAN_THREADS) {
        const int i = base + tid;
        const int32_t v = i < T ? counts[(size_t)i * EG_CNT_STRIDE] : 0;
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
        if (i < T) {
            offsets[i] = incl - v;
            counts[(size_t)i * EG_CNT_STRIDE + 1] = incl - v;  // append cursor of the tile
        }
        __syncthreads();
        if (tid == SCAN_THREADS - 1) carry_s = incl;
        __syncthreads();
    }
    if (tid == 0) {
        const int32_t total = carry_s;
        offsets[T] = total;
        status[EG_ST_NISECT] = total;
        if ((long long)total > capacity) status[EG_ST_OVERFLOW] = 1;
    }
}

__global__ void __launch_bounds__(256) emit_kernel(int n, const float4 *__restrict__ rec,
                                                   const int2 *__restrict__ gint, int32_t *__restrict__ counts,
                                                   unsigned long long *__restrict__ keys, long long capacity,
                                                   int tw, int th) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const int2 gi = __ldg(gint + g);
    if (gi.y <= 0) return;
    const float4 r0 = __ldg(rec + 2 * g);
    uint32_t x0, y0, x1, y1;
    eg_tile_rect(r0.x, r0.y, gi.x, tw, th, x0, y0, x1, y1);
    const unsigned long long key = ((unsigned long long)__float_as_uint(r0.w) << 32) | (unsigned int)g;
    for (uint32_t i = y0; i < y1; ++i)
        for (uint32_t j = x0; j < x1; ++j) {
            const long long pos = atomicAdd(counts + (size_t)(i * tw + j) * EG_CNT_STRIDE + 1, 1);
            if (pos < capacity) keys[pos] = key;
        }
}

}  // namespace

extern "C" int eg_bin(const eg_config *cfg, const float *rec, const int32_t *gint, int32_t *tile_counts,
                      int32_t *tile_offsets, uint64_t *keys, int32_t *status, void *stream) {
    if (cfg == nullptr || cfg->tile_size != EG_TILE) {
        eg_set_error("eg_bin: tile_size must be %d", EG_TILE);
        return 1;
    }
    int tw, th;
    eg_tile_grid(cfg->width, cfg->height, cfg->tile_size, &tw, &th);
    cudaStream_t s = (cudaStream_t)stream;
    scan_kernel<<<1, SCAN_THREADS, 0, s>>>(tile_counts, tile_offsets, tw * th, status,
                                           (long long)cfg->isect_capacity);
    if (int e = eg_check_launch("eg_bin/scan")) return e;
    if (cfg->n > 0) {
        emit_kernel<<<(cfg->n + 255) / 256, 256, 0, s>>>(cfg->n, (const float4 *)rec, (const int2 *)gint, tile_counts,
                                                         (unsigned long long *)keys, (long long)cfg->isect_capacity,
                                                         tw, th);
        if (int e = eg_check_launch("eg_bin/emit")) return e;
    }
    return 0;
}
