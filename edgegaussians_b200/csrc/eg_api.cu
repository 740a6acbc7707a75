// eg_api.cu -- error reporting and small host helpers of the C ABI (include/edgegs.h).
#include <cstdarg>
#include <cstdio>

#include "eg_common.cuh"

static thread_local char g_err[512] = "";

void eg_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int eg_check_launch(const char *what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        eg_set_error("%s: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

extern "C" const char *eg_last_error(void) { return g_err; }

extern "C" int eg_abi_version(void) { return EG_ABI_VERSION; }

extern "C" int eg_tile_grid(int width, int height, int tile_size, int *tile_w, int *tile_h) {
    if (tile_size <= 0 || width <= 0 || height <= 0) {
        eg_set_error("eg_tile_grid: bad arguments (%d, %d, %d)", width, height, tile_size);
        return 1;
    }
    if (tile_w) *tile_w = (width + tile_size - 1) / tile_size;
    if (tile_h) *tile_h = (height + tile_size - 1) / tile_size;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Sizing helpers (host arithmetic only): a caller that is not this package's Python can size every buffer of
// the fused iteration from C.
// ---------------------------------------------------------------------------------------------------------
extern "C" int eg_grad_layout(int n, int64_t *offsets5) {
    if (n < 0 || offsets5 == nullptr) {
        eg_set_error("eg_grad_layout: bad arguments");
        return 1;
    }
    const int64_t p = ((int64_t)n + 3) / 4 * 4;  // every segment starts on a 16-byte boundary for any n
    offsets5[0] = 0;
    offsets5[1] = 3 * p;
    offsets5[2] = 6 * p;
    offsets5[3] = 10 * p;
    offsets5[4] = 11 * p;
    return 0;
}

extern "C" int eg_tile_capacity_for(int64_t isect_capacity, int n_tiles, int max_tile) {
    // keys per tile bucket: 4x the mean tile load implied by the intersection capacity, at least 1.25x the
    // largest tile seen so far, at least 256, rounded up to a multiple of 64
    int64_t want = 256;
    const int64_t mean4 = 4 * isect_capacity / (n_tiles > 0 ? n_tiles : 1);
    if (mean4 > want) want = mean4;
    const int64_t seen = (int64_t)(max_tile * 1.25) + 1;
    if (seen > want) want = seen;
    want = (want + 63) / 64 * 64;
    return want > 0x7fffffc0 ? 0x7fffffc0 : (int)want;
}

extern "C" int eg_workspace_sizes_for(const eg_config *cfg, int pipeline, int max_tile, eg_workspace_sizes *out) {
    if (cfg == nullptr || out == nullptr || cfg->n < 0 || cfg->width <= 0 || cfg->height <= 0 ||
        cfg->tile_size != EG_TILE || pipeline < EG_PIPE_SPLAT || pipeline > EG_PIPE_TILES) {
        eg_set_error("eg_workspace_sizes_for: bad arguments");
        return 1;
    }
    int tw, th;
    eg_tile_grid(cfg->width, cfg->height, cfg->tile_size, &tw, &th);
    const size_t T = (size_t)tw * th, P = (size_t)cfg->width * cfg->height, N = (size_t)cfg->n;
    const size_t cap = (size_t)(cfg->isect_capacity > 0 ? cfg->isect_capacity : 0);
    eg_workspace_sizes s = {};
    s.tile_capacity = cfg->tile_capacity > 0 ? cfg->tile_capacity : eg_tile_capacity_for((int64_t)cap, (int)T, max_tile);
    s.compact_keys = T * (size_t)s.tile_capacity * 8 > ((size_t)1 << 30);  // buckets above 1 GiB: compact segments
    int64_t lay[5];
    eg_grad_layout(cfg->n, lay);
    s.rec = N * 32;
    s.gint = N * 8;
    s.head = (EG_ST_WORDS + 2 + 2 * T) * 4;  // status | loss accumulator (f64) | tile_stop | tile_cnt
    s.tile_counts = T * EG_CNT_STRIDE * 4;
    s.stop_list = T * 4;
    s.tile_offsets = (T + 1) * 4;
    s.wpix = P * 4;
    s.render0 = P * 4;
    s.last_depth = P * 4;
    s.last_gid = P * 4;
    s.grads = (size_t)lay[4] * 4;
    const size_t bucket_keys = s.compact_keys ? cap : T * (size_t)s.tile_capacity;
    if (pipeline == EG_PIPE_SPLAT) {
        s.logT = P * 4;
        // buckets / sorted ids of the tiles redone by the exact stop-rule fallback (any tile may be flagged)
        s.keys = T * (size_t)s.tile_capacity * 8;
        s.flatten_ids = T * (size_t)s.tile_capacity * 4;
    } else {
        s.keys = bucket_keys * 8;
        s.flatten_ids = (cap > bucket_keys ? cap : bucket_keys) * 4;
        if (pipeline == EG_PIPE_TILES) {
            s.cmask = cap * 32;
            s.grad2d = N * 32;
        }
    }
    s.total = s.rec + s.gint + s.head + s.tile_counts + s.stop_list + s.tile_offsets + s.keys + s.flatten_ids +
              s.cmask + s.logT + s.wpix + s.render0 + s.last_depth + s.last_gid + s.grad2d + s.grads;
    *out = s;
    return 0;
}

extern "C" size_t eg_workspace_bytes(const eg_config *cfg, int pipeline) {
    eg_workspace_sizes s;
    if (eg_workspace_sizes_for(cfg, pipeline, 0, &s)) return 0;
    return s.total;
}
