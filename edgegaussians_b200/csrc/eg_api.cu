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
