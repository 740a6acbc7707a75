/*
 * edgegs.h -- C ABI of the B200-native edge-Gaussian splat path (libedgegs.so).
 *
 * This is the drop-in boundary for the ONE third-party call on the reference's hot path:
 *     render, alpha, info = gsplat.rasterization(...)      /root/reference/edgegaussians/models/edge_gs.py:250-268
 * plus the reference-owned arithmetic that surrounds it (activations edge_gs.py:253-254, clamp and
 * channel select edge_gs.py:279 / train_gaussians.py:84, "whole" L1 loss edge_gs.py:290-296,
 * update_absgrads edge_gs.py:603-613, direction / ratio regularisers edge_gs.py:346-380).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked "host";
 *   - the caller owns every buffer, the library allocates nothing and keeps no state besides the
 *     thread-local error string; all launches are asynchronous on `stream` (a cudaStream_t);
 *   - return 0 = ok, non-zero = error, text via eg_last_error();
 *   - no host synchronisation anywhere: the number of tile intersections stays on the device
 *     (status[EG_ST_NISECT]); if it exceeds `isect_capacity` the flag status[EG_ST_OVERFLOW] is
 *     raised, the raster kernels become no-ops and the caller re-runs with a larger capacity.
 *
 * Layouts (fp32 unless stated), N Gaussians, T = tile_w*tile_h tiles, P = W*H pixels, cap = isect_capacity:
 *   means [N,3]  quats [N,4] (w,x,y,z; any norm)  scales [N,3]  opacities [N]
 *       raw_params = 1: scales are LOG-scales and opacities are LOGITS as the reference stores
 *       them (edge_gs.py:78-103), exp / sigmoid are fused; raw_params = 0: activated values
 *       (the gsplat signature).
 *   viewmat [16] row-major world->camera, K [9] row-major (cameras.py:84-96,129-135), on the device.
 *   rec   [N,8]  = (mean2d.x, mean2d.y, opacity*comp, depth | conic.a, conic.b, conic.c, comp)
 *                  gsplat's meta means2d/opacities/depths/conics are strided views of this record.
 *   gint  [N,2] i32 = (radius, tiles_per_gauss); radius 0 = culled.
 *   tile_counts [T, EG_CNT_STRIDE] i32 (zeroed by caller before eg_project_fwd): column 0 = number of
 *                  intersections of the tile, column 1 = append cursor (scratch of eg_bin); one 128-byte
 *                  line per tile so that the L2 atomic units do not serialise neighbouring tiles.
 *   tile_offsets [T+1] i32 (exclusive scan; [T] = n_isects).
 *   keys  [T, tile_capacity] u64 = depth_bits<<32 | gaussian_id: one fixed-capacity bucket per tile,
 *                  filled in arbitrary order by eg_project_fwd, consumed (sorted on chip) by eg_raster_fwd.
 *   flatten_ids [cap] i32 : gsplat's flatten_ids (sorted by tile, depth bits, id).
 *   isect_ids   [cap] i64 : gsplat's isect_ids (tile<<32 | depth bits), optional (may be NULL).
 *   alpha [P], render0 [P] (channel 0 of render; all three channels are equal because the reference
 *                  passes colors == 1, edge_gs.py:247), last_ids [P] i32.
 *   grad2d [N,8] = (v_mean2d.x, .y, absgrad.x, absgrad.y | v_conic.a, .b, .c, v_opacity_eff),
 *                  accumulated with atomics: zero it before eg_raster_bwd.
 *   status [EG_ST_WORDS] i32, zeroed by the caller before eg_project_fwd.
 */
#ifndef EDGEGS_H
#define EDGEGS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EG_ABI_VERSION 10
#define EG_CNT_STRIDE 32

enum {
    EG_ST_NISECT = 0,   /* number of tile intersections (sum of tiles_per_gauss)                               */
    EG_ST_OVERFLOW = 1, /* a key buffer was too small: raster / splat kernels are no-ops, caller re-runs       */
    EG_ST_BADCOLOR = 2, /* a colour differed from 1                                                            */
    EG_ST_MAXTILE = 3,  /* largest per-tile intersection count                                                 */
    EG_ST_REDO = 4,     /* tiles composited a second time in sorted order (lazy sort / splat fallback)         */
    EG_ST_STOPPED = 5,  /* tiles in which some pixel hit gsplat's transmittance stop (T * (1 - alpha) <= 1e-4) */
    EG_ST_NKEYS = 6,    /* keys emitted into the tile lists (eg_bin): == EG_ST_NISECT unless EG_FLAG_CULL_TILES   */
    EG_ST_WORDS = 8
};

enum { EG_GT_NONE = 0, EG_GT_F32 = 1, EG_GT_U8 = 2 };

/* eg_config.flags.  EG_FLAG_LAZY_SORT (fused training step only; ignored when isect_ids or last_ids are
 * requested): eg_raster_fwd composites each tile once in arbitrary order and sorts + redoes it only if some
 * pixel came near gsplat's transmittance stop threshold -- when none does, the blend result cannot depend on
 * the order.  flatten_ids is then unordered inside such tiles (cmask stays aligned with it).
 * status[EG_ST_REDO] counts the tiles that had to be redone. */
enum { EG_FLAG_LAZY_SORT = 1, EG_FLAG_COMPACT_KEYS = 2, EG_FLAG_NO_EMIT = 4, EG_FLAG_CULL_TILES = 8,
       EG_FLAG_FRONT_SORT = 16 };
/* EG_FLAG_CULL_TILES (fused training step only): eg_project_fwd / eg_bin emit a Gaussian's key only to the tiles
 * its alpha >= 1/255 footprint can reach (a conservative ellipse / tile-row test) instead of to every tile of
 * gsplat's bounding rectangle: the dropped (tile, Gaussian) pairs would fail the alpha test at every pixel, so the
 * render and the gradients are unchanged, but an elongated Gaussian's lists shrink by 2-3x.  tiles_per_gauss and
 * status[EG_ST_NISECT] stay gsplat's; tile_offsets / flatten_ids then describe the culled lists
 * (status[EG_ST_NKEYS] entries).
 * EG_FLAG_FRONT_SORT (fused training step only; ignored when isect_ids or last_ids are requested): eg_raster_fwd
 * sorts a long tile list front to back in depth slices of at most 2048 keys (histogram of the depth bits, slice by
 * slice selection) and stops as soon as every pixel of the tile has hit the transmittance stop -- the exact
 * gsplat order for everything that is composited, no work for what lies behind.  tile_done[t] = number of list
 * entries that were processed; flatten_ids / cmask are defined for those entries only. */
/* EG_FLAG_NO_EMIT: eg_project_fwd neither counts nor emits tile intersections (Gaussian-major forward); it
 * accumulates status[EG_ST_NISECT] = sum of tiles_per_gauss itself (no eg_bin in that pipeline). */
/* EG_FLAG_COMPACT_KEYS: keys is a compact [isect_capacity] array segmented by tile_offsets instead of T
 * fixed-capacity buckets (for views where a few tiles hold most intersections and T * tile_capacity keys
 * would not fit): eg_project_fwd then only counts, eg_bin scans AND emits (second pass over the Gaussians,
 * needs rec / gint).  tile_capacity is ignored. */

typedef struct eg_config {
    int32_t n;            /* number of Gaussians                                   */
    int32_t width;        /* image width  (camera.width,  edge_gs.py:239)          */
    int32_t height;       /* image height (camera.height)                          */
    int32_t tile_size;    /* must be 16 (BLOCK_WIDTH, edge_gs.py:233)              */
    float eps2d;          /* 0.3                                                   */
    float near_plane;     /* 0.01  (edge_gs.py:261)                                */
    float far_plane;      /* 1e10  (edge_gs.py:262)                                */
    float radius_clip;    /* 0.0                                                   */
    int32_t antialiased;  /* 1 = rasterize_mode "antialiased" (edge_gs.py:50)      */
    int32_t raw_params;   /* 1 = log-scales / logit-opacities, activations fused   */
    int64_t isect_capacity; /* elements available in flatten_ids / isect_ids          */
    int32_t tile_capacity;  /* keys per tile bucket (keys holds T * tile_capacity)    */
    int32_t flags;          /* EG_FLAG_*                                              */
} eg_config;

const char *eg_last_error(void);
int eg_abi_version(void);
/* tile grid for an image: replaces gsplat's tile_width / tile_height arithmetic (rendering.py) */
int eg_tile_grid(int width, int height, int tile_size, int *tile_w, int *tile_h);

/* Sizing helpers (host arithmetic, no CUDA call): everything a caller needs to allocate the buffers of the fused
 * iteration (SURVEY.md section 8b "scratch obtained via eg_workspace_bytes").
 *   eg_grad_layout        offsets5 = float offsets of v_means | v_scales | v_quats | v_opacities in the flat gradient
 *                         buffer and its total length; segment offsets use n rounded up to a multiple of 4, so every
 *                         segment is 16-byte aligned for any n (the pad floats are never written);
 *   eg_tile_capacity_for  keys per tile bucket for an intersection capacity (max_tile = largest tile load seen so
 *                         far, 0 if unknown);
 *   eg_workspace_sizes_for / eg_workspace_bytes   per-buffer / total bytes for cfg (n, width, height,
 *                         isect_capacity; tile_capacity 0 = derive it) and a pipeline EG_PIPE_*. */
enum { EG_PIPE_SPLAT = 0, EG_PIPE_TILES_SPLAT = 1, EG_PIPE_TILES = 2 };
typedef struct eg_workspace_sizes {
    size_t rec, gint, head /* status | loss_sum | tile_stop | tile_cnt */, tile_counts, stop_list, tile_offsets, keys,
        flatten_ids, cmask, logT, wpix, render0, last_depth, last_gid, grad2d, grads, total;
    int32_t tile_capacity;
    int32_t compact_keys; /* 1: EG_FLAG_COMPACT_KEYS layout (buckets would exceed 1 GiB) */
} eg_workspace_sizes;
int eg_grad_layout(int n, int64_t *offsets5);
int eg_tile_capacity_for(int64_t isect_capacity, int n_tiles, int max_tile);
int eg_workspace_sizes_for(const eg_config *cfg, int pipeline, int max_tile, eg_workspace_sizes *out);
size_t eg_workspace_bytes(const eg_config *cfg, int pipeline);

/* K1 (+ K2 pass 1): projection forward + per-tile intersection counts.
 * Replaces gsplat fully_fused_projection fwd + isect_tiles pass 1 behind edge_gs.py:250-268.
 * Also K2 emission: every Gaussian appends its (depth,id) key to the bucket of each tile it touches
 * (one atomic per intersection; no second pass, no global sort).
 * colors: [N,3] or NULL; when given it is only VERIFIED to be all-ones (status[EG_ST_BADCOLOR]). */
int eg_project_fwd(const eg_config *cfg, const float *means, const float *quats, const float *scales,
                   const float *opacities, const float *colors, const float *viewmat, const float *K,
                   float *rec, int32_t *gint, int32_t *tile_counts, uint64_t *keys, int32_t *status,
                   void *stream);

/* K2/K4: exclusive scan of the tile counts -> tile_offsets (== gsplat isect_offsets), n_isects and
 * the largest tile count, all kept on the device (status).  Replaces gsplat cumsum + the n_isects
 * D2H sync + isect_offset_encode (keys were already emitted into the tile buckets by eg_project_fwd;
 * the sort itself is per tile inside eg_raster_fwd). */
int eg_bin(const eg_config *cfg, int32_t *tile_counts, int32_t *tile_offsets, int32_t *status,
           const float *rec, const int32_t *gint, uint64_t *keys, void *stream);

/* K3 + K5 (+ a8 "whole" L1): per-tile sort, front-to-back compositing, optional fused edge-map loss.
 * Replaces cub radix sort + gsplat rasterize_to_pixels fwd; with gt != NULL also
 * clamp/select/L1 (edge_gs.py:279,290-296; train_gaussians.py:84-94):
 *   loss_sum[0] += sum_p |clamp(render0) - gt|      (one fp64 word; caller divides by P)
 *   wpix[p]      = sign(clamp(render0) - gt) * T_final   (the backward seed of pixel p, unscaled)
 * gt_kind: EG_GT_F32 (values in [0,1]) or EG_GT_U8 (value/255, train_gaussians.py:87).
 * cmask [cap,8] u32: per tile intersection (in flatten_ids order) the 256-bit mask of the tile's pixels
 * that composited that Gaussian; word w covers the 8x4 pixel block (x0 = 8*(w&1), y0 = 4*(w>>1)),
 * bit l the pixel (x0 + (l&7), y0 + (l>>3)).  It is the backward kernel's work list.
 * last_depth [P] u32 / last_gid [P] i32 (both or neither): sort key (depth bits, Gaussian id) of the last
 * Gaussian composited by each pixel that hit the transmittance stop, (0xffffffff, -1) elsewhere -- the
 * per-pixel cut-off eg_splat_bwd needs; status[EG_ST_STOPPED] counts the tiles that contain such pixels.
 * stop_list [T] i32 / tile_cnt [T] i32 (both or neither): fallback mode of the Gaussian-major forward -- only the
 * status[EG_ST_STOPPED] tiles listed in stop_list (eg_splat_resolve) are processed by a small persistent grid,
 * their keys are the first tile_cnt[t] entries of bucket t (eg_emit_flagged), always sorted; tile_offsets is then
 * unused, flatten_ids needs T * tile_capacity entries, last_depth / last_gid are required.
 * tile_done [T] i32 (optional): number of list entries of each tile that were processed -- all of them unless
 * EG_FLAG_FRONT_SORT stopped early; flatten_ids / cmask are defined for those entries only (eg_raster_bwd takes it).
 * loss_params (DEVICE, 4 floats: w_edge, w_bg, w_sel, threshold) / sel_mask [P] u8, both optional, need gt: the
 * per-pixel coefficient c_p of the fused loss, loss_sum[0] += sum_p c_p |clamp(render0) - gt|, wpix[p] *= c_p:
 *   c_p = (gt_p >= threshold ? w_edge : w_bg) + (sel_mask[p] ? w_sel : 0);   NULL loss_params: c_p = 1.
 * That covers the reference's three strategies (edge_gs.py:288-324): "whole" (NULL), "weighted" (w_edge = n_bg / P,
 * w_bg = n_edge / P, loss = loss_sum / P), "bg_edge_ratio" (w_edge = 1 / n_edge, w_bg = 0, sel_mask = the sampled
 * pixels, w_sel = 1 / n_sel, loss = loss_sum); the scalars live on the device so a captured graph can be re-used
 * when they change.
 * Any of render0 / alpha / last_ids / isect_ids / cmask / gt / loss_sum / wpix / last_* / tile_* may be NULL. */
int eg_raster_fwd(const eg_config *cfg, const float *rec, const int32_t *tile_offsets, uint64_t *keys,
                  int32_t *flatten_ids, int64_t *isect_ids, float *render0, float *alpha,
                  int32_t *last_ids, uint32_t *cmask, const void *gt, int gt_kind, double *loss_sum,
                  float *wpix, uint32_t *last_depth, int32_t *last_gid, const int32_t *stop_list,
                  const int32_t *tile_cnt, int32_t *tile_done, const float *loss_params, const uint8_t *sel_mask,
                  int32_t *status, void *stream);

/* K6: compositing backward with abs-grad.  Replaces gsplat rasterize_to_pixels bwd.
 * The seed of pixel p is  w_p = (sum_ch v_render[p,ch] + v_alpha[p]) * (1 - alpha[p])  when
 * v_render/v_alpha/alpha are given (generic autograd path; v_render has `v_render_channels`
 * interleaved channels), or  w_p = seed_scale * wpix[p]  (fused-loss path).  tile_done [T] (optional, written by
 * eg_raster_fwd): only the first tile_done[t] entries of tile t's list are walked. */
int eg_raster_bwd(const eg_config *cfg, const float *rec, const int32_t *tile_offsets,
                  const int32_t *flatten_ids, const uint32_t *cmask, const int32_t *tile_done, const float *alpha,
                  const float *v_render, int v_render_channels, const float *v_alpha,
                  const float *wpix, float seed_scale, float *grad2d, const int32_t *status,
                  void *stream);

/* K7 (+ activation VJPs + a9): projection backward.  Replaces gsplat fully_fused_projection bwd,
 * the opacity*compensation VJP, Exp/Sigmoid backward and update_absgrads (edge_gs.py:603-613).
 * grads are WRITTEN (not accumulated): v_means [N,3], v_quats [N,4], v_scales [N,3],
 * v_opacities [N] w.r.t. the inputs as given (raw or activated per cfg->raw_params).
 * v_depths may be NULL.  absgrad_accum [N] (may be NULL) += ||absgrad||_2 .
 * zero_grad2d != 0: grad2d is cleared after it has been consumed (ready for the next iteration). */
int eg_project_bwd(const eg_config *cfg, const float *means, const float *quats, const float *scales,
                   const float *opacities, const float *viewmat, const float *K, const float *rec,
                   const int32_t *gint, float *grad2d, int zero_grad2d, const float *v_depths,
                   float *v_means, float *v_quats, float *v_scales, float *v_opacities,
                   float *absgrad_accum, void *stream);

/* Gaussian-major forward of the fused training step (replaces isect_tiles + radix sort + rasterize_to_pixels fwd
 * behind edge_gs.py:250-268 wherever gsplat's transmittance stop cannot trigger; exact fallback otherwise):
 *   eg_splat_fwd      logT [P] fp32 (zero on entry) += log2(1 - alpha) of every (pixel, Gaussian) pair that passes
 *                     gsplat's tile-rectangle / sigma / alpha tests (red.global.add.v4.f32, no tile lists);
 *   eg_splat_resolve  per 16x16 tile: T = 2^logT, render = alpha = 1 - T, fused clamp + "whole" L1 loss + backward
 *                     seed exactly as eg_raster_fwd (loss_sum, wpix, loss_params, sel_mask), logT re-zeroed.  A tile in which some
 *                     pixel's T is within 0.1 % of gsplat's stop threshold 1e-4 is NOT resolved: tile_stop[t] = 1,
 *                     its id is appended to stop_list [T] and status[EG_ST_STOPPED] += 1 (tile_stop zeroed by
 *                     the caller);
 *   eg_emit_flagged   appends the keys of the Gaussians touching flagged tiles to keys [T, tile_capacity]
 *                     (cursor tile_cnt [T] i32, zeroed by the caller); then eg_raster_fwd(tile_stop, tile_cnt)
 *                     redoes exactly those tiles in sorted order.  Both return at once when nothing is flagged. */
int eg_splat_fwd(const eg_config *cfg, const float *rec, const int32_t *gint, float *logT, const int32_t *status,
                 void *stream);
int eg_splat_resolve(const eg_config *cfg, float *logT, const void *gt, int gt_kind, double *loss_sum, float *wpix,
                     float *render0, float *alpha, int32_t *tile_stop, int32_t *stop_list, const float *loss_params,
                     const uint8_t *sel_mask, int32_t *status, void *stream);
int eg_emit_flagged(const eg_config *cfg, const float *rec, const int32_t *gint, const int32_t *tile_stop,
                    int32_t *tile_cnt, uint64_t *keys, int32_t *status, void *stream);

/* K6 + K7 fused, Gaussian-major (the fused training step's backward): replaces gsplat rasterize_to_pixels bwd
 * + fully_fused_projection bwd + the opacity*compensation VJP + Exp/Sigmoid backward + update_absgrads
 * (edge_gs.py:250-268, 603-613) in one kernel without tile lists or atomics.  Each warp owns 32 Gaussians, walks
 * the pixel rows of their alpha >= 1/255 footprints (clipped to gsplat's tile rectangles) and re-applies the
 * forward's exact per-pixel test, so the set of (pixel, Gaussian) pairs is the one the forward composited.
 *   wpix [P]        per-pixel seed: seed_scale * wpix[p] = dL/d render(p) * T_final(p)   (eg_raster_fwd / eg_make_seed)
 *   last_depth [P] u32, last_gid [P] i32 (both or neither): sort key (depth bits, id) of the last Gaussian each
 *                   pixel composited, last_depth = 0xffffffff where the pixel never stopped; only read when
 *                   status[EG_ST_STOPPED] != 0, and, when tile_stop [T] i32 is given (eg_splat_resolve), only
 *                   inside the tiles it flags (the planes are undefined elsewhere).
 *   g_begin, g_end  only the Gaussians [g_begin, g_end) are processed (g_end < 0: all) -- every Gaussian has one
 *                   owner, so the gradients of a range are final as soon as its launch has drained.
 *   grad2d_out [N,8] optional: the 2D gradients (layout of grad2d above), WRITTEN.
 * Gradient outputs are WRITTEN (layout as eg_project_bwd); absgrad_accum [N] (may be NULL) += ||absgrad||_2. */
int eg_splat_bwd(const eg_config *cfg, const float *means, const float *quats, const float *scales,
                 const float *opacities, const float *viewmat, const float *K, const float *rec,
                 const int32_t *gint, const float *wpix, float seed_scale, const uint32_t *last_depth,
                 const int32_t *last_gid, const int32_t *tile_stop, const int32_t *status, int g_begin, int g_end,
                 float *grad2d_out, float *v_means,
                 float *v_quats, float *v_scales, float *v_opacities, float *absgrad_accum, void *stream);

/* View-sharded multi-GPU step (no reference counterpart: the reference is single-GPU, train_gaussians.py:311; the
 * sharding is SURVEY.md section 8e's: parameters replicated, one view per GPU per step, per-view gradients summed).
 *
 * eg_allreduce_symm -- the exchange as ONE kernel of this library over NVLink / NVSwitch: in-place fp32 sum of
 *   `count` floats (a multiple of 4; eg_grad_layout pads to that) that every rank holds at the same offset of a
 *   SYMMETRIC allocation.  peer_bufs [world] (HOST array of device pointers): the buffer as mapped into this
 *   process for every rank (peer_bufs[rank] = the local one).  mc_buf: the same buffer through an NVSwitch multicast
 *   mapping, or NULL -- with it each rank reduces its 1/world slice with multimem.ld_reduce (summed inside the
 *   switch) and broadcasts it with multimem.st; without it, 128-bit peer loads / stores.  peer_flags [world] (HOST
 *   array): per rank a zero-initialised symmetric area of eg_allreduce_flag_words(grid) u32 used by the in-kernel
 *   rank barriers (self-resetting: no host work between calls; capturable in a CUDA graph).  Every rank must
 *   enqueue the call with the same count / grid; `grid` <= 0 picks the default.  At most 8 ranks (one NVSwitch domain).
 * eg_comm_* -- the same sum through NCCL (resolved at run time with dlopen, no link dependency), kept as the A/B
 *   baseline: eg_comm_unique_id on rank 0 (128-byte id in host memory, broadcast by the caller), eg_comm_init
 *   collectively on the calling thread's current device, eg_comm_allreduce in place on `stream`. */
int eg_allreduce_flag_words(int grid);
int eg_allreduce_symm(float *const *peer_bufs, float *mc_buf, uint32_t *const *peer_flags, int64_t count, int rank,
                      int world, int grid, void *stream);
/* the same over 1..4 segments (offset, count in floats, multiples of 4; HOST arrays) of the buffer: the slices of
 * means | scales | quats | opacities that hold one Gaussian range -- eg_splat_bwd finishes its gradients range by
 * range, so the exchange of a finished range can run on a side stream while the next range is computed. */
int eg_allreduce_symm_segs(float *const *peer_bufs, float *mc_buf, uint32_t *const *peer_flags, int n_segs,
                           const int64_t *seg_offsets, const int64_t *seg_counts, int rank, int world, int grid,
                           void *stream);
/* Push form of the exchange -- the backward's gradient stores ARE the reduce-scatter ("K7
 * writing into the all-reduce buffer"; no reference counterpart).  Rank o owns the Gaussians [o * per, (o + 1) * per),
 * per = eg_exchange_push_per(n, world).  Every rank holds a SYMMETRIC staging area of eg_exchange_stage_floats(n, world)
 * floats (zero-initialised once): `world` slots, slot s = what rank s computed for the owned range, laid out
 *     means [3 per] | scales [3 per] | quats [4 per] | opacities [per].
 *   eg_splat_bwd_push / eg_project_bwd_push   the backward kernels of eg_splat_bwd / eg_project_bwd with their four
 *       gradient outputs redirected: the gradients of Gaussian g go to slot `rank` of stage[g / per] (peer stores over
 *       NVLink, issued as the Gaussians finish -- the transfer overlaps the backward itself);
 *   eg_exchange_reduce_bcast   one kernel behind it on the same stream: rank barrier, the owner sums its `world` local
 *       slots in rank order and broadcasts the result into the flat gradient buffer (eg_grad_layout) of EVERY rank
 *       (multimem.st through mc_buf when given, else peer stores), rank barrier.  peer_bufs / mc_buf / peer_flags /
 *       grid as eg_allreduce_symm.  Every rank receives bit-identical sums;
 *   eg_exchange_push_zero      a rank without a view (ragged last step) contributes zeros instead of a backward.
 * All calls are asynchronous on `stream`, self-resetting and CUDA-graph capturable. */
typedef struct eg_push_target {
    float *stage[8]; /* by owner rank: that rank's staging area as mapped into this process (stage[rank] = local) */
    int32_t per;     /* Gaussians per owner, a multiple of 128 */
    int32_t rank;    /* this rank: the slot it writes in every staging area */
    int32_t world;
    int32_t reserved;
} eg_push_target;
int eg_exchange_push_per(int n, int world);
int64_t eg_exchange_stage_floats(int n, int world);
int eg_splat_bwd_push(const eg_config *cfg, const float *means, const float *quats, const float *scales,
                      const float *opacities, const float *viewmat, const float *K, const float *rec,
                      const int32_t *gint, const float *wpix, float seed_scale, const uint32_t *last_depth,
                      const int32_t *last_gid, const int32_t *tile_stop, const int32_t *status, int g_begin, int g_end,
                      const eg_push_target *push, float *absgrad_accum, void *stream);
int eg_project_bwd_push(const eg_config *cfg, const float *means, const float *quats, const float *scales,
                        const float *opacities, const float *viewmat, const float *K, const float *rec,
                        const int32_t *gint, float *grad2d, int zero_grad2d, const eg_push_target *push,
                        float *absgrad_accum, void *stream);
int eg_exchange_reduce_bcast(const eg_push_target *push, float *const *peer_bufs, float *mc_buf,
                             uint32_t *const *peer_flags, int n, int grid, void *stream);
int eg_exchange_push_zero(const eg_push_target *push, int n, void *stream);
int eg_comm_unique_id(void *id128);
int eg_comm_init(const void *id128, int rank, int world, void **comm_out);
int eg_comm_destroy(void *comm);
int eg_comm_allreduce(float *buf, int64_t count, void *comm, void *stream);

/* Per-pixel seed of the gsplat-shaped autograd path:
 *   wpix[p] = (sum_ch v_render[p,ch] + v_alpha[p]) * (1 - alpha[p])      (v_render / v_alpha may be NULL) */
int eg_make_seed(int64_t n_pixels, const float *alpha, const float *v_render, int v_render_channels,
                 const float *v_alpha, float *wpix, void *stream);

/* a10 + a11: edge-direction and anisotropy regularisers, forward + backward in one pass.
 * Replaces compute_direction_loss / compute_ratio_loss + autograd (edge_gs.py:346-380).
 * nn_indices [N,nn_cols] i32 (nn_cols = k, or 2k for enforce_half); losses (2 fp64 words, accumulated):
 * losses[0] += sum_i mean_n |m.d|
 * (direction loss = 1 - losses[0]/N), losses[1] = sum_i ratio_i (ratio loss = losses[1]/N).
 * Gradients are ACCUMULATED (atomics): dir_weight * dL_dir and ratio_weight * dL_ratio into
 * v_means [N,3], v_quats [N,4], v_log_scales [N,3]. */
int eg_reg_fwd_bwd(int n, const float *means, const float *quats, const float *log_scales,
                   const int32_t *nn_indices, int nn_cols, int k, int enforce_half, float dir_weight,
                   float ratio_weight, double *losses, float *v_means, float *v_quats,
                   float *v_log_scales, void *stream);

/* a12 ("next"): exact k-nearest neighbours on the device.  Replaces
 * k_nearest_sklearn / update_nearest_neighbors (edge_gs.py:135-151, 326-344).  points [N,3] fp32;
 * out [N,kk] i32 = neighbours of rank skip .. skip+kk-1 by (distance, index), rank 0 being the point
 * itself; the reference's "ask for kk+2, drop the first column twice" is skip = 2.  kk + skip <= 48.
 * workspace: eg_knn_workspace_bytes(n) bytes of device memory (host helper, no CUDA call). */
size_t eg_knn_workspace_bytes(int n);
int eg_knn(int n, const float *points, int kk, int skip, int32_t *out, void *workspace,
           size_t workspace_bytes, void *stream);

/* "next", SURVEY.md section 8f-4: batched visibility filter.  Replaces the per-view CPU loop of
 * cull_gaussians_not_projecting (edge_gs.py:578-601) up to its threshold: fraction [N] fp32 = share of the
 * n_views views in which the Gaussian's mean projects (P = K @ viewmat[:3,:4], round half to even, no depth test --
 * the reference's arithmetic) inside the image and onto a non-zero pixel of that view's edge mask.
 * viewmats [V,16], Ks [V,9], sizes [V,2] i32 = (width, height), masks u8 = the views' [H,W] masks back to back,
 * mask_offsets [V] i64 = start of each view's mask in `masks`.  At most 1024 views per call.
 * mode 0: the hit fraction above.  mode 1: `masks` holds the uint8 EDGE MAPS and fraction = SUM over the views of
 * the edge value at the projected pixel (0 outside the image; an integer, exact in fp32) -- divided by 255 V it is the
 * statistic of the post-processing filter filter_by_projection (edge_extraction/filtering.py:80-123). */
int eg_projecting_fraction(int n, const float *means, int n_views, const float *viewmats, const float *Ks,
                           const int32_t *sizes, const uint8_t *masks, const int64_t *mask_offsets, int mode,
                           float *fraction, void *stream);

/* a13 ("next"): fused Adam update of one parameter tensor, torch.optim.Adam
 * semantics as configured at utils/train_utils.py:48-65 (no weight decay, no amsgrad);
 * bias_correction{1,2} = 1 - beta{1,2}^t.  Hyper-parameters are doubles (as Python floats are) and are
 * rounded to fp32 only after 1 - beta etc. have been formed, exactly like torch does.
 * zero_grad != 0 also clears the gradient. */
int eg_adam_step(int64_t n, float *param, float *grad, float *exp_avg, float *exp_avg_sq, double lr,
                 double beta1, double beta2, double eps, double bias_correction1, double bias_correction2,
                 int zero_grad, void *stream);

/* a13 / section 8f-3: the reference's four Adams as ONE launch over the fused step's flat gradient buffer.
 * segs [n_segs <= 8] (HOST structs): parameter tensor, its torch-layout moment tensors, where its gradients start in
 * `grads` (eg_grad_layout) and its element count.  hyper [n_segs] (DEVICE, 24 bytes each: f64 lr | i64 completed
 * steps | i64 enabled): what changes between steps lives on the device, so the launch can sit in a captured CUDA
 * graph while schedulers rewrite the learning rates (train_utils.py:15-37) and the regulariser steps disable the
 * opacity segment (train_gaussians.py:118-121); the kernel advances the step counts of the enabled segments itself.
 * ticket: one zero-initialised u32 of device scratch.  zero_grad != 0 also clears the consumed gradients.
 * Same arithmetic as eg_adam_step / torch.optim.Adam (bias corrections formed in double). */
typedef struct eg_adam_segment {
    float *param, *exp_avg, *exp_avg_sq;
    int64_t grad_offset, count;
} eg_adam_segment;
int eg_adam_multi(int n_segs, const eg_adam_segment *segs, float *grads, void *hyper, double beta1, double beta2,
                  double eps, int zero_grad, uint32_t *ticket, void *stream);

/* Section 8f-3: densify / cull compaction as one kernel.  For every array a (at most 16; HOST structs):
 *   dst[a][r, 0:width] = r < zero_from_row ? src[a][idx[r], 0:width] : 0      for r in [0, n_out)
 * (zero_from_row < 0: never zero).  idx [n_out] i32 on the device.  A cull passes the surviving row ids; a
 * duplication passes the old ids followed by the duplicated ones with zero_from_row = old N for the Adam moments
 * (edge_gs.py:384-475: parameters, exp_avg, exp_avg_sq and the abs-grad statistic move together). */
typedef struct eg_row_array {
    const float *src;
    float *dst;
    int32_t width;
    int64_t zero_from_row;
} eg_row_array;
int eg_gather_rows(int64_t n_out, const int32_t *idx, int n_arrays, const eg_row_array *arrays, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EDGEGS_H */
