// eg_raster_fwd.cu -- K3 + K5 (+ fused "whole" L1 edge-map loss): one CTA per 16x16 tile.
//
//   phase A  sort the tile's (depth_bits<<32 | id) keys.  Up to 2048 keys live in REGISTERS (1..8 per
//            thread): bitonic exchanges at distance < 32 are warp shuffles, distances 32..128 go
//            through a double-buffered shared-memory exchange (one barrier each), distances >= 256
//            are register-to-register inside the thread.  Longer segments (rare) fall back to a
//            chunked shared/global bitonic network.  gsplat's flatten_ids / isect_ids are written
//            from the sorted keys;
//   phase B  front-to-back alpha compositing, SURVEY.md Appendix A.3 (gsplat==1.0.0
//            rasterize_to_pixels fwd behind /root/reference/edgegaussians/models/edge_gs.py:250-268):
//            batches of 256 Gaussian records are staged in shared memory; each warp owns an 8x4
//            pixel sub-tile and only walks the Gaussians whose alpha >= 1/255 footprint can reach
//            that sub-tile (a conservative bounding test done once per (tile, Gaussian) by the
//            staging thread) -- the sequence of Gaussians each pixel composites is exactly gsplat's;
//            Every (warp, Gaussian) step also records WHICH of its 32 pixels composited the Gaussian
//            (one ballot): the 256-bit contribution mask per tile intersection is what the backward
//            kernel walks, so it never re-derives footprints or skip/stop decisions;
//   epilogue alpha = 1 - T, render channel 0, last_ids, and optionally the reference's
//            clamp -> channel 0 -> mean |render - gt| (edge_gs.py:279,290-296; train_gaussians.py:84-94)
//            as a per-CTA partial sum plus the per-pixel backward seed.
//
// colors == 1 on this path (edge_gs.py:247), so the three render channels are identical; one is stored.
#include "eg_common.cuh"

namespace {

constexpr int RF_THREADS = 256;
constexpr int SORT_CAP = 2048;  // keys sorted on chip (registers + 2 x 16 KB exchange buffers)
typedef unsigned long long u64;

__device__ __forceinline__ u64 u64min(u64 a, u64 b) { return a < b ? a : b; }
__device__ __forceinline__ u64 u64max(u64 a, u64 b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------------------
// Register-resident bitonic sort of n_pad = 256*E keys (E == 1: n_pad may be 32..256).
// Element index of key[r] in thread tid is e = r*256 + tid.  Standard network: step (k, j) pairs e
// with e^j, ascending where (e & k) == 0.
// ---------------------------------------------------------------------------------------------
template <int E>
__device__ __forceinline__ void sort_regs(u64 (&key)[E], const int n_pad, u64 *sbuf, const int tid) {
    int buf = 0;
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 256) {  // partner is another register of this thread (only when E > 1)
                const int jr = j >> 8;
#pragma unroll
                for (int jc = 1; jc < E; jc <<= 1) {
                    if (jr == jc) {
#pragma unroll
                        for (int r = 0; r < E; ++r) {
                            if ((r & jc) == 0) {
                                const bool up = (((r << 8) | tid) & k) == 0;
                                const u64 lo = u64min(key[r], key[r | jc]), hi = u64max(key[r], key[r | jc]);
                                key[r] = up ? lo : hi;
                                key[r | jc] = up ? hi : lo;
                            }
                        }
                    }
                }
            } else if (j >= 32) {  // partner lives in another warp: shared-memory exchange
                u64 *sb = sbuf + buf * SORT_CAP;
#pragma unroll
                for (int r = 0; r < E; ++r) sb[(r << 8) | tid] = key[r];
                __syncthreads();
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const int e = (r << 8) | tid;
                    const u64 p = sb[e ^ j];
                    const bool keep_min = ((e & k) == 0) == ((e & j) == 0);
                    key[r] = ((key[r] < p) == keep_min) ? key[r] : p;  // keys are unique (or both +inf)
                }
                buf ^= 1;
            } else {  // partner is a lane of the same warp
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const int e = (r << 8) | tid;
                    const u64 p = __shfl_xor_sync(0xffffffffu, key[r], j);
                    const bool keep_min = ((e & k) == 0) == ((e & j) == 0);
                    key[r] = ((key[r] < p) == keep_min) ? key[r] : p;
                }
            }
        }
    }
}

// FROM_SMEM: the keys sit in the exchange buffer itself (front-sort slices): every thread has loaded its keys
// before the network starts to overwrite the buffer.
template <int E, bool FROM_SMEM>
__device__ __forceinline__ void sort_segment_regs(const u64 *gkeys, const int L, const int n_pad, u64 *sbuf,
                                                  uint32_t *sids, int32_t *__restrict__ flat,
                                                  long long *__restrict__ isect, const long long tile_hi,
                                                  const int tid) {
    u64 key[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int e = (r << 8) | tid;
        key[r] = e < L ? gkeys[e] : ~0ull;
    }
    if (FROM_SMEM) __syncthreads();
    sort_regs<E>(key, n_pad, sbuf, tid);
    __syncthreads();  // exchange buffers are dead: reuse them for the sorted ids
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int e = (r << 8) | tid;
        if (e < L) {
            sids[e] = (uint32_t)key[r];
            flat[e] = (int32_t)(uint32_t)key[r];
            if (isect) isect[e] = tile_hi | (long long)(key[r] >> 32);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fallback for segments longer than SORT_CAP: ascending-only bitonic network, chunks of SORT_CAP in
// shared memory, the large-distance stages directly in global memory (virtual +inf padding: a
// comparator whose upper index is >= L is a no-op and is skipped).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cmpx(u64 &a, u64 &b) {
    if (a > b) { const u64 t = a; a = b; b = t; }
}

__device__ void bitonic_smem_full(u64 *s, int tid) {  // sorts SORT_CAP keys
    for (int lk = 1; (1 << lk) <= SORT_CAP; ++lk) {
        const int k = 1 << lk, hk = k >> 1;
        for (int idx = tid; idx < (SORT_CAP >> 1); idx += RF_THREADS) {
            const int blk = idx >> (lk - 1), off = idx & (hk - 1);
            cmpx(s[blk * k + off], s[blk * k + k - 1 - off]);
        }
        __syncthreads();
        for (int lj = lk - 2; lj >= 0; --lj) {
            const int j = 1 << lj;
            for (int idx = tid; idx < (SORT_CAP >> 1); idx += RF_THREADS) {
                const int i = ((idx >> lj) << (lj + 1)) | (idx & (j - 1));
                cmpx(s[i], s[i + j]);
            }
            __syncthreads();
        }
    }
}

__device__ void bitonic_tail_smem(u64 *s, int tid) {  // half-cleaners j = SORT_CAP/2 ... 1
    for (int lj = 10; lj >= 0; --lj) {
        const int j = 1 << lj;
        for (int idx = tid; idx < (SORT_CAP >> 1); idx += RF_THREADS) {
            const int i = ((idx >> lj) << (lj + 1)) | (idx & (j - 1));
            cmpx(s[i], s[i + j]);
        }
        __syncthreads();
    }
}

__device__ void sort_large(u64 *__restrict__ gk, int L, u64 *s, int tid) {
    static_assert(SORT_CAP == 2048, "bitonic_tail_smem assumes SORT_CAP == 2048");
    int n_pad = SORT_CAP;
    while (n_pad < L) n_pad <<= 1;
    const int n_chunks = (L + SORT_CAP - 1) / SORT_CAP;
    for (int c = 0; c < n_chunks; ++c) {
        const int base = c * SORT_CAP;
        for (int i = tid; i < SORT_CAP; i += RF_THREADS) s[i] = (base + i < L) ? gk[base + i] : ~0ull;
        __syncthreads();
        bitonic_smem_full(s, tid);
        for (int i = tid; i < SORT_CAP; i += RF_THREADS)
            if (base + i < L) gk[base + i] = s[i];
        __syncthreads();
    }
    for (int k = 2 * SORT_CAP; k <= n_pad; k <<= 1) {
        const int hk = k >> 1;
        for (int idx = tid; idx < (n_pad >> 1); idx += RF_THREADS) {
            const int blk = idx / hk, off = idx - blk * hk;
            const int i = blk * k + off, j = blk * k + k - 1 - off;
            if (j < L) {
                u64 a = gk[i], b = gk[j];
                if (a > b) { gk[i] = b; gk[j] = a; }
            }
        }
        __syncthreads();
        for (int j = k >> 2; j >= SORT_CAP; j >>= 1) {
            for (int idx = tid; idx < (n_pad >> 1); idx += RF_THREADS) {
                const int i = 2 * j * (idx / j) + (idx % j);
                if (i + j < L) {
                    u64 a = gk[i], b = gk[i + j];
                    if (a > b) { gk[i] = b; gk[i + j] = a; }
                }
            }
            __syncthreads();
        }
        for (int c = 0; c < n_chunks; ++c) {
            const int base = c * SORT_CAP;
            for (int i = tid; i < SORT_CAP; i += RF_THREADS) s[i] = (base + i < L) ? gk[base + i] : ~0ull;
            __syncthreads();
            bitonic_tail_smem(s, tid);
            for (int i = tid; i < SORT_CAP; i += RF_THREADS)
                if (base + i < L) gk[base + i] = s[i];
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Front-to-back compositing of list entries [0, n) of a tile (phase B).  ids: the entries' Gaussian ids, in shared
// memory (sids) or, for lists sorted through global memory, flat[0 .. n).  `base` = position of entry 0 in the tile's
// list (cmask / last index).  Stops early once every pixel of the tile is done.
// ---------------------------------------------------------------------------------------------
struct PixState {
    float T, out;
    int last, lastg;
    bool done;
};

template <bool WANT_LAST>
__device__ __forceinline__ void composite_entries(const int n, const int base, const uint32_t *sids,
                                                  const int32_t *__restrict__ flat, const float4 *__restrict__ rec,
                                                  float4 *sAB, uint32_t *s_cm, uint4 *__restrict__ cmask_out,
                                                  const float X0, const float Y0, const float px, const float py,
                                                  const float t_stop, const int tid, const int lane, const int warp,
                                                  PixState &st, int &b_done) {
    for (int b0 = 0; b0 < n; b0 += RF_THREADS) {
        // barrier doubles as "sorted ids / previous batch visible" and the all-pixels-done early exit
        if (__syncthreads_and(st.done)) break;
        const int k = b0 + tid;
        if (k < n) {
            const int gid = sids != nullptr ? (int)sids[k] : flat[k];
            const float4 r0 = __ldg(rec + 2 * gid), r1 = __ldg(rec + 2 * gid + 1);
            // sub-tiles (8x4 pixels, one per warp) the alpha >= 1/255 ellipse can reach: a warp only walks those
            const int mask = eg_subtile_mask(r0.x, r0.y, r1.x, r1.y, r1.z, r0.z, X0, Y0);
            const EgFold f = eg_fold(r1.x, r1.y, r1.z, r0.z);
            sAB[2 * tid] = make_float4(r0.x, r0.y, f.lo, __int_as_float(mask));
            sAB[2 * tid + 1] = make_float4(f.fa, f.fb, f.fc, __int_as_float(gid));
        }
        if (cmask_out != nullptr) {  // zero the [8 warps][256 Gaussians] contribution words of the batch
            reinterpret_cast<uint4 *>(s_cm)[tid] = make_uint4(0u, 0u, 0u, 0u);
            reinterpret_cast<uint4 *>(s_cm)[tid + RF_THREADS] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        const int nb = min(RF_THREADS, n - b0);
        for (int c = 0; c < nb && !__all_sync(0xffffffffu, st.done); c += 32) {
            const int m = (c + lane < nb) ? __float_as_int(sAB[2 * (c + lane)].w) : 0;
            unsigned bits = __ballot_sync(0xffffffffu, (m >> warp) & 1);
            unsigned myword = 0;  // lane j keeps the contribution mask of Gaussian c + j for this warp's pixels
            while (bits) {
                const int j = __ffs(bits) - 1;
                const int t = c + j;
                bits &= bits - 1;
                const float4 a = sAB[2 * t];
                const float4 cn = sAB[2 * t + 1];
                const float dx = a.x - px, dy = a.y - py;
                const float pw2 = eg_pow2arg(cn.x, cn.y, cn.z, a.z, dx, dy);  // log2(opacity * exp(-sigma))
                const float al = fminf(EG_ALPHA_MAX, eg_ex2(pw2));
                const bool valid = !st.done && pw2 <= a.z && al >= EG_ALPHA_MIN;  // sigma >= 0 and alpha >= 1/255
                const float nT = st.T * (1.0f - al);
                const bool stop = valid && nT <= t_stop;
                const bool take = valid && !stop;
                st.done = st.done || stop;
                st.out = take ? fmaf(al, st.T, st.out) : st.out;
                st.T = take ? nT : st.T;
                if (WANT_LAST) {
                    st.last = take ? (base + b0 + t) : st.last;
                    st.lastg = take ? __float_as_int(cn.w) : st.lastg;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, take);
                myword = (lane == j) ? bal : myword;
            }
            s_cm[warp * RF_THREADS + c + lane] = myword;  // [warp][Gaussian]: conflict-free
        }
        if (cmask_out != nullptr) {  // contribution masks of this batch -> global (32 B per intersection)
            __syncthreads();
            if (k < n) {
                cmask_out[2 * (size_t)(base + k)] = make_uint4(s_cm[tid], s_cm[RF_THREADS + tid], s_cm[2 * RF_THREADS + tid],
                                                               s_cm[3 * RF_THREADS + tid]);
                cmask_out[2 * (size_t)(base + k) + 1] = make_uint4(s_cm[4 * RF_THREADS + tid], s_cm[5 * RF_THREADS + tid],
                                                                   s_cm[6 * RF_THREADS + tid], s_cm[7 * RF_THREADS + tid]);
            }
            b_done = base + min(n, b0 + RF_THREADS);
        }
    }
}

// EG_FLAG_FRONT_SORT: depth slices of at most FS_SLICE keys (one histogram bin is never split, so a slice can be
// larger when many keys share a bin: up to SORT_CAP, beyond that the tile takes the full sort)
constexpr int FS_BINS = 256;
constexpr int FS_SLICE = 512;

template <int GT_KIND, bool WANT_LAST>
__device__ __forceinline__ void raster_tile(
    const eg_config &cfg, const int tw, const int tile, const bool flagged_only, const float4 *__restrict__ rec,
    const int32_t *__restrict__ tile_offsets, u64 *__restrict__ keys, int32_t *__restrict__ flatten_ids,
    long long *__restrict__ isect_ids, float *__restrict__ render0, float *__restrict__ alpha_out,
    int32_t *__restrict__ last_ids, uint4 *__restrict__ cmask, const void *__restrict__ gt,
    double *__restrict__ loss_sum, float *__restrict__ wpix, uint32_t *__restrict__ last_depth,
    int32_t *__restrict__ last_gid, const int32_t *__restrict__ tile_cnt, int32_t *__restrict__ tile_done,
    const float *__restrict__ loss_params, const unsigned char *__restrict__ sel_mask, int32_t *__restrict__ status) {
    __shared__ __align__(16) u64 sbuf[2 * SORT_CAP];  // sort exchange buffers, then the sorted ids
    __shared__ __align__(16) float4 sAB[2 * RF_THREADS];  // per Gaussian: (mean2d.x, mean2d.y, log2 opacity,
                                                          // sub-tile mask bits) , (folded conic fa, fb, fc, id)
    __shared__ float s_red[RF_THREADS / 32];
    __shared__ int s_hist[FS_BINS + 1];
    __shared__ unsigned s_drange[2];
    __shared__ int s_cursor;
    __shared__ int s_wtot[RF_THREADS / 32];
    uint32_t *sids = reinterpret_cast<uint32_t *>(sbuf);             // first 8 KB of the (dead) exchange buffers
    uint32_t *s_cm = reinterpret_cast<uint32_t *>(sbuf + SORT_CAP);  // second half: contribution masks [256][8]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_y = tile / tw, tile_x = tile - tile_y * tw;
    // flagged_only: fallback of the Gaussian-major forward (eg_splat_fwd.cu) -- the tile's keys were appended by
    // eg_emit_flagged to its fixed-capacity bucket (count in tile_cnt, no scan)
    const int start = flagged_only ? tile * cfg.tile_capacity : tile_offsets[tile];
    const int L = flagged_only ? min(tile_cnt[tile], cfg.tile_capacity) : tile_offsets[tile + 1] - start;

    const int sub_x = tile_x * EG_TILE + 8 * (warp & 1), sub_y = tile_y * EG_TILE + 4 * (warp >> 1);
    const int pxi = sub_x + (lane & 7), pyi = sub_y + (lane >> 3);
    const bool inside = pxi < cfg.width && pyi < cfg.height;
    const float px = (float)pxi + 0.5f, py = (float)pyi + 0.5f;
    const float X0 = (float)(tile_x * EG_TILE), Y0 = (float)(tile_y * EG_TILE);
    const bool on_chip = L <= SORT_CAP;
    u64 *bucket = ((cfg.flags & EG_FLAG_COMPACT_KEYS) && !flagged_only) ? keys + start
                                                                       : keys + (size_t)tile * (size_t)cfg.tile_capacity;
    uint4 *cm_tile = cmask != nullptr ? cmask + 2 * (size_t)start : nullptr;

    // EG_FLAG_LAZY_SORT: composite once in bucket (arbitrary) order.  If no pixel of the tile comes near the
    // transmittance stop threshold, no prefix product in ANY order can cross it, so gsplat's result is the
    // order-free product and the sort is skipped (flatten_ids then holds the tile's ids unsorted).  Otherwise
    // the tile is redone in sorted order (pass 1), which is always exact.
    const bool ids_free = isect_ids == nullptr && last_ids == nullptr;  // nobody asked for gsplat's full sorted lists
    const bool lazy = (cfg.flags & EG_FLAG_LAZY_SORT) != 0 && ids_free && !flagged_only;
    // EG_FLAG_FRONT_SORT: sort and composite the list front to back in depth slices, stop when the tile is done
    bool front = (cfg.flags & EG_FLAG_FRONT_SORT) != 0 && ids_free && L > FS_SLICE;
    PixState ps;
    ps.T = 1.0f; ps.out = 0.0f; ps.last = -start; ps.lastg = -1; ps.done = !inside;
    int n_processed = L;  // list entries with defined flatten_ids / cmask
    for (int pass = lazy ? 0 : 1; pass < 2; ++pass) {
    const bool sorted = pass == 1;
    const float t_stop = sorted ? EG_T_MIN : EG_T_MIN * 1.0002f;
    ps.T = 1.0f; ps.out = 0.0f; ps.done = !inside; ps.lastg = -1;
    ps.last = -start;  // relative to the segment start; gsplat initialises the absolute index to 0
    int b_done = 0;  // list entries [0, b_done) have their contribution masks written
    n_processed = L;

    unsigned dmin = 0u, sh = 0u;
    if (sorted && front) {
        // ---- depth range and histogram of the tile's keys ----
        if (tid < 2) s_drange[tid] = tid == 0 ? 0xffffffffu : 0u;
        s_hist[tid] = 0;
        if (tid == 0) s_hist[FS_BINS] = 0;
        __syncthreads();
        unsigned lo = 0xffffffffu, hi = 0u;
        for (int i = tid; i < L; i += RF_THREADS) {
            const unsigned d = (unsigned)(bucket[i] >> 32);
            lo = min(lo, d);
            hi = max(hi, d);
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) {
            atomicMin(&s_drange[0], lo);
            atomicMax(&s_drange[1], hi);
        }
        __syncthreads();
        dmin = s_drange[0];
        const unsigned range = s_drange[1] - dmin;
        sh = range >= (unsigned)FS_BINS ? (unsigned)(32 - __clz(range) - 8) : 0u;  // (d - dmin) >> sh < 256, monotone in d
        for (int i = tid; i < L; i += RF_THREADS) atomicAdd(&s_hist[((unsigned)(bucket[i] >> 32) - dmin) >> sh], 1);
        __syncthreads();
        // exclusive scan of the 256 bin counts (thread = bin) -> s_hist[b] = keys in bins < b, s_hist[256] = L
        const int cnt = s_hist[tid];
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_wtot[warp] = incl;
        const bool too_big = __syncthreads_or(cnt > SORT_CAP);  // a single bin that cannot be sorted on chip
        int wbase = 0;
#pragma unroll
        for (int w = 0; w < RF_THREADS / 32; ++w)
            if (w < warp) wbase += s_wtot[w];
        s_hist[tid] = wbase + incl - cnt;
        if (tid == RF_THREADS - 1) s_hist[FS_BINS] = wbase + incl;
        __syncthreads();
        if (too_big) front = false;  // uniform: fall through to the full sort below
    }

    if (sorted && front) {
        // ---- slices of whole bins, front to back ----
        int bin_lo = 0, done_cnt = 0;
        while (bin_lo < FS_BINS) {
            const int off_lo = s_hist[bin_lo];
            // bins [bin_lo, bin_hi]: as many as fit FS_SLICE keys, at least one (cumulative counts are monotone)
            const int fit = __syncthreads_count(tid >= bin_lo && s_hist[tid + 1] - off_lo <= FS_SLICE);
            const int bin_hi = bin_lo + max(fit, 1) - 1;
            const int n = s_hist[bin_hi + 1] - off_lo;
            if (tid == 0) s_cursor = 0;
            __syncthreads();
            if (n > 0) {
                // gather the slice's keys (arbitrary order) into the first half of the exchange buffer
                for (int i = tid; i < L; i += RF_THREADS) {
                    const u64 key = bucket[i];
                    const int b = (int)(((unsigned)(key >> 32) - dmin) >> sh);
                    if (b >= bin_lo && b <= bin_hi) sbuf[atomicAdd(&s_cursor, 1)] = key;
                }
                __syncthreads();
                int32_t *flat = flatten_ids + start + done_cnt;
                if (n <= 256) {
                    int n_pad = 32;
                    while (n_pad < n) n_pad <<= 1;
                    sort_segment_regs<1, true>(sbuf, n, n_pad, sbuf, sids, flat, nullptr, 0, tid);
                } else if (n <= 512) {
                    sort_segment_regs<2, true>(sbuf, n, 512, sbuf, sids, flat, nullptr, 0, tid);
                } else if (n <= 1024) {
                    sort_segment_regs<4, true>(sbuf, n, 1024, sbuf, sids, flat, nullptr, 0, tid);
                } else {
                    sort_segment_regs<8, true>(sbuf, n, 2048, sbuf, sids, flat, nullptr, 0, tid);
                }
                composite_entries<WANT_LAST>(n, done_cnt, sids, nullptr, rec, sAB, s_cm, cm_tile, X0, Y0, px, py, t_stop,
                                             tid, lane, warp, ps, b_done);
                done_cnt += n;
            }
            bin_lo = bin_hi + 1;
            if (__syncthreads_and(ps.done)) break;  // everything behind this slice is invisible
        }
        n_processed = done_cnt;
        if (cmask != nullptr)  // entries of the last slice skipped by the early exit inside composite_entries
            for (int k = b_done + tid; k < done_cnt; k += RF_THREADS) {
                cm_tile[2 * (size_t)k] = make_uint4(0u, 0u, 0u, 0u);
                cm_tile[2 * (size_t)k + 1] = make_uint4(0u, 0u, 0u, 0u);
            }
    } else {
    // ---------------- phase A: sort the whole segment ----------------
    if (L > 0 && !sorted) {
        for (int i = tid; i < L; i += RF_THREADS) {
            const uint32_t id = (uint32_t)bucket[i];
            flatten_ids[start + i] = (int32_t)id;
            if (on_chip) sids[i] = id;
        }
    } else if (L > 0) {
        const long long tile_hi = (long long)tile << 32;
        long long *isect = isect_ids ? isect_ids + start : nullptr;
        if (L <= 256) {
            int n_pad = 32;
            while (n_pad < L) n_pad <<= 1;
            sort_segment_regs<1, false>(bucket, L, n_pad, sbuf, sids, flatten_ids + start, isect, tile_hi, tid);
        } else if (L <= 512) {
            sort_segment_regs<2, false>(bucket, L, 512, sbuf, sids, flatten_ids + start, isect, tile_hi, tid);
        } else if (L <= 1024) {
            sort_segment_regs<4, false>(bucket, L, 1024, sbuf, sids, flatten_ids + start, isect, tile_hi, tid);
        } else if (L <= 2048) {
            sort_segment_regs<8, false>(bucket, L, 2048, sbuf, sids, flatten_ids + start, isect, tile_hi, tid);
        } else {
            sort_large(bucket, L, sbuf, tid);
            for (int i = tid; i < L; i += RF_THREADS) {
                const u64 k = bucket[i];
                flatten_ids[start + i] = (int32_t)(uint32_t)k;
                if (isect) isect[i] = tile_hi | (long long)(k >> 32);
            }
        }
    }
    // ---------------- phase B: compositing ----------------
    composite_entries<WANT_LAST>(L, 0, on_chip ? sids : nullptr, flatten_ids + start, rec, sAB, s_cm, cm_tile, X0, Y0, px, py,
                                 t_stop, tid, lane, warp, ps, b_done);
    if (cmask != nullptr)  // batches skipped by the all-pixels-done early exit contribute nothing
        for (int k = b_done + tid; k < L; k += RF_THREADS) {
            cm_tile[2 * (size_t)k] = make_uint4(0u, 0u, 0u, 0u);
            cm_tile[2 * (size_t)k + 1] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    // `done` of an in-image pixel can only have been set by the stop rule
    if (sorted || !__syncthreads_or(ps.done && inside)) break;
    if (tid == 0) atomicAdd(status + EG_ST_REDO, 1);  // lets the host switch lazy sorting off when it stops paying
    }  // pass

    // ---------------- epilogue ----------------
    if (tile_done != nullptr && tid == 0 && !flagged_only) tile_done[tile] = n_processed;
    if (__syncthreads_or(ps.done && inside) && tid == 0 && !flagged_only) atomicAdd(status + EG_ST_STOPPED, 1);
    float absd = 0.0f;
    if (inside) {
        const long long pix = (long long)pyi * cfg.width + pxi;
        const float T = ps.T, out = ps.out;
        if (WANT_LAST && last_depth != nullptr) {
            // sort key of the last Gaussian a STOPPED pixel composited (what eg_splat_bwd compares against);
            // a pixel that never stopped composited every Gaussian that passed the alpha test
            uint32_t ld = 0xffffffffu;
            int lg = -1;
            if (ps.done && ps.lastg >= 0) {
                lg = ps.lastg;
                ld = __float_as_uint(__ldg(rec + 2 * lg).w);
            } else if (ps.done) {
                ld = 0u;  // stopped on its very first Gaussian: nothing was composited, no key is <= (0, -1)
            }
            last_depth[pix] = ld;
            last_gid[pix] = lg;
        }
        if (alpha_out) alpha_out[pix] = 1.0f - T;
        if (render0) render0[pix] = out;
        if (WANT_LAST && last_ids) last_ids[pix] = start + ps.last;
        if (GT_KIND != EG_GT_NONE) {
            float g;
            if (GT_KIND == EG_GT_F32) g = __ldg(reinterpret_cast<const float *>(gt) + pix);
            else g = __fdiv_rn((float)__ldg(reinterpret_cast<const unsigned char *>(gt) + pix), 255.0f);
            const float rc = fminf(fmaxf(out, 0.0f), 1.0f);
            const float d = rc - g;
            const float coef = eg_loss_coef(loss_params, sel_mask, g, pix);
            absd = coef * fabsf(d);
            const float sgn = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
            const float pass = (out >= 0.0f && out <= 1.0f) ? 1.0f : 0.0f;
            if (wpix) wpix[pix] = sgn * pass * T * coef;
        }
    }
    if (GT_KIND != EG_GT_NONE && loss_sum != nullptr) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) absd += __shfl_xor_sync(0xffffffffu, absd, d);
        if (lane == 0) s_red[warp] = absd;
        __syncthreads();
        if (tid == 0) {
            float tsum = 0.0f;
#pragma unroll
            for (int w = 0; w < RF_THREADS / 32; ++w) tsum += s_red[w];
            if (tsum != 0.0f) atomicAdd(loss_sum, (double)tsum);
        }
    }
}

template <int GT_KIND, bool WANT_LAST>
__global__ void __launch_bounds__(RF_THREADS, 5) raster_fwd_kernel(
    const eg_config cfg, int tw, const float4 *__restrict__ rec, const int32_t *__restrict__ tile_offsets,
    u64 *__restrict__ keys, int32_t *__restrict__ flatten_ids, long long *__restrict__ isect_ids,
    float *__restrict__ render0, float *__restrict__ alpha_out, int32_t *__restrict__ last_ids,
    uint4 *__restrict__ cmask, const void *__restrict__ gt, double *__restrict__ loss_sum,
    float *__restrict__ wpix, uint32_t *__restrict__ last_depth, int32_t *__restrict__ last_gid,
    int32_t *__restrict__ tile_done, const float *__restrict__ loss_params,
    const unsigned char *__restrict__ sel_mask, int32_t *__restrict__ status) {
    if (status[EG_ST_OVERFLOW]) return;
    raster_tile<GT_KIND, WANT_LAST>(cfg, tw, blockIdx.x, false, rec, tile_offsets, keys, flatten_ids, isect_ids, render0,
                                    alpha_out, last_ids, cmask, gt, loss_sum, wpix, last_depth, last_gid, nullptr,
                                    tile_done, loss_params, sel_mask, status);
}

// Fallback of the Gaussian-major forward: a small persistent grid walks the list of flagged tiles (usually empty:
// the kernel then costs a launch and nothing else).
template <int GT_KIND>
__global__ void __launch_bounds__(RF_THREADS, 5) raster_fwd_flagged_kernel(
    const eg_config cfg, int tw, const float4 *__restrict__ rec, u64 *__restrict__ keys,
    int32_t *__restrict__ flatten_ids, float *__restrict__ render0, float *__restrict__ alpha_out,
    const void *__restrict__ gt, double *__restrict__ loss_sum, float *__restrict__ wpix,
    uint32_t *__restrict__ last_depth, int32_t *__restrict__ last_gid, const int32_t *__restrict__ stop_list,
    const int32_t *__restrict__ tile_cnt, const float *__restrict__ loss_params,
    const unsigned char *__restrict__ sel_mask, int32_t *__restrict__ status) {
    if (status[EG_ST_OVERFLOW]) return;
    const int n = status[EG_ST_STOPPED];
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        __syncthreads();  // shared memory of the previous tile is dead
        raster_tile<GT_KIND, true>(cfg, tw, stop_list[i], true, rec, nullptr, keys, flatten_ids, nullptr, render0,
                                   alpha_out, nullptr, nullptr, gt, loss_sum, wpix, last_depth, last_gid, tile_cnt,
                                   nullptr, loss_params, sel_mask, status);
    }
}

}  // namespace

extern "C" int eg_raster_fwd(const eg_config *cfg, const float *rec, const int32_t *tile_offsets, uint64_t *keys,
                             int32_t *flatten_ids, int64_t *isect_ids, float *render0, float *alpha,
                             int32_t *last_ids, uint32_t *cmask, const void *gt, int gt_kind, double *loss_sum,
                             float *wpix, uint32_t *last_depth, int32_t *last_gid, const int32_t *stop_list,
                             const int32_t *tile_cnt, int32_t *tile_done, const float *loss_params,
                             const uint8_t *sel_mask, int32_t *status, void *stream) {
    if (cfg == nullptr || cfg->tile_size != EG_TILE) {
        eg_set_error("eg_raster_fwd: tile_size must be %d", EG_TILE);
        return 1;
    }
    if (flatten_ids == nullptr) {
        eg_set_error("eg_raster_fwd: flatten_ids is required");
        return 1;
    }
    if ((last_depth == nullptr) != (last_gid == nullptr)) {
        eg_set_error("eg_raster_fwd: last_depth and last_gid go together");
        return 1;
    }
    if ((stop_list == nullptr) != (tile_cnt == nullptr)) {
        eg_set_error("eg_raster_fwd: stop_list and tile_cnt go together");
        return 1;
    }
    if (stop_list == nullptr && tile_offsets == nullptr) {
        eg_set_error("eg_raster_fwd: tile_offsets is required");
        return 1;
    }
    if (gt == nullptr) gt_kind = EG_GT_NONE;
    if (gt_kind == EG_GT_NONE && (loss_params != nullptr || sel_mask != nullptr)) {
        eg_set_error("eg_raster_fwd: loss_params / sel_mask need the edge map gt");
        return 1;
    }
    if (sel_mask != nullptr && loss_params == nullptr) {
        eg_set_error("eg_raster_fwd: sel_mask needs loss_params");
        return 1;
    }
    if ((cfg->flags & EG_FLAG_FRONT_SORT) && cmask != nullptr && tile_done == nullptr && isect_ids == nullptr &&
        last_ids == nullptr && stop_list == nullptr) {
        eg_set_error("eg_raster_fwd: EG_FLAG_FRONT_SORT with cmask needs tile_done (entries behind it are undefined)");
        return 1;
    }
    int tw, th;
    eg_tile_grid(cfg->width, cfg->height, cfg->tile_size, &tw, &th);
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = tw * th;
    if (stop_list != nullptr) {  // flagged tiles only
        if (last_depth == nullptr || isect_ids != nullptr || last_ids != nullptr || cmask != nullptr) {
            eg_set_error("eg_raster_fwd: the flagged-tile mode writes last_depth / last_gid and no ids or masks");
            return 1;
        }
        const int pgrid = grid < 148 * 5 ? grid : 148 * 5;
#define EG_LAUNCHF(KIND)                                                                                          \
    raster_fwd_flagged_kernel<KIND><<<pgrid, RF_THREADS, 0, s>>>(*cfg, tw, (const float4 *)rec, (u64 *)keys,       \
                                                                 flatten_ids, render0, alpha, gt, loss_sum, wpix, \
                                                                 last_depth, last_gid, stop_list, tile_cnt,       \
                                                                 loss_params, sel_mask, status)
        switch (gt_kind) {
            case EG_GT_NONE: EG_LAUNCHF(EG_GT_NONE); break;
            case EG_GT_F32: EG_LAUNCHF(EG_GT_F32); break;
            case EG_GT_U8: EG_LAUNCHF(EG_GT_U8); break;
            default: eg_set_error("eg_raster_fwd: bad gt_kind %d", gt_kind); return 1;
        }
#undef EG_LAUNCHF
        return eg_check_launch("eg_raster_fwd/flagged");
    }
#define EG_LAUNCH2(KIND, WL)                                                                                     \
    raster_fwd_kernel<KIND, WL><<<grid, RF_THREADS, 0, s>>>(*cfg, tw, (const float4 *)rec, tile_offsets,         \
                                                            (u64 *)keys, flatten_ids, (long long *)isect_ids,    \
                                                            render0, alpha, last_ids, (uint4 *)cmask, gt,        \
                                                            loss_sum, wpix, last_depth, last_gid, tile_done,     \
                                                            loss_params, sel_mask, status)
#define EG_LAUNCH(KIND)                       \
    do {                                      \
        if (last_ids != nullptr || last_depth != nullptr) EG_LAUNCH2(KIND, true); \
        else EG_LAUNCH2(KIND, false);         \
    } while (0)
    switch (gt_kind) {
        case EG_GT_NONE: EG_LAUNCH(EG_GT_NONE); break;
        case EG_GT_F32: EG_LAUNCH(EG_GT_F32); break;
        case EG_GT_U8: EG_LAUNCH(EG_GT_U8); break;
        default: eg_set_error("eg_raster_fwd: bad gt_kind %d", gt_kind); return 1;
    }
#undef EG_LAUNCH
#undef EG_LAUNCH2
    return eg_check_launch("eg_raster_fwd");
}
