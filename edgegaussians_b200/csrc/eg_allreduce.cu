// eg_allreduce.cu -- the view-sharded step's gradient exchange as ONE kernel of this library over NVLink 5 /
// NVSwitch (no NCCL on the data path).
//
// The reference is single-GPU (train_gaussians.py:311); the sharding is SURVEY.md section 8e's: parameters
// replicated, one view per GPU per step, the per-view gradients of the flat buffer means | scales | quats |
// opacities (eg_grad_layout) summed over the ranks.  Every rank holds that buffer at the same offset of a SYMMETRIC
// allocation (peer-mapped on every rank, and bound to an NVSwitch multicast object where the fabric has one).
//
//   barrier   every CTA b of every rank signals flag[b][my rank] on each peer and waits for its own flags --
//             the peers' backward kernels (same stream, earlier) have then written their gradients;
//   reduce    rank r owns the r-th slice.  Multicast path: multimem.ld_reduce.add.v4.f32 pulls the slice through
//             the switch, which sums the G copies in flight (one 16-byte response instead of G), and
//             multimem.st.v4.f32 writes the sum back to ALL ranks' buffers in one store.  Per GPU and direction
//             that is one buffer length on the links instead of 2 (G-1)/G lengths for a ring, and two latencies
//             instead of 2 (G-1).  Peer path (no multicast object): plain 128-bit loads from the G peer buffers in
//             rank order, 128-bit stores to all of them;
//   barrier   the slices written by the peers are visible before anything on this stream reads the buffer.
//
// The barriers use monotonic arrival counters plus per-CTA local epochs (rank_barrier_* below): the kernel is re-entrant
// without host resets and can be captured in a CUDA graph -- the whole iteration, exchange included, is then one graph
// replay.
#include <cstdlib>

#include "eg_common.cuh"

namespace {

constexpr int AR_THREADS = 512;
constexpr int AR_MAX_RANKS = 8;

struct ArPeers {
    float *buf[AR_MAX_RANKS];       // peer-mapped gradient buffers, by rank
    uint32_t *flags[AR_MAX_RANKS];  // peer-mapped flag areas, by rank: (2 + grid) blocks of 16 words (rank_barrier_*)
    int leader;                     // 0: every CTA runs its own rank barrier; 1: one leader CTA per rank
};

// The reduced index space: up to 4 segments of the buffer (the slices of means | scales | quats | opacities that hold a
// Gaussian range), concatenated, in float4 units.
struct ArSegs {
    long long off4[4];    // first float4 of each segment in the buffer
    long long first4[5];  // prefix of the segment lengths (float4)
    int n;
};
__device__ __forceinline__ long long seg_index(const ArSegs &sg, const long long i) {
    int s = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k)
        if (k < sg.n && i >= sg.first4[k]) s = k;
    return sg.off4[s] + (i - sg.first4[s]);
}

// Rank barrier of CTA `blockIdx.x` with the same CTA of every other rank.  Threads 0 .. world-1 each handle one peer:
//   signal   red.release.sys.add of 1 on MY arrival counter in the peer's flag area -- fire-and-forget (no round trip);
//            the release orders this CTA's earlier stores (cumulative through the __syncthreads) before it;
//   wait     ld.acquire.sys polling of the PEER's counter in my own (local) flag area until it reaches this barrier's
//            epoch.  Counters only grow; the epoch each (CTA, peer) pair has reached lives in a second, purely local
//            word, so the barrier needs no reset between launches and can sit in a captured CUDA graph.
// One barrier costs about one NVLink one-way latency (the compare-and-swap toggles this replaces paid a remote
// round trip per attempt: 12 us per barrier measured on 2 x B200, now the data phase dominates).
__device__ __forceinline__ void red_release_sys_add(uint32_t *addr, uint32_t v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *addr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
// flag area (u32 words, 16 per block): block 0 = arrival counters of the LEADER barrier, block 1 = its local words
// (launch count, arrivals of this grid's CTAs, go epoch), blocks 2 + b = per-CTA barrier of CTA b (8 arrival counters, 8 epochs)
constexpr int AR_BLK = 2 * AR_MAX_RANKS;
enum { AR_L_LAUNCH = 0, AR_L_ARRIVED = 1, AR_L_GO = 2 };

__device__ __forceinline__ void rank_barrier_cta(const ArPeers &peers, const int rank, const int world) {
    __syncthreads();
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
        const int peer = threadIdx.x;
        const size_t blk = (size_t)(blockIdx.x + 2) * AR_BLK;
        uint32_t *epoch = peers.flags[rank] + blk + AR_MAX_RANKS + peer;   // local, this thread's alone
        const uint32_t target = *epoch + 1u;
        red_release_sys_add(peers.flags[peer] + blk + rank, 1u);
        const uint32_t *get = peers.flags[rank] + blk + peer;
        while ((int32_t)(ld_acquire_sys(get) - target) < 0) {}
        *epoch = target;
    }
    __syncthreads();
}

// Leader form (peers.leader != 0): ONE CTA per rank talks to the peers, the others synchronise with it through local
// (gpu-scope) words -- 7 system-scope releases per barrier and rank instead of 7 x grid.  The launch count L kept in the
// local block gives both barriers of launch L their epochs (2L + 1, 2L + 2) for any grid size.
__device__ __forceinline__ void st_release_gpu(uint32_t *addr, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *addr) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t atom_add_acq_rel_gpu(uint32_t *addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void leader_exchange(const ArPeers &peers, const int rank, const int world, const uint32_t target) {
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
        const int peer = threadIdx.x;
        red_release_sys_add(peers.flags[peer] + rank, 1u);
        const uint32_t *get = peers.flags[rank] + peer;
        while ((int32_t)(ld_acquire_sys(get) - target) < 0) {}
    }
    __syncthreads();
}
// returns the launch count (to be handed to rank_barrier_end)
__device__ __forceinline__ uint32_t rank_barrier_begin(const ArPeers &peers, const int rank, const int world) {
    if (!peers.leader) {
        rank_barrier_cta(peers, rank, world);
        return 0u;
    }
    __shared__ uint32_t s_launch;
    uint32_t *loc = peers.flags[rank] + AR_BLK;
    if (threadIdx.x == 0) s_launch = *reinterpret_cast<volatile uint32_t *>(loc + AR_L_LAUNCH);
    __syncthreads();
    const uint32_t L = s_launch, target = 2u * L + 1u;
    if (blockIdx.x == 0) {
        leader_exchange(peers, rank, world, target);
        if (threadIdx.x == 0) st_release_gpu(loc + AR_L_GO, target);
    } else if (threadIdx.x == 0) {
        while ((int32_t)(ld_acquire_gpu(loc + AR_L_GO) - target) < 0) {}
    }
    __syncthreads();
    return L;
}
__device__ __forceinline__ void rank_barrier_end(const ArPeers &peers, const int rank, const int world, const uint32_t L) {
    if (!peers.leader) {
        rank_barrier_cta(peers, rank, world);
        return;
    }
    __shared__ uint32_t s_last;
    uint32_t *loc = peers.flags[rank] + AR_BLK;
    __syncthreads();   // this CTA's stores happen before its arrival below
    if (threadIdx.x == 0) s_last = atom_add_acq_rel_gpu(loc + AR_L_ARRIVED, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!s_last) return;   // only the last CTA to arrive holds the kernel open until every peer has arrived too
    leader_exchange(peers, rank, world, 2u * L + 2u);
    if (threadIdx.x == 0) {
        loc[AR_L_ARRIVED] = 0u;
        loc[AR_L_LAUNCH] = L + 1u;
    }
}

__device__ __forceinline__ float4 mc_ld_reduce(const float *mc_addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc_addr) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float *mc_addr, const float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// n4 = number of float4 elements of the whole buffer; rank r owns the float4 range [r * per, min(n4, (r+1) * per)).
// UNROLL independent 16-byte requests per peer are in flight per thread: one round trip over NVLink takes
// microseconds, so the bytes in flight (threads x UNROLL x 16 B) are what sets the rate, not the instruction count.
template <bool MULTICAST, int UNROLL>
__global__ void __launch_bounds__(AR_THREADS) allreduce_kernel(const ArPeers peers, float *__restrict__ mc_buf,
                                                               const ArSegs sg, const int rank, const int world) {
    const long long n4 = sg.first4[sg.n];
    const uint32_t launch = rank_barrier_begin(peers, rank, world);
    const long long per = (n4 + world - 1) / world;
    const long long lo = (long long)rank * per, hi = (lo + per < n4) ? lo + per : n4;
    const long long stride = (long long)gridDim.x * AR_THREADS;
    if (MULTICAST) {
        float4 *mc4 = reinterpret_cast<float4 *>(mc_buf);
        for (long long i0 = lo + (long long)blockIdx.x * AR_THREADS + threadIdx.x; i0 < hi; i0 += UNROLL * stride) {
            float4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (i0 + u * stride < hi) v[u] = mc_ld_reduce(reinterpret_cast<const float *>(mc4 + seg_index(sg, i0 + u * stride)));
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (i0 + u * stride < hi) mc_st(reinterpret_cast<float *>(mc4 + seg_index(sg, i0 + u * stride)), v[u]);
        }
    } else {
        constexpr int PU = UNROLL >= 4 ? 2 : 1;  // elements per thread and iteration: PU x world loads in flight
        for (long long i0 = lo + (long long)blockIdx.x * AR_THREADS + threadIdx.x; i0 < hi; i0 += PU * stride) {
            float4 v[PU][AR_MAX_RANKS];
#pragma unroll
            for (int u = 0; u < PU; ++u)
#pragma unroll
                for (int r = 0; r < AR_MAX_RANKS; ++r)
                    if (r < world && i0 + u * stride < hi)
                        v[u][r] = __ldcg(reinterpret_cast<const float4 *>(peers.buf[r]) + seg_index(sg, i0 + u * stride));
#pragma unroll
            for (int u = 0; u < PU; ++u) {
                if (i0 + u * stride >= hi) continue;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int r = 0; r < AR_MAX_RANKS; ++r)   // fixed rank order: every rank computes the same bits
                    if (r < world) { acc.x += v[u][r].x; acc.y += v[u][r].y; acc.z += v[u][r].z; acc.w += v[u][r].w; }
#pragma unroll
                for (int r = 0; r < AR_MAX_RANKS; ++r)
                    if (r < world) __stcg(reinterpret_cast<float4 *>(peers.buf[r]) + seg_index(sg, i0 + u * stride), acc);
            }
        }
    }
    rank_barrier_end(peers, rank, world, launch);   // (its release covers every thread's stores above: cumulative through __syncthreads)
}

// ---------------------------------------------------------------------------------------------------------------
// Push form (eg_push_target, include/edgegs.h): the backward's own gradient stores carry the reduce-scatter.
// Rank o OWNS the Gaussians [o * per, (o + 1) * per); its staging area holds one slot per source rank,
//     slot s = means [3 per] | scales [3 per] | quats [4 per] | opacities [per]   (floats, Gaussian index relative to o * per),
// and eg_splat_bwd_push / eg_project_bwd_push of rank s store the gradients of an owned Gaussian straight into slot s of
// its owner (plain 128-byte-coalesced peer stores, fire-and-forget, spread over the whole backward).  What is left for
// this kernel, after a rank barrier: the owner adds its `world` LOCAL slots in rank order (one rank computes each sum:
// every replica receives the same bits) and broadcasts the sum into the flat gradient buffer of every rank
// (multimem.st through the switch, or `world` peer stores).  Per GPU the links carry one buffer length inbound
// during the backward (hidden) and one inbound here, against two exposed lengths each way for the pull form above.
// ---------------------------------------------------------------------------------------------------------------
struct PushSegs {
    long long src4[4];    // first float4 of each segment inside a slot
    long long dst4[4];    // first float4 of the owner's part of each segment in the flat buffer
    long long first4[5];  // prefix of the segment lengths (float4) of the owner's range
    long long slot4;      // slot stride (float4)
};

template <bool MULTICAST, int PU>
__global__ void __launch_bounds__(AR_THREADS) push_reduce_kernel(const ArPeers peers, float *__restrict__ mc_buf,
                                                                 const float *__restrict__ stage, const PushSegs sg,
                                                                 const int rank, const int world) {
    const uint32_t launch = rank_barrier_begin(peers, rank, world);   // every peer's backward (earlier on its stream) has stored into my slots
    const long long n4 = sg.first4[4];
    const long long stride = (long long)gridDim.x * AR_THREADS;
    const float4 *st4 = reinterpret_cast<const float4 *>(stage);
    for (long long i0 = (long long)blockIdx.x * AR_THREADS + threadIdx.x; i0 < n4; i0 += PU * stride) {
        float4 v[PU][AR_MAX_RANKS];
        long long dst[PU];
#pragma unroll
        for (int u = 0; u < PU; ++u) {
            const long long i = i0 + u * stride;
            if (i >= n4) continue;
            int s = 0;
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (i >= sg.first4[k]) s = k;
            const long long j = i - sg.first4[s];
            dst[u] = sg.dst4[s] + j;
#pragma unroll
            for (int r = 0; r < AR_MAX_RANKS; ++r)
                if (r < world) v[u][r] = __ldcg(st4 + r * sg.slot4 + sg.src4[s] + j);
        }
#pragma unroll
        for (int u = 0; u < PU; ++u) {
            if (i0 + u * stride >= n4) continue;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < AR_MAX_RANKS; ++r)   // fixed rank order, summed once (by the owner)
                if (r < world) { acc.x += v[u][r].x; acc.y += v[u][r].y; acc.z += v[u][r].z; acc.w += v[u][r].w; }
            if (MULTICAST) {
                mc_st(reinterpret_cast<float *>(reinterpret_cast<float4 *>(mc_buf) + dst[u]), acc);
            } else {
#pragma unroll
                for (int r = 0; r < AR_MAX_RANKS; ++r)
                    if (r < world) __stcg(reinterpret_cast<float4 *>(peers.buf[r]) + dst[u], acc);
            }
        }
    }
    rank_barrier_end(peers, rank, world, launch);   // the slices the peers own have arrived in my buffer (release: see rank_barrier_cta)
}

// an idle rank of a ragged step contributes zeros: clear my slot in every owner's staging area
__global__ void __launch_bounds__(AR_THREADS) push_zero_kernel(const eg_push_target push, const long long slot4) {
    const long long stride = (long long)gridDim.x * AR_THREADS;
    for (int o = 0; o < push.world; ++o) {
        float4 *slot = reinterpret_cast<float4 *>(push.stage[o]) + (long long)push.rank * slot4;
        for (long long i = (long long)blockIdx.x * AR_THREADS + threadIdx.x; i < slot4; i += stride)
            __stcg(slot + i, make_float4(0.f, 0.f, 0.f, 0.f));
    }
}

}  // namespace

// 2 leader blocks + one block per CTA (a launch may use any grid up to the one the area was sized for)
extern "C" int eg_allreduce_flag_words(int grid) { return ((grid > 0 ? grid : 0) + 2) * AR_BLK; }
// EG_AR_BARRIER=cta switches back to one rank barrier per CTA (A/B)
static int ar_leader_mode() {
    static const int mode = (getenv("EG_AR_BARRIER") != nullptr && getenv("EG_AR_BARRIER")[0] == 'c') ? 0 : 1;
    return mode;
}

extern "C" int eg_allreduce_symm_segs(float *const *peer_bufs, float *mc_buf, uint32_t *const *peer_flags, int n_segs,
                                      const int64_t *seg_offsets, const int64_t *seg_counts, int rank, int world, int grid,
                                      void *stream) {
    if (peer_bufs == nullptr || peer_flags == nullptr || world < 1 || world > AR_MAX_RANKS || rank < 0 || rank >= world) {
        eg_set_error("eg_allreduce_symm: bad arguments (world %d, rank %d; at most %d ranks)", world, rank, AR_MAX_RANKS);
        return 1;
    }
    if (n_segs < 1 || n_segs > 4 || seg_offsets == nullptr || seg_counts == nullptr) {
        eg_set_error("eg_allreduce_symm: 1..4 segments");
        return 1;
    }
    ArSegs sg;
    sg.n = n_segs;
    sg.first4[0] = 0;
    for (int k = 0; k < 4; ++k) {
        const long long off = k < n_segs ? seg_offsets[k] : 0, cnt = k < n_segs ? seg_counts[k] : 0;
        if (off < 0 || cnt < 0 || (off & 3) != 0 || (cnt & 3) != 0) {
            eg_set_error("eg_allreduce_symm: segment offsets and counts must be multiples of 4 floats (eg_grad_layout pads to that)");
            return 1;
        }
        sg.off4[k] = off / 4;
        sg.first4[k + 1] = sg.first4[k] + cnt / 4;
    }
    if (world == 1 || sg.first4[n_segs] == 0) return 0;
    if (grid <= 0) grid = 148;  // nothing else runs at the tail of a step: one CTA per SM
    static const int unroll = getenv("EG_AR_UNROLL") ? atoi(getenv("EG_AR_UNROLL")) : 8;  // tuning knob (2 / 4 / 8)
    ArPeers peers;
    peers.leader = ar_leader_mode();
    for (int r = 0; r < AR_MAX_RANKS; ++r) {
        peers.buf[r] = r < world ? peer_bufs[r] : nullptr;
        peers.flags[r] = r < world ? peer_flags[r] : nullptr;
        if (r < world && (peers.buf[r] == nullptr || peers.flags[r] == nullptr || ((uintptr_t)peers.buf[r] & 15))) {
            eg_set_error("eg_allreduce_symm: peer pointer %d missing or not 16-byte aligned", r);
            return 1;
        }
    }
    if (mc_buf != nullptr && ((uintptr_t)mc_buf & 15)) {
        eg_set_error("eg_allreduce_symm: multicast pointer not 16-byte aligned");
        return 1;
    }
    cudaStream_t s = (cudaStream_t)stream;
#define EG_AR_LAUNCH(MC, U) allreduce_kernel<MC, U><<<grid, AR_THREADS, 0, s>>>(peers, mc_buf, sg, rank, world)
    if (mc_buf != nullptr) {
        if (unroll <= 2) EG_AR_LAUNCH(true, 2); else if (unroll <= 4) EG_AR_LAUNCH(true, 4); else EG_AR_LAUNCH(true, 8);
    } else {
        if (unroll <= 2) EG_AR_LAUNCH(false, 2); else EG_AR_LAUNCH(false, 8);
    }
#undef EG_AR_LAUNCH
    return eg_check_launch("eg_allreduce_symm");
}

extern "C" int eg_allreduce_symm(float *const *peer_bufs, float *mc_buf, uint32_t *const *peer_flags, int64_t count,
                                 int rank, int world, int grid, void *stream) {
    const int64_t off = 0;
    return eg_allreduce_symm_segs(peer_bufs, mc_buf, peer_flags, 1, &off, &count, rank, world, grid, stream);
}

// ---- push form: host side ----
extern "C" int eg_exchange_push_per(int n, int world) {
    if (n <= 0 || world <= 0) return 0;
    const long long each = ((long long)n + world - 1) / world;
    return (int)((each + 127) / 128 * 128);   // whole CTAs of eg_splat_bwd (128 Gaussians): one owner per CTA, 512-byte runs
}

extern "C" int64_t eg_exchange_stage_floats(int n, int world) {
    return (int64_t)world * 11 * (int64_t)eg_exchange_push_per(n, world);
}

static int push_args_ok(const char *what, const eg_push_target *push, int n) {
    if (push == nullptr || push->world < 1 || push->world > AR_MAX_RANKS || push->rank < 0 || push->rank >= push->world ||
        push->per <= 0 || (push->per & 3) != 0 || (long long)push->per * push->world < n) {
        eg_set_error("%s: bad push target (world, rank, per must cover n = %d Gaussians; per a multiple of 4)", what, n);
        return 0;
    }
    for (int r = 0; r < push->world; ++r)
        if (push->stage[r] == nullptr || ((uintptr_t)push->stage[r] & 15)) {
            eg_set_error("%s: staging pointer %d missing or not 16-byte aligned", what, r);
            return 0;
        }
    return 1;
}
int eg_push_target_ok(const char *what, const eg_push_target *push, int n) { return push_args_ok(what, push, n); }

extern "C" int eg_exchange_reduce_bcast(const eg_push_target *push, float *const *peer_bufs, float *mc_buf,
                                        uint32_t *const *peer_flags, int n, int grid, void *stream) {
    if (!push_args_ok("eg_exchange_reduce_bcast", push, n)) return 1;
    const int world = push->world, rank = push->rank;
    if (peer_bufs == nullptr || peer_flags == nullptr) {
        eg_set_error("eg_exchange_reduce_bcast: peer_bufs and peer_flags are required");
        return 1;
    }
    if (world == 1) return 0;
    if (grid <= 0) grid = 148;
    ArPeers peers;
    peers.leader = ar_leader_mode();
    for (int r = 0; r < AR_MAX_RANKS; ++r) {
        peers.buf[r] = r < world ? peer_bufs[r] : nullptr;
        peers.flags[r] = r < world ? peer_flags[r] : nullptr;
        if (r < world && (peers.buf[r] == nullptr || peers.flags[r] == nullptr || ((uintptr_t)peers.buf[r] & 15))) {
            eg_set_error("eg_exchange_reduce_bcast: peer pointer %d missing or not 16-byte aligned", r);
            return 1;
        }
    }
    if (mc_buf != nullptr && ((uintptr_t)mc_buf & 15)) {
        eg_set_error("eg_exchange_reduce_bcast: multicast pointer not 16-byte aligned");
        return 1;
    }
    int64_t offs[5];
    eg_grad_layout(n, offs);
    const long long per = push->per, np = (n + 3) / 4 * 4;
    long long cnt = np - (long long)rank * per;   // Gaussians (padded to 4) of my range
    cnt = cnt < 0 ? 0 : (cnt > per ? per : cnt);
    const int w[4] = {3, 3, 4, 1};
    const long long src[4] = {0, 3 * per, 6 * per, 10 * per};
    PushSegs sg;
    sg.first4[0] = 0;
    for (int k = 0; k < 4; ++k) {
        sg.src4[k] = src[k] / 4;
        sg.dst4[k] = (offs[k] + (long long)w[k] * rank * per) / 4;
        sg.first4[k + 1] = sg.first4[k] + (long long)w[k] * cnt / 4;
    }
    sg.slot4 = 11 * per / 4;
    cudaStream_t s = (cudaStream_t)stream;
    if (mc_buf != nullptr)
        push_reduce_kernel<true, 2><<<grid, AR_THREADS, 0, s>>>(peers, mc_buf, push->stage[rank], sg, rank, world);
    else
        push_reduce_kernel<false, 2><<<grid, AR_THREADS, 0, s>>>(peers, mc_buf, push->stage[rank], sg, rank, world);
    return eg_check_launch("eg_exchange_reduce_bcast");
}

extern "C" int eg_exchange_push_zero(const eg_push_target *push, int n, void *stream) {
    if (!push_args_ok("eg_exchange_push_zero", push, n)) return 1;
    push_zero_kernel<<<148, AR_THREADS, 0, (cudaStream_t)stream>>>(*push, 11ll * push->per / 4);
    return eg_check_launch("eg_exchange_push_zero");
}
