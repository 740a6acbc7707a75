// eg_comm.cu -- the view-sharded step's exchange below the C ABI: the Gaussian-major backward launched in Gaussian
// ranges, each range's gradients all-reduced (NCCL, NVLink / NVSwitch) on a side stream while the next range is
// still being computed.
//
// The reference has no multi-GPU path; the sharding is by VIEW: parameters are
// replicated, every rank renders its own camera and the per-view gradients of the flat buffer
// means | scales | quats | opacities (11 N floats) are summed.  eg_splat_bwd gives every Gaussian exactly one
// owner and writes its gradients once, so the gradients of Gaussians [g0, g1) are final as soon as that range's
// launch has drained -- nothing of the remaining backward touches them.  Issuing range launches, event
// fork/joins and grouped collectives from Python costs more host time than the exchange it hides (measured), so
// the whole sequence is one C call.
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 the process already carries -- PyTorch bundles it), so
// libedgegs.so has no link-time dependency on it and still loads on a box without NCCL.
#include <dlfcn.h>

#include "eg_common.cuh"

namespace {

struct NcclId {
    char internal[128];  // NCCL_UNIQUE_ID_BYTES
};
typedef void *nccl_comm_t;
enum { kNcclSum = 0, kNcclFloat32 = 7 };  // ncclRedOp_t / ncclDataType_t values, stable across NCCL 2.x

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, NcclId, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

NcclApi g_nccl;

bool load_nccl() {
    if (g_nccl.handle != nullptr) return true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) {
        eg_set_error("eg_comm: cannot load libnccl.so.2 (%s)", dlerror());
        return false;
    }
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = (int (*)(NcclId *))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(nccl_comm_t *, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (int (*)(nccl_comm_t))dlsym(h, "ncclCommDestroy");
    a.AllReduce = (int (*)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    a.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    a.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    a.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GroupStart || !a.GroupEnd) {
        eg_set_error("eg_comm: libnccl.so.2 lacks a required symbol");
        return false;
    }
    g_nccl = a;
    return true;
}

int nccl_fail(const char *what, int rc) {
    eg_set_error("%s: NCCL error %d (%s)", what, rc, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return 3;
}

constexpr int kMaxRanges = 16;

struct EgComm {
    nccl_comm_t comm;
    int rank, world;
    cudaEvent_t ev[kMaxRanges + 1];
};

}  // namespace

extern "C" int eg_comm_unique_id(void *id128) {
    if (id128 == nullptr) {
        eg_set_error("eg_comm_unique_id: null buffer");
        return 1;
    }
    if (!load_nccl()) return 1;
    if (int rc = g_nccl.GetUniqueId(reinterpret_cast<NcclId *>(id128))) return nccl_fail("ncclGetUniqueId", rc);
    return 0;
}

extern "C" int eg_comm_init(const void *id128, int rank, int world, void **comm_out) {
    if (id128 == nullptr || comm_out == nullptr || world < 1 || rank < 0 || rank >= world) {
        eg_set_error("eg_comm_init: bad arguments");
        return 1;
    }
    if (!load_nccl()) return 1;
    EgComm *c = new EgComm();
    c->rank = rank;
    c->world = world;
    NcclId id = *reinterpret_cast<const NcclId *>(id128);
    if (int rc = g_nccl.CommInitRank(&c->comm, world, id, rank)) {
        delete c;
        return nccl_fail("ncclCommInitRank", rc);
    }
    for (int i = 0; i <= kMaxRanges; ++i) {
        if (cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming) != cudaSuccess) {
            eg_set_error("eg_comm_init: cudaEventCreate failed");
            return 2;
        }
    }
    *comm_out = c;
    return 0;
}

extern "C" int eg_comm_destroy(void *comm) {
    if (comm == nullptr) return 0;
    EgComm *c = reinterpret_cast<EgComm *>(comm);
    for (int i = 0; i <= kMaxRanges; ++i) cudaEventDestroy(c->ev[i]);
    if (g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
    return 0;
}

extern "C" int eg_comm_allreduce(float *buf, int64_t count, void *comm, void *stream) {
    if (buf == nullptr || comm == nullptr || count < 0) {
        eg_set_error("eg_comm_allreduce: bad arguments");
        return 1;
    }
    EgComm *c = reinterpret_cast<EgComm *>(comm);
    if (c->world == 1 || count == 0) return 0;
    if (int rc = g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat32, kNcclSum, c->comm, (cudaStream_t)stream))
        return nccl_fail("ncclAllReduce", rc);
    return 0;
}

extern "C" int eg_splat_bwd_allreduce(const eg_config *cfg, const float *means, const float *quats,
                                      const float *scales, const float *opacities, const float *viewmat,
                                      const float *K, const float *rec, const int32_t *gint, const float *wpix,
                                      float seed_scale, const uint32_t *last_depth, const int32_t *last_gid,
                                      const int32_t *tile_stop, const int32_t *status, float *grads,
                                      float *absgrad_accum, int n_ranges, void *comm, void *comm_stream,
                                      void *stream) {
    if (cfg == nullptr || grads == nullptr || comm == nullptr) {
        eg_set_error("eg_splat_bwd_allreduce: cfg, grads and comm are required");
        return 1;
    }
    EgComm *c = reinterpret_cast<EgComm *>(comm);
    const int n = cfg->n;
    if (n <= 0) return 0;
    if (n_ranges < 1) n_ranges = 1;
    if (n_ranges > kMaxRanges) n_ranges = kMaxRanges;
    // ranges whose boundaries are multiples of the 128 Gaussians one CTA of eg_splat_bwd owns
    int per = (n + n_ranges - 1) / n_ranges;
    per = (per + 127) / 128 * 128;
    float *v_means = grads, *v_scales = grads + 3ll * n, *v_quats = grads + 6ll * n, *v_opac = grads + 10ll * n;
    cudaStream_t main_s = (cudaStream_t)stream, comm_s = (cudaStream_t)comm_stream;
    int r = 0;
    for (int g0 = 0; g0 < n; g0 += per, ++r) {
        const int g1 = g0 + per < n ? g0 + per : n;
        if (int rc = eg_splat_bwd(cfg, means, quats, scales, opacities, viewmat, K, rec, gint, wpix, seed_scale,
                                  last_depth, last_gid, tile_stop, status, g0, g1, nullptr, v_means, v_quats, v_scales,
                                  v_opac, absgrad_accum, stream))
            return rc;
        if (c->world == 1) continue;
        if (cudaEventRecord(c->ev[r], main_s) != cudaSuccess || cudaStreamWaitEvent(comm_s, c->ev[r], 0) != cudaSuccess) {
            eg_set_error("eg_splat_bwd_allreduce: event fork failed");
            return 2;
        }
        // the four slices of the range, one grouped launch
        const size_t cnt = (size_t)(g1 - g0);
        int rc = g_nccl.GroupStart();
        if (!rc) rc = g_nccl.AllReduce(v_means + 3ll * g0, v_means + 3ll * g0, 3 * cnt, kNcclFloat32, kNcclSum, c->comm, comm_s);
        if (!rc) rc = g_nccl.AllReduce(v_scales + 3ll * g0, v_scales + 3ll * g0, 3 * cnt, kNcclFloat32, kNcclSum, c->comm, comm_s);
        if (!rc) rc = g_nccl.AllReduce(v_quats + 4ll * g0, v_quats + 4ll * g0, 4 * cnt, kNcclFloat32, kNcclSum, c->comm, comm_s);
        if (!rc) rc = g_nccl.AllReduce(v_opac + g0, v_opac + g0, cnt, kNcclFloat32, kNcclSum, c->comm, comm_s);
        const int rc2 = g_nccl.GroupEnd();
        if (rc || rc2) return nccl_fail("ncclAllReduce", rc ? rc : rc2);
    }
    if (c->world > 1) {  // the optimizer / next step needs the reduced gradients
        if (cudaEventRecord(c->ev[kMaxRanges], comm_s) != cudaSuccess ||
            cudaStreamWaitEvent(main_s, c->ev[kMaxRanges], 0) != cudaSuccess) {
            eg_set_error("eg_splat_bwd_allreduce: event join failed");
            return 2;
        }
    }
    return 0;
}
