// eg_comm.cu -- an NCCL communicator below the C ABI: the A/B baseline of the view-sharded step's exchange.
//
// The reference has no multi-GPU path; the sharding is by VIEW: parameters are
// replicated, every rank renders its own camera and the per-view gradients of the flat buffer
// means | scales | quats | opacities (eg_grad_layout) are summed.  The product exchange is this library's own
// kernel over symmetric memory (eg_allreduce.cu: switch-side reduction through NVLink multicast);
// eg_comm_allreduce issues the same sum through NCCL on the caller's stream and is what that kernel is measured
// against.  (Round 1 also had the backward launched in Gaussian ranges with each range's NCCL all-reduce
// overlapped on a side stream; it measured slower than one all-reduce per step at every range count -- DESIGN.md
// section 5 -- and was removed.)
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

struct EgComm {
    nccl_comm_t comm;
    int rank, world;
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
    *comm_out = c;
    return 0;
}

extern "C" int eg_comm_destroy(void *comm) {
    if (comm == nullptr) return 0;
    EgComm *c = reinterpret_cast<EgComm *>(comm);
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
