// comm.cu -- cnvb_comm: NCCL communicator of the multi-GPU entry points.
//
// The reference's only parallelism is rayon inside one process (lib-model-wis/src/lib.rs:144-193,
// cnvetti/src/quick_wis_build_model.rs:148-184); the sharded entry points keep ITS host (one process per
// GPU, any language) and put the exchanges the data flow needs behind the C ABI: the per-pass histogram
// all-reduce of the GC-binned medians, the double-double partial sums of TotalCovSum, the panel
// all-gather and the nearest-distance all-gather of build-model-wis.
//
// NCCL is not linked: the few entry points used are resolved with dlopen / dlsym at the first use, so the
// library loads on a box without NCCL and a torch process shares torch's own copy of it.
#include <dlfcn.h>

#include <mutex>

#include "comm.cuh"

namespace {

// the stable part of nccl.h (2.x) that is needed here
typedef void *nccl_comm_t;
typedef int nccl_result_t;  // ncclSuccess == 0
enum { kNcclChar = 0, kNcclUint32 = 3, kNcclFloat64 = 8, kNcclSum = 0 };
struct nccl_unique_id {
    char internal[128];
};
static_assert(sizeof(nccl_unique_id) == sizeof(cnvb_comm_id), "cnvb_comm_id carries an ncclUniqueId");

struct NcclApi {
    void *handle = nullptr;
    std::string error;
    nccl_result_t (*GetUniqueId)(nccl_unique_id *) = nullptr;
    nccl_result_t (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
    nccl_result_t (*CommDestroy)(nccl_comm_t) = nullptr;
    nccl_result_t (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    nccl_result_t (*Broadcast)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    nccl_result_t (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    nccl_result_t (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    nccl_result_t (*GroupStart)() = nullptr;
    nccl_result_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(nccl_result_t) = nullptr;
    nccl_result_t (*GetVersion)(int *) = nullptr;
};

NcclApi *nccl_api() {  // process-wide, immutable once resolved
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("CNVB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
            api.error = dlerror();
        }
        if (!api.handle) return;
        bool ok = true;
        auto sym = [&](const char *name) {
            void *p = dlsym(api.handle, name);
            if (!p) {
                ok = false;
                api.error = std::string("missing symbol ") + name;
            }
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
        if (!ok) {
            dlclose(api.handle);
            api.handle = nullptr;
        }
    });
    return &api;
}

int fail_nccl(cnvb_ctx *ctx, nccl_result_t r, const char *what) {
    NcclApi *a = nccl_api();
    return cnvb::fail(ctx, CNVB_ERR_COMM, "%s\ncaused by: NCCL error %d: %s", what, (int)r,
                      a->GetErrorString ? a->GetErrorString(r) : "?");
}

#define CNVB_NCCL(ctx, call, what)                             \
    do {                                                       \
        nccl_result_t r__ = (call);                            \
        if (r__ != 0) return fail_nccl((ctx), r__, what);      \
    } while (0)

thread_local std::string g_comm_error;

}  // namespace

namespace cnvb {

int comm_all_reduce_u32(cnvb_comm *c, uint32_t *buf, size_t count) {
    if (!c || c->world == 1 || count == 0) return 0;
    CNVB_NCCL(c->ctx, nccl_api()->AllReduce(buf, buf, count, kNcclUint32, kNcclSum, c->nccl, c->ctx->stream),
              "all-reduce (u32 sum)");
    return 0;
}

int comm_all_reduce_f64(cnvb_comm *c, double *buf, size_t count) {
    if (!c || c->world == 1 || count == 0) return 0;
    CNVB_NCCL(c->ctx, nccl_api()->AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, c->nccl, c->ctx->stream),
              "all-reduce (f64 sum)");
    return 0;
}

int comm_all_gather_v(cnvb_comm *c, const void *send, void *recv, const uint64_t *bytes) {
    cnvb_ctx *ctx = c->ctx;
    uint64_t off = 0;
    if (c->world == 1) {
        if (send != recv && bytes[0])
            CNVB_CUDA(ctx, cudaMemcpyAsync(recv, send, bytes[0], cudaMemcpyDeviceToDevice, ctx->stream), "all-gather");
        return 0;
    }
    NcclApi *a = nccl_api();
    CNVB_NCCL(ctx, a->GroupStart(), "all-gather");
    for (int r = 0; r < c->world; ++r) {
        char *slot = static_cast<char *>(recv) + off;
        if (bytes[r]) {
            const nccl_result_t rc = a->Broadcast(r == c->rank ? send : slot, slot, bytes[r], kNcclChar, r, c->nccl, ctx->stream);
            if (rc != 0) {
                a->GroupEnd();
                return fail_nccl(ctx, rc, "all-gather (broadcast of one rank's block)");
            }
        }
        off += bytes[r];
    }
    CNVB_NCCL(ctx, a->GroupEnd(), "all-gather");
    return 0;
}

int comm_all_to_all_v(cnvb_comm *c, const void *send, const uint64_t *send_off, void *recv, const uint64_t *recv_off) {
    cnvb_ctx *ctx = c->ctx;
    const char *s = static_cast<const char *>(send);
    char *d = static_cast<char *>(recv);
    const int me = c->rank;
    if (send_off[me + 1] - send_off[me] != recv_off[me + 1] - recv_off[me])
        return fail(ctx, CNVB_ERR_INVALID, "all-to-all: the block a rank keeps differs in size on the two sides");
    if (send_off[me + 1] > send_off[me])
        CNVB_CUDA(ctx, cudaMemcpyAsync(d + recv_off[me], s + send_off[me], send_off[me + 1] - send_off[me],
                                       cudaMemcpyDeviceToDevice, ctx->stream), "all-to-all");
    if (c->world == 1) return 0;
    NcclApi *a = nccl_api();
    CNVB_NCCL(ctx, a->GroupStart(), "all-to-all");
    nccl_result_t rc = 0;
    for (int q = 0; q < c->world && rc == 0; ++q) {
        if (q == me) continue;
        if (send_off[q + 1] > send_off[q])
            rc = a->Send(s + send_off[q], send_off[q + 1] - send_off[q], kNcclChar, q, c->nccl, ctx->stream);
        if (rc == 0 && recv_off[q + 1] > recv_off[q])
            rc = a->Recv(d + recv_off[q], recv_off[q + 1] - recv_off[q], kNcclChar, q, c->nccl, ctx->stream);
    }
    if (rc != 0) {
        a->GroupEnd();
        return fail_nccl(ctx, rc, "all-to-all (send / receive)");
    }
    CNVB_NCCL(ctx, a->GroupEnd(), "all-to-all");
    return 0;
}

}  // namespace cnvb

using namespace cnvb;

extern "C" const char *cnvb_comm_create_error(void) { return g_comm_error.c_str(); }

extern "C" int cnvb_comm_nccl_version(int *version) {
    NcclApi *a = nccl_api();
    if (!a->handle) {
        g_comm_error = "NCCL is not available: " + a->error;
        return CNVB_ERR_COMM;
    }
    int v = 0;
    if (a->GetVersion(&v) != 0) return CNVB_ERR_COMM;
    if (version) *version = v;
    return CNVB_OK;
}

extern "C" int cnvb_comm_unique_id(cnvb_comm_id *id) {
    if (!id) return CNVB_ERR_INVALID;
    NcclApi *a = nccl_api();
    if (!a->handle) {
        g_comm_error = "NCCL is not available: " + a->error;
        return CNVB_ERR_COMM;
    }
    nccl_unique_id u;
    const nccl_result_t r = a->GetUniqueId(&u);
    if (r != 0) {
        g_comm_error = std::string("ncclGetUniqueId failed: ") + a->GetErrorString(r);
        return CNVB_ERR_COMM;
    }
    memcpy(id->bytes, u.internal, sizeof(u.internal));
    return CNVB_OK;
}

extern "C" int cnvb_comm_init_rank(cnvb_ctx *ctx, const cnvb_comm_id *id, int rank, int world, cnvb_comm **out) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return fail(ctx, CNVB_ERR_INVALID, "cnvb_comm_init_rank: bad argument");
    *out = nullptr;
    NcclApi *a = nccl_api();
    if (!a->handle) return fail(ctx, CNVB_ERR_COMM, "NCCL is not available\ncaused by: %s", a->error.c_str());
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    nccl_unique_id u;
    memcpy(u.internal, id->bytes, sizeof(u.internal));
    nccl_comm_t comm = nullptr;
    CNVB_NCCL(ctx, a->CommInitRank(&comm, world, u, rank), "ncclCommInitRank");
    cnvb_comm *c = new cnvb_comm();
    c->ctx = ctx;
    c->nccl = comm;
    c->rank = rank;
    c->world = world;
    c->own = true;
    *out = c;
    return CNVB_OK;
}

extern "C" int cnvb_comm_from_nccl(cnvb_ctx *ctx, void *nccl_comm, int rank, int world, cnvb_comm **out) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_comm))
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_comm_from_nccl: bad argument");
    if (world > 1 && !nccl_api()->handle)
        return fail(ctx, CNVB_ERR_COMM, "NCCL is not available\ncaused by: %s", nccl_api()->error.c_str());
    cnvb_comm *c = new cnvb_comm();
    c->ctx = ctx;
    c->nccl = nccl_comm;
    c->rank = rank;
    c->world = world;
    c->own = false;
    *out = c;
    return CNVB_OK;
}

extern "C" void cnvb_comm_destroy(cnvb_comm *c) {
    if (!c) return;
    if (c->own && c->nccl) {
        cudaSetDevice(c->ctx->device);
        cudaStreamSynchronize(c->ctx->stream);
        nccl_api()->CommDestroy(c->nccl);
    }
    delete c;
}

extern "C" int cnvb_comm_rank(const cnvb_comm *c) { return c ? c->rank : 0; }
extern "C" int cnvb_comm_world(const cnvb_comm *c) { return c ? c->world : 1; }

extern "C" int cnvb_comm_all_gather_v(cnvb_comm *c, const void *send, void *recv, const uint64_t *bytes_per_rank) {
    if (!c) return CNVB_ERR_INVALID;
    if (!recv || !bytes_per_rank || (bytes_per_rank[c->rank] && !send))
        return fail(c->ctx, CNVB_ERR_INVALID, "cnvb_comm_all_gather_v: null argument");
    CNVB_CUDA(c->ctx, cudaSetDevice(c->ctx->device), "selecting device");
    return comm_all_gather_v(c, send, recv, bytes_per_rank);
}

extern "C" int cnvb_comm_all_to_all_v(cnvb_comm *c, const void *send, const uint64_t *send_off, void *recv,
                                      const uint64_t *recv_off) {
    if (!c) return CNVB_ERR_INVALID;
    if (!send_off || !recv_off || !send || !recv) return fail(c->ctx, CNVB_ERR_INVALID, "cnvb_comm_all_to_all_v: null argument");
    CNVB_CUDA(c->ctx, cudaSetDevice(c->ctx->device), "selecting device");
    return comm_all_to_all_v(c, send, send_off, recv, recv_off);
}

extern "C" int cnvb_comm_all_reduce_u32(cnvb_comm *c, uint32_t *buf, uint64_t count) {
    if (!c) return CNVB_ERR_INVALID;
    if (count && !buf) return fail(c->ctx, CNVB_ERR_INVALID, "cnvb_comm_all_reduce_u32: null argument");
    CNVB_CUDA(c->ctx, cudaSetDevice(c->ctx->device), "selecting device");
    return comm_all_reduce_u32(c, buf, count);
}
