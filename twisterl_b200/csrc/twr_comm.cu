// twr_comm.cu -- multi-GPU plumbing of the C ABI: one engine per GPU, envs sharded with no data-path
// collective; NCCL over NVLink carries only the per-iteration weight broadcast and the statistics reduction.
//
// NCCL is bound at RUN time (dlopen of libnccl.so.2, or the path in TWISTERL_B200_NCCL_LIB): a host process that already
// loaded NCCL (PyTorch ships one) shares that copy, and a single-GPU host needs no NCCL at all.  Only the prototypes of
// <nccl.h> are used at compile time.
#include "twr_private.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <mutex>
#include <string>

extern "C" int twr_set_error(int code, const char* msg);

namespace {

struct Nccl {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string why;
};
Nccl g_nccl;
std::once_flag g_nccl_once;

const Nccl& nccl() {
    std::call_once(g_nccl_once, [] {
        const char* names[] = {getenv("TWISTERL_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.handle) break;
            g_nccl.why = dlerror();
        }
        if (!g_nccl.handle) return;
        auto sym = [&](const char* s) { void* p = dlsym(g_nccl.handle, s); if (!p) g_nccl.why = std::string("missing symbol ") + s; return p; };
        g_nccl.GetUniqueId = reinterpret_cast<decltype(&ncclGetUniqueId)>(sym("ncclGetUniqueId"));
        g_nccl.CommInitRank = reinterpret_cast<decltype(&ncclCommInitRank)>(sym("ncclCommInitRank"));
        g_nccl.CommDestroy = reinterpret_cast<decltype(&ncclCommDestroy)>(sym("ncclCommDestroy"));
        g_nccl.Broadcast = reinterpret_cast<decltype(&ncclBroadcast)>(sym("ncclBroadcast"));
        g_nccl.AllReduce = reinterpret_cast<decltype(&ncclAllReduce)>(sym("ncclAllReduce"));
        g_nccl.GetErrorString = reinterpret_cast<decltype(&ncclGetErrorString)>(sym("ncclGetErrorString"));
        g_nccl.GetVersion = reinterpret_cast<decltype(&ncclGetVersion)>(sym("ncclGetVersion"));
        if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.Broadcast || !g_nccl.AllReduce ||
            !g_nccl.GetErrorString) {
            dlclose(g_nccl.handle);
            g_nccl.handle = nullptr;
        }
    });
    return g_nccl;
}

int no_nccl() { return twr_set_error(TWR_ERR_UNSUPPORTED, ("NCCL is not available: " + nccl().why).c_str()); }
int nccl_fail(const char* what, ncclResult_t r) {
    return twr_set_error(TWR_ERR_CUDA, (std::string(what) + ": " + nccl().GetErrorString(r)).c_str());
}
#define CU_TRY(expr)                                                                                        \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess) return twr_set_error(TWR_ERR_CUDA, (std::string(#expr) + ": " + cudaGetErrorString(_e)).c_str()); \
    } while (0)

static_assert(sizeof(ncclUniqueId) == TWR_COMM_ID_BYTES, "TWR_COMM_ID_BYTES must be sizeof(ncclUniqueId)");

}  // namespace

extern "C" {

int twr_comm_version(void) {
    int v = 0;
    if (!nccl().handle || !nccl().GetVersion || nccl().GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

int twr_comm_unique_id(uint8_t* id) {
    if (!id) return twr_set_error(TWR_ERR_INVALID, "id is NULL");
    if (!nccl().handle) return no_nccl();
    ncclUniqueId u;
    const ncclResult_t r = nccl().GetUniqueId(&u);
    if (r != ncclSuccess) return nccl_fail("ncclGetUniqueId", r);
    memcpy(id, &u, sizeof(u));
    return TWR_OK;
}

int twr_comm_init(twr_engine* e, const uint8_t* id) {
    if (!e || !id) return twr_set_error(TWR_ERR_INVALID, "NULL argument");
    if (e->comm) return twr_set_error(TWR_ERR_STATE, "twr_comm_init: this engine already has a communicator");
    if (!nccl().handle) return no_nccl();
    CU_TRY(cudaSetDevice(e->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t c = nullptr;
    const ncclResult_t r = nccl().CommInitRank(&c, e->world, u, e->rank);
    if (r != ncclSuccess) return nccl_fail("ncclCommInitRank", r);
    e->comm = c;
    if (!e->d_stats) CU_TRY(cudaMalloc(reinterpret_cast<void**>(&e->d_stats), TWR_COMM_MAX_STATS * sizeof(double)));
    return TWR_OK;
}

void twr_comm_destroy(twr_engine* e) {
    if (!e) return;
    if (e->comm && nccl().handle) {
        cudaSetDevice(e->device);
        cudaStreamSynchronize(e->stream);
        nccl().CommDestroy(static_cast<ncclComm_t>(e->comm));
    }
    e->comm = nullptr;
    if (e->d_stats) { cudaFree(e->d_stats); e->d_stats = nullptr; }
}

int twr_broadcast_weights(twr_engine* e, twr_policy* p, int32_t root) {
    if (!e || !p) return twr_set_error(TWR_ERR_INVALID, "NULL argument");
    if (p->eng != e) return twr_set_error(TWR_ERR_INVALID, "policy belongs to another engine");
    if (root < 0 || root >= e->world) return twr_set_error(TWR_ERR_INVALID, "root out of range");
    CU_TRY(cudaSetDevice(e->device));
    if (e->world > 1) {
        if (!e->comm) return twr_set_error(TWR_ERR_STATE, "twr_broadcast_weights: call twr_comm_init first");
        const ncclResult_t r = nccl().Broadcast(p->d_blob, p->d_blob, (size_t)p->blob_floats, ncclFloat32, root,
                                                static_cast<ncclComm_t>(e->comm), e->stream);
        if (r != ncclSuccess) return nccl_fail("ncclBroadcast", r);
    }
    // every rank (the root included) rebuilds the kernels' operand layouts from the blob it now holds
    return twr_policy_update_from_device(p, p->d_blob);
}

int twr_allreduce_stats(twr_engine* e, double* stats, int32_t n, int32_t op) {
    if (!e || !stats) return twr_set_error(TWR_ERR_INVALID, "NULL argument");
    if (n < 1 || n > TWR_COMM_MAX_STATS) return twr_set_error(TWR_ERR_INVALID, "n must be in 1..TWR_COMM_MAX_STATS");
    if (op != TWR_REDUCE_SUM && op != TWR_REDUCE_MAX) return twr_set_error(TWR_ERR_INVALID, "op must be TWR_REDUCE_SUM or TWR_REDUCE_MAX");
    if (e->world == 1) return TWR_OK;
    if (!e->comm) return twr_set_error(TWR_ERR_STATE, "twr_allreduce_stats: call twr_comm_init first");
    CU_TRY(cudaSetDevice(e->device));
    CU_TRY(cudaMemcpyAsync(e->d_stats, stats, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, e->stream));
    const ncclResult_t r = nccl().AllReduce(e->d_stats, e->d_stats, (size_t)n, ncclFloat64, op == TWR_REDUCE_MAX ? ncclMax : ncclSum,
                                            static_cast<ncclComm_t>(e->comm), e->stream);
    if (r != ncclSuccess) return nccl_fail("ncclAllReduce", r);
    CU_TRY(cudaMemcpyAsync(stats, e->d_stats, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));       // the reduced numbers are read by the host (what a trainer logs); this
                                                    // also keeps a rank from queueing the next step behind a slower peer
    return TWR_OK;
}

}  // extern "C"
