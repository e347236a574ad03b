// collective.cpp -- the one collective of the path through the C ABI: an NCCL sum all-reduce of the
// int64[4] counter block {n_records, n_low_score, n_pass, n_cg} (records shard by contig,
// nothing is exchanged while they are filtered).  NCCL is bound at run time (dlsym on the process first --
// a torch process has already loaded its own libnccl -- then dlopen("libnccl.so.2")), so the library links
// and loads on a box without NCCL and a host that never calls these entry points needs none.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <string>

#include "../../include/ugvc_b200.h"

struct ugvc_ctx;
int ugvc_ctx_device(const ugvc_ctx* ctx);
int ugvc_ctx_fail(ugvc_ctx* ctx, int code, const char* msg);
cudaStream_t ugvc_ctx_default_stream(ugvc_ctx* ctx);

namespace {
// the few NCCL declarations used (nccl.h: ncclUniqueId is 128 opaque bytes, ncclInt64 = 4, ncclSum = 0)
typedef int (*fn_get_unique_id)(void* id128);
typedef int (*fn_comm_init_rank)(void** comm, int nranks, struct Id128 id, int rank);
struct Id128 { char b[128]; };
typedef int (*fn_comm_destroy)(void* comm);
typedef int (*fn_all_reduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t st);
typedef const char* (*fn_error_string)(int);

struct Nccl {
    bool tried = false, ok = false;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_error_string error_string = nullptr;
    std::string why;
} g_nccl;
std::mutex g_mu;

void* sym(void* h, const char* name) { return dlsym(h ? h : RTLD_DEFAULT, name); }

bool load_nccl() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_nccl.tried) return g_nccl.ok;
    g_nccl.tried = true;
    void* h = nullptr;
    if (!sym(nullptr, "ncclAllReduce")) {
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            g_nccl.why = std::string("NCCL is not loaded in this process and libnccl.so.2 cannot be opened: ") + dlerror();
            return false;
        }
    }
    g_nccl.get_unique_id = (fn_get_unique_id)sym(h, "ncclGetUniqueId");
    g_nccl.comm_init_rank = (fn_comm_init_rank)sym(h, "ncclCommInitRank");
    g_nccl.comm_destroy = (fn_comm_destroy)sym(h, "ncclCommDestroy");
    g_nccl.all_reduce = (fn_all_reduce)sym(h, "ncclAllReduce");
    g_nccl.error_string = (fn_error_string)sym(h, "ncclGetErrorString");
    g_nccl.ok = g_nccl.get_unique_id && g_nccl.comm_init_rank && g_nccl.comm_destroy && g_nccl.all_reduce;
    if (!g_nccl.ok) g_nccl.why = "libnccl lacks one of ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllReduce";
    return g_nccl.ok;
}

int nccl_fail(ugvc_ctx* ctx, const char* what, int rc) {
    std::string m = std::string(what) + ": " + (g_nccl.error_string ? g_nccl.error_string(rc) : "NCCL error") + " (" +
                    std::to_string(rc) + ")";
    return ugvc_ctx_fail(ctx, UGVC_E_CUDA, m.c_str());
}
}  // namespace

extern "C" int ugvc_nccl_unique_id(uint8_t id[UGVC_NCCL_ID_BYTES]) {
    if (!id) return UGVC_E_ARG;
    if (!load_nccl()) return ugvc_ctx_fail(nullptr, UGVC_E_STATE, g_nccl.why.c_str());
    const int rc = g_nccl.get_unique_id(id);
    return rc ? nccl_fail(nullptr, "ncclGetUniqueId", rc) : UGVC_OK;
}

extern "C" int ugvc_nccl_comm_init(ugvc_ctx* ctx, const uint8_t id[UGVC_NCCL_ID_BYTES], int world_size, int rank, void** out_comm) {
    if (!ctx || !id || !out_comm || world_size < 1 || rank < 0 || rank >= world_size) return ugvc_ctx_fail(ctx, UGVC_E_ARG, "nccl_comm_init: bad arguments");
    if (!load_nccl()) return ugvc_ctx_fail(ctx, UGVC_E_STATE, g_nccl.why.c_str());
    if (cudaSetDevice(ugvc_ctx_device(ctx)) != cudaSuccess) return ugvc_ctx_fail(ctx, UGVC_E_CUDA, "cudaSetDevice failed");
    Id128 u;
    memcpy(u.b, id, sizeof(u.b));
    void* comm = nullptr;
    const int rc = g_nccl.comm_init_rank(&comm, world_size, u, rank);
    if (rc) return nccl_fail(ctx, "ncclCommInitRank", rc);
    *out_comm = comm;
    return UGVC_OK;
}

extern "C" int ugvc_nccl_comm_destroy(void* comm) {
    if (!comm) return UGVC_OK;
    if (!load_nccl()) return UGVC_E_STATE;
    return g_nccl.comm_destroy(comm) ? UGVC_E_CUDA : UGVC_OK;
}

extern "C" int ugvc_counts_allreduce(ugvc_ctx* ctx, void* nccl_comm, int64_t out_counts[4], void* stream) {
    if (!ctx || !nccl_comm) return ugvc_ctx_fail(ctx, UGVC_E_ARG, "counts_allreduce: NULL context or communicator");
    if (!load_nccl()) return ugvc_ctx_fail(ctx, UGVC_E_STATE, g_nccl.why.c_str());
    if (cudaSetDevice(ugvc_ctx_device(ctx)) != cudaSuccess) return ugvc_ctx_fail(ctx, UGVC_E_CUDA, "cudaSetDevice failed");
    int64_t* d_counts = nullptr;
    int rc = ugvc_counts_device_ptr(ctx, &d_counts);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : ugvc_ctx_default_stream(ctx);
    if (!stream) {  // the counters are written by kernels of every lane: order the collective after all of them
        if (cudaDeviceSynchronize() != cudaSuccess) return ugvc_ctx_fail(ctx, UGVC_E_CUDA, "cudaDeviceSynchronize failed");
    }
    rc = g_nccl.all_reduce(d_counts, d_counts, 4, /*ncclInt64*/ 4, /*ncclSum*/ 0, nccl_comm, st);
    if (rc) return nccl_fail(ctx, "ncclAllReduce", rc);
    if (out_counts) {  // blocking read of the reduced block
        if (cudaMemcpyAsync(out_counts, d_counts, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            return ugvc_ctx_fail(ctx, UGVC_E_CUDA, "counts_allreduce: reading the reduced counters failed");
    }
    return UGVC_OK;
}
