#include "comm.h"

#include <dlfcn.h>

#include <cstring>
#include <stdexcept>
#include <string>

namespace cslam {
namespace {
// Minimal NCCL ABI (nccl.h 2.x): ncclUniqueId is 128 opaque bytes, results are ints,
// ncclDouble = 8 (ncclFloat64), ncclSum = 0, ncclMax = 2.
struct NcclId {
    char internal[128];
};
typedef int (*GetUniqueIdFn)(NcclId*);
typedef int (*CommInitRankFn)(void**, int, NcclId, int);
typedef int (*CommDestroyFn)(void*);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*BroadcastFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);

struct Api {
    void* handle = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    AllReduceFn all_reduce = nullptr;
    BroadcastFn broadcast = nullptr;
    AllGatherFn all_gather = nullptr;
    GetErrorStringFn error_string = nullptr;
};

Api& api() {
    static Api a;
    if (a.handle) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) throw std::runtime_error(std::string("cannot load libnccl: ") + dlerror());
    a.get_unique_id = (GetUniqueIdFn)dlsym(a.handle, "ncclGetUniqueId");
    a.comm_init_rank = (CommInitRankFn)dlsym(a.handle, "ncclCommInitRank");
    a.comm_destroy = (CommDestroyFn)dlsym(a.handle, "ncclCommDestroy");
    a.all_reduce = (AllReduceFn)dlsym(a.handle, "ncclAllReduce");
    a.broadcast = (BroadcastFn)dlsym(a.handle, "ncclBroadcast");
    a.all_gather = (AllGatherFn)dlsym(a.handle, "ncclAllGather");
    a.error_string = (GetErrorStringFn)dlsym(a.handle, "ncclGetErrorString");
    if (!a.get_unique_id || !a.comm_init_rank || !a.comm_destroy || !a.all_reduce || !a.broadcast || !a.all_gather)
        throw std::runtime_error("libnccl is missing required symbols");
    return a;
}
void check(int rc, const char* what) {
    if (rc != 0) {
        Api& a = api();
        throw std::runtime_error(std::string(what) + ": " + (a.error_string ? a.error_string(rc) : "nccl error"));
    }
}
}  // namespace

void comm_unique_id(uint8_t id[128]) {
    NcclId nid;
    check(api().get_unique_id(&nid), "ncclGetUniqueId");
    std::memcpy(id, nid.internal, 128);
}
void* comm_create(int n_ranks, int rank, const uint8_t id[128]) {
    NcclId nid;
    std::memcpy(nid.internal, id, 128);
    void* comm = nullptr;
    check(api().comm_init_rank(&comm, n_ranks, nid, rank), "ncclCommInitRank");
    return comm;
}
void comm_destroy(void* comm) {
    if (comm) api().comm_destroy(comm);
}
void comm_allreduce_sum(void* comm, double* buf, size_t count, cudaStream_t s) {
    check(api().all_reduce(buf, buf, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm, s), "ncclAllReduce(sum)");
}
void comm_allreduce_max(void* comm, double* buf, size_t count, cudaStream_t s) {
    check(api().all_reduce(buf, buf, count, /*ncclFloat64*/ 8, /*ncclMax*/ 2, comm, s), "ncclAllReduce(max)");
}
void comm_broadcast(void* comm, double* buf, size_t count, int root, cudaStream_t s) {
    check(api().broadcast(buf, buf, count, /*ncclFloat64*/ 8, root, comm, s), "ncclBroadcast");
}
void comm_allgather_bytes(void* comm, void* buf, size_t chunk_bytes, int rank, cudaStream_t s) {
    check(api().all_gather(static_cast<const char*>(buf) + size_t(rank) * chunk_bytes, buf, chunk_bytes, /*ncclUint8*/ 1, comm, s),
          "ncclAllGather");
}
}  // namespace cslam
