#include "comm.h"

#include <dlfcn.h>

#include <mutex>
#include <stdexcept>
#include <string>

namespace btrc {

namespace {

// the handful of NCCL entry points used, with the ABI of nccl.h (2.x): opaque comm pointer, 128-byte id by value
struct NcclId { char internal[kNcclIdBytes]; };
using NcclComm = void *;
enum { kNcclSuccess = 0, kNcclFloat = 7, kNcclSum = 0 };

struct NcclApi {
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

NcclApi &api() {
    static NcclApi a;
    static std::once_flag once;
    static std::string error;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { error = std::string("NCCL is not available: ") + dlerror(); return; }
        auto sym = [&](const char *n) { void *p = dlsym(h, n); if (!p) error = std::string("NCCL symbol missing: ") + n; return p; };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.Reduce = reinterpret_cast<decltype(a.Reduce)>(sym("ncclReduce"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    });
    if (!error.empty()) throw std::runtime_error(error);
    return a;
}

void check(int rc, const char *what) {
    if (rc != kNcclSuccess) throw std::runtime_error(std::string(what) + ": " + api().GetErrorString(rc));
}

}  // namespace

struct Comm {
    NcclComm comm = nullptr;
    int n_ranks = 1, rank = 0;
};

void nccl_unique_id(uint8_t out[kNcclIdBytes]) {
    NcclId id;
    check(api().GetUniqueId(&id), "ncclGetUniqueId");
    for (size_t i = 0; i < kNcclIdBytes; ++i) out[i] = (uint8_t)id.internal[i];
}

Comm *comm_create(int n_ranks, int rank, const uint8_t id_bytes[kNcclIdBytes]) {
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) throw std::runtime_error("comm_create: rank out of range");
    Comm *c = new Comm();
    c->n_ranks = n_ranks; c->rank = rank;
    if (n_ranks == 1) return c;  // a single rank reduces onto itself: no communicator needed
    NcclId id;
    for (size_t i = 0; i < kNcclIdBytes; ++i) id.internal[i] = (char)id_bytes[i];
    try {
        check(api().CommInitRank(&c->comm, n_ranks, id, rank), "ncclCommInitRank");
    } catch (...) { delete c; throw; }
    return c;
}

void comm_destroy(Comm *c) {
    if (!c) return;
    if (c->comm) api().CommDestroy(c->comm);
    delete c;
}

int comm_rank(const Comm *c) { return c->rank; }
int comm_size(const Comm *c) { return c->n_ranks; }

void comm_reduce_sum(Comm *c, float *const *bufs, const size_t *counts, int n_bufs, int root, cudaStream_t stream) {
    if (!c) throw std::runtime_error("comm_reduce_sum: null communicator");
    if (root < 0 || root >= c->n_ranks) throw std::runtime_error("comm_reduce_sum: root out of range");
    if (c->n_ranks == 1) return;
    NcclApi &a = api();
    check(a.GroupStart(), "ncclGroupStart");
    for (int k = 0; k < n_bufs; ++k)
        if (bufs[k] && counts[k]) check(a.Reduce(bufs[k], bufs[k], counts[k], kNcclFloat, kNcclSum, root, c->comm, stream), "ncclReduce");
    check(a.GroupEnd(), "ncclGroupEnd");
}

}  // namespace btrc
