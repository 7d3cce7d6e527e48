// Row-sharded runs: one NCCL allreduce of the d-vector per transpose sweep.
// NCCL is bound at run time with dlopen so that libkcrn.so has no link-time dependency on it: the
// copy already loaded by the host process (torch's bundled libnccl.so.2) is preferred, which keeps a
// single NCCL in the process.
#include <dlfcn.h>

#include "kcrn_internal.cuh"

namespace kcrn {

namespace {

typedef struct { char internal[128]; } NcclUniqueId;
typedef void *NcclComm;
typedef int (*fn_get_unique_id)(NcclUniqueId *);
typedef int (*fn_comm_init_rank)(NcclComm *, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(NcclComm);
typedef int (*fn_all_reduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t);
typedef const char *(*fn_get_error_string)(int);

constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;      // ncclSum

struct Api {
    void *lib = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_get_error_string get_error_string = nullptr;
};

Api *api() {
    static Api a;
    if (a.lib) return &a;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *nm : names) {
        lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process?
        if (lib) break;
    }
    if (!lib) {
        for (const char *nm : names) {
            lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
    }
    if (!lib) {
        set_error("NCCL: cannot dlopen libnccl.so.2 (%s)", dlerror());
        return nullptr;
    }
    a.get_unique_id = (fn_get_unique_id)dlsym(lib, "ncclGetUniqueId");
    a.comm_init_rank = (fn_comm_init_rank)dlsym(lib, "ncclCommInitRank");
    a.comm_destroy = (fn_comm_destroy)dlsym(lib, "ncclCommDestroy");
    a.all_reduce = (fn_all_reduce)dlsym(lib, "ncclAllReduce");
    a.get_error_string = (fn_get_error_string)dlsym(lib, "ncclGetErrorString");
    if (!a.get_unique_id || !a.comm_init_rank || !a.comm_destroy || !a.all_reduce) {
        set_error("NCCL: required symbols missing in the loaded library");
        return nullptr;
    }
    a.lib = lib;
    return &a;
}

int fail(Api *a, const char *what, int code) {
    set_error("NCCL %s failed: %s", what, a->get_error_string ? a->get_error_string(code) : "?");
    return KCRN_E_NCCL;
}

}  // namespace

int nccl_unique_id(void *out128) {
    Api *a = api();
    if (!a) return KCRN_E_NCCL;
    NcclUniqueId id;
    const int rc = a->get_unique_id(&id);
    if (rc != 0) return fail(a, "ncclGetUniqueId", rc);
    memcpy(out128, &id, sizeof(id));
    return KCRN_OK;
}

int nccl_init(kcrn_problem *h, const void *id128, int rank, int world) {
    Api *a = api();
    if (!a) return KCRN_E_NCCL;
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NcclComm comm = nullptr;
    const int rc = a->comm_init_rank(&comm, world, id, rank);
    if (rc != 0) return fail(a, "ncclCommInitRank", rc);
    h->comm = comm;
    h->rank = rank;
    h->world = world;
    return KCRN_OK;
}

int nccl_allreduce_sum(kcrn_problem *h, double *buf, int64_t count) {
    if (h->world == 1 && h->emulate_sharded) return KCRN_OK;      // one rank: the sum over the ranks is the identity
    Api *a = api();
    if (!a || !h->comm) {
        set_error("NCCL: communicator not initialised");
        return KCRN_E_NCCL;
    }
    const int rc = a->all_reduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, h->comm, h->stream);
    if (rc != 0) return fail(a, "ncclAllReduce", rc);
    h->counters[5] += 1;
    return KCRN_OK;
}

void nccl_destroy(kcrn_problem *h) {
    Api *a = api();
    if (a && h->comm) a->comm_destroy(h->comm);
    h->comm = nullptr;
}

}  // namespace kcrn
