// nccl_dyn.h -- NCCL bound at run time (dlopen), so that libb200lp.so has no link-time NCCL
// dependency: a single-GPU Lisp image does not need NCCL at all, and inside a Python process the
// already-loaded torch-bundled libnccl.so.2 is reused instead of a second copy.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace b200lp {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;

    bool load(const char **why)
    {
        if (handle) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { *why = "dlopen(libnccl.so.2) failed"; return false; }
#define B200LP_SYM(field, name)                                           \
        field = reinterpret_cast<decltype(field)>(dlsym(handle, name));   \
        if (!field) { *why = "missing NCCL symbol " name; return false; }
        B200LP_SYM(GetUniqueId, "ncclGetUniqueId")
        B200LP_SYM(CommInitRank, "ncclCommInitRank")
        B200LP_SYM(CommInitAll, "ncclCommInitAll")
        B200LP_SYM(CommDestroy, "ncclCommDestroy")
        B200LP_SYM(AllGather, "ncclAllGather")
        B200LP_SYM(GroupStart, "ncclGroupStart")
        B200LP_SYM(GroupEnd, "ncclGroupEnd")
        B200LP_SYM(GetErrorString, "ncclGetErrorString")
#undef B200LP_SYM
        return true;
    }
};

} // namespace b200lp
