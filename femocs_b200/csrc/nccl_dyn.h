// NCCL bound at run time (dlopen by SONAME): when the host process has already loaded a libnccl.so.2 (e.g. the
// one bundled with PyTorch under torchrun) that copy is reused, otherwise the system library is loaded.  Only
// the handful of entry points the partitioned CG needs.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace fb {

struct Nccl {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    const char* why = "";

    static Nccl& get() {
        static Nccl n;
        static bool tried = false;
        if (tried) return n;
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { n.why = "libnccl.so.2 not found"; return n; }
#define FB_SYM(field, name) *(void**) (&n.field) = dlsym(h, name); if (!n.field) { n.why = "missing symbol " name; return n; }
        FB_SYM(GetUniqueId, "ncclGetUniqueId") FB_SYM(CommInitRank, "ncclCommInitRank") FB_SYM(CommDestroy, "ncclCommDestroy")
        FB_SYM(AllReduce, "ncclAllReduce") FB_SYM(AllGather, "ncclAllGather") FB_SYM(Send, "ncclSend") FB_SYM(Recv, "ncclRecv")
        FB_SYM(GroupStart, "ncclGroupStart") FB_SYM(GroupEnd, "ncclGroupEnd") FB_SYM(GetErrorString, "ncclGetErrorString")
#undef FB_SYM
        n.ok = true;
        return n;
    }
};

}  // namespace fb
