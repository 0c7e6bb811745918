// misa_md_b200/csrc/nccl_dl.cuh -- NCCL entry points resolved with dlopen at comm-init time.
#pragma once
#include <dlfcn.h>
#include <stdlib.h>
#include <string>
#include "util.cuh"

// -------------------------------------------------------------------------------------------------
// NCCL (loaded lazily with dlopen so the library itself carries no link-time NCCL dependency)
// -------------------------------------------------------------------------------------------------
typedef struct { char internal[128]; } nccl_uid;
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(nccl_uid *) = nullptr;
    int (*CommInitRank)(void **, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static const int kNcclDouble = 8, kNcclInt32 = 2, kNcclUint64 = 5, kNcclInt8 = 0, kNcclSum = 0, kNcclMax = 2, kNcclMin = 3;

static int nccl_load() {
    if (g_nccl.h) return 0;
    const char *names[] = {getenv("MISA_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm) continue;
        g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    REQ(g_nccl.h, MISA_B200_ENCCL, "cannot dlopen libnccl.so.2 (set MISA_B200_NCCL_LIB)");
#define SYM(field, name)                                                   \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name);                     \
    REQ(g_nccl.field, MISA_B200_ENCCL, std::string("missing NCCL symbol ") + name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return 0;
}
#define NC(call)                                                                                                   \
    do {                                                                                                           \
        int _r = (call);                                                                                           \
        if (_r != 0) return fail(MISA_B200_ENCCL, std::string(#call) + ": " + g_nccl.GetErrorString(_r));          \
    } while (0)

