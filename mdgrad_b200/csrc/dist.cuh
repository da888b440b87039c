// dist.cuh - NCCL function table (resolved at run time from the NCCL torch loaded) - see dist.cu
#pragma once
#include <stddef.h>

struct NcclUid { char b[128]; };

// enums as in nccl.h (stable ABI values)
#define MDG_NCCL_SUM 0
#define MDG_NCCL_MAX 2
#define MDG_NCCL_INT32 2
#define MDG_NCCL_FLOAT32 7
#define MDG_NCCL_FLOAT64 8

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUid*) = nullptr;
    int (*CommInitRank)(void** comm, int nranks, NcclUid id, int rank) = nullptr;
    int (*CommDestroy)(void* comm) = nullptr;
    int (*Send)(const void* buf, size_t count, int dtype, int peer, void* comm, cudaStream_t st) = nullptr;
    int (*Recv)(void* buf, size_t count, int dtype, int peer, void* comm, cudaStream_t st) = nullptr;
    int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t st) = nullptr;
    int (*Broadcast)(const void* send, void* recv, size_t count, int dtype, int root, void* comm, cudaStream_t st) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* mdg_nccl();
int mdg_nccl_check(int r, const char* what);
