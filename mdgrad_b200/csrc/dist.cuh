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
    int (*AllGather)(const void* send, void* recv, size_t sendcount, int dtype, void* comm, cudaStream_t st) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* mdg_nccl();
int mdg_nccl_check(int r, const char* what);

// ---------------------------------------------------------------------------------------------------------------------
// Peer-to-peer step path (NVLink loads / stores into the neighbours' memory, no NCCL kernel on the per-step critical path).
// One DistSync block per rank, in that rank's own memory, mapped into every peer with CUDA IPC; peers write, the owner spins.
//   ke[p][r][0..1]   kinetic energies (v, v + vh) of rank r's slab for the step with sequence parity p
//   ke_flag[p][r]    sequence number of that entry (written after the values, system-scope fence in between)
//   halo_flag[s]     sequence number of the ghost positions the neighbour below (s = 0) / above (s = 1) stored into my array
//   ack_flag[s]      the neighbour below / above has finished reading the ghosts I stored for that sequence number
// ---------------------------------------------------------------------------------------------------------------------
#define MDG_DIST_MAXW 16
struct DistSync {
    double ke[2][MDG_DIST_MAXW][2];
    int    ke_flag[2][MDG_DIST_MAXW];
    int    halo_flag[2];
    int    ack_flag[2];
    int    ticket;             // block counter of k_dist_push (self-resetting)
    int    pad[3];
};
struct mdg_ctx;
int mdg_i_dist_p2p_setup(mdg_ctx* c, cudaStream_t st);
void mdg_i_dist_p2p_release(mdg_ctx* c);
