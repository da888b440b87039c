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
#define MDG_DIST_MAXLAY 2048
struct DistSync {
    double ke[2][MDG_DIST_MAXW][2];
    int    ke_flag[2][MDG_DIST_MAXW];
    int    halo_flag[2];
    int    ack_flag[2];
    int    ticket;             // block counter of k_dist_push / k_dist_push_state (self-resetting)
    int    pad[3];             // [0] set-up agreement, [1] latched spin time-out
    int    pos_flag[2];        // pull path: the positions of the neighbour below / above are final for the step with this sequence number
    int    pull_flag[2];       // pull path: the neighbour below / above has copied MY boundary layer of that step (I may overwrite positions)
    int    ba_ticket, pull_ticket;   // block counters of the integrator kernels / k_dist_pull (self-resetting)
    int    push_started;       // (local) sequence number of the push kernel that has started on this GPU - gates the interior force launch
    int    pad3;
    int    state_flag[2];      // rebuild: the neighbour below / above stored its two boundary layers of (q, v, vh) for this sequence number
    int    lay_flag[2][MDG_DIST_MAXW];   // rebuild: rank r's per-layer atom totals for the rebuild with parity p have arrived
    int    lay[2][MDG_DIST_MAXLAY];      // ... the totals, each layer written by its owner into every rank's table
    unsigned long long dbg[16];          // (MDG_TIMELINE) %globaltimer stamps of the last push / wait kernels on this GPU
};
struct PeerTab { DistSync* s[MDG_DIST_MAXW]; };
struct mdg_ctx;
int mdg_i_dist_p2p_setup(mdg_ctx* c, cudaStream_t st);
void mdg_i_dist_p2p_release(mdg_ctx* c);

#if defined(__CUDACC__) || defined(MDG_EMU)
// load from a peer's memory that a remote kernel has just written: never from a stale L1 line
__device__ __forceinline__ float4 mdg_ld_peer(const float4* p) {
#ifdef MDG_EMU
    return *p;
#else
    return __ldcv(p);
#endif
}
__device__ __forceinline__ int vload_i(const int* p) { return *(const volatile int*)p; }
__device__ __forceinline__ void vstore_i(int* p, int v) { *(volatile int*)p = v; }
// Bounded spin (a peer that died must not hang this GPU): after 60 s the wait gives up and latches *timeout_flag, which the
// host turns into an error at the end of the epoch.
__device__ __forceinline__ unsigned long long mdg_globaltimer_ns() {
#ifdef MDG_EMU
    return 0ull;
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
#endif
}
__device__ __forceinline__ void spin_until_ge(const int* p, int v, int* timeout_flag) {
    if (vload_i(p) >= v) return;
    const unsigned long long t0 = mdg_globaltimer_ns();
    unsigned ns = 32;
    while (vload_i(p) < v) {
#ifndef MDG_EMU
        __nanosleep(ns);
        if (ns < 256) ns <<= 1;
#endif
        if (mdg_globaltimer_ns() - t0 > 60ull * 1000000000ull) { *(volatile int*)timeout_flag = 1; break; }   // 60 s
    }
}
#endif
