// dist.cu - multi-GPU plumbing of the MD engine: NCCL through a dlopen shim (the process must use the
// ONE NCCL that torch already loaded - SURVEY A9) and the 1-D slab plan.
//
// Decomposition (new design - the reference has no distributed code): the box is cut
// into slabs of whole z-layers of cells IN THE GLOBAL CELL-SORTED INDEX SPACE.  All ranks sort the
// same global state identically at every list rebuild, so a rank's atoms are one contiguous range of
// the sorted arrays and its ghosts (the cell layer below and the layer above) are two more contiguous
// ranges of the same arrays:
//   every step     : ghost-position halo  = 2 x (ncclSend + ncclRecv) of float4 ranges, in place
//                    kinetic energies      = one ncclAllReduce of 2 doubles (the NHC bath is global)
//   every rebuild  : state all-gather      = grouped ncclBroadcast of each rank's range (q, v, vh), then
//                    every rank re-sorts identically and builds the rows of its own slab only.
// Forces on owned atoms use the full (both-direction) list -> no reverse force communication, and
// the per-atom summation order is the single-GPU one -> forces are bit-identical to 1 GPU.
#include <dlfcn.h>
#include <vector>
#include "common.cuh"
#include "dist.cuh"

static NcclApi g_nccl;

static int load_nccl(const char* path) {
    if (g_nccl.handle) return MDG_OK;
    void* h = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { mdg_set_error("dlopen(%s) failed: %s", path ? path : "libnccl.so.2", dlerror()); return MDG_E_NCCL; }
#define SYM(field, name)                                                                    \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                              \
    if (!g_nccl.field) { mdg_set_error("NCCL symbol %s not found", name); return MDG_E_NCCL; }
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(AllReduce, "ncclAllReduce");
    SYM(Broadcast, "ncclBroadcast");
    *(void**)(&g_nccl.AllGather) = dlsym(h, "ncclAllGather");      // optional (peer-to-peer set-up only)
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.handle = h;
    return MDG_OK;
}

NcclApi* mdg_nccl() { return &g_nccl; }

int mdg_nccl_check(int r, const char* what) {
    if (r == 0) return MDG_OK;
    mdg_set_error("NCCL %s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return MDG_E_NCCL;
}

// Layers [zlo, zhi) of rank `rank` when ncz layers are split over `world` ranks (first ncz % world
// ranks get one extra layer).  Pure host logic - callable without a GPU (tests/test_dist_cpu.py).
extern "C" int mdg_slab_plan(int ncz, int world, int rank, int* out4) {
    if (world < 1 || rank < 0 || rank >= world || !out4) { mdg_set_error("mdg_slab_plan: bad arguments"); return MDG_E_BADARG; }
    if (ncz < world) { mdg_set_error("mdg_slab_plan: %d cell layers cannot be split over %d ranks", ncz, world); return MDG_E_BADARG; }
    int base = ncz / world, extra = ncz % world;
    int zlo = rank * base + (rank < extra ? rank : extra);
    int zhi = zlo + base + (rank < extra ? 1 : 0);
    out4[0] = zlo;
    out4[1] = zhi;
    out4[2] = (rank - 1 + world) % world;   // owner of the layer below zlo (periodic)
    out4[3] = (rank + 1) % world;           // owner of the layer zhi (periodic)
    return MDG_OK;
}

extern "C" int mdg_dist_unique_id(const char* nccl_lib_path, char* out128) {
    if (!out128) { mdg_set_error("null output"); return MDG_E_BADARG; }
    MDG_TRY(load_nccl(nccl_lib_path));
    NcclUid id;
    MDG_TRY(mdg_nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId"));
    memcpy(out128, id.b, 128);
    return MDG_OK;
}

extern "C" int mdg_dist_init(mdg_ctx* c, const char* nccl_lib_path, const char* id128, int rank, int world) {
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) { mdg_set_error("mdg_dist_init: bad arguments"); return MDG_E_BADARG; }
    MDG_CUDA(cudaSetDevice(c->device));
    c->dist_rank = rank;
    c->dist_world = world;
    if (world == 1) return MDG_OK;
    MDG_TRY(load_nccl(nccl_lib_path));
    NcclUid id;
    memcpy(id.b, id128, 128);
    void* comm = nullptr;
    MDG_TRY(mdg_nccl_check(g_nccl.CommInitRank(&comm, world, id, rank), "ncclCommInitRank"));
    c->dist_comm = comm;
    const char* pe = getenv("MDG_DIST_P2P");
    c->dist_p2p_off = pe && pe[0] == '0';
    return MDG_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Peer-to-peer set-up: export my two position buffers and my DistSync block with CUDA IPC, all-gather the handles over
// NCCL, map the neighbours' position buffers and every rank's DistSync.  Called by the engine after the first build of an
// epoch (the buffers exist then); repeated when a position buffer was re-allocated (every rank does so at the same call).
// Any failure leaves the NCCL path in place (dist_p2p_off).
// ---------------------------------------------------------------------------------------------------------------------
void mdg_i_dist_p2p_release(mdg_ctx* c) {
#ifndef MDG_EMU
    for (int k = 0; k < c->p2p_n_opened; ++k)
        if (c->p2p_opened[k]) cudaIpcCloseMemHandle(c->p2p_opened[k]);
#endif
    c->p2p_n_opened = 0;
    c->dist_p2p = false;
    for (int r = 0; r < MDG_DIST_MAXW; ++r) c->peer_sync[r] = nullptr;
    for (int s = 0; s < 2; ++s) c->peer_qs[s][0] = c->peer_qs[s][1] = c->peer_v[s] = c->peer_vh[s] = nullptr;
}

int mdg_i_dist_p2p_setup(mdg_ctx* c, cudaStream_t st) {
#ifdef MDG_EMU
    (void)st;
    c->dist_p2p_off = true;      // (the CPU emulation runs ranks as processes over a fake NCCL: no CUDA IPC there)
    return MDG_OK;
#else
    const int W = c->dist_world, me = c->dist_rank;
    if (c->dist_p2p_off || W < 2 || W > MDG_DIST_MAXW || !g_nccl.AllGather) return MDG_OK;
    if (c->dist_p2p && c->p2p_exported[0] == c->qs_buf[0].p && c->p2p_exported[1] == c->qs_buf[1].p && c->p2p_exported[2] == c->v4.p &&
        c->p2p_exported[3] == c->vh4.p)
        return MDG_OK;
    const bool first = c->dsync.p == nullptr;
    mdg_i_dist_p2p_release(c);
    if (first) {
        MDG_TRY(c->dsync.reserve(sizeof(DistSync)));
        MDG_CUDA(cudaMemsetAsync(c->dsync.p, 0, sizeof(DistSync), st));
        MDG_CUDA(cudaEventCreateWithFlags(&c->ev_push, cudaEventDisableTiming));
    }
    struct Pack { cudaIpcMemHandle_t h[5]; int ok; int pad[15]; };      // 5 x 64 + 64 bytes
    Pack mine;
    memset(&mine, 0, sizeof(mine));
    void* ptrs[5] = {c->qs_buf[0].p, c->qs_buf[1].p, c->dsync.p, c->v4.p, c->vh4.p};     // (v4 / vh4: the rebuild's state push)
    mine.ok = 1;
    for (int k = 0; k < 5; ++k)
        if (cudaIpcGetMemHandle(&mine.h[k], ptrs[k]) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
    DevBuf xfer;
    MDG_TRY(xfer.reserve(sizeof(Pack) * (size_t)(W + 1)));
    Pack* d_all = xfer.as<Pack>();
    MDG_CUDA(cudaMemcpyAsync(d_all + W, &mine, sizeof(Pack), cudaMemcpyHostToDevice, st));
    int r = mdg_nccl_check(g_nccl.AllGather(d_all + W, d_all, sizeof(Pack), 0 /* ncclInt8 */, c->dist_comm, st), "AllGather");
    if (r != MDG_OK) { xfer.release(); return r; }
    std::vector<Pack> all((size_t)W);
    MDG_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(Pack) * (size_t)W, cudaMemcpyDeviceToHost, st));
    MDG_CUDA(cudaStreamSynchronize(st));
    xfer.release();
    bool ok = true;
    for (int q = 0; q < W; ++q) ok = ok && all[q].ok == 1;
    const int below = (me - 1 + W) % W, above = (me + 1) % W;
    auto open = [&](const cudaIpcMemHandle_t& h) -> void* {
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (c->p2p_n_opened < 40) c->p2p_opened[c->p2p_n_opened++] = p;
        return p;
    };
    if (ok) {
        for (int q = 0; q < W && ok; ++q) {
            c->peer_sync[q] = (q == me) ? c->dsync.p : open(all[q].h[2]);
            ok = ok && c->peer_sync[q] != nullptr;
        }
        for (int k = 0; k < 2 && ok; ++k) {
            c->peer_qs[0][k] = open(all[below].h[k]);
            c->peer_qs[1][k] = (above == below) ? c->peer_qs[0][k] : open(all[above].h[k]);
            ok = ok && c->peer_qs[0][k] && c->peer_qs[1][k];
        }
        if (ok) {
            c->peer_v[0] = open(all[below].h[3]);
            c->peer_vh[0] = open(all[below].h[4]);
            c->peer_v[1] = (above == below) ? c->peer_v[0] : open(all[above].h[3]);
            c->peer_vh[1] = (above == below) ? c->peer_vh[0] : open(all[above].h[4]);
            ok = c->peer_v[0] && c->peer_vh[0] && c->peer_v[1] && c->peer_vh[1];
        }
    }
    // every rank must take the same path: agree through a 1-int max all-reduce of the failure flag
    int* d_flag = c->dsync.as<int>() + offsetof(DistSync, pad) / sizeof(int);
    int bad = ok ? 0 : 1;
    MDG_CUDA(cudaMemcpyAsync(d_flag, &bad, sizeof(int), cudaMemcpyHostToDevice, st));
    MDG_TRY(mdg_nccl_check(g_nccl.AllReduce(d_flag, d_flag, 1, MDG_NCCL_INT32, MDG_NCCL_MAX, c->dist_comm, st), "AllReduce"));
    MDG_CUDA(cudaMemcpyAsync(&bad, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    MDG_CUDA(cudaStreamSynchronize(st));
    if (bad) {
        mdg_i_dist_p2p_release(c);
        c->dist_p2p_off = true;
        return MDG_OK;
    }
    c->p2p_exported[0] = c->qs_buf[0].p;
    c->p2p_exported[1] = c->qs_buf[1].p;
    c->p2p_exported[2] = c->v4.p;
    c->p2p_exported[3] = c->vh4.p;
    c->dist_p2p = true;
    return MDG_OK;
#endif
}

extern "C" int mdg_dist_finalize(mdg_ctx* c) {
    if (!c) return MDG_OK;
    mdg_i_dist_p2p_release(c);
    if (c->dist_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->dist_comm);
    c->dist_comm = nullptr;
    c->dist_world = 1;
    c->dist_rank = 0;
    return MDG_OK;
}
