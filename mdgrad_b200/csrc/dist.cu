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
    return MDG_OK;
}

extern "C" int mdg_dist_finalize(mdg_ctx* c) {
    if (!c) return MDG_OK;
    if (c->dist_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->dist_comm);
    c->dist_comm = nullptr;
    c->dist_world = 1;
    c->dist_rank = 0;
    return MDG_OK;
}
