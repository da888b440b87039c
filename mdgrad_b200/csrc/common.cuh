// common.cuh - context, error plumbing, exact-arithmetic helpers and analytic pair potentials
// shared by the kernels of libmdgrad_b200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/mdgrad_b200.h"

#define MDG_VERSION 100

// ---------------------------------------------------------------------------------------------
// error handling (thread-local message, never throws)
// ---------------------------------------------------------------------------------------------
void mdg_set_error(const char* fmt, ...);

#define MDG_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            mdg_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return MDG_E_CUDA;                                                               \
        }                                                                                    \
    } while (0)

#define MDG_TRY(call)                \
    do {                             \
        int _s = (call);             \
        if (_s != MDG_OK) return _s; \
    } while (0)

#define MDG_KERNEL_CHECK() MDG_CUDA(cudaGetLastError())

// grow-only device buffer
struct DevBuf {
    void*  p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return MDG_OK;
        if (p) cudaFree(p);
        p = nullptr;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cap = 0;
            mdg_set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return MDG_E_CUDA;
        }
        cap = want;
        return MDG_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return (T*)p; }
};

// ---------------------------------------------------------------------------------------------
// list entry encoding: bits 0..25 = neighbor index (sorted order), bits 26..31 = image code,
// 2 bits per axis holding off_k + 1 in {0,1,2} (off = reference `offsets` row-relative:
// computed from d = x_j - x_i of the row atom i, topology.py:35,59-62)
// ---------------------------------------------------------------------------------------------
#define MDG_STREAM_CAP 1024   // stream-index row form (unused, see build_fast.cuh): max atoms of a 27-cell stencil stream
#define MDG_IDX_BITS 26
#define MDG_IDX_MASK ((1u << MDG_IDX_BITS) - 1u)
#define MDG_MAX_ATOMS (1 << MDG_IDX_BITS)
#define MDG_ROW_PURE 0x40000000        // row_len flag set by k_build_fast: entries are bare indices, no image shifts
#define MDG_ROW_LEN_MASK 0x3fffffff

struct Box {
    float L[3];
    float invL[3];  // fl32(1/L): torch `.inverse()` of a diagonal cell is the correctly rounded reciprocal
};

// Exact restatement of the reference membership arithmetic (SURVEY Appendix A1,
// reference torchmd/topology.py:35,59-67) for ONE axis: returns d + off*L and the off code.
// No FMA contraction anywhere: every product and sum is individually rounded like the
// reference's separate ATen ops.
__device__ __forceinline__ float mdg_min_image_axis(float xi, float xj, float L, float invL, int& code) {
    float d = __fsub_rn(xj, xi);
    float red = __fmul_rn(d, invL);
    float off = 0.0f;
    code = 1;
    if (red > 0.5f) { off = -1.0f; code = 0; }
    else if (red < -0.5f) { off = 1.0f; code = 2; }
    return __fadd_rn(d, __fmul_rn(off, L));
}

// d2 = (dx*dx + dy*dy) + dz*dz, left to right, each op rounded (torch .pow(2).sum(-1) on CPU)
__device__ __forceinline__ float mdg_d2_exact(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ uint32_t mdg_pack_entry(int j, int cx, int cy, int cz) {
    return (uint32_t)j | ((uint32_t)(cx | (cy << 2) | (cz << 4)) << MDG_IDX_BITS);
}

// Rows are padded to a multiple of 32 entries with SELF entries (index of the row atom, no image shift):
// their d2 is exactly 0 and the reference's `d2 != 0` test drops them, so the force kernels can stream whole
// 32-entry blocks without bounds guards.  (row_len keeps the true length for the export.)
__device__ __forceinline__ void mdg_pad_row(uint32_t* row, int cnt, int cap, uint32_t self_ref) {
    int end = (cnt + 31) & ~31;
    if (end > cap) end = cap;
    const uint32_t e = self_ref | ((1u | (1u << 2) | (1u << 4)) << MDG_IDX_BITS);
    for (int k = cnt; k < end; ++k) row[k] = e;
}

// image shift (off*L) for a packed code, per axis
__device__ __forceinline__ float mdg_code_shift(uint32_t code2, float L) {
    // code2 in {0,1,2} -> off in {-1,0,+1}
    return code2 == 1u ? 0.0f : (code2 == 0u ? -L : L);
}

// ---------------------------------------------------------------------------------------------
// analytic pair potentials (reference torchmd/potentials.py). pair_eval returns
//   e = u(r),  g = -u'(r)/r  (so that F_i = -g * (x_j - x_i + off*L)),
// and optionally the parameter derivatives du/dparam in dp[0..3].
// ---------------------------------------------------------------------------------------------
struct PotParams {
    int   kind;
    float p[MDG_MAX_POT_PARAMS];
    float aux;  // Morse: A0 ; LJ: sigma^2
    float se, sg;                       // post-loop scales of the per-atom energy / force sums
    float sdp[MDG_MAX_POT_PARAMS];      // post-loop scales of the parameter-gradient sums
};

__device__ __forceinline__ float mdg_ipow(float x, int n) {
    float r = 1.0f;
    float b = x;
    while (n > 0) {
        if (n & 1) r *= b;
        b *= b;
        n >>= 1;
    }
    return r;
}

// 1-ulp reciprocal on the SFU (MUFU.RCP); energies/forces are tolerance-parity (1e-5), membership is not affected
__device__ __forceinline__ float mdg_rcp(float x) {
#ifdef MDG_EMU       // CPU emulation harness (tests/cuemu): no PTX
    return 1.0f / x;
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

// Per-pair values are returned UNSCALED for the hot kinds; the per-atom sums are multiplied once by
// (se, sg, sdp[]) after the neighbor loop (mdg_pot_scales).  LJ: e_raw = s6 (s6 - 1), g_raw = s6 (2 s6 - 1) / r^2
// with s6 = (sigma^2 / r^2)^3, so u = 4 eps e_raw and -u'/r = 24 eps g_raw.
template <int KIND, bool WITH_DP>
__device__ __forceinline__ void pair_eval(const PotParams& P, float d2, float& e, float& g, float* dp) {
    if (KIND == MDG_POT_LJ) {
        float r2i = mdg_rcp(d2);
        float s2 = P.aux * r2i;               // aux = sigma^2
        float s6 = s2 * s2 * s2;
        float w = s6 * (2.0f * s6 - 1.0f);
        e = s6 * (s6 - 1.0f);
        g = w * r2i;
        if (WITH_DP) {
            dp[0] = w;                        // * 24 eps / sigma
            dp[1] = e;                        // * 4
        }
    } else if (KIND == MDG_POT_LJFAM || KIND == MDG_POT_LJ69) {
        float sigma = P.p[0], eps = P.p[1];
        float rp = (KIND == MDG_POT_LJ69) ? 9.0f : P.p[2];
        float ap = (KIND == MDG_POT_LJ69) ? 6.0f : P.p[3];
        float r = sqrtf(d2);
        float s = sigma / r;
        float srp, sap;
        if (rp == floorf(rp) && ap == floorf(ap) && rp >= 0.f && ap >= 0.f && rp < 64.f && ap < 64.f) {
            srp = mdg_ipow(s, (int)rp);
            sap = mdg_ipow(s, (int)ap);
        } else {
            srp = powf(s, rp);
            sap = powf(s, ap);
        }
        e = 4.0f * eps * (srp - sap);
        float w = 4.0f * eps * (rp * srp - ap * sap);
        g = w / d2;
        if (WITH_DP) {
            dp[0] = w / sigma;
            dp[1] = 4.0f * (srp - sap);
        }
    } else if (KIND == MDG_POT_EXV) {
        float sigma = P.p[0], eps = P.p[1], pw = P.p[2];
        float r = sqrtf(d2);
        float s = sigma / r;
        float sp = (pw == floorf(pw) && pw >= 0.f && pw < 64.f) ? mdg_ipow(s, (int)pw) : powf(s, pw);
        e = 4.0f * eps * sp;
        float w = 4.0f * eps * pw * sp;
        g = w / d2;
        if (WITH_DP) {
            dp[0] = w / sigma;
            dp[1] = 4.0f * sp;
        }
    } else if (KIND == MDG_POT_BUCK) {
        float A = P.p[0], B = P.p[1], C = P.p[2];
        float r = sqrtf(d2);
        float ex = expf(-B * r);
        float r2i = 1.0f / d2;
        float r6i = r2i * r2i * r2i;
        e = A * ex - C * r6i;
        g = A * B * ex / r - 6.0f * C * r6i * r2i;
        if (WITH_DP) {
            dp[0] = ex;
            dp[1] = -A * r * ex;
            dp[2] = -r6i;
        }
    } else {  // MDG_POT_MORSE
        float a = P.p[0], phi = P.p[1], A0 = P.aux;
        float r = sqrtf(d2);
        float rphi = powf(r, phi);
        float x = a * (1.0f - rphi) / phi;
        float ex = expf(x);
        float inv = 1.0f / (1.0f + A0);
        e = (ex * ex - 2.0f * ex - A0) * inv;
        // du/dr = (2 e^{2x} - 2 e^{x}) * dx/dr / (1+A0), dx/dr = -a r^{phi-1} = -a rphi / r
        float dudr = (2.0f * ex * ex - 2.0f * ex) * (-a * rphi / r) * inv;
        g = -dudr / r;
    }
}

// geometry of the engine's tile list (tiles.cuh)
struct TileGeom {
    int ncx, ncy, ncz;
    int nblk;                      // blocks per x-row (balanced widths <= MDG_TILE_MAXW, each width + 2 <= ncx)
    int wbase, wrem;               // ncx / nblk, ncx % nblk: block bi spans cells [bi * wbase + min(bi, wrem), ... + wbase + (bi < wrem))
    int capc;                      // row capacity in chunks
    int scap;                      // staged-atom capacity (dynamic shared memory = 16 * scap bytes)
    int g_base;                    // group index of the first stored group (rows are allocated for the own range only)
    int b_base;                    // block index that group offset counts from
};

// ---------------------------------------------------------------------------------------------
// the context
// ---------------------------------------------------------------------------------------------
struct mdg_ctx {
    int device = 0;
    int sm_count = 148;

    // geometry of the current build
    int    n = 0;
    Box    box;
    int    nc[3] = {1, 1, 1};
    int    ncell = 0;
    int    path = 1;          // 0 = cell list, 1 = all-pairs
    float  rlist2 = 0.f;      // fl32(double(rlist)^2): list membership threshold
    float  rc2 = 0.f;         // fl32(double(cutoff)^2): force re-test threshold
    int    cap = 0;           // row capacity (entries), multiple of 32
    bool   built = false;
    bool   has_sel = false;
    // atom / cell range this context computes (whole box unless a multi-GPU slab plan is active)
    int    own_s0 = 0, own_s1 = 0;   // sorted-atom range [own_s0, own_s1) of rows / forces / integration
    int    own_c0 = 0, own_c1 = 0;   // cell range whose rows are built
    int    force_c0 = 0, force_c1 = 0;    // matching cell sub-range (whole z-layers)
    int    force_s0 = -1, force_s1 = 0;   // >= 0: explicit row sub-range for the next force launch
    int    force_gap_at = 0, force_gap = 0;   // force_gap > 0: the sub-range skips rows [force_gap_at, force_gap_at + force_gap)
    int    rows_s0 = 0;              // first row held in `rows` (rows are allocated for the own range only)
    bool   slab = false;             // true: own_* are set by the distributed engine after the sort
    bool   slab_local = false;       // rebuild touches own +- 2 layers only (engine refreshed them by a 2-layer exchange)
    bool   layers_fresh = false;     // h_layers already holds the offsets of the sort in progress
    DevBuf lay_tot;                  // per-layer atom totals / offsets (distributed local rebuild)
    int    slab_zlo = 0, slab_zhi = 0;
    // distributed state (dist.cu): one NCCL communicator per context
    int    dist_rank = 0, dist_world = 1;
    void*  dist_comm = nullptr;
    cudaStream_t comm_stream = nullptr;            // NCCL side stream: halo + KE all-reduce overlap the interior forces
    cudaEvent_t  ev_a = nullptr, ev_halo = nullptr, ev_ke = nullptr;
    cudaStream_t bnd_stream = nullptr;             // boundary-layer forces of a slab step: wait for the ghosts + two layers of rows, BESIDE the interior rows
    cudaEvent_t  ev_bnd = nullptr;
    // peer-to-peer step path (dist.cuh DistSync): IPC mappings of the neighbours' position buffers and of every rank's sync block
    bool   dist_p2p = false;         // mappings valid -> the per-step halo / kinetic-energy exchange uses NVLink stores + flags
    bool   dist_p2p_off = false;     // MDG_DIST_P2P=0, or the IPC set-up failed once: stay on the NCCL path
    DevBuf dsync;                    // my DistSync block
    void*  peer_sync[16] = {nullptr};        // every rank's DistSync as mapped here (own entry = dsync.p)
    void*  peer_qs[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [below, above][qs_buf 0, 1] of the two neighbours
    void*  peer_v[2] = {nullptr, nullptr}, *peer_vh[2] = {nullptr, nullptr};   // [below, above] v4 / vh4 buffers of the neighbours (rebuild state push)
    void*  p2p_exported[4] = {nullptr, nullptr, nullptr, nullptr};   // my qs_buf / v4 / vh4 pointers at export time (re-export when they change)
    int    dist_rebuilds = 0;        // rebuilds of the slab engine so far (parity of the layer-total tables; identical on all ranks)
    void*  p2p_opened[40] = {nullptr};       // everything cudaIpcOpenMemHandle returned (closed on release)
    int    p2p_n_opened = 0;
    int    dist_seq = 0;             // running sequence number of the distributed steps (identical on all ranks)
    cudaEvent_t ev_push = nullptr;
    int*   h_layers = nullptr;       // pinned: atom offset of every z-layer of cells (ncz + 1 entries)
    int    n_layers = 0;
    bool   fast_build = false; // engine skin lists: approximate (FMA) membership at the list radius is allowed
    bool   rows_wanted = true;// false: sort into cells only (RDF traversal needs no stored rows)

    // cell / sort tables
    DevBuf cell_of, slot_of, cell_count, cell_start, perm, perm_tmp, stencil;
    DevBuf qs_buf[2];         // float4 sorted positions (double-buffered: the sorter's input may be
    float4* qs_ptr = nullptr; //   the previous output), w = original index (int bits)
    // list
    DevBuf rows, row_len;     // uint32 [n*cap], int [n]
    int    force_group = 4;            // lanes per row in k_force_rows (MDG_FORCE_GROUP=2|4|8)
    bool   force_energy = true;        // false: the next force launches skip the per-atom energy (engine inner steps)
    DevBuf work_ctr;          // int: cell counter of k_build_fast
    DevBuf flags;             // int[8]: 0 = capacity overflow, 1 = skin violation
    // tile list (tiles.cuh): the engine's skin list in block-local 16-bit form
    bool     tiles = false;            // the last build produced tile rows (force launches use k_force_tiles)
    TileGeom tile;
    int      tile_warps = 12;          // warps per CTA of k_force_tiles
    int      tile_warps_env = 0;       // MDG_TILE_WARPS
    int      tile_ctas_env = 0;        // MDG_TILE_CTAS: resident CTAs per SM of the persistent k_force_tiles
    bool     tiles_off = true;         // MDG_TILES=1 switches the engine's skin list to the tile form (opt-in, see api.cu)
    bool     flags_sticky = false;     // mdg_i_build_list must not clear the overflow flags (engine epochs)
    int      tile_scap_min = 0;        // staged-atom capacity demanded by a previous overflow
    DevBuf   tile_rows, tile_len;      // uint16 [groups * capc * 128], uint32 [n]
    DevBuf   tile_desc;                // int [blocks * MDG_TILE_DESC]: block descriptors written by k_build_tiles
    // export scratch
    DevBuf up_cnt, up_off, scan_tmp;
    int64_t npairs = 0;
    // force scratch
    DevBuf fs;                // float4 sorted force+energy
    DevBuf partials;          // double partial sums
    // masks (kept for the build)
    const uint8_t* sel_a = nullptr;
    const uint8_t* sel_b = nullptr;
    const int64_t* ex_keys = nullptr;
    int n_ex = 0;
    // persistent filter used by mdg_md_run (mdg_set_pair_filter)
    const uint8_t* eng_sel_a = nullptr;
    const uint8_t* eng_sel_b = nullptr;
    const int64_t* eng_ex_keys = nullptr;
    int eng_n_ex = 0;

    // SchNet graph (graph.cu): node -> incident-edge CSR of the last mdg_graph_build
    DevBuf g_off, g_cnt, g_edge, g_other;
    DevBuf sn_ws;             // SchNet activations / workspace (schnet.cu)
    DevBuf sn_wcache;         // cached transposed filter weights of the last model (schnet.cu), valid for sn_wtag
    uint64_t sn_wtag = 0;
    const float* sn_wkey[2 * MDG_SCHNET_MAX_LAYERS] = {nullptr};
    DevBuf sn_wt;             // transposed weight scratch of the tensor-core dense layers (schnet_tc.cuh)
    DevBuf gnn_nbr, gnn_off, gnn_xyz, gnn_f3, gnn_fp3;   // GNN epoch (engine.cu): exported list, xyz / force staging
    cudaStream_t gnn_stream = nullptr;   // private stream of the graph-replay GNN epochs (capture is illegal on torch's legacy default stream)
    cudaEvent_t  ev_gnn = nullptr;
    int64_t stat_graph_replays = 0;
    int64_t gnn_last_cap = 0;            // edge capacity a completed asynchronous epoch ran with (0: none) - lets the next epoch
    int     gnn_last_n = 0;              //   of the same system start without any read-back
    DevBuf bd_slots, bd_part;  // bonded terms (bonded.cu): per-term gradient slots, block partial sums
    int     g_n = -1;
    int64_t g_edges = 0;
    const int64_t* g_nbr = nullptr;

    // engine state (sorted order)
    DevBuf v4, vh4, q4b, f4b, qref, mass_sorted, pvbuf, kebuf, dtbuf;

    // stats
    int64_t stat_launches = 0, stat_rebuilds = 0, stat_entries = 0, stat_maxrow = 0;
    int64_t stat_async_retries = 0;   // GNN epochs repeated on the synchronous path after a latched overflow (engine.cu)
    // optional per-kernel timing of the engine's force launches (mdg_set_profile)
    int     prof_enable = 0;
    void*   prof_events = nullptr;   // std::vector<cudaEvent_t>* (pairs)
    int     prof_used = 0;
    double  prof_force_ms = 0.0;
    int64_t prof_force_launches = 0;
    int* h_pinned = nullptr;  // pinned int[16] for read-backs
};

// internal entry points shared across translation units ------------------------------------
int mdg_i_build_list(mdg_ctx* c, const float* d_xyz, const float4* d_q4_sorted_in, int n,
                     const float* h_cell3, double rlist, double cutoff, cudaStream_t st);
int mdg_i_scan_exclusive(mdg_ctx* c, const int* d_in, int* d_out, int n, int* d_total, cudaStream_t st);
int mdg_i_force_sorted(mdg_ctx* c, const PotParams& P, const float4* d_qs, float4* d_fs, bool retest,
                       bool with_dp, double* d_dp_partials, cudaStream_t st);
int mdg_i_force_range(mdg_ctx* c, const PotParams& P, const float4* d_qs, float4* d_fs, bool retest, int s0, int s1,
                      int c0, int c1, cudaStream_t st);
int mdg_i_force_range2(mdg_ctx* c, const PotParams& P, const float4* d_qs, float4* d_fs, bool retest, int s0a, int s1a, int s0b,
                       int s1b, cudaStream_t st);
PotParams mdg_make_pot(int kind, const float* h_params, int n_params);
int mdg_i_check_flags(mdg_ctx* c, cudaStream_t st, bool sync);
