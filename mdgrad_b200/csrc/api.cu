// api.cu - C-ABI entry points of libmdgrad_b200.so (context, neighbor list, pair force, stats).
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void mdg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int mdg_i_export_count(mdg_ctx* c, cudaStream_t st, int64_t* h_npairs);
void mdg_i_release_profile(mdg_ctx* c);
int mdg_i_export_fill(mdg_ctx* c, int64_t* d_nbr, float* d_offsets, float* d_dis, cudaStream_t st);
int mdg_i_pair_force_op(mdg_ctx* c, const PotParams& P, const float* d_xyz, int n, float* d_energy, float* d_force,
                        float* d_dparams, cudaStream_t st);

extern "C" int mdg_version(void) { return MDG_VERSION; }
extern "C" const char* mdg_last_error(void) { return g_err; }

extern "C" int mdg_create(int device, mdg_ctx** out) {
    if (!out) { mdg_set_error("mdg_create: out is NULL"); return MDG_E_BADARG; }
    *out = nullptr;
    int ndev = 0;
    MDG_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { mdg_set_error("mdg_create: device %d of %d", device, ndev); return MDG_E_BADARG; }
    MDG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MDG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        mdg_set_error("mdg_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return MDG_E_CUDA;
    }
    mdg_ctx* c = new mdg_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cudaMallocHost((void**)&c->h_pinned, sizeof(int) * 16) != cudaSuccess) {
        delete c;
        mdg_set_error("cudaMallocHost failed");
        return MDG_E_CUDA;
    }
    memset(c->h_pinned, 0, sizeof(int) * 16);
    int s = c->flags.reserve(sizeof(int) * 8);
    if (s != MDG_OK) { cudaFreeHost(c->h_pinned); delete c; return s; }
    cudaMemset(c->flags.p, 0, sizeof(int) * 8);
    const char* fg = getenv("MDG_FORCE_GROUP");
    c->force_group = (fg && (atoi(fg) == 8 || atoi(fg) == 2)) ? atoi(fg) : 4;
    // MDG_TILES=1: engine skin list in the block-local tile form (tiles.cuh: TMA-staged stencil, 16-bit rows).  Measured on the
    // 256k-atom box (profiles/r02_force_kernel.md) it halves the kernel's DRAM traffic but is slower than the row form
    // (52-57 vs 47 us), so the row form stays the default.
    const char* tl = getenv("MDG_TILES");
    c->tiles_off = !(tl && tl[0] == '1');
    const char* tw = getenv("MDG_TILE_WARPS");
    c->tile_warps_env = (tw && atoi(tw) >= 1 && atoi(tw) <= 16) ? atoi(tw) : 0;
    const char* tc = getenv("MDG_TILE_CTAS");
    c->tile_ctas_env = (tc && atoi(tc) >= 1 && atoi(tc) <= 16) ? atoi(tc) : 0;
    *out = c;
    return MDG_OK;
}

extern "C" int mdg_destroy(mdg_ctx* c) {
    if (!c) return MDG_OK;
    cudaSetDevice(c->device);
    DevBuf* bufs[] = {&c->cell_of, &c->slot_of, &c->cell_count, &c->cell_start, &c->perm, &c->perm_tmp, &c->stencil,
                      &c->qs_buf[0], &c->qs_buf[1], &c->rows, &c->row_len, &c->tile_rows, &c->tile_len, &c->tile_desc, &c->dsync, &c->flags, &c->up_cnt, &c->up_off,
                      &c->scan_tmp, &c->fs, &c->partials, &c->v4, &c->vh4, &c->q4b, &c->f4b, &c->qref,
                      &c->mass_sorted, &c->pvbuf, &c->kebuf, &c->dtbuf, &c->g_off, &c->g_cnt, &c->g_edge,
                      &c->g_other, &c->sn_ws, &c->sn_wt, &c->sn_wcache, &c->gnn_nbr, &c->gnn_off, &c->gnn_xyz, &c->gnn_f3, &c->gnn_fp3,
                      &c->bd_slots, &c->bd_part, &c->work_ctr};
    for (DevBuf* b : bufs) b->release();
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_layers) cudaFreeHost(c->h_layers);
    mdg_i_release_profile(c);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    if (c->bnd_stream) cudaStreamDestroy(c->bnd_stream);
    if (c->ev_bnd) cudaEventDestroy(c->ev_bnd);
    if (c->gnn_stream) cudaStreamDestroy(c->gnn_stream);
    if (c->ev_gnn) cudaEventDestroy(c->ev_gnn);
    if (c->ev_a) cudaEventDestroy(c->ev_a);
    if (c->ev_halo) cudaEventDestroy(c->ev_halo);
    if (c->ev_ke) cudaEventDestroy(c->ev_ke);
    delete c;
    return MDG_OK;
}

extern "C" int mdg_nbr_build(mdg_ctx* c, const float* d_xyz, int n, const float* h_cell3, double cutoff,
                             const uint8_t* d_sel_a, const uint8_t* d_sel_b, const int64_t* d_ex_keys, int n_ex,
                             void* stream, int64_t* h_npairs) {
    if (!c || !h_cell3 || !h_npairs || (n > 0 && !d_xyz)) { mdg_set_error("mdg_nbr_build: null argument"); return MDG_E_BADARG; }
    if ((d_sel_a == nullptr) != (d_sel_b == nullptr)) { mdg_set_error("mdg_nbr_build: give both sel_a and sel_b or neither"); return MDG_E_BADARG; }
    if (!(cutoff > 0)) { mdg_set_error("mdg_nbr_build: cutoff must be > 0"); return MDG_E_BADARG; }
    MDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    c->sel_a = d_sel_a;
    c->sel_b = d_sel_b;
    c->ex_keys = d_ex_keys;
    c->n_ex = d_ex_keys ? n_ex : 0;
    c->rows_wanted = true;
    c->fast_build = false;       // the exported list must be bit-exact
    c->stat_launches = 0;
    for (int attempt = 0; attempt < 10; ++attempt) {
        MDG_TRY(mdg_i_build_list(c, d_xyz, nullptr, n, h_cell3, cutoff, cutoff, st));
        int s = mdg_i_export_count(c, st, h_npairs);
        if (s == MDG_E_CAPACITY) {
            int need = c->h_pinned[2];
            int cap = ((need + need / 8 + 31) / 32) * 32;
            if (cap <= c->cap) cap = c->cap + 32;
            c->cap = cap;
            continue;
        }
        if (s == MDG_OK) { c->stat_entries = 2 * c->npairs; c->stat_maxrow = c->h_pinned[2]; }
        return s;
    }
    mdg_set_error("mdg_nbr_build: row capacity could not be satisfied");
    return MDG_E_CAPACITY;
}

extern "C" int mdg_nbr_export(mdg_ctx* c, int64_t* d_nbr, float* d_offsets, float* d_dis, void* stream) {
    if (!c) { mdg_set_error("null ctx"); return MDG_E_BADARG; }
    if (!c->built) { mdg_set_error("mdg_nbr_export: call mdg_nbr_build first"); return MDG_E_STATE; }
    if (c->npairs > 0 && (!d_nbr || !d_offsets)) { mdg_set_error("mdg_nbr_export: null output"); return MDG_E_BADARG; }
    MDG_CUDA(cudaSetDevice(c->device));
    return mdg_i_export_fill(c, d_nbr, d_offsets, d_dis, (cudaStream_t)stream);
}

extern "C" int mdg_pair_force(mdg_ctx* c, int kind, const float* h_params, int n_params, const float* d_xyz, int n,
                              float* d_energy, float* d_force, float* d_dparams, void* stream) {
    if (!c || !h_params) { mdg_set_error("mdg_pair_force: null argument"); return MDG_E_BADARG; }
    if (kind < MDG_POT_LJ || kind > MDG_POT_MORSE) { mdg_set_error("mdg_pair_force: unknown kind %d", kind); return MDG_E_BADARG; }
    if (n_params < 0 || n_params > MDG_MAX_POT_PARAMS) { mdg_set_error("mdg_pair_force: n_params=%d", n_params); return MDG_E_BADARG; }
    MDG_CUDA(cudaSetDevice(c->device));
    PotParams P = mdg_make_pot(kind, h_params, n_params);
    return mdg_i_pair_force_op(c, P, d_xyz, n, d_energy, d_force, d_dparams, (cudaStream_t)stream);
}

extern "C" int mdg_set_pair_filter(mdg_ctx* c, const uint8_t* d_sel_a, const uint8_t* d_sel_b, const int64_t* d_ex_keys,
                                   int n_ex) {
    if (!c) { mdg_set_error("null ctx"); return MDG_E_BADARG; }
    if ((d_sel_a == nullptr) != (d_sel_b == nullptr)) { mdg_set_error("give both sel_a and sel_b or neither"); return MDG_E_BADARG; }
    c->eng_sel_a = d_sel_a;
    c->eng_sel_b = d_sel_b;
    c->eng_ex_keys = d_ex_keys;
    c->eng_n_ex = d_ex_keys ? n_ex : 0;
    return MDG_OK;
}

extern "C" int mdg_get_stats(mdg_ctx* c, int64_t* o) {
    if (!c || !o) { mdg_set_error("null argument"); return MDG_E_BADARG; }
    o[0] = c->stat_launches;
    o[1] = c->stat_rebuilds;
    o[2] = c->stat_entries;
    o[3] = c->stat_maxrow;
    o[4] = c->nc[0];
    o[5] = c->nc[1];
    o[6] = c->nc[2];
    o[7] = c->tiles ? 2 : c->path;      // 0 = cell list (rows), 1 = all-pairs, 2 = cell list in tile form (tiles.cuh)
    return MDG_OK;
}
