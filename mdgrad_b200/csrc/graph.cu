// graph.cu - SchNet continuous-filter convolution, aggregation part (K5 in SURVEY.md 2c).
//
// Replaces the gather-multiply-scatter of the reference message passing:
//   message  = (h[a0] * W, h[a1] * W)                     nff/nn/modules.py:568-572 (SchNetConv.message)
//   agg      = scatter_add(m0 -> a1) + scatter_add(m1 -> a0)   nff/nn/graphconv.py:43-53, nff/utils/scatter.py:24-45
// i.e. agg[k] = sum over edges e incident to node k of h[other(e,k)] * W[e]   (features elementwise),
// with an atomics-free, deterministic segment reduction over a node->incident-edge CSR that is built
// once per topology from the reference-layout (E,2) int64 neighbor list.
// Backward: d/dh is the SAME operator applied to the upstream gradient (the incidence structure is
// symmetric); d/dW[e] = h[a0]*g[a1] + h[a1]*g[a0].
// The dense layers around it (filter MLP, node / update GEMMs) go through cuBLAS this round; the
// tcgen05 path is the planned replacement.
#include "common.cuh"

// dE != nullptr: the edge count lives on the device (asynchronous engine steps); E is then only the launch bound
__global__ void k_inc_count(const int64_t* __restrict__ nbr, int64_t E, const int* __restrict__ dE, int n, int* __restrict__ cnt) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (dE) E = min(E, (int64_t)*dE);
    if (e >= E) return;
    int a0 = (int)nbr[2 * e], a1 = (int)nbr[2 * e + 1];
    if ((unsigned)a0 < (unsigned)n) atomicAdd(&cnt[a0], 1);
    if ((unsigned)a1 < (unsigned)n) atomicAdd(&cnt[a1], 1);
}

__global__ void k_inc_fill(const int64_t* __restrict__ nbr, int64_t E, const int* __restrict__ dE, int n, const int* __restrict__ off,
                           int* __restrict__ cursor, int* __restrict__ inc_edge, int* __restrict__ inc_other) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (dE) E = min(E, (int64_t)*dE);
    if (e >= E) return;
    int a0 = (int)nbr[2 * e], a1 = (int)nbr[2 * e + 1];
    if ((unsigned)a0 < (unsigned)n) { int p = off[a0] + atomicAdd(&cursor[a0], 1); inc_edge[p] = (int)e; inc_other[p] = a1; }
    if ((unsigned)a1 < (unsigned)n) { int p = off[a1] + atomicAdd(&cursor[a1], 1); inc_edge[p] = (int)e; inc_other[p] = a0; }
}

// deterministic order: each node's incident entries sorted by edge id.  One WARP per node: the entries are staged in shared
// memory and every lane RANKS its entries (edge ids are unique: rank = number of smaller ids) - O(m^2 / 32) broadcast reads
// instead of a per-thread insertion sort in global memory (553 us for the 64-water box, m ~ 83: 39% of a SchNet MD step).
// Nodes with more than INC_SORT_CAP incident edges fall back to the insertion sort on lane 0.
#define INC_SORT_WARPS 4
#define INC_SORT_CAP 768
__global__ void __launch_bounds__(INC_SORT_WARPS * 32) k_inc_sort(int n, const int* __restrict__ off, int* __restrict__ inc_edge,
                                                                 int* __restrict__ inc_other) {
    __shared__ int s_e[INC_SORT_WARPS][INC_SORT_CAP], s_o[INC_SORT_WARPS][INC_SORT_CAP];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * INC_SORT_WARPS + w;
    if (k >= n) return;
    const int b = off[k], m = off[k + 1] - b;
    if (m > INC_SORT_CAP) {
        if (lane == 0)
            for (int a = 1; a < m; ++a) {
                int e = inc_edge[b + a], o = inc_other[b + a];
                int j = a - 1;
                while (j >= 0 && inc_edge[b + j] > e) { inc_edge[b + j + 1] = inc_edge[b + j]; inc_other[b + j + 1] = inc_other[b + j]; --j; }
                inc_edge[b + j + 1] = e;
                inc_other[b + j + 1] = o;
            }
        return;
    }
    for (int a = lane; a < m; a += 32) { s_e[w][a] = inc_edge[b + a]; s_o[w][a] = inc_other[b + a]; }
    __syncwarp();
    for (int a = lane; a < m; a += 32) {
        const int e = s_e[w][a];
        int rank = 0;
        for (int j = 0; j < m; ++j) rank += (s_e[w][j] < e);
        inc_edge[b + rank] = e;
        inc_other[b + rank] = s_o[w][a];
    }
}

// one warp per node; lanes stride the feature dimension in float4 (F % 4 == 0) or scalar
template <bool VEC4>
__global__ void __launch_bounds__(256) k_cfconv_agg(int n, int F, const int* __restrict__ off, const int* __restrict__ inc_edge,
                                                    const int* __restrict__ inc_other, const float* __restrict__ h,
                                                    const float* __restrict__ W, float* __restrict__ out) {
    int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (k >= n) return;
    int b = off[k], e1 = off[k + 1];
    if (VEC4) {
        // the (edge, other) indices of 32 entries are fetched by the lanes at once and broadcast with shuffles, so the row
        // gathers of consecutive entries are independent and overlap (the first version chained index load -> row load per
        // entry: 45 us for 192 nodes x 83 entries); the summation order over the entries is unchanged
        int F4 = F >> 2;
        for (int f0 = 0; f0 < F4; f0 += 32) {
            const int f = f0 + lane;
            const bool fa = f < F4;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p0 = b; p0 < e1; p0 += 32) {
                const int mine = p0 + lane;
                const int my_e = mine < e1 ? inc_edge[mine] : 0, my_o = mine < e1 ? inc_other[mine] : 0;
                const int cnt = min(32, e1 - p0);
                for (int j = 0; j < cnt; j += 4) {
                    // four entries' rows are requested before any is used (explicitly: the compiler kept them serial - 45 us
                    // for 192 nodes x 83 entries was 83 exposed L2 latencies per warp)
                    float4 w[4], x[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = __shfl_sync(0xffffffffu, my_e, (j + u) & 31), o = __shfl_sync(0xffffffffu, my_o, (j + u) & 31);
                        w[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        x[u] = w[u];
                        if (fa && j + u < cnt) {
                            w[u] = __ldg(reinterpret_cast<const float4*>(W + (size_t)e * F) + f);
                            x[u] = __ldg(reinterpret_cast<const float4*>(h + (size_t)o * F) + f);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (j + u < cnt) {       // (same order and the same operations as the entry-by-entry loop)
                            acc.x += x[u].x * w[u].x; acc.y += x[u].y * w[u].y; acc.z += x[u].z * w[u].z; acc.w += x[u].w * w[u].w;
                        }
                    }
                }
            }
            if (fa) reinterpret_cast<float4*>(out + (size_t)k * F)[f] = acc;
        }
    } else {
        for (int f = lane; f < F; f += 32) {
            float acc = 0.f;
            for (int p = b; p < e1; ++p) acc += h[(size_t)inc_other[p] * F + f] * W[(size_t)inc_edge[p] * F + f];
            out[(size_t)k * F + f] = acc;
        }
    }
}

__global__ void k_cfconv_edge_grad(int64_t E, int F, const int64_t* __restrict__ nbr, const float* __restrict__ h,
                                   const float* __restrict__ g, float* __restrict__ gW) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= E * F) return;
    int64_t e = idx / F;
    int f = (int)(idx - e * F);
    int64_t a0 = nbr[2 * e], a1 = nbr[2 * e + 1];
    gW[idx] = h[a0 * F + f] * g[a1 * F + f] + h[a1 * F + f] * g[a0 * F + f];
}

int mdg_i_graph_build(mdg_ctx* c, const int64_t* d_nbr, int64_t n_edges, int n, cudaStream_t st, const int* d_n_edges = nullptr);
int mdg_i_cfconv_agg(mdg_ctx* c, const float* d_h, const float* d_W, int n, int F, float* d_out, cudaStream_t st);

extern "C" int mdg_graph_build(mdg_ctx* c, const int64_t* d_nbr, int64_t n_edges, int n, void* stream) {
    if (!c) { mdg_set_error("mdg_graph_build: null ctx"); return MDG_E_BADARG; }
    MDG_CUDA(cudaSetDevice(c->device));
    return mdg_i_graph_build(c, d_nbr, n_edges, n, (cudaStream_t)stream);
}

// d_n_edges != nullptr: n_edges is an upper bound (buffer capacity), the actual count is read on the device
int mdg_i_graph_build(mdg_ctx* c, const int64_t* d_nbr, int64_t n_edges, int n, cudaStream_t st, const int* d_n_edges) {
    if ((n_edges > 0 && !d_nbr) || n < 0 || n_edges < 0 || 2 * n_edges > 0x7fffffffLL) { mdg_set_error("mdg_graph_build: bad arguments"); return MDG_E_BADARG; }
    c->g_n = n;
    c->g_edges = n_edges;
    c->g_nbr = d_nbr;
    MDG_TRY(c->g_off.reserve(sizeof(int) * (size_t)(n + 2)));
    MDG_TRY(c->g_cnt.reserve(sizeof(int) * (size_t)(n + 2)));
    MDG_TRY(c->g_edge.reserve(sizeof(int) * (size_t)(2 * n_edges + 1)));
    MDG_TRY(c->g_other.reserve(sizeof(int) * (size_t)(2 * n_edges + 1)));
    MDG_CUDA(cudaMemsetAsync(c->g_cnt.p, 0, sizeof(int) * (size_t)(n + 1), st));
    if (n_edges > 0) k_inc_count<<<(unsigned)((n_edges + 255) / 256), 256, 0, st>>>(d_nbr, n_edges, d_n_edges, n, c->g_cnt.as<int>());
    MDG_TRY(mdg_i_scan_exclusive(c, c->g_cnt.as<int>(), c->g_off.as<int>(), n + 1, nullptr, st));
    MDG_CUDA(cudaMemsetAsync(c->g_cnt.p, 0, sizeof(int) * (size_t)(n + 1), st));
    if (n_edges > 0) {
        k_inc_fill<<<(unsigned)((n_edges + 255) / 256), 256, 0, st>>>(d_nbr, n_edges, d_n_edges, n, c->g_off.as<int>(), c->g_cnt.as<int>(),
                                                                   c->g_edge.as<int>(), c->g_other.as<int>());
        k_inc_sort<<<(n + INC_SORT_WARPS - 1) / INC_SORT_WARPS, INC_SORT_WARPS * 32, 0, st>>>(n, c->g_off.as<int>(), c->g_edge.as<int>(), c->g_other.as<int>());
    }
    MDG_KERNEL_CHECK();
    return MDG_OK;
}

extern "C" int mdg_cfconv_agg(mdg_ctx* c, const float* d_h, const float* d_W, int n, int F, float* d_out, void* stream) {
    if (!c || !d_h || !d_out || F <= 0) { mdg_set_error("mdg_cfconv_agg: bad arguments"); return MDG_E_BADARG; }
    if (n != c->g_n) { mdg_set_error("mdg_cfconv_agg: graph was built for %d nodes, got %d", c->g_n, n); return MDG_E_STATE; }
    MDG_CUDA(cudaSetDevice(c->device));
    return mdg_i_cfconv_agg(c, d_h, d_W, n, F, d_out, (cudaStream_t)stream);
}

int mdg_i_cfconv_agg(mdg_ctx* c, const float* d_h, const float* d_W, int n, int F, float* d_out, cudaStream_t st) {
    if (n == 0) return MDG_OK;
    int nb = (int)(((int64_t)n * 32 + 255) / 256);
    if ((F & 3) == 0 && (((uintptr_t)d_h | (uintptr_t)d_W | (uintptr_t)d_out) & 15) == 0)
        k_cfconv_agg<true><<<nb, 256, 0, st>>>(n, F, c->g_off.as<int>(), c->g_edge.as<int>(), c->g_other.as<int>(), d_h, d_W, d_out);
    else
        k_cfconv_agg<false><<<nb, 256, 0, st>>>(n, F, c->g_off.as<int>(), c->g_edge.as<int>(), c->g_other.as<int>(), d_h, d_W, d_out);
    MDG_KERNEL_CHECK();
    return MDG_OK;
}

extern "C" int mdg_cfconv_edge_grad(mdg_ctx* c, const float* d_h, const float* d_g, int n, int F, float* d_gW, void* stream) {
    if (!c || !d_h || !d_g || F <= 0) { mdg_set_error("mdg_cfconv_edge_grad: bad arguments"); return MDG_E_BADARG; }
    if (n != c->g_n) { mdg_set_error("mdg_cfconv_edge_grad: graph was built for %d nodes, got %d", c->g_n, n); return MDG_E_STATE; }
    MDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    int64_t tot = c->g_edges * (int64_t)F;
    if (tot == 0) return MDG_OK;
    k_cfconv_edge_grad<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(c->g_edges, F, c->g_nbr, d_h, d_g, d_gW);
    MDG_KERNEL_CHECK();
    return MDG_OK;
}
