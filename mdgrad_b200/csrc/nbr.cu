// nbr.cu - cell-list build (device-wide counting sort) + neighbor-list construction and the
// export in the reference's layout/order.  Replaces generate_nbr_list, reference
// torchmd/topology.py:30-73 (K1 in SURVEY.md 2c).
#include "common.cuh"
#include "dist.cuh"

extern "C" int mdg_slab_plan(int ncz, int world, int rank, int* out4);

#define ALLPAIRS_MAX_ATOMS 3072   // below this (or with < 3 cells on an axis) use the all-pairs search

// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------
__global__ void k_pack_xyz(const float* __restrict__ xyz, float4* __restrict__ q, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    q[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], __int_as_float(i));
}

__global__ void k_iota(int* __restrict__ p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// exclusive scan, 3 phases -------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem, int& total) {
    // smem: SCAN_THREADS/32 ints
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) smem[w] = x;
    __syncthreads();
    if (w == 0) {
        int s = lane < (blockDim.x >> 5) ? smem[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        smem[lane] = s;  // inclusive warp totals (32 slots)
    }
    __syncthreads();
    int wbase = w ? smem[w - 1] : 0;
    total = smem[(blockDim.x >> 5) - 1];
    __syncthreads();
    return wbase + x - v;
}

// single_total != nullptr (one-tile scans only): the grand total is written here and no second kernel is needed
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const int* __restrict__ in, int* __restrict__ out,
                                                            int* __restrict__ tile_tot, int n, int* __restrict__ single_total) {
    __shared__ int sm[32];
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int tot;
    int ex = block_exclusive_scan(s, sm, tot);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) {
        tile_tot[blockIdx.x] = tot;
        if (single_total) *single_total = tot;
    }
}

__global__ void __launch_bounds__(1024) k_scan_totals(int* __restrict__ tile_tot, int ntiles, int* __restrict__ grand) {
    // single block; sequential over chunks of 1024
    __shared__ int sm[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < ntiles; c0 += 1024) {
        int i = c0 + threadIdx.x;
        int v = i < ntiles ? tile_tot[i] : 0;
        int tot;
        int ex = block_exclusive_scan(v, sm, tot);
        int carry = carry_s;
        if (i < ntiles) tile_tot[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && grand) *grand = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(int* __restrict__ out, const int* __restrict__ tile_tot, int n) {
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int add = tile_tot[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += add;
}

// Mid-size arrays (the cell counts of a rebuild: ~1e4 cells): ONE block walks the array in chunks of 1024 x SCAN_ONE_ITEMS with a
// running carry - one launch instead of three (tiles / totals / add: 11 us of launch-bound work per rebuild in the r02 launch list).
#define SCAN_ONE_ITEMS 16
__global__ void __launch_bounds__(1024) k_scan_one(const int* __restrict__ in, int* __restrict__ out, int n, int* __restrict__ grand) {
    __shared__ int sm[32];
    int carry = 0;
    for (int c0 = 0; c0 < n; c0 += 1024 * SCAN_ONE_ITEMS) {
        const int base = c0 + threadIdx.x * SCAN_ONE_ITEMS;
        int v[SCAN_ONE_ITEMS];
        int s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ONE_ITEMS; ++k) {
            v[k] = (base + k < n) ? in[base + k] : 0;
            s += v[k];
        }
        int tot;
        int ex = carry + block_exclusive_scan(s, sm, tot);
#pragma unroll
        for (int k = 0; k < SCAN_ONE_ITEMS; ++k) {
            if (base + k < n) out[base + k] = ex;
            ex += v[k];
        }
        carry += tot;
    }
    if (threadIdx.x == 0 && grand) *grand = carry;
}

// exclusive scan of d_in[0..n) into d_out[0..n); *d_total (device, optional) = sum
int mdg_i_scan_exclusive(mdg_ctx* c, const int* d_in, int* d_out, int n, int* d_total, cudaStream_t st) {
    if (n <= 0) {
        if (d_total) MDG_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int), st));
        return MDG_OK;
    }
    int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    MDG_TRY(c->scan_tmp.reserve(sizeof(int) * (size_t)(ntiles + 1)));
    int* tt = c->scan_tmp.as<int>();
    if (ntiles > 1 && n <= 2 * 1024 * SCAN_ONE_ITEMS) {
        k_scan_one<<<1, 1024, 0, st>>>(d_in, d_out, n, d_total);
        c->stat_launches += 1;
    } else if (ntiles == 1) {   // small arrays (SchNet-sized systems): one launch
        k_scan_tiles<<<1, SCAN_THREADS, 0, st>>>(d_in, d_out, tt, n, d_total);
        c->stat_launches += 1;
    } else {
        k_scan_tiles<<<ntiles, SCAN_THREADS, 0, st>>>(d_in, d_out, tt, n, nullptr);
        k_scan_totals<<<1, 1024, 0, st>>>(tt, ntiles, d_total);
        k_scan_add<<<ntiles, SCAN_THREADS, 0, st>>>(d_out, tt, n);
        c->stat_launches += 3;
    }
    MDG_KERNEL_CHECK();
    return MDG_OK;
}

// ---------------------------------------------------------------------------------------------
// cell binning + counting sort
// ---------------------------------------------------------------------------------------------
struct Grid {
    int   nc[3];
    float L[3], invL[3];
};

__device__ __forceinline__ int cell_coord(float x, float L, float invL, int nc) {
    // bin by the wrapped fractional coordinate; atoms need not be inside the box
    float f = x * invL;
    f -= floorf(f);               // [0,1]
    int c = (int)(f * (float)nc);
    return min(max(c, 0), nc - 1);
}

__global__ void k_bin(const float4* __restrict__ q, int i0, int n, Grid g, int* __restrict__ cell_of,
                      int* __restrict__ slot_of, int* __restrict__ cell_count, int* __restrict__ flags) {
    int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;          // atoms [i0, n)
    if (i >= n) return;
    float4 p = q[i];
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) flags[6] = 1;   // diverged dynamics: reported, never binned blindly
    int cx = cell_coord(p.x, g.L[0], g.invL[0], g.nc[0]);
    int cy = cell_coord(p.y, g.L[1], g.invL[1], g.nc[1]);
    int cz = cell_coord(p.z, g.L[2], g.invL[2], g.nc[2]);
    int c = (cz * g.nc[1] + cy) * g.nc[0] + cx;
    cell_of[i] = c;
    slot_of[i] = atomicAdd(&cell_count[c], 1);
}

// zwin0/zwin_n: only atoms whose new cell lies in the z-layer window [zwin0, zwin0 + zwin_n) (periodic) are placed
// (distributed rebuild: a rank places the atoms of its slab + ghost layers only); zwin_n <= 0 = everything
__global__ void k_scatter(int i0, int n, const int* __restrict__ cell_of, const int* __restrict__ slot_of,
                          const int* __restrict__ cell_start, int* __restrict__ perm_tmp, int nxy, int ncz, int zwin0, int zwin_n) {
    int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cell_of[i];
    if (zwin_n > 0) {
        int lz = (c / nxy - zwin0 + ncz) % ncz;
        if (lz >= zwin_n) return;
    }
    perm_tmp[cell_start[c] + slot_of[i]] = i;
}

// ---- distributed local rebuild helpers -----------------------------------------------------------
// totals of the layers this rank owns (complete after binning own +- 2 layers); zero elsewhere -> all-reduce
__global__ void k_layer_totals(const int* __restrict__ cell_count, int nxy, int ncz, int zlo, int zhi, int* __restrict__ tot) {
    __shared__ int sm[32];
    int z = blockIdx.x;
    int v = 0;
    if (z >= zlo && z < zhi)
        for (int k = threadIdx.x; k < nxy; k += blockDim.x) v += cell_count[z * nxy + k];
    int dummy;
    int ex = block_exclusive_scan(v, sm, dummy);
    (void)ex;
    if (threadIdx.x == 0) tot[z] = dummy;
}

__global__ void k_layer_offsets(const int* __restrict__ tot, int ncz, int* __restrict__ off) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int acc = 0;
        for (int z = 0; z < ncz; ++z) { off[z] = acc; acc += tot[z]; }
        off[ncz] = acc;
    }
}

// Peer-to-peer form of "all-reduce the layer totals" (dist.cuh): every layer has ONE owner, which stores its total into every
// rank's table, then raises its flag there; the reader waits for all flags and prefix-sums the table.  Replaces a latency-bound
// NCCL all-reduce of ncz ints on the rebuild path.  Tables are double-buffered by the parity of the rebuild counter.
__global__ void __launch_bounds__(256) k_dist_layers_push(const int* __restrict__ lay_tot, int zlo, int zhi, PeerTab T, int me, int world,
                                                          int par, int seq) {
    for (int k = threadIdx.x; k < (zhi - zlo) * world; k += blockDim.x) {
        const int r = k / (zhi - zlo), z = zlo + k % (zhi - zlo);
        vstore_i(&T.s[r]->lay[par][z], lay_tot[z]);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        vstore_i(&T.s[threadIdx.x]->lay_flag[par][me], seq);
    }
}
__global__ void k_layer_offsets_p2p(DistSync* mine, int world, int par, int seq, int ncz, int* __restrict__ off) {
    if ((int)threadIdx.x < world) spin_until_ge(&mine->lay_flag[par][threadIdx.x], seq, &mine->pad[1]);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int z = 0; z < ncz; ++z) { off[z] = acc; acc += vload_i(&mine->lay[par][z]); }
        off[ncz] = acc;
    }
}

// cell_start of the cells of layer (zwin0 + blockIdx.x) % ncz = layer offset + exclusive scan of the cell counts
__global__ void __launch_bounds__(SCAN_THREADS) k_cellstart_layer(const int* __restrict__ cell_count, const int* __restrict__ lay_off,
                                                                 int nxy, int ncz, int zwin0, int* __restrict__ cell_start) {
    __shared__ int sm[32];
    __shared__ int carry_s;
    int z = (zwin0 + blockIdx.x) % ncz;
    if (threadIdx.x == 0) carry_s = lay_off[z];
    __syncthreads();
    for (int k0 = 0; k0 < nxy; k0 += blockDim.x) {
        int k = k0 + threadIdx.x;
        int v = k < nxy ? cell_count[z * nxy + k] : 0;
        int tot;
        int ex = block_exclusive_scan(v, sm, tot);
        int carry = carry_s;
        if (k < nxy) cell_start[z * nxy + k] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) cell_start[(z + 1) * nxy] = lay_off[z + 1];    // end sentinel of the layer (same value from any block)
}

// one thread per cell: order the cell's atoms by ORIGINAL id (deterministic, history-free),
// then emit the sorted positions and the permutation (index into the input array)
#define MAX_CELL_SORT 4096
__global__ void k_cellsort_gather(int ncell, const int* __restrict__ cell_start, const int* __restrict__ cell_count,
                                  int* __restrict__ perm_tmp, const float4* __restrict__ qin,
                                  float4* __restrict__ qs, int* __restrict__ perm, int* __restrict__ flags) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    int b = cell_start[c], m = cell_count[c];
    if (m > MAX_CELL_SORT) { flags[7] = m; m = 0; }   // absurd occupancy (collapsed / diverged system): flag, do not spin
    for (int a = 1; a < m; ++a) {
        int pa = perm_tmp[b + a];
        int ka = __float_as_int(qin[pa].w);
        int k = a - 1;
        while (k >= 0) {
            int pk = perm_tmp[b + k];
            if (__float_as_int(qin[pk].w) <= ka) break;
            perm_tmp[b + k + 1] = pk;
            --k;
        }
        perm_tmp[b + k + 1] = pa;
    }
    for (int a = 0; a < m; ++a) {
        int pa = perm_tmp[b + a];
        qs[b + a] = qin[pa];
        perm[b + a] = pa;
    }
}

// Warp-per-cell variant: each lane holds one atom of the cell, its rank = number of atoms in the cell
// with a smaller original id (ids are unique), computed with shuffles; writes the sorted positions,
// the permutation and the per-atom cell id in one pass.  Cells with more than 32 atoms fall back to
// the serial insertion sort on lane 0 (rare: > 1.5x the mean occupancy at liquid density).
__global__ void __launch_bounds__(256) k_cellsort_warp(int cell0, int ncell, const int* __restrict__ cell_start, const int* __restrict__ cell_count,
                                                       int* __restrict__ perm_tmp, const float4* __restrict__ qin,
                                                       float4* __restrict__ qs, int* __restrict__ perm,
                                                       int* __restrict__ cell_sorted, int* __restrict__ flags) {
    int c = cell0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // cells [cell0, ncell)
    int lane = threadIdx.x & 31;
    if (c >= ncell) return;
    int b = cell_start[c], m = cell_count[c];
    if (m <= 32) {
        int pa = lane < m ? perm_tmp[b + lane] : 0;
        float4 q = lane < m ? qin[pa] : make_float4(0, 0, 0, 0);
        int key = lane < m ? __float_as_int(q.w) : 0x7fffffff;
        int rank = 0;
        for (int o = 0; o < m; ++o) rank += (__shfl_sync(0xffffffffu, key, o) < key);
        if (lane < m) {
            qs[b + rank] = q;
            perm[b + rank] = pa;
            cell_sorted[b + rank] = c;
        }
        return;
    }
    if (m > MAX_CELL_SORT) { if (lane == 0) flags[7] = m; m = 0; }
    if (lane == 0) {
        for (int a = 1; a < m; ++a) {
            int pa = perm_tmp[b + a];
            int ka = __float_as_int(qin[pa].w);
            int k = a - 1;
            while (k >= 0) {
                int pk = perm_tmp[b + k];
                if (__float_as_int(qin[pk].w) <= ka) break;
                perm_tmp[b + k + 1] = pk;
                --k;
            }
            perm_tmp[b + k + 1] = pa;
        }
    }
    __syncwarp();
    for (int a = lane; a < m; a += 32) {
        int pa = perm_tmp[b + a];
        qs[b + a] = qin[pa];
        perm[b + a] = pa;
        cell_sorted[b + a] = c;
    }
}

// 27-cell stencil per cell, ascending linear id (so that row entries come out ascending)
__global__ void k_stencil(Grid g, int* __restrict__ stencil) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int ncell = g.nc[0] * g.nc[1] * g.nc[2];
    if (c >= ncell) return;
    int cx = c % g.nc[0], cy = (c / g.nc[0]) % g.nc[1], cz = c / (g.nc[0] * g.nc[1]);
    int s[27];
    int m = 0;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                int x = (cx + dx + g.nc[0]) % g.nc[0];
                int y = (cy + dy + g.nc[1]) % g.nc[1];
                int z = (cz + dz + g.nc[2]) % g.nc[2];
                int v = (z * g.nc[1] + y) * g.nc[0] + x;
                int k = m - 1;
                while (k >= 0 && s[k] > v) { s[k + 1] = s[k]; --k; }
                s[k + 1] = v;
                ++m;
            }
    for (int k = 0; k < 27; ++k) stencil[c * 27 + k] = s[k];
}

// ---------------------------------------------------------------------------------------------
// pair filters (species selection / exclusions), reference topology.py:15-27,37-53
// ---------------------------------------------------------------------------------------------
struct PairFilter {
    const uint8_t* sel_a;
    const uint8_t* sel_b;
    const int64_t* ex_keys;
    int n_ex;
    int n;
};

__device__ __forceinline__ bool pair_allowed(const PairFilter& F, int ida, int idb) {
    if (F.sel_a) {
        bool ok = (F.sel_a[ida] && F.sel_b[idb]) || (F.sel_b[ida] && F.sel_a[idb]);
        if (!ok) return false;
    }
    if (F.n_ex > 0) {
        int lo = min(ida, idb), hi = max(ida, idb);
        int64_t key = (int64_t)lo * F.n + hi;
        int a = 0, b = F.n_ex - 1;
        while (a <= b) {
            int mid = (a + b) >> 1;
            int64_t v = F.ex_keys[mid];
            if (v == key) return false;
            if (v < key) a = mid + 1; else b = mid - 1;
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// list construction.  One thread per (sorted) atom; row entries in ascending neighbor index.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool test_pair(const float4& qi, const float4& qj, const Box& bx, float r2max,
                                          uint32_t& entry_code) {
    int cx, cy, cz;
    float dx = mdg_min_image_axis(qi.x, qj.x, bx.L[0], bx.invL[0], cx);
    float dy = mdg_min_image_axis(qi.y, qj.y, bx.L[1], bx.invL[1], cy);
    float dz = mdg_min_image_axis(qi.z, qj.z, bx.L[2], bx.invL[2], cz);
    float d2 = mdg_d2_exact(dx, dy, dz);
    entry_code = (uint32_t)(cx | (cy << 2) | (cz << 4)) << MDG_IDX_BITS;
    return (d2 < r2max) && (d2 != 0.0f);
}

__global__ void __launch_bounds__(128) k_build_cells(int s0, int n, const float4* __restrict__ qs, const int* __restrict__ cell_sorted,
                                                     const int* __restrict__ cell_start, const int* __restrict__ stencil,
                                                     Box bx, float r2max, int cap, PairFilter F,
                                                     uint32_t* __restrict__ rows, int* __restrict__ row_len,
                                                     int* __restrict__ flags) {
    int s = s0 + blockIdx.x * blockDim.x + threadIdx.x;      // rows [s0, n)
    if (s >= n) return;
    if (flags[6] | flags[7]) { row_len[s] = 0; return; }   // non-finite / collapsed input: empty list, error reported by the host
    float4 qi = qs[s];
    int idi = __float_as_int(qi.w);
    int c = cell_sorted[s];
    uint32_t* row = rows + (size_t)s * cap;
    int cnt = 0;
    bool filt = (F.sel_a != nullptr) || (F.n_ex > 0);
    for (int k = 0; k < 27; ++k) {
        int cc = stencil[c * 27 + k];
        int t0 = cell_start[cc], t1 = cell_start[cc + 1];
        for (int t = t0; t < t1; ++t) {
            if (t == s) continue;
            float4 qj = qs[t];
            uint32_t code;
            if (!test_pair(qi, qj, bx, r2max, code)) continue;
            if (filt && !pair_allowed(F, idi, __float_as_int(qj.w))) continue;
            if (cnt < cap) row[cnt] = (uint32_t)t | code;
            ++cnt;
        }
    }
    if (cnt > cap) { atomicMax(&flags[2], cnt); flags[0] = 1; cnt = cap; }
    row_len[s] = cnt;
    mdg_pad_row(row, cnt, cap, (uint32_t)s);
}

// Warp-per-atom form of k_build_cells for small systems (SchNet boxes: a few thousand atoms): the 27 stencil cells of the
// atom are scanned 32 candidates at a time, hits are appended in ascending order through a ballot + prefix popcount - the same
// rows as the thread-per-atom kernel, whose serial scan of ~200 exact tests is pure latency there (81 us for 4096 atoms).
__global__ void __launch_bounds__(128) k_build_cells_warp(int s0, int n, const float4* __restrict__ qs, const int* __restrict__ cell_sorted,
                                                          const int* __restrict__ cell_start, const int* __restrict__ stencil,
                                                          Box bx, float r2max, int cap, PairFilter F,
                                                          uint32_t* __restrict__ rows, int* __restrict__ row_len,
                                                          int* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int s = s0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // rows [s0, n)
    if (s >= n) return;
    if (flags[6] | flags[7]) { if (lane == 0) row_len[s] = 0; return; }
    const float4 qi = qs[s];
    const int idi = __float_as_int(qi.w);
    const int c = cell_sorted[s];
    uint32_t* row = rows + (size_t)s * cap;
    int cnt = 0;
    const bool filt = (F.sel_a != nullptr) || (F.n_ex > 0);
    for (int k = 0; k < 27; ++k) {
        const int cc = stencil[c * 27 + k];
        const int t0 = cell_start[cc], t1 = cell_start[cc + 1];
        for (int tb = t0; tb < t1; tb += 32) {
            const int t = tb + lane;
            bool hit = false;
            uint32_t code = 0;
            if (t < t1 && t != s) {
                const float4 qj = qs[t];
                hit = test_pair(qi, qj, bx, r2max, code);
                if (hit && filt) hit = pair_allowed(F, idi, __float_as_int(qj.w));
            }
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const int kk = cnt + __popc(m & ((1u << lane) - 1u));
                if (kk < cap) row[kk] = (uint32_t)t | code;
            }
            cnt += __popc(m);
        }
    }
    if (lane == 0) {
        if (cnt > cap) { atomicMax(&flags[2], cnt); flags[0] = 1; cnt = cap; }
        row_len[s] = cnt;
        mdg_pad_row(row, cnt, cap, (uint32_t)s);
    }
}

// ---------------------------------------------------------------------------------------------
// Fast builder for the engine's Verlet-SKIN list.  Membership at the list radius does not have to
// be bit-exact (the force kernel re-tests every entry against rc^2 with the reference arithmetic),
// so candidates are screened with a 7-instruction FMA distance test on cell-local coordinates that
// are staged once per cell in shared memory.  One warp per cell, one lane per atom of the cell; all
// lanes scan the same staged candidate stream (broadcast LDS.128, no bank conflicts), rows come
// out in ascending neighbor index.  The image code of an accepted pair is the integer image
// difference of the two atoms (x = (frac + n) L): off = -(I_j - I_i), identical to the reference's
// -(red > 0.5) + (red < -0.5) for every pair that can be within the cutoff; |m| >= 2 pairs are the
// ones the reference's single +-1 correction loses (SURVEY 7 "unwrapped positions") and are dropped.
// ---------------------------------------------------------------------------------------------
#include "build_fast.cuh"
#include "build_tiles.cuh"

// written per sorted atom: its cell id (needed by k_build_cells)
__global__ void k_cell_sorted(int ncell, const int* __restrict__ cell_start, int* __restrict__ cell_sorted) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    for (int t = cell_start[c]; t < cell_start[c + 1]; ++t) cell_sorted[t] = c;
}

// All-pairs builder for small boxes (path 1): one WARP per atom.  Lanes stride the candidates in ascending index, the exact
// reference membership test (test_pair) runs per lane, and a ballot + prefix popcount appends the hits in ascending order -
// the same rows as the first version (one thread per atom, 27 - 65 us for 108 - 192 atoms: a serial chain of n exact tests),
// at 1/32 of the dependent-chain length.
#define AP_TILE 128
__global__ void __launch_bounds__(AP_TILE) k_build_allpairs(int n, const float4* __restrict__ qs, Box bx, float r2max,
                                                            int cap, PairFilter F, uint32_t* __restrict__ rows,
                                                            int* __restrict__ row_len, int* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n) return;
    const float4 qi = qs[s];
    if (!(isfinite(qi.x) && isfinite(qi.y) && isfinite(qi.z))) flags[6] = 1;
    const int idi = __float_as_int(qi.w);
    uint32_t* row = rows + (size_t)s * cap;
    int cnt = 0;
    const bool filt = (F.sel_a != nullptr) || (F.n_ex > 0);
    for (int t0 = 0; t0 < n; t0 += 32) {
        const int t = t0 + lane;
        bool hit = false;
        uint32_t code = 0;
        if (t < n && t != s) {
            const float4 qj = qs[t];
            hit = test_pair(qi, qj, bx, r2max, code);
            if (hit && filt) hit = pair_allowed(F, idi, __float_as_int(qj.w));
        }
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const int k = cnt + __popc(m & ((1u << lane) - 1u));
            if (k < cap) row[k] = (uint32_t)t | code;
        }
        cnt += __popc(m);
    }
    if (lane == 0) {
        if (cnt > cap) { atomicMax(&flags[2], cnt); flags[0] = 1; cnt = cap; }
        row_len[s] = cnt;
        mdg_pad_row(row, cnt, cap, (uint32_t)s);
    }
}

// ---------------------------------------------------------------------------------------------
// host: build
// ---------------------------------------------------------------------------------------------
static inline int roundup(int x, int m) { return (x + m - 1) / m * m; }

// d_xyz (n x 3, original order) XOR d_q4_in (float4 with ids in .w, any order).
int mdg_i_build_list(mdg_ctx* c, const float* d_xyz, const float4* d_q4_in, int n, const float* h_cell3,
                     double rlist, double cutoff, cudaStream_t st) {
    if (n < 0 || n >= MDG_MAX_ATOMS) { mdg_set_error("n=%d out of range", n); return MDG_E_BADARG; }
    for (int k = 0; k < 3; ++k)
        if (!(h_cell3[k] > 0.f)) { mdg_set_error("cell length %d must be > 0", k); return MDG_E_BADARG; }
    c->n = n;
    c->built = false;
    Grid g;
    for (int k = 0; k < 3; ++k) {
        c->box.L[k] = h_cell3[k];
        c->box.invL[k] = 1.0f / h_cell3[k];  // correctly rounded fp32 reciprocal
        g.L[k] = c->box.L[k];
        g.invL[k] = c->box.invL[k];
        int nck = (int)((double)h_cell3[k] / (rlist * 1.0001));
        g.nc[k] = nck < 1 ? 1 : nck;
    }
    c->rlist2 = (float)(rlist * rlist);
    c->rc2 = (float)(cutoff * cutoff);
    bool cells_ok = g.nc[0] >= 3 && g.nc[1] >= 3 && g.nc[2] >= 3;
    int path = (cells_ok && n > ALLPAIRS_MAX_ATOMS) ? 0 : 1;
    // limit cell count relative to atoms (very dilute boxes): keep >= ~2 atoms per cell on average
    if (path == 0) {
        while ((int64_t)g.nc[0] * g.nc[1] * g.nc[2] > (int64_t)n && g.nc[0] > 3 && g.nc[1] > 3 && g.nc[2] > 3) {
            for (int k = 0; k < 3; ++k) g.nc[k] = g.nc[k] > 3 ? g.nc[k] - 1 : 3;
        }
    }
    bool grid_changed = (path != c->path) || g.nc[0] != c->nc[0] || g.nc[1] != c->nc[1] || g.nc[2] != c->nc[2] || c->ncell == 0;
    c->path = path;
    for (int k = 0; k < 3; ++k) c->nc[k] = g.nc[k];
    int ncell = g.nc[0] * g.nc[1] * g.nc[2];
    c->ncell = ncell;
    if (n == 0) { c->built = true; c->npairs = 0; c->stat_entries = 0; return MDG_OK; }

    // capacity estimate from the mean density (grown by the caller on overflow)
    if (c->cap == 0) {
        double vol = (double)h_cell3[0] * h_cell3[1] * h_cell3[2];
        double avg = (double)n / vol * 4.18879 * rlist * rlist * rlist;
        int want = roundup((int)(avg * 1.5) + 24, 32);
        if (want > roundup(n, 32)) want = roundup(n, 32);
        if (want < 32) want = 32;
        c->cap = want;
    }
    MDG_TRY(c->qs_buf[0].reserve(sizeof(float4) * (size_t)n));
    MDG_TRY(c->qs_buf[1].reserve(sizeof(float4) * (size_t)n));
    MDG_TRY(c->perm.reserve(sizeof(int) * (size_t)n));
    if (c->rows_wanted) MDG_TRY(c->row_len.reserve(sizeof(int) * (size_t)n));
    if (c->slab && (path != 0 || g.nc[2] < 2)) {
        mdg_set_error("distributed slab decomposition needs the cell-list path (N > %d, >= 3 cells per axis)", ALLPAIRS_MAX_ATOMS);
        return MDG_E_BADARG;
    }
    MDG_TRY(c->flags.reserve(sizeof(int) * 8));
    MDG_TRY(c->cell_of.reserve(sizeof(int) * (size_t)n));   // reused as cell_sorted after the scatter
    const int T = 256;
    int nb = (n + T - 1) / T;

    // input as float4 + id
    const float4* qin = d_q4_in;
    if (!qin) {
        MDG_TRY(c->q4b.reserve(sizeof(float4) * (size_t)n));
        k_pack_xyz<<<nb, T, 0, st>>>(d_xyz, c->q4b.as<float4>(), n);
        c->stat_launches++;
        qin = c->q4b.as<float4>();
    }
    PairFilter F{c->sel_a, c->sel_b, c->ex_keys, c->n_ex, n};
    float4* qs = (qin == c->qs_buf[0].as<float4>()) ? c->qs_buf[1].as<float4>() : c->qs_buf[0].as<float4>();
    c->qs_ptr = qs;
    if (!c->flags_sticky) {      // (the engine clears the flags once per epoch: an overflow of ANY rebuild must survive to its end)
        MDG_CUDA(cudaMemsetAsync(c->flags.p, 0, sizeof(int) * 3, st));   // [0] overflow, [1] staged-atom demand (tiles), [2] max row count
        MDG_CUDA(cudaMemsetAsync(c->flags.as<int>() + 6, 0, sizeof(int) * 2, st));   // [6] non-finite, [7] cell overflow
    }
    if (path == 0) {
        MDG_TRY(c->slot_of.reserve(sizeof(int) * (size_t)n));
        MDG_TRY(c->perm_tmp.reserve(sizeof(int) * (size_t)n));
        MDG_TRY(c->cell_count.reserve(sizeof(int) * (size_t)(ncell + 1)));
        MDG_TRY(c->cell_start.reserve(sizeof(int) * (size_t)(ncell + 2)));
        MDG_TRY(c->stencil.reserve(sizeof(int) * (size_t)ncell * 27));
        int ncb = (ncell + T - 1) / T;
        if (grid_changed) { k_stencil<<<ncb, T, 0, st>>>(g, c->stencil.as<int>()); c->stat_launches++; }
        MDG_CUDA(cudaMemsetAsync(c->cell_count.p, 0, sizeof(int) * (size_t)(ncell + 1), st));
        const bool local_rebuild = c->slab && c->slab_local && qin == d_q4_in;
        if (!local_rebuild) {
            k_bin<<<nb, T, 0, st>>>(qin, 0, n, g, c->cell_of.as<int>(), c->slot_of.as<int>(), c->cell_count.as<int>(), c->flags.as<int>());
            MDG_TRY(mdg_i_scan_exclusive(c, c->cell_count.as<int>(), c->cell_start.as<int>(), ncell + 1, nullptr, st));
            k_scatter<<<nb, T, 0, st>>>(0, n, c->cell_of.as<int>(), c->slot_of.as<int>(), c->cell_start.as<int>(), c->perm_tmp.as<int>(),
                                        g.nc[0] * g.nc[1], g.nc[2], 0, 0);
        } else {
            // ---- distributed LOCAL rebuild: nothing here scales with the total atom count --------------------------
            // The engine has just refreshed q (and v, vh) on the old layers [zlo-2, zhi+2) through a two-layer halo
            // exchange.  An atom moves far less than a layer between rebuilds, so every atom of the new layers
            // [zlo-1, zhi+1) is among them: bin those, all-reduce the per-layer totals (ncz ints) for the global layer
            // offsets, scan the counts of the window's cells, place the window's atoms.
            NcclApi* N = mdg_nccl();
            const int nxy = g.nc[0] * g.nc[1], ncz = g.nc[2];
            int plan[4];
            MDG_TRY(mdg_slab_plan(ncz, c->dist_world, c->dist_rank, plan));
            const int zlo = plan[0], zhi = plan[1];
            if (zhi - zlo < 2 || (zhi - zlo) + 4 > ncz) {
                mdg_set_error("distributed run: every rank needs >= 2 cell layers and the box >= own + 4 layers (ncz=%d, world=%d)", ncz, c->dist_world);
                return MDG_E_BADARG;
            }
            if (ncz + 1 != c->n_layers) { mdg_set_error("cell grid changed during a distributed run"); return MDG_E_STATE; }
            const int* Lold = c->h_layers;                        // offsets of the PREVIOUS sort
            int pieces[3][2];
            int np = 0;
            {   // old layers zlo-2 .. zhi+1 (periodic) as contiguous index ranges
                int za = zlo - 2, zb = zhi + 2;
                if (za < 0) { pieces[np][0] = Lold[za + ncz]; pieces[np][1] = Lold[ncz]; ++np; za = 0; }
                int zb_in = zb > ncz ? ncz : zb;
                pieces[np][0] = Lold[za]; pieces[np][1] = Lold[zb_in]; ++np;
                if (zb > ncz) { pieces[np][0] = Lold[0]; pieces[np][1] = Lold[zb - ncz]; ++np; }
            }
            for (int k = 0; k < np; ++k) {
                int cnt = pieces[k][1] - pieces[k][0];
                if (cnt > 0)
                    k_bin<<<(cnt + T - 1) / T, T, 0, st>>>(qin, pieces[k][0], pieces[k][1], g, c->cell_of.as<int>(), c->slot_of.as<int>(),
                                                           c->cell_count.as<int>(), c->flags.as<int>());
            }
            MDG_TRY(c->lay_tot.reserve(sizeof(int) * (size_t)(2 * ncz + 4)));
            int* lay_tot = c->lay_tot.as<int>();
            int* lay_off = lay_tot + ncz + 1;
            k_layer_totals<<<ncz, SCAN_THREADS, 0, st>>>(c->cell_count.as<int>(), nxy, ncz, zlo, zhi, lay_tot);
            if (c->dist_p2p && ncz <= MDG_DIST_MAXLAY) {
                const int par = c->dist_rebuilds & 1, rs = ++c->dist_rebuilds;
                PeerTab PT;
                for (int r = 0; r < MDG_DIST_MAXW; ++r) PT.s[r] = (DistSync*)c->peer_sync[r < c->dist_world ? r : c->dist_rank];
                k_dist_layers_push<<<1, 256, 0, st>>>(lay_tot, zlo, zhi, PT, c->dist_rank, c->dist_world, par, rs);
                k_layer_offsets_p2p<<<1, 32, 0, st>>>((DistSync*)c->dsync.p, c->dist_world, par, rs, ncz, lay_off);
            } else {
                MDG_TRY(mdg_nccl_check(N->AllReduce(lay_tot, lay_tot, (size_t)ncz, MDG_NCCL_INT32, MDG_NCCL_SUM, c->dist_comm, st), "AllReduce"));
                k_layer_offsets<<<1, 32, 0, st>>>(lay_tot, ncz, lay_off);
            }
            const int zwin0 = (zlo - 1 + ncz) % ncz, zwin_n = zhi - zlo + 2;
            k_cellstart_layer<<<zwin_n, SCAN_THREADS, 0, st>>>(c->cell_count.as<int>(), lay_off, nxy, ncz, zwin0, c->cell_start.as<int>());
            for (int k = 0; k < np; ++k) {
                int cnt = pieces[k][1] - pieces[k][0];
                if (cnt > 0)
                    k_scatter<<<(cnt + T - 1) / T, T, 0, st>>>(pieces[k][0], pieces[k][1], c->cell_of.as<int>(), c->slot_of.as<int>(),
                                                               c->cell_start.as<int>(), c->perm_tmp.as<int>(), nxy, ncz, zwin0, zwin_n);
            }
            // new layer offsets -> host (SYNC): sizes the halo messages until the next rebuild
            MDG_CUDA(cudaMemcpyAsync(c->h_layers, lay_off, sizeof(int) * (size_t)(ncz + 1), cudaMemcpyDeviceToHost, st));
            MDG_CUDA(cudaStreamSynchronize(st));
            c->layers_fresh = true;
            c->stat_launches += 6;
        }
        if (!c->slab) {
            k_cellsort_warp<<<(ncell + 7) / 8, 256, 0, st>>>(0, ncell, c->cell_start.as<int>(), c->cell_count.as<int>(), c->perm_tmp.as<int>(),
                                                             qin, qs, c->perm.as<int>(), c->cell_of.as<int>(), c->flags.as<int>());
        } else {
            // distributed: only the cells of this rank's slab and of its two ghost layers are ever read -> sort just those
            // (the binning / scan above stay global so that every rank keeps the same sorted index space)
            int nxy = g.nc[0] * g.nc[1], ncz = g.nc[2];
            int plan[4];
            MDG_TRY(mdg_slab_plan(ncz, c->dist_world, c->dist_rank, plan));
            int zl = (plan[0] - 1 + ncz) % ncz, zu = plan[1] % ncz;
            int ranges[3][2] = {{plan[0] * nxy, plan[1] * nxy}, {zl * nxy, (zl + 1) * nxy}, {zu * nxy, (zu + 1) * nxy}};
            for (int k = 0; k < 3; ++k) {
                if (k == 2 && zu == zl) break;                       // single ghost layer shared by both sides
                int nc_k = ranges[k][1] - ranges[k][0];
                if (nc_k <= 0) continue;
                k_cellsort_warp<<<(nc_k + 7) / 8, 256, 0, st>>>(ranges[k][0], ranges[k][1], c->cell_start.as<int>(), c->cell_count.as<int>(),
                                                                c->perm_tmp.as<int>(), qin, qs, c->perm.as<int>(), c->cell_of.as<int>(),
                                                                c->flags.as<int>());
            }
        }
        // ---- range of rows this context owns: everything, or (distributed) whole z-layers of cells -------
        c->own_s0 = 0; c->own_s1 = n; c->own_c0 = 0; c->own_c1 = ncell; c->rows_s0 = 0;
        if (c->slab) {
            // atom offset of every z-layer of cells -> host (SYNC, once per rebuild): sizes the halo / all-gather messages
            int nxy = g.nc[0] * g.nc[1], ncz = g.nc[2];
            int plan[4];
            MDG_TRY(mdg_slab_plan(ncz, c->dist_world, c->dist_rank, plan));
            c->slab_zlo = plan[0]; c->slab_zhi = plan[1];
            if (!c->h_layers || c->n_layers < ncz + 1) {
                if (c->h_layers) cudaFreeHost(c->h_layers);
                MDG_CUDA(cudaMallocHost((void**)&c->h_layers, sizeof(int) * (size_t)(ncz + 1)));
            }
            c->n_layers = ncz + 1;
            if (!c->layers_fresh) {
                MDG_CUDA(cudaMemcpy2DAsync(c->h_layers, sizeof(int), c->cell_start.as<int>(), sizeof(int) * (size_t)nxy, sizeof(int),
                                           (size_t)(ncz + 1), cudaMemcpyDeviceToHost, st));
                MDG_CUDA(cudaStreamSynchronize(st));
            }
            c->layers_fresh = false;
            c->own_c0 = c->slab_zlo * nxy; c->own_c1 = c->slab_zhi * nxy;
            c->own_s0 = c->h_layers[c->slab_zlo]; c->own_s1 = c->h_layers[c->slab_zhi];
            c->rows_s0 = c->own_s0;
        }
        c->tiles = false;
        bool roomy = g.nc[0] >= 5 && g.nc[1] >= 5 && g.nc[2] >= 5;   // stencil extent < half a box
        if (c->rows_wanted && c->fast_build && rlist > cutoff && roomy && !c->tiles_off) {
            // ---- tile list (tiles.cuh): block-local 16-bit rows for k_force_tiles --------------------------------
            const double occ = (double)n / (double)ncell;
            int maxw = g.nc[0] - 2 < MDG_TILE_MAXW ? g.nc[0] - 2 : MDG_TILE_MAXW;
            int scap = 0;
            for (; maxw >= 1; --maxw) {
                scap = roundup((int)(9.0 * (maxw + 2) * occ * 1.2) + 64, 32);
                if (scap < c->tile_scap_min) scap = roundup(c->tile_scap_min, 64);
                if (scap <= MDG_TILE_MAXSCAP && maxw * occ * 1.1 <= 24.0 * MDG_TILE_GROUP) break;
            }
            if (maxw >= 1 && c->cap / MDG_TILE_CHUNK <= 255) {
                TileGeom& G = c->tile;
                G.ncx = g.nc[0]; G.ncy = g.nc[1]; G.ncz = g.nc[2];
                G.nblk = (g.nc[0] + maxw - 1) / maxw;
                G.wbase = g.nc[0] / G.nblk;
                G.wrem = g.nc[0] % G.nblk;
                G.capc = c->cap / MDG_TILE_CHUNK;
                G.scap = scap;
                const int xr0 = c->own_c0 / g.nc[0], nxr = (c->own_c1 - c->own_c0) / g.nc[0];
                G.b_base = xr0 * G.nblk;
                G.g_base = c->own_s0 >> 3;
                const int nblocks = nxr * G.nblk;
                const size_t ngroups = (size_t)((c->own_s1 - c->own_s0) >> 3) + (size_t)nblocks + 2;
                MDG_TRY(c->tile_rows.reserve(sizeof(uint16_t) * ngroups * (size_t)G.capc * MDG_TILE_GCHUNK));
                MDG_TRY(c->tile_len.reserve(sizeof(uint32_t) * (size_t)n));
                MDG_TRY(c->tile_desc.reserve(sizeof(int) * (size_t)(nblocks + 1) * MDG_TILE_DESC));
                if (c->tile_warps_env > 0) c->tile_warps = c->tile_warps_env > 11 ? 11 : c->tile_warps_env;
                else {
                    int tw = (int)((occ * g.nc[0] / G.nblk * 1.1 + MDG_TILE_GROUP - 1) / MDG_TILE_GROUP);
                    c->tile_warps = tw < 2 ? 2 : (tw > 11 ? 11 : tw);      // consumer warps (+ 1 producer warp <= 384 threads)
                }
                if (nblocks > 0)
                    k_build_tiles<<<dim3(G.nblk, g.nc[1], nxr / g.nc[1]), MDG_TILE_MAXW * 32, 0, st>>>(xr0 / g.nc[1], G, qs, c->cell_start.as<int>(), c->box, c->rlist2, F,
                                                                         c->tile_rows.as<uint16_t>(), c->tile_len.as<uint32_t>(),
                                                                         c->tile_desc.as<int>(), c->flags.as<int>());
                c->tiles = true;
            }
        }
        if (c->rows_wanted && !c->tiles) {
            MDG_TRY(c->rows.reserve(sizeof(uint32_t) * (size_t)(c->own_s1 - c->own_s0 + 1) * c->cap));
            uint32_t* rows_base = c->rows.as<uint32_t>() - (size_t)c->rows_s0 * c->cap;
            if (c->fast_build && rlist > cutoff && roomy) {
                int ncl = c->own_c1 - c->own_c0;
                if (ncl > 0) {
                    MDG_TRY(c->work_ctr.reserve(sizeof(int) * 4));
                    MDG_CUDA(cudaMemsetAsync(c->work_ctr.p, 0, sizeof(int), st));          // (k_build_fast hands out cells through it)
                    k_build_fast<<<(ncl + FB_WARPS - 1) / FB_WARPS, FB_WARPS * 32, 0, st>>>(
                        c->own_c0, c->own_c1, qs, c->cell_start.as<int>(), c->stencil.as<int>(), c->box, g.nc[0], g.nc[1], g.nc[2],
                        c->rlist2, c->cap, F, rows_base, c->row_len.as<int>(), c->flags.as<int>(),
                        c->work_ctr.as<int>());
                }
            } else {
                int nl = c->own_s1 - c->own_s0;
                if (nl > 0 && nl <= 65536)       // small systems: a warp per atom (the thread-per-atom scan is latency-bound there)
                    k_build_cells_warp<<<(nl + 3) / 4, 128, 0, st>>>(c->own_s0, c->own_s1, qs, c->cell_of.as<int>(), c->cell_start.as<int>(),
                                                                    c->stencil.as<int>(), c->box, c->rlist2, c->cap, F, rows_base,
                                                                    c->row_len.as<int>(), c->flags.as<int>());
                else if (nl > 0)
                    k_build_cells<<<(nl + 127) / 128, 128, 0, st>>>(c->own_s0, c->own_s1, qs, c->cell_of.as<int>(), c->cell_start.as<int>(),
                                                                  c->stencil.as<int>(), c->box, c->rlist2, c->cap, F, rows_base,
                                                                  c->row_len.as<int>(), c->flags.as<int>());
            }
        }
        c->stat_launches += 3 + (c->rows_wanted ? 1 : 0);
    } else {
        MDG_CUDA(cudaMemcpyAsync(qs, qin, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToDevice, st));
        k_iota<<<nb, T, 0, st>>>(c->perm.as<int>(), n);
        c->own_s0 = 0; c->own_s1 = n; c->own_c0 = 0; c->own_c1 = ncell; c->rows_s0 = 0;
        c->tiles = false;
        if (c->rows_wanted) MDG_TRY(c->rows.reserve(sizeof(uint32_t) * (size_t)n * c->cap));
        if (c->rows_wanted)
            k_build_allpairs<<<(n + AP_TILE / 32 - 1) / (AP_TILE / 32), AP_TILE, 0, st>>>(n, qs, c->box, c->rlist2, c->cap, F,
                                                                            c->rows.as<uint32_t>(), c->row_len.as<int>(),
                                                                            c->flags.as<int>());
        c->stat_launches += 1 + (c->rows_wanted ? 1 : 0);
    }
    MDG_KERNEL_CHECK();
    c->built = c->rows_wanted && !c->tiles;      // (tile rows only serve the engine's force kernel)
    c->stat_rebuilds++;
    return MDG_OK;
}

// ---------------------------------------------------------------------------------------------
// export in the reference layout and order (topology.py:66-73): rows i<j by ORIGINAL ids,
// sorted (i, then j)
// ---------------------------------------------------------------------------------------------
__global__ void k_export_count(int n, const float4* __restrict__ qs, const uint32_t* __restrict__ rows,
                               const int* __restrict__ row_len, int cap, int* __restrict__ up_cnt) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int idi = __float_as_int(qs[s].w);
    const uint32_t* row = rows + (size_t)s * cap;
    int m = row_len[s], cnt = 0;
    for (int k = 0; k < m; ++k) {
        int t = row[k] & MDG_IDX_MASK;
        cnt += (__float_as_int(qs[t].w) > idi);
    }
    up_cnt[idi] = cnt;
}

// One WARP per row: the original ids of the row's neighbors are staged in shared memory once, then every lane ranks its
// entries against them (broadcast LDS) - O(m^2 / 32) shared-memory reads per row instead of the O(m^2) dependent global
// gathers per THREAD of the first version (178 us for 108 atoms: the top kernel of the small-box regime).
#define EXPORT_WARPS 4
__global__ void __launch_bounds__(EXPORT_WARPS * 32) k_export_fill(int n, const float4* __restrict__ qs, const uint32_t* __restrict__ rows,
                              const int* __restrict__ row_len, int cap, const int* __restrict__ up_off, Box bx,
                              int64_t* __restrict__ nbr, float* __restrict__ offsets, float* __restrict__ dis,
                              int64_t cap_pairs) {
    extern __shared__ int s_ids[];                       // [warps in the block][cap]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int s = blockIdx.x * (blockDim.x >> 5) + w;
    if (s >= n) return;
    int* ids = s_ids + (size_t)w * cap;
    const float4 qi = qs[s];
    const int idi = __float_as_int(qi.w);
    const uint32_t* row = rows + (size_t)s * cap;
    const int m = row_len[s];
    for (int k = lane; k < m; k += 32) ids[k] = __float_as_int(qs[row[k] & MDG_IDX_MASK].w);
    __syncwarp();
    const int64_t base = up_off[idi];
    for (int k = lane; k < m; k += 32) {
        const int idj = ids[k];
        if (idj <= idi) continue;
        int rank = 0;
        for (int k2 = 0; k2 < m; ++k2) {
            const int id2 = ids[k2];
            rank += (id2 > idi && id2 < idj);
        }
        const int64_t p = base + rank;
        if (p >= cap_pairs) continue;      // asynchronous export into a buffer sized from an earlier count (k_latch reports it)
        const uint32_t e = row[k];
        const float4 qj = qs[e & MDG_IDX_MASK];
        nbr[2 * p] = idi;
        nbr[2 * p + 1] = idj;
        const uint32_t code = e >> MDG_IDX_BITS;
        const float ox = (float)((int)(code & 3u) - 1), oy = (float)((int)((code >> 2) & 3u) - 1),
                    oz = (float)((int)((code >> 4) & 3u) - 1);
        offsets[3 * p] = ox;
        offsets[3 * p + 1] = oy;
        offsets[3 * p + 2] = oz;
        if (dis) {
            float dx = __fadd_rn(__fsub_rn(qj.x, qi.x), __fmul_rn(ox, bx.L[0]));
            float dy = __fadd_rn(__fsub_rn(qj.y, qi.y), __fmul_rn(oy, bx.L[1]));
            float dz = __fadd_rn(__fsub_rn(qj.z, qi.z), __fmul_rn(oz, bx.L[2]));
            dis[p] = __fsqrt_rn(mdg_d2_exact(dx, dy, dz));
        }
    }
}

// launch shape of k_export_fill for a row capacity: warps per block limited by 48 KB of dynamic shared memory
static inline int export_warps(int cap) {
    int w = (int)((48 * 1024) / (sizeof(int) * (size_t)(cap > 0 ? cap : 1)));
    return w < 1 ? 1 : (w > EXPORT_WARPS ? EXPORT_WARPS : w);
}

int mdg_i_export_count(mdg_ctx* c, cudaStream_t st, int64_t* h_npairs) {
    int n = c->n;
    if (n == 0) { *h_npairs = 0; c->npairs = 0; return MDG_OK; }
    MDG_TRY(c->up_cnt.reserve(sizeof(int) * (size_t)(n + 1)));
    MDG_TRY(c->up_off.reserve(sizeof(int) * (size_t)(n + 2)));
    const int T = 256;
    int nb = (n + T - 1) / T;
    k_export_count<<<nb, T, 0, st>>>(n, c->qs_ptr, c->rows.as<uint32_t>(), c->row_len.as<int>(), c->cap,
                                     c->up_cnt.as<int>());
    c->stat_launches++;
    int* d_total = c->flags.as<int>() + 4;
    MDG_TRY(mdg_i_scan_exclusive(c, c->up_cnt.as<int>(), c->up_off.as<int>(), n, d_total, st));
    MDG_CUDA(cudaMemcpyAsync(c->h_pinned, c->flags.p, sizeof(int) * 8, cudaMemcpyDeviceToHost, st));
    MDG_CUDA(cudaStreamSynchronize(st));
    if (c->h_pinned[6] || c->h_pinned[7]) {
        mdg_set_error("neighbor list: non-finite coordinates or collapsed cell (occupancy %d)", c->h_pinned[7]);
        return MDG_E_NUMERIC;
    }
    if (c->h_pinned[0]) return MDG_E_CAPACITY;
    c->npairs = c->h_pinned[4];
    *h_npairs = c->npairs;
    return MDG_OK;
}

int mdg_i_export_fill(mdg_ctx* c, int64_t* d_nbr, float* d_offsets, float* d_dis, cudaStream_t st) {
    int n = c->n;
    if (n == 0 || c->npairs == 0) return MDG_OK;
    const int T = 128;
    const int ew = export_warps(c->cap);
    if ((size_t)c->cap * sizeof(int) > 48 * 1024) { mdg_set_error("neighbor-list export: row capacity %d too large", c->cap); return MDG_E_CAPACITY; }
    k_export_fill<<<(n + ew - 1) / ew, ew * 32, sizeof(int) * (size_t)ew * c->cap, st>>>(n, c->qs_ptr, c->rows.as<uint32_t>(), c->row_len.as<int>(),
                                                 c->cap, c->up_off.as<int>(), c->box, d_nbr, d_offsets, d_dis, INT64_MAX);
    c->stat_launches++;
    MDG_KERNEL_CHECK();
    return MDG_OK;
}

// ---------------------------------------------------------------------------------------------
// Asynchronous list build for the device engine's GNN / Stack epochs (engine.cu): the same exact list as
// mdg_nbr_build, but NOTHING is read back - the host keeps enqueueing the step while the list is being built.
//   * the pair count stays on the device (flags[4]); consumers take it from there, bounded by `cap_pairs`, the size of
//     the buffers the host allocated from an EARLIER step's count;
//   * a row-capacity overflow, non-finite input or a pair count above cap_pairs is LATCHED in flags[3] (bit 0 / 1 / 2),
//     which survives later builds; the engine reads it once, at the end of the epoch, and repeats the epoch on the
//     synchronous path (which grows the capacities) if it is set.
// ---------------------------------------------------------------------------------------------
__global__ void k_latch(int* __restrict__ flags, int check_total, int64_t cap_pairs) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int v = 0;
        if (flags[0]) v |= 1;
        if (flags[6] | flags[7]) v |= 2;
        if (check_total && (int64_t)flags[4] > cap_pairs) v |= 4;
        if (v) flags[3] |= v;
    }
}

// d_q4 (optional): the same positions already packed as float4 with w = atom id (the engine's state) - saves the pack kernel
int mdg_i_nbr_build_async(mdg_ctx* c, const float* d_xyz, const float4* d_q4, int n, const float* h_cell3, double cutoff, const uint8_t* d_sel_a,
                          const uint8_t* d_sel_b, const int64_t* d_ex_keys, int n_ex, bool want_export, int64_t cap_pairs,
                          int64_t* d_nbr, float* d_offsets, cudaStream_t st) {
    c->sel_a = d_sel_a;
    c->sel_b = d_sel_b;
    c->ex_keys = d_ex_keys;
    c->n_ex = d_ex_keys ? n_ex : 0;
    c->rows_wanted = true;
    c->fast_build = false;       // exact membership, as mdg_nbr_build
    c->stat_launches = 0;
    MDG_TRY(mdg_i_build_list(c, d_xyz, d_q4, n, h_cell3, cutoff, cutoff, st));
    if (n == 0) return MDG_OK;
    if (want_export) {
        MDG_TRY(c->up_cnt.reserve(sizeof(int) * (size_t)(n + 1)));
        MDG_TRY(c->up_off.reserve(sizeof(int) * (size_t)(n + 2)));
        const int T = 256;
        k_export_count<<<(n + T - 1) / T, T, 0, st>>>(n, c->qs_ptr, c->rows.as<uint32_t>(), c->row_len.as<int>(), c->cap,
                                                     c->up_cnt.as<int>());
        MDG_TRY(mdg_i_scan_exclusive(c, c->up_cnt.as<int>(), c->up_off.as<int>(), n, c->flags.as<int>() + 4, st));
        const int ew = export_warps(c->cap);
        k_export_fill<<<(n + ew - 1) / ew, ew * 32, sizeof(int) * (size_t)ew * c->cap, st>>>(n, c->qs_ptr, c->rows.as<uint32_t>(), c->row_len.as<int>(), c->cap,
                                                      c->up_off.as<int>(), c->box, d_nbr, d_offsets, nullptr, cap_pairs);
        c->stat_launches += 2;
    }
    k_latch<<<1, 32, 0, st>>>(c->flags.as<int>(), want_export ? 1 : 0, cap_pairs);
    c->stat_launches++;
    MDG_KERNEL_CHECK();
    return MDG_OK;
}

// ---------------------------------------------------------------------------------------------
// K6  RDF: Gaussian-smeared pair-distance histogram (reference torchmd/observable.py:62-76,
// nff/nn/layers.py:14-31).  Traverses the cell structure directly (no stored rows); every
// undirected pair (id_j > id_i) with d < end + 0.5 contributes exp(-0.5/w^2 (d - mu_k)^2) to
// bin k.  Terms beyond +-RDF_NSIG widths (< 3e-10 of a unit term) are skipped.
// ---------------------------------------------------------------------------------------------
#define RDF_NSIG 6.6f
#define RDF_MAX_BINS 2048

struct RdfArgs {
    float start, dmu, inv_dmu, coeff, win;  // mu_k = start + k*dmu ; coeff = -0.5/w^2 ; win = NSIG*w
    int   nbins;
};

__device__ __forceinline__ void rdf_add(const RdfArgs& R, float d, float* sh) {
    int k0 = (int)ceilf((d - R.win - R.start) * R.inv_dmu);
    int k1 = (int)floorf((d + R.win - R.start) * R.inv_dmu);
    k0 = max(k0, 0);
    k1 = min(k1, R.nbins - 1);
    for (int k = k0; k <= k1; ++k) {
        float x = d - (R.start + (float)k * R.dmu);
        atomicAdd(&sh[k], expf(R.coeff * x * x));
    }
}

template <bool CELLS>
__global__ void __launch_bounds__(128) k_rdf(int n, const float4* __restrict__ qs, const int* __restrict__ cell_sorted,
                                             const int* __restrict__ cell_start, const int* __restrict__ stencil, Box bx,
                                             float r2max, PairFilter F, RdfArgs R, float* __restrict__ block_hist) {
    extern __shared__ float sh[];
    for (int k = threadIdx.x; k < R.nbins; k += blockDim.x) sh[k] = 0.f;
    __syncthreads();
    bool filt = (F.sel_a != nullptr) || (F.n_ex > 0);
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        float4 qi = qs[s];
        int idi = __float_as_int(qi.w);
        if (CELLS) {
            int c = cell_sorted[s];
            for (int k = 0; k < 27; ++k) {
                int cc = stencil[c * 27 + k];
                for (int t = cell_start[cc]; t < cell_start[cc + 1]; ++t) {
                    float4 qj = qs[t];
                    int idj = __float_as_int(qj.w);
                    if (idj <= idi) continue;
                    int cx, cy, cz;
                    float dx = mdg_min_image_axis(qi.x, qj.x, bx.L[0], bx.invL[0], cx);
                    float dy = mdg_min_image_axis(qi.y, qj.y, bx.L[1], bx.invL[1], cy);
                    float dz = mdg_min_image_axis(qi.z, qj.z, bx.L[2], bx.invL[2], cz);
                    float d2 = mdg_d2_exact(dx, dy, dz);
                    if (!(d2 < r2max) || d2 == 0.f) continue;
                    if (filt && !pair_allowed(F, idi, idj)) continue;
                    rdf_add(R, __fsqrt_rn(d2), sh);
                }
            }
        } else {
            for (int t = 0; t < n; ++t) {
                float4 qj = qs[t];
                int idj = __float_as_int(qj.w);
                if (idj <= idi) continue;
                int cx, cy, cz;
                float dx = mdg_min_image_axis(qi.x, qj.x, bx.L[0], bx.invL[0], cx);
                float dy = mdg_min_image_axis(qi.y, qj.y, bx.L[1], bx.invL[1], cy);
                float dz = mdg_min_image_axis(qi.z, qj.z, bx.L[2], bx.invL[2], cz);
                float d2 = mdg_d2_exact(dx, dy, dz);
                if (!(d2 < r2max) || d2 == 0.f) continue;
                if (filt && !pair_allowed(F, idi, idj)) continue;
                rdf_add(R, __fsqrt_rn(d2), sh);
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < R.nbins; k += blockDim.x) block_hist[(size_t)blockIdx.x * R.nbins + k] = sh[k];
}

__global__ void k_rdf_reduce(int nblocks, int nbins, const float* __restrict__ block_hist, float* __restrict__ count) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nbins) return;
    double t = 0;
    for (int b = 0; b < nblocks; ++b) t += (double)block_hist[(size_t)b * nbins + k];
    count[k] += (float)t;
}

extern "C" int mdg_rdf_accumulate(mdg_ctx* c, const float* d_xyz, int n, const float* h_cell3, double start, double end,
                                  int nbins, double width, const uint8_t* d_sel_a, const uint8_t* d_sel_b,
                                  float* d_count, void* stream) {
    if (!c || !h_cell3 || !d_count || (n > 0 && !d_xyz)) { mdg_set_error("mdg_rdf_accumulate: null argument"); return MDG_E_BADARG; }
    if (nbins < 2 || nbins > RDF_MAX_BINS) { mdg_set_error("mdg_rdf_accumulate: nbins=%d not in [2,%d]", nbins, RDF_MAX_BINS); return MDG_E_BADARG; }
    if ((d_sel_a == nullptr) != (d_sel_b == nullptr)) { mdg_set_error("give both sel_a and sel_b or neither"); return MDG_E_BADARG; }
    MDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return MDG_OK;
    double rcut = end + 0.5;                 // observable.py:59
    c->sel_a = d_sel_a; c->sel_b = d_sel_b; c->ex_keys = nullptr; c->n_ex = 0;
    c->rows_wanted = false;
    int s = mdg_i_build_list(c, d_xyz, nullptr, n, h_cell3, rcut, rcut, st);
    c->rows_wanted = true;
    MDG_TRY(s);
    // mu = torch.linspace(start, end, nbins) in fp32; width default mu[1]-mu[0] (layers.py:55-59)
    RdfArgs R;
    float fstart = (float)start, fend = (float)end;
    float step = (fend - fstart) / (float)(nbins - 1);
    R.start = fstart;
    R.dmu = step;
    R.inv_dmu = 1.0f / step;
    float w = width > 0 ? (float)width : step;
    R.coeff = -0.5f / (w * w);
    R.win = RDF_NSIG * w;
    R.nbins = nbins;
    int nblocks = c->sm_count * 4;
    int maxb = (n + 127) / 128;
    if (nblocks > maxb) nblocks = maxb;
    MDG_TRY(c->partials.reserve(sizeof(float) * (size_t)nblocks * nbins + 64));
    float* bh = c->partials.as<float>();
    PairFilter F{d_sel_a, d_sel_b, nullptr, 0, n};
    size_t shb = sizeof(float) * nbins;
    if (c->path == 0)
        k_rdf<true><<<nblocks, 128, shb, st>>>(n, c->qs_ptr, c->cell_of.as<int>(), c->cell_start.as<int>(),
                                               c->stencil.as<int>(), c->box, c->rc2, F, R, bh);
    else
        k_rdf<false><<<nblocks, 128, shb, st>>>(n, c->qs_ptr, nullptr, nullptr, nullptr, c->box, c->rc2, F, R, bh);
    k_rdf_reduce<<<(nbins + 127) / 128, 128, 0, st>>>(nblocks, nbins, bh, d_count);
    c->stat_launches += 2;
    MDG_KERNEL_CHECK();
    return MDG_OK;
}

// ---------------------------------------------------------------------------------------------
// velocity autocorrelation (reference torchmd/observable.py:153-163): out[0] = mean(v * v),
// out[t] = mean(v[t:] * v[:-t]) over ALL elements of the (frames - t, N, 3) product, un-normalised as the reference.
// grid = (blocks, lags): fp64 block partials, fixed-order final sum (deterministic).
// ---------------------------------------------------------------------------------------------
#define VACF_BLOCKS 296
__global__ void __launch_bounds__(256) k_vacf_partial(const float* __restrict__ vel, int64_t frame_elems, int n_frames,
                                                      double* __restrict__ part) {
    __shared__ double sm[8];
    const int t = blockIdx.y;
    const int64_t cnt = (int64_t)(n_frames - t) * frame_elems, shift = (int64_t)t * frame_elems;
    double acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x)
        acc += (double)(vel[i + shift] * vel[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += sm[w];
        part[(size_t)t * gridDim.x + blockIdx.x] = s;
    }
}

__global__ void k_vacf_final(const double* __restrict__ part, int nblocks, int64_t frame_elems, int n_frames, int t_range,
                             float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_range) return;
    double s = 0;
    for (int b = 0; b < nblocks; ++b) s += part[(size_t)t * nblocks + b];
    const double cnt = (double)(n_frames - t) * (double)frame_elems;
    out[t] = (float)(s / cnt);
}

extern "C" int mdg_vacf(mdg_ctx* c, const float* d_vel, int n_frames, int n_atoms, int dim, int t_range, float* d_out, void* stream) {
    if (!c || !d_vel || !d_out) { mdg_set_error("mdg_vacf: null argument"); return MDG_E_BADARG; }
    if (n_frames < 1 || n_atoms < 1 || dim < 1 || t_range < 1 || t_range > n_frames) {
        mdg_set_error("mdg_vacf: need 1 <= t_range <= n_frames (frames=%d atoms=%d t_range=%d)", n_frames, n_atoms, t_range);
        return MDG_E_BADARG;
    }
    MDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    MDG_TRY(c->partials.reserve(sizeof(double) * (size_t)VACF_BLOCKS * t_range));
    const int64_t fe = (int64_t)n_atoms * dim;
    k_vacf_partial<<<dim3(VACF_BLOCKS, t_range), 256, 0, st>>>(d_vel, fe, n_frames, c->partials.as<double>());
    k_vacf_final<<<(t_range + 127) / 128, 128, 0, st>>>(c->partials.as<double>(), VACF_BLOCKS, fe, n_frames, t_range, d_out);
    c->stat_launches += 2;
    MDG_KERNEL_CHECK();
    return MDG_OK;
}
