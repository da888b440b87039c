// force.cu - listed-pair energy/force kernels (K2+K3 in SURVEY.md 2c).
// Replaces PairPotentials.forward (reference torchmd/interface.py:284-300: compute_dis
// topology.py:5-12 -> u(r).sum()) and the autograd force F = -dE/dxyz (torchmd/md.py:227-228).
#include "common.cuh"
#include "force_tiles.cuh"

PotParams mdg_make_pot(int kind, const float* h_params, int n_params) {
    PotParams P;
    P.kind = kind;
    for (int k = 0; k < MDG_MAX_POT_PARAMS; ++k) P.p[k] = (k < n_params) ? h_params[k] : 0.f;
    P.aux = 0.f;
    P.se = P.sg = 1.f;
    for (int k = 0; k < MDG_MAX_POT_PARAMS; ++k) P.sdp[k] = 1.f;
    if (kind == MDG_POT_LJ) {
        float sigma = P.p[0], eps = P.p[1];
        P.aux = sigma * sigma;
        P.se = 4.0f * eps;
        P.sg = 24.0f * eps;
        P.sdp[0] = 24.0f * eps / sigma;
        P.sdp[1] = 4.0f;
    }
    if (kind == MDG_POT_MORSE) {
        // A = 0 if phi >= 0 else exp(2a/phi) - 2 exp(a/phi)  (potentials.py:82-85, numpy double)
        double a = P.p[0], phi = P.p[1];
        P.aux = (phi >= 0) ? 0.f : (float)(exp(2 * a / phi) - 2 * exp(a / phi));
    }
    return P;
}

// ---------------------------------------------------------------------------------------------
// List-streaming force kernel: GROUP (4) lanes cooperate on one row (one atom).  Every lane streams whole
// 16-entry blocks of the row with one 16-byte evict-first load per block (rows are 128-byte aligned, cap is a
// multiple of 32, rows are padded to 32 with self entries whose d2 == 0 is dropped - no bounds guards),
// issues its 4 position gathers (sorted float4 array, L1/L2 resident) back to back, and the row sums are
// reduced with warp shuffles: one float4 store per atom.
//  * RETEST: re-apply the reference membership test d2 < rc2 (exact arithmetic) to a skin list.
//  * PURE rows: k_build_fast marks rows in which no entry carries an image shift (interior cells) with
//    MDG_ROW_PURE in row_len and stores their entries as BARE indices - the loop over such a row has no
//    index mask and no image-code test (28 instead of ~38 instructions per entry).
//  * The LJ pair evaluation is if-converted: evaluated for every lane, only the four accumulations are
//    predicated (with 32 lanes at ~61% acceptance a branch was never skipped anyway).
//  * The loop is layout-agnostic inside a 16-entry block, which lets k_build_fast store the blocks transposed
//    (build_fast.cuh fb_slot): the four lanes of a row then gather four CONSECUTIVE neighbors per load.
// Measured on the 256k-atom box (profiles/r01_ab_*.json): 59.9 us -> 57.8 (pure rows + if-conversion) -> 53.4
// (transposed blocks) -> 51.8 (evict-first row stream) -> 51.6 us (32 registers, 8 CTAs per SM).
// ---------------------------------------------------------------------------------------------
// float4 gather with a single mad.wide address computation
__device__ __forceinline__ float4 mdg_gather4(const float4* __restrict__ base, uint32_t idx) {
#ifdef MDG_EMU       // CPU emulation harness (tests/cuemu): no PTX
    return base[idx];
#else
    const float4* p;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(p) : "r"(idx), "l"(base));
    return __ldg(p);
#endif
}

#ifndef MDG_FORCE_PREFETCH
#define MDG_FORCE_PREFETCH 0
#endif
#ifndef MDG_FORCE_UNROLL
#define MDG_FORCE_UNROLL 1
#endif
#if MDG_FORCE_PREFETCH && MDG_FORCE_UNROLL != 1
#error "MDG_FORCE_PREFETCH needs MDG_FORCE_UNROLL == 1"
#endif
#ifndef MDG_FORCE_MINBLOCKS
#define MDG_FORCE_MINBLOCKS 8      // 32-register budget (measured best of the round); build variants mb6 / mb4 relax it
#endif
template <int KIND, bool RETEST, bool WITH_DP, bool PURE, int GROUP, bool WITH_E>
__device__ __forceinline__ void mdg_row_stream(const float4* __restrict__ qs, const uint32_t* __restrict__ row, int m,
                                               int lane_in_group, const float4 qi, const Box& bx, float rc2,
                                               const PotParams& P, float& fx, float& fy, float& fz, float& en, float* dpa) {
    constexpr bool IFCONV = (KIND == MDG_POT_LJ) && !WITH_DP;
    // (down-counting loop: no loop-bound register - at the 32-register budget the bound was spilled to local memory)
    // MDG_FORCE_UNROLL = 2 (build variants u2*): two 16-entry blocks per iteration, 8 gathers in flight per lane; rows are
    // padded to 32 entries, so this is valid for GROUP * 4 * 2 <= 32 only.
    constexpr int U = (GROUP * 4 * MDG_FORCE_UNROLL <= 32) ? MDG_FORCE_UNROLL : 1;
    const uint32_t* rp = row + lane_in_group * 4;
#if MDG_FORCE_PREFETCH
    // build variant pf*: the index block of the NEXT iteration is requested before this iteration's gathers are consumed
    // (register double buffer), so the row-stream latency overlaps a whole iteration instead of heading its dependency
    // chain (index load -> gather -> arithmetic).  U = 1 only.
    uint4 nxt = __ldcs(reinterpret_cast<const uint4*>(rp));
#endif
    for (int rem = m; rem > 0; rem -= GROUP * 4 * U, rp += GROUP * 4 * U) {
        // the row stream (~100 MB per launch) is read once: evict-first, so that it does not push the gathered
        // neighbor positions (4 MB, re-read ~90 times) out of L1/L2
        uint32_t es[4 * U];
#if MDG_FORCE_PREFETCH
        es[0] = nxt.x; es[1] = nxt.y; es[2] = nxt.z; es[3] = nxt.w;
        if (rem > GROUP * 4) nxt = __ldcs(reinterpret_cast<const uint4*>(rp + GROUP * 4));
#else
#pragma unroll
        for (int b = 0; b < U; ++b) {
            const uint4 e4 = __ldcs(reinterpret_cast<const uint4*>(rp + b * GROUP * 4));
            es[4 * b] = e4.x; es[4 * b + 1] = e4.y; es[4 * b + 2] = e4.z; es[4 * b + 3] = e4.w;
        }
#endif
        float4 qj[4 * U];
#pragma unroll
        for (int u = 0; u < 4 * U; ++u) qj[u] = mdg_gather4(qs, PURE ? es[u] : (es[u] & MDG_IDX_MASK));
#pragma unroll
        for (int u = 0; u < 4 * U; ++u) {
            const uint32_t e = es[u];
            float dx = __fsub_rn(qj[u].x, qi.x), dy = __fsub_rn(qj[u].y, qi.y), dz = __fsub_rn(qj[u].z, qi.z);
            if (!PURE) {
                // Image shift, branch-free: off * L = (code - 1) * L evaluated as fma(code, L, -L), which is EXACT for
                // code in {0, 1, 2} (-L, 0, 2L - L = L), so d = fl(fl(xj - xi) + off * L) keeps the reference's bits.
                // (A test "does this entry cross the boundary?" is true for some lane of nearly every warp of a
                // boundary cell - ~1/3 of their entries cross - so the branchy form ran its ~45-instruction body with its
                // three nested per-axis branches almost always: 73 instead of 28 instructions per entry on 25% of the rows.)
                const uint32_t code = e >> MDG_IDX_BITS;
                dx = __fadd_rn(dx, __fmaf_rn((float)(code & 3u), bx.L[0], -bx.L[0]));
                dy = __fadd_rn(dy, __fmaf_rn((float)((code >> 2) & 3u), bx.L[1], -bx.L[1]));
                dz = __fadd_rn(dz, __fmaf_rn((float)(code >> 4), bx.L[2], -bx.L[2]));
            }
            float d2;
            bool in;
            if (RETEST) {
                d2 = mdg_d2_exact(dx, dy, dz);
                in = (d2 < rc2) && (d2 != 0.0f);
            } else {
                d2 = dx * dx + dy * dy + dz * dz;
                in = d2 != 0.0f;
            }
            if (IFCONV) {
                float e_p, g, dp[MDG_MAX_POT_PARAMS];
                pair_eval<KIND, false>(P, d2, e_p, g, dp);      // d2 == 0 gives inf/nan here; discarded by the predicate below
                if (in) {
                    fx -= g * dx;
                    fy -= g * dy;
                    fz -= g * dz;
                    if (WITH_E) en += e_p;
                }
            } else if (in) {
                float e_p, g, dp[MDG_MAX_POT_PARAMS];
                pair_eval<KIND, WITH_DP>(P, d2, e_p, g, dp);
                fx -= g * dx;
                fy -= g * dy;
                fz -= g * dz;
                if (WITH_E) en += e_p;
                if (WITH_DP) {
#pragma unroll
                    for (int q = 0; q < MDG_MAX_POT_PARAMS; ++q) dpa[q] += dp[q];
                }
            }
        }
    }
}

// WITH_E = false: the per-atom energy (fs.w) is not accumulated (written as 0) - the MD loop only needs it after
// the last step of an epoch, and the energy costs 3 of the 28 instructions of a pure-row entry.
template <int KIND, bool RETEST, bool WITH_DP, int GROUP, bool WITH_E>
__global__ void __launch_bounds__(256, MDG_FORCE_MINBLOCKS) k_force_rows(int s0, int n, int gap_at, int gap, const float4* __restrict__ qs,
                                                                    const uint32_t* __restrict__ rows,
                                                                    const int* __restrict__ row_len, int cap, Box bx, float rc2,
                                                                    PotParams P, float4* __restrict__ fs,
                                                                    double* __restrict__ dp_partials) {
    const int lane_in_group = threadIdx.x % GROUP;
    int s = s0 + (blockIdx.x * blockDim.x + threadIdx.x) / GROUP;
    if (s >= gap_at) s += gap;       // two row ranges in one launch (multi-GPU: bottom + top boundary layer): [s0, gap_at) and [gap_at + gap, n)
    float fx = 0.f, fy = 0.f, fz = 0.f, en = 0.f;
    float dpa[MDG_MAX_POT_PARAMS] = {0.f, 0.f, 0.f, 0.f};
    if (s < n) {
        const float4 qi = qs[s];
        const uint32_t* row = rows + (size_t)s * cap;
        const int ml = row_len[s];
        const int m = ml & MDG_ROW_LEN_MASK;
        if (ml & MDG_ROW_PURE)
            mdg_row_stream<KIND, RETEST, WITH_DP, true, GROUP, WITH_E>(qs, row, m, lane_in_group, qi, bx, rc2, P, fx, fy, fz, en, dpa);
        else
            mdg_row_stream<KIND, RETEST, WITH_DP, false, GROUP, WITH_E>(qs, row, m, lane_in_group, qi, bx, rc2, P, fx, fy, fz, en, dpa);
        fx *= P.sg; fy *= P.sg; fz *= P.sg;
        en *= 0.5f * P.se;
        if (WITH_DP) {
#pragma unroll
            for (int q = 0; q < MDG_MAX_POT_PARAMS; ++q) dpa[q] *= 0.5f * P.sdp[q];
        }
    }
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, o);
        fy += __shfl_xor_sync(0xffffffffu, fy, o);
        fz += __shfl_xor_sync(0xffffffffu, fz, o);
        if (WITH_E) en += __shfl_xor_sync(0xffffffffu, en, o);
    }
    if (s < n && lane_in_group == 0) fs[s] = make_float4(fx, fy, fz, en);
    if (WITH_DP) {
        __shared__ double sm[8][MDG_MAX_POT_PARAMS];
        double v[MDG_MAX_POT_PARAMS];
#pragma unroll
        for (int q = 0; q < MDG_MAX_POT_PARAMS; ++q) {
            v[q] = (double)dpa[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        }
        int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0)
            for (int q = 0; q < MDG_MAX_POT_PARAMS; ++q) sm[w][q] = v[q];
        __syncthreads();
        if (threadIdx.x < MDG_MAX_POT_PARAMS) {
            double t = 0;
            for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) t += sm[ww][threadIdx.x];
            dp_partials[(size_t)blockIdx.x * MDG_MAX_POT_PARAMS + threadIdx.x] = t;
        }
    }
}


template <bool RETEST, bool WITH_DP, int GROUP, bool WITH_E>
static int launch_force_g(mdg_ctx* c, const PotParams& P, const float4* qs, float4* fs, double* dpp, cudaStream_t st) {
    const int T = 256;
    int s0 = c->force_s0 >= 0 ? c->force_s0 : c->own_s0, n = c->force_s0 >= 0 ? c->force_s1 : c->own_s1;
    const int gap_at = (c->force_s0 >= 0 && c->force_gap > 0) ? c->force_gap_at : 0x7fffffff, gap = c->force_gap;
    const int nrows = (n - s0) - ((c->force_s0 >= 0 && c->force_gap > 0) ? gap : 0);
    int nb = (int)(((int64_t)nrows * GROUP + T - 1) / T);
    if (nb <= 0) return MDG_OK;
    // rows are allocated for the own range only: address them by the global sorted index
    const uint32_t* rows_base = c->rows.as<uint32_t>() - (size_t)c->rows_s0 * c->cap;
#define LF(K)                                                                                                  \
    k_force_rows<K, RETEST, WITH_DP, GROUP, WITH_E><<<nb, T, 0, st>>>(s0, n, gap_at, gap, qs, rows_base, c->row_len.as<int>(),      \
                                                              c->cap, c->box, c->rc2, P, fs, dpp)
    switch (P.kind) {
        case MDG_POT_LJ: LF(MDG_POT_LJ); break;
        case MDG_POT_LJFAM: LF(MDG_POT_LJFAM); break;
        case MDG_POT_LJ69: LF(MDG_POT_LJ69); break;
        case MDG_POT_EXV: LF(MDG_POT_EXV); break;
        case MDG_POT_BUCK: LF(MDG_POT_BUCK); break;
        case MDG_POT_MORSE: LF(MDG_POT_MORSE); break;
        default: mdg_set_error("unknown potential kind %d", P.kind); return MDG_E_BADARG;
    }
#undef LF
    c->stat_launches++;
    MDG_KERNEL_CHECK();
    return MDG_OK;
}

// lanes per row: 4 (default; measured 60.0 us vs 65.8 us with 8 on the 256k-atom box) - MDG_FORCE_GROUP=2|4|8
template <bool RETEST, bool WITH_DP>
static int launch_force(mdg_ctx* c, const PotParams& P, const float4* qs, float4* fs, double* dpp, cudaStream_t st) {
    if (c->force_group == 8) return launch_force_g<RETEST, WITH_DP, 8, true>(c, P, qs, fs, dpp, st);
    if (c->force_group == 2 && RETEST && !WITH_DP) return launch_force_g<RETEST, WITH_DP, 2, true>(c, P, qs, fs, dpp, st);
    // engine steps whose energy nobody reads (all but the last of an epoch): force-only specialisation
    if (RETEST && !WITH_DP && !c->force_energy) return launch_force_g<RETEST, WITH_DP, 4, false>(c, P, qs, fs, dpp, st);
    return launch_force_g<RETEST, WITH_DP, 4, true>(c, P, qs, fs, dpp, st);
}

int mdg_i_force_blocks(mdg_ctx* c) { return (int)(((int64_t)(c->own_s1 - c->own_s0) * c->force_group + 255) / 256); }

// explicit sub-range of the own rows (multi-GPU: interior layers first, boundary layers after the halo arrived)
int mdg_i_force_range(mdg_ctx* c, const PotParams& P, const float4* d_qs, float4* d_fs, bool retest, int s0, int s1,
                      int c0, int c1, cudaStream_t st) {
    if (s1 <= s0) return MDG_OK;
    c->force_s0 = s0;
    c->force_s1 = s1;
    c->force_c0 = c0;
    c->force_c1 = c1;
    int r = mdg_i_force_sorted(c, P, d_qs, d_fs, retest, false, nullptr, st);
    c->force_s0 = -1;
    return r;
}

// two row ranges [s0a, s1a) and [s0b, s1b) (s1a <= s0b) in ONE launch: the bottom and the top boundary layer of a slab
int mdg_i_force_range2(mdg_ctx* c, const PotParams& P, const float4* d_qs, float4* d_fs, bool retest, int s0a, int s1a, int s0b,
                       int s1b, cudaStream_t st) {
    if (c->tiles || s1a > s0b) { mdg_set_error("mdg_i_force_range2: row-list ranges in ascending order only"); return MDG_E_STATE; }
    if (s1a <= s0a && s1b <= s0b) return MDG_OK;
    c->force_s0 = s0a;
    c->force_s1 = s1b;
    c->force_gap_at = s1a;
    c->force_gap = s0b - s1a;
    int r = mdg_i_force_sorted(c, P, d_qs, d_fs, retest, false, nullptr, st);
    c->force_s0 = -1;
    c->force_gap = 0;
    return r;
}

int mdg_i_force_sorted(mdg_ctx* c, const PotParams& P, const float4* d_qs, float4* d_fs, bool retest, bool with_dp,
                       double* d_dp_partials, cudaStream_t st) {
    if (c->n == 0) return MDG_OK;
    if (c->tiles) {       // engine skin list in tile form (tiles.cuh): always re-tested, no parameter gradients
        if (!retest || with_dp) { mdg_set_error("tile list: only the engine's re-tested force evaluation is available"); return MDG_E_STATE; }
        const int c0 = c->force_s0 >= 0 ? c->force_c0 : c->own_c0, c1 = c->force_s0 >= 0 ? c->force_c1 : c->own_c1;
        if (c->force_energy) return launch_force_tiles<true>(c, P, d_qs, d_fs, c0, c1, st);
        return launch_force_tiles<false>(c, P, d_qs, d_fs, c0, c1, st);
    }
    if (retest) {
        if (with_dp) return launch_force<true, true>(c, P, d_qs, d_fs, d_dp_partials, st);
        return launch_force<true, false>(c, P, d_qs, d_fs, d_dp_partials, st);
    }
    if (with_dp) return launch_force<false, true>(c, P, d_qs, d_fs, d_dp_partials, st);
    return launch_force<false, false>(c, P, d_qs, d_fs, d_dp_partials, st);
}

// ---------------------------------------------------------------------------------------------
// op-level plumbing: gather current positions into sorted order, scatter forces back, reduce E
// ---------------------------------------------------------------------------------------------
__global__ void k_gather_sorted(int n, const float* __restrict__ xyz, const int* __restrict__ perm,
                                float4* __restrict__ qs) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int i = perm[s];
    qs[s] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], __int_as_float(i));
}

__global__ void k_scatter_force(int n, const float4* __restrict__ fs, const int* __restrict__ perm,
                                float* __restrict__ force) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int i = perm[s];
    float4 f = fs[s];
    force[3 * i] = f.x;
    force[3 * i + 1] = f.y;
    force[3 * i + 2] = f.z;
}

__global__ void __launch_bounds__(256) k_energy_partials(int n, const float4* __restrict__ fs, double* __restrict__ part) {
    __shared__ double sm[8];
    double v = 0;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) v += (double)fs[s].w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += sm[w];
        part[blockIdx.x] = t;
    }
}

// single block: energy = sum(part[0..npart)), dparams[q] = sum over blocks of dp_part[b*4+q]
__global__ void __launch_bounds__(256) k_finalize(int npart, const double* __restrict__ part, int ndp_blocks,
                                                  const double* __restrict__ dp_part, float* __restrict__ energy,
                                                  float* __restrict__ dparams) {
    __shared__ double sm[8];
    for (int what = 0; what < 1 + MDG_MAX_POT_PARAMS; ++what) {
        if (what == 0 && !energy) continue;
        if (what > 0 && !dparams) continue;
        double v = 0;
        if (what == 0)
            for (int i = threadIdx.x; i < npart; i += blockDim.x) v += part[i];
        else
            for (int i = threadIdx.x; i < ndp_blocks; i += blockDim.x) v += dp_part[(size_t)i * MDG_MAX_POT_PARAMS + what - 1];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0;
            for (int w = 0; w < 8; ++w) t += sm[w];
            if (what == 0) *energy = (float)t; else dparams[what - 1] = (float)t;
        }
    }
}

int mdg_i_pair_force_op(mdg_ctx* c, const PotParams& P, const float* d_xyz, int n, float* d_energy, float* d_force,
                        float* d_dparams, cudaStream_t st) {
    if (!c->built || n != c->n) { mdg_set_error("mdg_pair_force: no list built for n=%d", n); return MDG_E_STATE; }
    if (n == 0) {
        if (d_energy) MDG_CUDA(cudaMemsetAsync(d_energy, 0, sizeof(float), st));
        if (d_dparams) MDG_CUDA(cudaMemsetAsync(d_dparams, 0, sizeof(float) * MDG_MAX_POT_PARAMS, st));
        return MDG_OK;
    }
    const int T = 256;
    int nb = (n + T - 1) / T;
    MDG_TRY(c->fs.reserve(sizeof(float4) * (size_t)n));
    int fblocks = mdg_i_force_blocks(c);
    const int EPART = 296;
    MDG_TRY(c->partials.reserve(sizeof(double) * ((size_t)EPART + (size_t)fblocks * MDG_MAX_POT_PARAMS)));
    double* epart = c->partials.as<double>();
    double* dpp = epart + EPART;
    float4* qs = c->qs_ptr;
    k_gather_sorted<<<nb, T, 0, st>>>(n, d_xyz, c->perm.as<int>(), qs);
    c->stat_launches++;
    MDG_TRY(mdg_i_force_sorted(c, P, qs, c->fs.as<float4>(), false, d_dparams != nullptr, dpp, st));
    if (d_force) { k_scatter_force<<<nb, T, 0, st>>>(n, c->fs.as<float4>(), c->perm.as<int>(), d_force); c->stat_launches++; }
    if (d_energy || d_dparams) {
        int npart = nb < EPART ? nb : EPART;
        if (d_energy) { k_energy_partials<<<npart, T, 0, st>>>(n, c->fs.as<float4>(), epart); c->stat_launches++; }
        k_finalize<<<1, 256, 0, st>>>(npart, epart, fblocks, dpp, d_energy, d_dparams);
        c->stat_launches++;
    }
    MDG_KERNEL_CHECK();
    return MDG_OK;
}

// ---------------------------------------------------------------------------------------------
// generic listed-pair distance op (compute_dis, reference torchmd/topology.py:5-12) fwd/bwd
// ---------------------------------------------------------------------------------------------
// dP != nullptr: the pair count lives on the device (asynchronous engine steps); P is then only the launch bound
__global__ void k_pair_dis_fwd(const float* __restrict__ xyz, const int64_t* __restrict__ nbr,
                               const float* __restrict__ off, int64_t P, const int* __restrict__ dP, Box bx,
                               float* __restrict__ dis) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (dP) P = min(P, (int64_t)*dP);
    if (p >= P) return;
    int64_t i = nbr[2 * p], j = nbr[2 * p + 1];
    float dx = (xyz[3 * i] - xyz[3 * j]) - off[3 * p] * bx.L[0];
    float dy = (xyz[3 * i + 1] - xyz[3 * j + 1]) - off[3 * p + 1] * bx.L[1];
    float dz = (xyz[3 * i + 2] - xyz[3 * j + 2]) - off[3 * p + 2] * bx.L[2];
    dis[p] = sqrtf(dx * dx + dy * dy + dz * dz);
}

__global__ void k_pair_dis_bwd(const float* __restrict__ xyz, const int64_t* __restrict__ nbr,
                               const float* __restrict__ off, int64_t P, Box bx, const float* __restrict__ dis,
                               const float* __restrict__ gdis, float* __restrict__ gxyz) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    int64_t i = nbr[2 * p], j = nbr[2 * p + 1];
    float dx = (xyz[3 * i] - xyz[3 * j]) - off[3 * p] * bx.L[0];
    float dy = (xyz[3 * i + 1] - xyz[3 * j + 1]) - off[3 * p + 1] * bx.L[1];
    float dz = (xyz[3 * i + 2] - xyz[3 * j + 2]) - off[3 * p + 2] * bx.L[2];
    float r = dis ? dis[p] : sqrtf(dx * dx + dy * dy + dz * dz);
    float w = r > 0.f ? gdis[p] / r : 0.f;
    atomicAdd(&gxyz[3 * i], w * dx);
    atomicAdd(&gxyz[3 * i + 1], w * dy);
    atomicAdd(&gxyz[3 * i + 2], w * dz);
    atomicAdd(&gxyz[3 * j], -w * dx);
    atomicAdd(&gxyz[3 * j + 1], -w * dy);
    atomicAdd(&gxyz[3 * j + 2], -w * dz);
}

static Box make_box(const float* h_cell3) {
    Box b;
    for (int k = 0; k < 3; ++k) { b.L[k] = h_cell3[k]; b.invL[k] = 1.0f / h_cell3[k]; }
    return b;
}

int mdg_i_pair_dis_fwd(const float* d_xyz, const int64_t* d_nbr, const float* d_offsets, int64_t n_pairs, const int* d_n_pairs,
                       const float* h_cell3, float* d_dis, cudaStream_t st) {
    if (n_pairs <= 0) return MDG_OK;
    k_pair_dis_fwd<<<(unsigned)((n_pairs + 255) / 256), 256, 0, st>>>(d_xyz, d_nbr, d_offsets, n_pairs, d_n_pairs, make_box(h_cell3),
                                                                     d_dis);
    MDG_KERNEL_CHECK();
    return MDG_OK;
}

extern "C" int mdg_pair_dis_fwd(const float* d_xyz, int n, const int64_t* d_nbr, const float* d_offsets,
                                int64_t n_pairs, const float* h_cell3, float* d_dis, void* stream) {
    (void)n;
    return mdg_i_pair_dis_fwd(d_xyz, d_nbr, d_offsets, n_pairs, nullptr, h_cell3, d_dis, (cudaStream_t)stream);
}

extern "C" int mdg_pair_dis_bwd(const float* d_xyz, int n, const int64_t* d_nbr, const float* d_offsets,
                                int64_t n_pairs, const float* h_cell3, const float* d_dis, const float* d_grad_dis,
                                float* d_grad_xyz, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MDG_CUDA(cudaMemsetAsync(d_grad_xyz, 0, sizeof(float) * 3 * (size_t)n, st));
    if (n_pairs <= 0) return MDG_OK;
    k_pair_dis_bwd<<<(unsigned)((n_pairs + 255) / 256), 256, 0, st>>>(d_xyz, d_nbr, d_offsets, n_pairs, make_box(h_cell3),
                                                                     d_dis, d_grad_dis, d_grad_xyz);
    MDG_KERNEL_CHECK();
    return MDG_OK;
}
