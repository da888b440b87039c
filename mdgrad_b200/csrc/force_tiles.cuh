// force_tiles.cuh - pair force over the engine's TILE list (included by force.cu; layout in tiles.cuh).
// Replaces, like k_force_rows, PairPotentials.forward (reference torchmd/interface.py:284-300: compute_dis
// topology.py:5-12 -> u(r).sum()) and the autograd force F = -dE/dxyz (torchmd/md.py:227-228) for the MD engine.
//
// One CTA per block of <= 4 cells.  Prologue: warp 0 scans the cell counts of the block's 9 x (w + 2) stencil cells, lane 0 arms
// an mbarrier with the byte count and the lanes issue one TMA bulk copy (cp.async.bulk.shared.global) per contiguous piece of
// the stencil (<= 18); everybody waits on the mbarrier.  Main loop: warp = 8 rows x 4 lanes; per iteration a lane loads 8 bytes
// of the interleaved row stream (4 local byte offsets), reads the 4 neighbor float4 with LDS.128, and evaluates the pairs with the
// reference's exact membership arithmetic; warp-shuffle reduction over the 4 lanes, one float4 store per atom.
#pragma once
#include "tiles.cuh"

#ifndef MDG_EMU
__device__ __forceinline__ uint32_t mdg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mdg_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mdg_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mdg_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mdg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mdg_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mdg_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(mdg_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mdg_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(mdg_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
#endif

// Stage the block's stencil stream into s_q.  s_cs / s_cn / s_off as filled by tile_stencil_prefix (synchronised by the caller).
__device__ __forceinline__ void tile_stage(const TileGeom& G, int bx0, int w, const float4* __restrict__ qs, const int* s_cs,
                                           const int* s_off, float4* s_q, uint64_t* s_bar) {
    const int kw = w + 2, nst = 9 * kw;
#ifdef MDG_EMU
    (void)s_bar;
    (void)bx0;
    for (int t = 0; t < nst; ++t) {
        const int cnt = s_off[t + 1] - s_off[t];
        for (int a = threadIdx.x; a < cnt; a += blockDim.x) s_q[s_off[t] + a] = qs[s_cs[t] + a];
    }
    __syncthreads();
#else
    if ((threadIdx.x >> 5) == 0) {
        const int lane = threadIdx.x & 31;
        if (lane == 0) mdg_mbar_expect_tx(s_bar, (uint32_t)s_off[nst] * 16u);
        __syncwarp();
        for (int t = lane; t < nst; t += 32) {
            const int k = t - tile_div_kw(t, kw) * kw;
            int x = bx0 - 1 + k;
            x = x < 0 ? x + G.ncx : (x >= G.ncx ? x - G.ncx : x);
            if (k == 0 || x == 0) {            // first cell of a piece that is contiguous in the sorted array
                int e = t + 1;
                for (int ke = k + 1; ke < kw; ++ke, ++e) {
                    int xe = bx0 - 1 + ke;
                    xe = xe >= G.ncx ? xe - G.ncx : xe;
                    if (xe == 0) break;
                }
                const int cnt = s_off[e] - s_off[t];
                if (cnt > 0) mdg_bulk_g2s(s_q + s_off[t], qs + s_cs[t], (uint32_t)cnt * 16u, s_bar);
            }
        }
    }
    mdg_mbar_wait(s_bar, 0);
#endif
}

template <int KIND, bool WITH_E, bool SHIFT>
__device__ __forceinline__ void tile_pair(const float4 qi, const float4 qj, uint32_t code, const Box& bx, float rc2,
                                          const PotParams& P, float& fx, float& fy, float& fz, float& en) {
    float dx = __fsub_rn(qj.x, qi.x), dy = __fsub_rn(qj.y, qi.y), dz = __fsub_rn(qj.z, qi.z);
    if (SHIFT) {   // off * L = (code - 1) * L as fma(code, L, -L): exact for code in {0, 1, 2} (see k_force_rows)
        dx = __fadd_rn(dx, __fmaf_rn((float)(code & 3u), bx.L[0], -bx.L[0]));
        dy = __fadd_rn(dy, __fmaf_rn((float)((code >> 2) & 3u), bx.L[1], -bx.L[1]));
        dz = __fadd_rn(dz, __fmaf_rn((float)((code >> 4) & 3u), bx.L[2], -bx.L[2]));
    }
    const float d2 = mdg_d2_exact(dx, dy, dz);
    const bool in = (d2 < rc2) && (d2 != 0.0f);
    float e_p, g, dp[MDG_MAX_POT_PARAMS];
    if (KIND == MDG_POT_LJ) {
        pair_eval<KIND, false>(P, d2, e_p, g, dp);      // if-converted: d2 == 0 gives inf/nan, discarded by the predicate
        if (in) {
            fx -= g * dx;
            fy -= g * dy;
            fz -= g * dz;
            if (WITH_E) en += e_p;
        }
    } else if (in) {
        pair_eval<KIND, false>(P, d2, e_p, g, dp);
        fx -= g * dx;
        fy -= g * dy;
        fz -= g * dz;
        if (WITH_E) en += e_p;
    }
}

#ifndef MDG_TILE_MINBLOCKS
#define MDG_TILE_MINBLOCKS 4
#endif
template <int KIND, bool WITH_E>
__global__ void __launch_bounds__(512, MDG_TILE_MINBLOCKS) k_force_tiles(int z0, TileGeom G, const float4* __restrict__ qs,
                                                                      const int* __restrict__ cell_start,
                                                                      const uint16_t* __restrict__ trows,
                                                                      const uint32_t* __restrict__ tlen, Box bx, float rc2,
                                                                      PotParams P, float4* __restrict__ fs, int* __restrict__ flags) {
    extern __shared__ float4 s_q[];
    __shared__ int s_cs[MDG_TILE_MAXST], s_cn[MDG_TILE_MAXST], s_off[MDG_TILE_MAXST + 1];
    __shared__ uint64_t s_bar;
    const int bi = blockIdx.x, cy = blockIdx.y, cz = z0 + blockIdx.z;       // grid = (blocks per x-row, ncy, z-layers)
    const int b = (cz * G.ncy + cy) * G.nblk + bi;
    const int bx0 = tile_bx0(G, bi), w = tile_bx0(G, bi + 1) - bx0, kw = w + 2, nst = 9 * kw;
    tile_stencil_prefix(G, bx0, w, cy, cz, cell_start, s_cs, s_cn, s_off);
#ifndef MDG_EMU
    if (threadIdx.x == 0) mdg_mbar_init(&s_bar, 1);
#endif
    __syncthreads();
    const int hcell = 4 * kw + 1;                       // first home cell in the stencil order (r = 4: dy = dz = 0)
    const int a0 = s_cs[hcell], hoff = s_off[hcell], na = s_off[hcell + w] - hoff;
    if (na == 0) return;
    if (s_off[nst] > G.scap) {                          // the builder flagged this already (rows of such a block are empty)
        if (threadIdx.x == 0) { flags[0] = 1; atomicMax(&flags[1], s_off[nst]); }
        return;
    }
    tile_stage(G, bx0, w, qs, s_cs, s_off, s_q, &s_bar);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int rr = lane >> 2, q = lane & 3;
    const int g0 = tile_group0(G, b, a0);
    const unsigned char* s_raw = reinterpret_cast<const unsigned char*>(s_q);
    for (int lg = warp; lg * MDG_TILE_GROUP < na; lg += nwarps) {
        const int r = lg * MDG_TILE_GROUP + rr;
        float fx = 0.f, fy = 0.f, fz = 0.f, en = 0.f;
        if (r < na) {
            const float4 qi = s_q[hoff + r];
            const uint32_t len = tlen[a0 + r];
            int nA = (int)(len & 255u), nB = (int)((len >> 8) & 255u);
            const uint16_t* p = trows + ((size_t)(g0 + lg) * G.capc) * MDG_TILE_GCHUNK + rr * MDG_TILE_CHUNK + q * 4;
            for (; nA > 0; --nA, p += MDG_TILE_GCHUNK) {
                const uint2 e = __ldcs(reinterpret_cast<const uint2*>(p));      // read once: evict-first
                const float4 q0 = *reinterpret_cast<const float4*>(s_raw + (e.x & 0xffffu));
                const float4 q1 = *reinterpret_cast<const float4*>(s_raw + (e.x >> 16));
                const float4 q2 = *reinterpret_cast<const float4*>(s_raw + (e.y & 0xffffu));
                const float4 q3 = *reinterpret_cast<const float4*>(s_raw + (e.y >> 16));
                tile_pair<KIND, WITH_E, false>(qi, q0, 0u, bx, rc2, P, fx, fy, fz, en);
                tile_pair<KIND, WITH_E, false>(qi, q1, 0u, bx, rc2, P, fx, fy, fz, en);
                tile_pair<KIND, WITH_E, false>(qi, q2, 0u, bx, rc2, P, fx, fy, fz, en);
                tile_pair<KIND, WITH_E, false>(qi, q3, 0u, bx, rc2, P, fx, fy, fz, en);
            }
            for (; nB > 0; --nB, p += MDG_TILE_GCHUNK) {                        // entries with an image shift: (offset, code) pairs
                const uint2 e = __ldcs(reinterpret_cast<const uint2*>(p));
                const float4 q0 = *reinterpret_cast<const float4*>(s_raw + (e.x & 0xffffu));
                const float4 q1 = *reinterpret_cast<const float4*>(s_raw + (e.y & 0xffffu));
                tile_pair<KIND, WITH_E, true>(qi, q0, e.x >> 16, bx, rc2, P, fx, fy, fz, en);
                tile_pair<KIND, WITH_E, true>(qi, q1, e.y >> 16, bx, rc2, P, fx, fy, fz, en);
            }
            fx *= P.sg; fy *= P.sg; fz *= P.sg;
            en *= 0.5f * P.se;
        }
#pragma unroll
        for (int o = MDG_TILE_LANES / 2; o > 0; o >>= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o);
            fy += __shfl_xor_sync(0xffffffffu, fy, o);
            fz += __shfl_xor_sync(0xffffffffu, fz, o);
            if (WITH_E) en += __shfl_xor_sync(0xffffffffu, en, o);
        }
        if (r < na && q == 0) fs[a0 + r] = make_float4(fx, fy, fz, en);
    }
}

template <bool WITH_E>
static int launch_force_tiles(mdg_ctx* c, const PotParams& P, const float4* qs, float4* fs, int c0, int c1, cudaStream_t st) {
    TileGeom G = c->tile;
    const int nxy = G.ncx * G.ncy, z0 = c0 / nxy, nz = (c1 - c0) / nxy;     // (cell ranges are whole z-layers)
    if (nz <= 0) return MDG_OK;
    const dim3 grid(G.nblk, G.ncy, nz);
    const size_t smem = sizeof(float4) * (size_t)G.scap;
    const int T = 32 * c->tile_warps;
#define LT(K)                                                                                                              \
    do {                                                                                                                   \
        static bool attr_set = false;                                                                                      \
        if (!attr_set) {                                                                                                   \
            MDG_CUDA(cudaFuncSetAttribute(k_force_tiles<K, WITH_E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));  \
            attr_set = true;                                                                                               \
        }                                                                                                                  \
        k_force_tiles<K, WITH_E><<<grid, T, smem, st>>>(z0, G, qs, c->cell_start.as<int>(), c->tile_rows.as<uint16_t>(),     \
                                                      c->tile_len.as<uint32_t>(), c->box, c->rc2, P, fs, c->flags.as<int>()); \
    } while (0)
    switch (P.kind) {
        case MDG_POT_LJ: LT(MDG_POT_LJ); break;
        case MDG_POT_LJFAM: LT(MDG_POT_LJFAM); break;
        case MDG_POT_LJ69: LT(MDG_POT_LJ69); break;
        case MDG_POT_EXV: LT(MDG_POT_EXV); break;
        case MDG_POT_BUCK: LT(MDG_POT_BUCK); break;
        case MDG_POT_MORSE: LT(MDG_POT_MORSE); break;
        default: mdg_set_error("unknown potential kind %d", P.kind); return MDG_E_BADARG;
    }
#undef LT
    c->stat_launches++;
    MDG_KERNEL_CHECK();
    return MDG_OK;
}
