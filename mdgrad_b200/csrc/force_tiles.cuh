// force_tiles.cuh - pair force over the engine's TILE list (included by force.cu; layout in tiles.cuh).
// Replaces, like k_force_rows, PairPotentials.forward (reference torchmd/interface.py:284-300: compute_dis
// topology.py:5-12 -> u(r).sum()) and the autograd force F = -dE/dxyz (torchmd/md.py:227-228) for the MD engine.
//
// One CTA per block of <= 4 cells.  Prologue: every thread reads the block header the builder left (first row, row count, home
// offset, staged atoms), lane 0 arms an mbarrier with the byte count and the lanes of warp 0 issue one TMA bulk copy
// (cp.async.bulk.shared.global) per contiguous piece of the stencil (<= 18, listed by the builder); while the copies fly every
// warp fetches the lengths and the first chunk of its rows; warp 0 polls the mbarrier, the others sleep in the block barrier.
// Main loop (software-pipelined: the next 8-byte chunk is requested before the current one is evaluated): warp = 8 rows x 4 lanes; per iteration a lane loads 8 bytes
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
__device__ __forceinline__ void mdg_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mdg_smem_u32(bar)) : "memory");
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

// Four unshifted entries at once, written stage by stage (differences and d2 of all four, then the reciprocals, then the
// potential) so that four independent dependency chains are in flight per warp: the kernel is latency-bound, not issue-bound
// (ncu: ~12 warps per scheduler, each a single chain of LDS -> 20 dependent FP instructions).
template <int KIND, bool WITH_E>
__device__ __forceinline__ void tile_pair4(const float4 qi, const unsigned char* s_raw, const uint2 e, const Box& bx, float rc2,
                                           const PotParams& P, float& fx, float& fy, float& fz, float& en) {
    if (KIND != MDG_POT_LJ) {
        const float4 q0 = *reinterpret_cast<const float4*>(s_raw + (e.x & 0xffffu));
        const float4 q1 = *reinterpret_cast<const float4*>(s_raw + (e.x >> 16));
        const float4 q2 = *reinterpret_cast<const float4*>(s_raw + (e.y & 0xffffu));
        const float4 q3 = *reinterpret_cast<const float4*>(s_raw + (e.y >> 16));
        tile_pair<KIND, WITH_E, false>(qi, q0, 0u, bx, rc2, P, fx, fy, fz, en);
        tile_pair<KIND, WITH_E, false>(qi, q1, 0u, bx, rc2, P, fx, fy, fz, en);
        tile_pair<KIND, WITH_E, false>(qi, q2, 0u, bx, rc2, P, fx, fy, fz, en);
        tile_pair<KIND, WITH_E, false>(qi, q3, 0u, bx, rc2, P, fx, fy, fz, en);
        return;
    }
    const uint32_t off[4] = {e.x & 0xffffu, e.x >> 16, e.y & 0xffffu, e.y >> 16};
    float dx[4], dy[4], dz[4], d2[4], r2i[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float4 qj = *reinterpret_cast<const float4*>(s_raw + off[u]);
        dx[u] = __fsub_rn(qj.x, qi.x);
        dy[u] = __fsub_rn(qj.y, qi.y);
        dz[u] = __fsub_rn(qj.z, qi.z);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) d2[u] = mdg_d2_exact(dx[u], dy[u], dz[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) r2i[u] = mdg_rcp(d2[u]);           // d2 == 0 gives inf: discarded by the predicate below
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const bool in = (d2[u] < rc2) && (d2[u] != 0.0f);
        const float s2 = P.aux * r2i[u];
        const float s6 = s2 * s2 * s2;
        const float g = s6 * (2.0f * s6 - 1.0f) * r2i[u];
        if (in) {
            fx -= g * dx[u];
            fy -= g * dy[u];
            fz -= g * dz[u];
            if (WITH_E) en += s6 * (s6 - 1.0f);
        }
    }
}

#ifndef MDG_TILE_MINBLOCKS
#define MDG_TILE_MINBLOCKS 4
#endif
#ifndef MDG_TILE_MAXTHREADS
#define MDG_TILE_MAXTHREADS 384
#endif
// warp 0: arm the buffer's mbarrier and issue the TMA bulk copies of block descriptor D into `dst`
__device__ __forceinline__ void tile_issue(const int* __restrict__ D, int total, const float4* __restrict__ qs, float4* dst,
                                           uint64_t* bar, int lane) {
#ifdef MDG_EMU
    (void)bar; (void)total;
    for (int k = 0; k < D[4]; ++k) {
        const int src = D[8 + 3 * k], off = D[9 + 3 * k], cnt = D[10 + 3 * k];
        for (int a = lane; a < cnt; a += 32) dst[off + a] = qs[src + a];
    }
#else
    if (lane == 0) mdg_mbar_expect_tx(bar, (uint32_t)total * 16u);
    __syncwarp();
    if (lane < D[4]) {
        const int src = D[8 + 3 * lane], off = D[9 + 3 * lane], cnt = D[10 + 3 * lane];
        mdg_bulk_g2s(dst + off, qs + src, (uint32_t)cnt * 16u, bar);
    }
#endif
}

// PERSISTENT + WARP-SPECIALISED: grid = min(blocks, resident CTAs); CTA c works on blocks c, c + gridDim.x, ...  The LAST warp of
// the CTA is the producer: for every block it waits until the consumers have released the ring slot (EMPTY mbarrier, one arrival
// per consumer warp), arms the slot's FULL mbarrier with the byte count and issues the TMA bulk copies of the block's stencil.
// The other warps are consumers: wait for FULL, evaluate their rows of the block, arrive on EMPTY.  Nobody waits for a sibling
// warp: a consumer that finishes early runs ahead into the next block (ring of MDG_TILE_NBUF slots), so neither the staging
// latency nor the ~30% spread between the warps' row lengths costs issue slots.
#ifndef MDG_TILE_NBUF
#define MDG_TILE_NBUF 2
#endif
template <int KIND, bool WITH_E>
__global__ void __launch_bounds__(MDG_TILE_MAXTHREADS, MDG_TILE_MINBLOCKS) k_force_tiles(int blk0, int nblocks, TileGeom G, const float4* __restrict__ qs,
                                                                      const int* __restrict__ bdesc,
                                                                      const uint16_t* __restrict__ trows,
                                                                      const uint32_t* __restrict__ tlen, Box bx, float rc2,
                                                                      PotParams P, float4* __restrict__ fs, int* __restrict__ flags) {
    extern __shared__ float4 s_q[];                     // MDG_TILE_NBUF slots of G.scap float4
    __shared__ uint64_t s_full[MDG_TILE_NBUF], s_empty[MDG_TILE_NBUF];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncons = (blockDim.x >> 5) - 1;            // consumer warps; warp `ncons` is the producer
    const int rr = lane >> 2, q = lane & 3;
#ifndef MDG_EMU
    if (threadIdx.x == 0) {
        for (int k = 0; k < MDG_TILE_NBUF; ++k) { mdg_mbar_init(&s_full[k], 1); mdg_mbar_init(&s_empty[k], ncons); }
    }
#endif
    __syncthreads();
#ifndef MDG_EMU
    if (warp == ncons) {
        // ------------------------------------------------ producer -------------------------------------------------------
        int it = 0;
        for (int bl = blockIdx.x; bl < nblocks; bl += gridDim.x, ++it) {
            const int* D = bdesc + (size_t)(blk0 + bl - G.b_base) * MDG_TILE_DESC;
            const int4 hd = __ldg(reinterpret_cast<const int4*>(D));
            const int slot = it % MDG_TILE_NBUF, use = it / MDG_TILE_NBUF;
            if (use > 0) mdg_mbar_wait(&s_empty[slot], (uint32_t)(use - 1) & 1u);     // consumers are done with block it - NBUF
            if (hd.y > 0 && hd.w <= G.scap) tile_issue(D, hd.w, qs, s_q + (size_t)slot * G.scap, &s_full[slot], lane);
            else if (lane == 0) mdg_mbar_expect_tx(&s_full[slot], 0u);                // nothing to stage: complete the phase
        }
        return;
    }
#endif
    // ---------------------------------------------------- consumers ------------------------------------------------------
    int it = 0;
    for (int bl = blockIdx.x; bl < nblocks; bl += gridDim.x, ++it) {
        const int b = blk0 + bl, slot = it % MDG_TILE_NBUF;
        const int4 hd = __ldg(reinterpret_cast<const int4*>(bdesc + (size_t)(b - G.b_base) * MDG_TILE_DESC));
        const int a0 = hd.x, na = hd.y, hoff = hd.z, total = hd.w;
        const bool live = na > 0 && total <= G.scap;
        if (na > 0 && total > G.scap && threadIdx.x == 0) { flags[0] = 1; atomicMax(&flags[1], total); }   // (builder flagged it too)
        const float4* buf = s_q + (size_t)slot * G.scap;
        // ---- rows of this warp: lengths and the first two chunks (independent of the staged positions) -------------------
        const int g0 = tile_group0(G, b, a0);
        int lg = warp + it;                              // rotate the assignment: the warp that gets the extra group changes per block
        lg -= (lg / ncons) * ncons;
#ifdef MDG_EMU
        if (warp >= ncons) lg = 1 << 24;                 // (the emulation has no producer role: the last warp only joins the barriers)
#endif
        int r = lg * MDG_TILE_GROUP + rr;
        uint32_t len = (live && r < na) ? tlen[a0 + r] : 0u;
        const uint16_t* p = trows + ((size_t)(g0 + lg) * G.capc) * MDG_TILE_GCHUNK + rr * MDG_TILE_CHUNK + q * 4;
        int nch = (int)(len & 255u) + (int)((len >> 8) & 255u);
        uint2 e0 = make_uint2(0u, 0u), e1 = make_uint2(0u, 0u);
        if (nch > 0) e0 = __ldcs(reinterpret_cast<const uint2*>(p));
        if (nch > 1) e1 = __ldcs(reinterpret_cast<const uint2*>(p + MDG_TILE_GCHUNK));
#ifdef MDG_EMU
        __syncthreads();
        if (live && warp == 0) tile_issue(bdesc + (size_t)(b - G.b_base) * MDG_TILE_DESC, total, qs, s_q + (size_t)slot * G.scap, &s_full[slot], lane);
        __syncthreads();
#else
        mdg_mbar_wait(&s_full[slot], (uint32_t)(it / MDG_TILE_NBUF) & 1u);
#endif
        const unsigned char* s_raw = reinterpret_cast<const unsigned char*>(buf);
        if (live) {
            for (; lg * MDG_TILE_GROUP < na; lg += ncons) {
                float fx = 0.f, fy = 0.f, fz = 0.f, en = 0.f;
                int nA = (int)(len & 255u), nB = (int)((len >> 8) & 255u);
                if (nA + nB > 0) {
                    const float4 qi = buf[hoff + r];
                    p += 2 * MDG_TILE_GCHUNK;                                                // chunk after the two in flight
                    for (; nA > 0; --nA, p += MDG_TILE_GCHUNK) {
                        uint2 nx = e1;
                        if (nA + nB > 2) nx = __ldcs(reinterpret_cast<const uint2*>(p));     // two chunks ahead of the one evaluated now
                        tile_pair4<KIND, WITH_E>(qi, s_raw, e0, bx, rc2, P, fx, fy, fz, en);
                        e0 = e1;
                        e1 = nx;
                    }
                    for (; nB > 0; --nB, p += MDG_TILE_GCHUNK) {                             // entries with an image shift: (offset, code) pairs
                        uint2 nx = e1;
                        if (nB > 2) nx = __ldcs(reinterpret_cast<const uint2*>(p));
                        const float4 q0 = *reinterpret_cast<const float4*>(s_raw + (e0.x & 0xffffu));
                        const float4 q1 = *reinterpret_cast<const float4*>(s_raw + (e0.y & 0xffffu));
                        tile_pair<KIND, WITH_E, true>(qi, q0, e0.x >> 16, bx, rc2, P, fx, fy, fz, en);
                        tile_pair<KIND, WITH_E, true>(qi, q1, e0.y >> 16, bx, rc2, P, fx, fy, fz, en);
                        e0 = e1;
                        e1 = nx;
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
                // next group of this warp (blocks with more than 8 * ncons rows)
                const int lgn = lg + ncons;
                if (lgn * MDG_TILE_GROUP < na) {
                    r = lgn * MDG_TILE_GROUP + rr;
                    len = (r < na) ? tlen[a0 + r] : 0u;
                    p = trows + ((size_t)(g0 + lgn) * G.capc) * MDG_TILE_GCHUNK + rr * MDG_TILE_CHUNK + q * 4;
                    nch = (int)(len & 255u) + (int)((len >> 8) & 255u);
                    e0 = make_uint2(0u, 0u); e1 = make_uint2(0u, 0u);
                    if (nch > 0) e0 = __ldcs(reinterpret_cast<const uint2*>(p));
                    if (nch > 1) e1 = __ldcs(reinterpret_cast<const uint2*>(p + MDG_TILE_GCHUNK));
                }
            }
        }
#ifndef MDG_EMU
        __syncwarp();
        if (lane == 0) mdg_mbar_arrive(&s_empty[slot]);                                      // this warp is done with the slot
#endif
    }
}

template <int K, bool WITH_E>
static int launch_force_tiles_k(mdg_ctx* c, const PotParams& P, const float4* qs, float4* fs, int blk0, int nb, int per_sm, int T,
                                size_t smem, cudaStream_t st) {
    static bool attr_set = false;       // (same spelling of the kernel in both statements: the CPU emulation keys on the text)
    if (!attr_set) {
        MDG_CUDA(cudaFuncSetAttribute(k_force_tiles<K, WITH_E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
#ifndef MDG_EMU
    if (c->tile_ctas_env <= 0) {        // persistent grid = what is resident at once (registers, threads, shared memory)
        int occ = 0;
        MDG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_force_tiles<K, WITH_E>, T, smem));
        if (occ >= 1) per_sm = occ;
    }
#endif
    int grid = (c->sm_count > 0 ? c->sm_count : 1) * per_sm;
    if (grid > nb) grid = nb;
    k_force_tiles<K, WITH_E><<<grid, T, smem, st>>>(blk0, nb, c->tile, qs, c->tile_desc.as<int>(), c->tile_rows.as<uint16_t>(),
                                                    c->tile_len.as<uint32_t>(), c->box, c->rc2, P, fs, c->flags.as<int>());
    return MDG_OK;
}

template <bool WITH_E>
static int launch_force_tiles(mdg_ctx* c, const PotParams& P, const float4* qs, float4* fs, int c0, int c1, cudaStream_t st) {
    const TileGeom& G = c->tile;
    const int blk0 = (c0 / G.ncx) * G.nblk, nb = ((c1 - c0) / G.ncx) * G.nblk;     // (cell ranges are whole x-rows)
    if (nb <= 0) return MDG_OK;
    const size_t smem = MDG_TILE_NBUF * sizeof(float4) * (size_t)G.scap;
    const int T = 32 * (c->tile_warps + 1);         // consumer warps + the producer warp
    int per_sm = 2048 / T;
    const int by_smem = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > by_smem) per_sm = by_smem;
    if (per_sm < 1) per_sm = 1;
    if (c->tile_ctas_env > 0) per_sm = c->tile_ctas_env;
#define LT(K) MDG_TRY((launch_force_tiles_k<K, WITH_E>(c, P, qs, fs, blk0, nb, per_sm, T, smem, st)))
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
