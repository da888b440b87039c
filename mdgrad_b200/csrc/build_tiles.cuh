// build_tiles.cuh - builder of the engine's TILE list (included by nbr.cu after build_fast.cuh; layout in tiles.cuh).
//
// Same two-phase scheme as k_build_fast (phase 1: lane = candidate, FMA screening at the list radius + ballots into a
// shared-memory bitmask; phase 2: lane = atom, walk the set bits), with three differences:
//   * one CTA per BLOCK of <= 4 cells of an x-row, one warp per cell; the block's stencil prefix is computed once and gives
//     every candidate its 16-bit LOCAL index in the stream k_force_tiles stages (tiles.cuh);
//   * entries are written into the interleaved group / chunk layout, unshifted entries first (16 bits each: local index << 4),
//     entries with a periodic image shift in a second walk behind them (32 bits each: local index << 4, image code);
//   * every segment is padded to whole chunks with the row's own slot (d2 == 0 is dropped by the reference's `d2 != 0` test).
#pragma once
#include "tiles.cuh"

__global__ void __launch_bounds__(MDG_TILE_MAXW * 32) k_build_tiles(int z0, TileGeom G, const float4* __restrict__ qs,
                                                                    const int* __restrict__ cell_start, Box bx, float r2list,
                                                                    PairFilter F, uint16_t* __restrict__ trows,
                                                                    uint32_t* __restrict__ tlen, int* __restrict__ bdesc, int* __restrict__ flags) {
    __shared__ uint32_t s_mask[MDG_TILE_MAXW][32][FB_CHUNKS + 1];   // [atom][chunk], padded: conflict-free for lane = atom
    __shared__ uint32_t s_img[MDG_TILE_MAXW][FB_BATCH];
    __shared__ uint16_t s_loc[MDG_TILE_MAXW][FB_BATCH];             // local (stream) index << 4 of each staged candidate
    __shared__ float4 s_ctr[MDG_TILE_MAXW][32];                     // local coords of the cell's atoms
    __shared__ int s_pre[MDG_TILE_MAXW][28];                        // candidate-index prefix over the cell's 27 stencil cells
    __shared__ int s_cs[MDG_TILE_MAXST], s_cn[MDG_TILE_MAXST], s_off[MDG_TILE_MAXST + 1];
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bi = blockIdx.x, cy = blockIdx.y, cz = z0 + blockIdx.z;       // grid = (blocks per x-row, ncy, z-layers)
    const int b = (cz * G.ncy + cy) * G.nblk + bi;
    const int bx0 = tile_bx0(G, bi), w = tile_bx0(G, bi + 1) - bx0, kw = w + 2, nst = 9 * kw;
    tile_stencil_prefix(G, bx0, w, cy, cz, cell_start, s_cs, s_cn, s_off);
    __syncthreads();
    const int hcell = 4 * kw + 1;
    if (wi == 0) {       // block descriptor for k_force_tiles: header + the contiguous pieces of the stencil stream
        int* D = bdesc + (size_t)(b - G.b_base) * MDG_TILE_DESC;
        int npieces = 0;
        for (int base = 0; base < nst; base += 32) {
            const int t = base + lane;
            bool start = false;
            int cnt = 0;
            if (t < nst) {
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
                    cnt = s_off[e] - s_off[t];
                    start = cnt > 0;
                }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, start);
            if (start) {
                const int k = npieces + __popc(m & ((1u << lane) - 1u));
                D[8 + 3 * k] = s_cs[t];
                D[9 + 3 * k] = s_off[t];
                D[10 + 3 * k] = cnt;
            }
            npieces += __popc(m);
        }
        if (lane == 0) {
            D[0] = s_cs[hcell];
            D[1] = s_off[hcell + w] - s_off[hcell];
            D[2] = s_off[hcell];
            D[3] = s_off[nst];
            D[4] = npieces;
        }
    }
    if (wi >= w) return;
    const int blk_a0 = s_cs[hcell];                              // first atom of the block
    const int a0 = s_cs[hcell + wi], na = s_cn[hcell + wi];      // this warp's cell
    if (na == 0) return;
    const bool bad = (flags[6] | flags[7]) != 0, over = s_off[nst] > G.scap;
    if (bad || over) {
        for (int a = lane; a < na; a += 32) tlen[a0 + a] = 0;
        if (over && lane == 0) { flags[0] = 1; atomicMax(&flags[1], s_off[nst]); }
        return;
    }
    // candidate prefix over this cell's 27 stencil cells in STREAM order: kk = r * 3 + j  <->  block stencil cell r * kw + wi + j
    {
        const int t = lane < 27 ? (lane / 3) * kw + wi + (lane % 3) : 0;
        const int cnt = lane < 27 ? s_cn[t] : 0;
        int x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane < 27) s_pre[wi][lane + 1] = x;
        if (lane == 0) s_pre[wi][0] = 0;
    }
    __syncwarp();
    const int total = s_pre[wi][27];
    const int cx = bx0 + wi;
    const float ox = (float)cx / (float)G.ncx, oy = (float)cy / (float)G.ncy, oz = (float)cz / (float)G.ncz;
    const bool filt = (F.sel_a != nullptr) || (F.n_ex > 0);
    const int g0 = tile_group0(G, b, blk_a0);
    const int cap_slots = G.capc * MDG_TILE_CHUNK;

    for (int pass = 0; pass < na; pass += 32) {
        const int np = min(32, na - pass);
        const int s = a0 + pass + lane;                        // this lane's atom (phase 2)
        const bool act = lane < np;
        int Iix = 0, Iiy = 0, Iiz = 0, idi = 0;
        uint32_t imc = 0;
        if (act) {
            float4 qi = qs[s];
            float lx, ly, lz;
            idi = __float_as_int(qi.w);
            local_coord(qi.x, bx.L[0], bx.invL[0], ox, lx, Iix);
            local_coord(qi.y, bx.L[1], bx.invL[1], oy, ly, Iiy);
            local_coord(qi.z, bx.L[2], bx.invL[2], oz, lz, Iiz);
            s_ctr[wi][lane] = make_float4(lx, ly, lz, 0.f);
            imc = pack_img(Iix, Iiy, Iiz);
        }
        const uint32_t im0 = __shfl_sync(0xffffffffu, imc, 0);
        const bool ctr_uniform = __all_sync(0xffffffffu, !act || imc == im0);
        // this row's place in the interleaved layout
        const int rl = s - blk_a0;                             // row index inside the block
        uint16_t* rowp = trows + ((size_t)(g0 + (rl >> 3)) * G.capc) * MDG_TILE_GCHUNK + (rl & 7) * MDG_TILE_CHUNK;
        const uint32_t self_off = (uint32_t)(s_off[hcell + wi] + pass + lane) << 4;
        int nA = 0, nBtot = 0;
        for (int B = 0; B < total; B += FB_BATCH) {
            const int nb = min(FB_BATCH, total - B);
            const int nch = (nb + 31) >> 5;
            __syncwarp();
            // ---------------- phase 1: lane = candidate -------------------------------------------
            int kk = 0;
            bool cand_uniform = true;
            for (int ch = 0; ch < nch; ++ch) {
                const int a = B + (ch << 5) + lane;
                const bool valid = a < B + nb;
                float lx = 1e30f, ly = 1e30f, lz = 1e30f;
                if (valid) {
                    while (a >= s_pre[wi][kk + 1]) ++kk;
                    const int t = (kk / 3) * kw + wi + (kk % 3);       // block stencil cell of this candidate
                    const int ia = a - s_pre[wi][kk];
                    float4 qj = qs[s_cs[t] + ia];
                    int Ix, Iy, Iz;
                    local_coord(qj.x, bx.L[0], bx.invL[0], ox, lx, Ix);
                    local_coord(qj.y, bx.L[1], bx.invL[1], oy, ly, Iy);
                    local_coord(qj.z, bx.L[2], bx.invL[2], oz, lz, Iz);
                    uint32_t imj = pack_img(Ix, Iy, Iz);
                    s_img[wi][a - B] = imj;
                    s_loc[wi][a - B] = (uint16_t)((s_off[t] + ia) << 4);         // byte offset of the staged float4
                    cand_uniform = cand_uniform && (imj == im0);
                }
#pragma unroll 4
                for (int i = 0; i < np; ++i) {          // (the self pair passes here and is dropped in phase 2)
                    float4 ci = s_ctr[wi][i];
                    float dx = lx - ci.x, dy = ly - ci.y, dz = lz - ci.z;
                    float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    uint32_t m = __ballot_sync(0xffffffffu, d2 < r2list);
                    if (lane == 0) s_mask[wi][i][ch] = m;
                }
            }
            const bool uniform = ctr_uniform && __all_sync(0xffffffffu, cand_uniform) && !filt;
            __syncwarp();
            // ---------------- phase 2: lane = atom ------------------------------------------------
            // Entry k of the unshifted segment lives at chunk k / 16, slot 4 (k % 4) + (k / 4) % 4 (lane q of the row's four lanes
            // reads slots 4 q .. 4 q + 3 with one 8-byte load); pair b of the shifted segment at chunk nAc + b / 8, slots 2 (b % 8), + 1.
            // walk 0: unshifted entries (and the count of the shifted ones); walk 1 (rows that have shifted entries): the shifted
            // entries behind the padded unshifted segment.  Rows with several batches (total > FB_BATCH: cells far above liquid
            // density) keep every entry of a later batch in the shifted form.
            if (act) {
                {   // the row's own atom is one of the candidates: clear its bit once instead of testing every entry
                    const int a_self = s_pre[wi][13] + pass + lane - B;     // stencil slot 13 = the cell itself (r = 4, j = 1)
                    if (a_self >= 0 && a_self < nb) s_mask[wi][lane][a_self >> 5] &= ~(1u << (a_self & 31));
                }
                int nBb = 0;
                if (uniform && B == 0) {
                    // interior cells: no pair crosses a periodic boundary, no filter - the common, lean loop
                    // One flattened loop over ALL set bits of the batch: lanes drift apart across the 32-candidate words, but
                    // (nearly) every iteration of every lane emits an entry - a per-word loop would make all lanes wait for the
                    // largest popcount of each word (measured: 199 instead of ~125 iterations per cell).
                    uint16_t* wp = rowp;
                    int k = 0, ch = 0;
                    uint32_t m = s_mask[wi][lane][0];
                    const uint16_t* lp = &s_loc[wi][0];
                    while (true) {
                        if (m == 0) {
                            if (++ch >= nch) break;
                            m = s_mask[wi][lane][ch];
                            lp += 32;
                            continue;
                        }
                        const int bit = __ffs(m) - 1;
                        m &= m - 1;
                        if (k < cap_slots) *wp = lp[bit];
                        // transposed inside a chunk (entry k at slot 4 (k % 4) + (k / 4) % 4): the four lanes of the row
                        // then read four CONSECUTIVE neighbors with one LDS.128 each - adjacent shared-memory slots, fewer
                        // bank conflicts (ncu: 2.9 M vs 4.0 M conflict wavefronts per launch with the plain order)
                        wp += ((k & 3) != 3) ? 4 : (((k & 15) == 15) ? (MDG_TILE_GCHUNK - 15) : -11);
                        ++k;
                    }
                    nA = k;
                } else {
                    int ch = 0;
                    uint32_t m = s_mask[wi][lane][0];
                    while (true) {
                        if (m == 0) {
                            if (++ch >= nch) break;
                            m = s_mask[wi][lane][ch];
                            continue;
                        }
                        const int bit = __ffs(m) - 1;
                        m &= m - 1;
                        const int al = (ch << 5) + bit;
                        const uint32_t lo = s_loc[wi][al];
                        bool plain = uniform;
                        if (!uniform) {
                            const uint32_t im = s_img[wi][al];
                            const int mx = (int)(im & 1023u) - 512 - Iix;
                            const int my = (int)((im >> 10) & 1023u) - 512 - Iiy;
                            const int mz = (int)((im >> 20) & 1023u) - 512 - Iiz;
                            if ((unsigned)(mx + 1) > 2u || (unsigned)(my + 1) > 2u || (unsigned)(mz + 1) > 2u) continue;
                            if (filt && !pair_allowed(F, idi, __float_as_int(qs[tile_global_of(s_cs, s_off, nst, (int)(lo >> 4))].w))) continue;
                            plain = (mx | my | mz) == 0;
                        }
                        if (plain && B == 0) {
                            if (nA < cap_slots) rowp[(nA >> 4) * MDG_TILE_GCHUNK + ((nA & 3) << 2) + ((nA >> 2) & 3)] = (uint16_t)lo;
                            ++nA;
                        } else {
                            ++nBb;
                        }
                    }
                }
                // pad the unshifted segment to whole chunks (after the first batch only: later batches add shifted-form entries)
                int nAc = (nA + 15) >> 4;
                if (B == 0) {
                    const int endA = min(nAc << 4, cap_slots);
                    for (int k = nA; k < endA; ++k) rowp[(k >> 4) * MDG_TILE_GCHUNK + ((k & 3) << 2) + ((k >> 2) & 3)] = (uint16_t)self_off;
                }
                if (nBb > 0) {
                    int ch = 0;
                    uint32_t m = s_mask[wi][lane][0];
                    int nBw = nBtot;
                    while (true) {
                        if (m == 0) {
                            if (++ch >= nch) break;
                            m = s_mask[wi][lane][ch];
                            continue;
                        }
                        const int bit = __ffs(m) - 1;
                        m &= m - 1;
                        const int al = (ch << 5) + bit;
                        const uint32_t lo = s_loc[wi][al];
                        int mx = 0, my = 0, mz = 0;
                        if (!uniform) {
                            const uint32_t im = s_img[wi][al];
                            mx = (int)(im & 1023u) - 512 - Iix;
                            my = (int)((im >> 10) & 1023u) - 512 - Iiy;
                            mz = (int)((im >> 20) & 1023u) - 512 - Iiz;
                            if ((unsigned)(mx + 1) > 2u || (unsigned)(my + 1) > 2u || (unsigned)(mz + 1) > 2u) continue;
                            if (filt && !pair_allowed(F, idi, __float_as_int(qs[tile_global_of(s_cs, s_off, nst, (int)(lo >> 4))].w))) continue;
                            if (B == 0 && (mx | my | mz) == 0) continue;           // written in walk 0
                        }
                        const int slot = (nAc << 4) + 2 * nBw;                      // 16-bit slot of this (offset, code) pair
                        if (slot + 1 < cap_slots) {
                            uint16_t* e = rowp + (slot >> 4) * MDG_TILE_GCHUNK + (slot & 15);
                            e[0] = (uint16_t)lo;
                            e[1] = (uint16_t)((1 - mx) | ((1 - my) << 2) | ((1 - mz) << 4));
                        }
                        ++nBw;
                    }
                    nBtot = nBw;
                }
            }
        }
        if (act) {
            const int nAc = (nA + 15) >> 4, nBc = (nBtot + 7) >> 3;
            {   // pad the shifted segment to whole chunks with (self, no shift)
                const int endB = min(nBc << 3, (cap_slots >> 1) - (nAc << 3));
                for (int k = nBtot; k < endB; ++k) {
                    const int slot = (nAc << 4) + 2 * k;
                    uint16_t* e = rowp + (slot >> 4) * MDG_TILE_GCHUNK + (slot & 15);
                    e[0] = (uint16_t)self_off;
                    e[1] = (uint16_t)(1 | (1 << 2) | (1 << 4));
                }
            }
            int a_c = nAc, b_c = nBc;
            if (nAc + nBc > G.capc) {
                atomicMax(&flags[2], (nAc + nBc) * MDG_TILE_CHUNK);
                flags[0] = 1;
                a_c = min(nAc, G.capc);
                b_c = min(nBc, G.capc - a_c);
            }
            tlen[s] = (uint32_t)a_c | ((uint32_t)b_c << 8);
        }
        __syncwarp();
    }
}
