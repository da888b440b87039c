// build_fast.cuh - fast builder for the engine's Verlet-SKIN list (included by nbr.cu).
//
// Membership at the LIST radius does not have to be bit-exact: the force kernel re-tests every
// entry against rc^2 with the reference arithmetic, and the skin criterion (no atom moved more than
// skin/2 since the build) has orders of magnitude more slack than the ~1e-5 coordinate error of the
// cell-local frame used here.  So candidates are screened with a 7-instruction FMA distance test.
//
// One warp per cell:
//   stage   lane = candidate: the atoms of the 27 stencil cells are loaded once (coalesced float4),
//           converted to cell-local coordinates + integer image counts and kept in shared memory;
//   scan    for every atom i of the cell (serial), lane = candidate again: 32 candidates per
//           iteration are tested against i, one ballot, accepted lanes write their entry straight
//           to row_i at cnt + popc(lower lanes)  -> rows in ascending neighbor index, 128-byte
//           coalesced stores, no divergent accept path.
// Image code of an accepted pair = integer image difference of the two atoms (x = (frac + I) L in
// the cell's frame): off = -(I_j - I_i), equal to the reference's -(red > 0.5) + (red < -0.5) for
// every pair that can be inside the cutoff; pairs with |m| >= 2 are the ones the reference's single
// +-1 correction loses (SURVEY 7 "unwrapped positions"): dropped.
// Requires >= 5 cells per axis (stencil extent < half a box), else the exact builder is used.
#pragma once

#define FB_WARPS 2
#define FB_CAP 768               // staged candidates per cell (27 cells x ~24 atoms = ~650 at liquid density)

__device__ __forceinline__ void local_coord(float x, float L, float invL, float origin, float& l, int& I) {
    float f = x * invL;
    float nf = floorf(f);
    float u = (f - nf) - origin;        // fractional position relative to the cell origin
    float r = rintf(u);                 // nearest periodic image of the cell frame
    l = (u - r) * L;
    I = (int)nf + (int)r;
}

__device__ __forceinline__ uint32_t pack_img(int Ix, int Iy, int Iz) {
    return (uint32_t)((Ix + 512) & 1023) | ((uint32_t)((Iy + 512) & 1023) << 10) | ((uint32_t)((Iz + 512) & 1023) << 20);
}

__device__ __forceinline__ uint32_t img_code(uint32_t imj, uint32_t imi, bool& ok) {
    int mx = (int)(imj & 1023u) - (int)(imi & 1023u);
    int my = (int)((imj >> 10) & 1023u) - (int)((imi >> 10) & 1023u);
    int mz = (int)((imj >> 20) & 1023u) - (int)((imi >> 20) & 1023u);
    ok = ((unsigned)(mx + 1) <= 2u) && ((unsigned)(my + 1) <= 2u) && ((unsigned)(mz + 1) <= 2u);
    return (uint32_t)((1 - mx) | ((1 - my) << 2) | ((1 - mz) << 4)) << MDG_IDX_BITS;
}

__global__ void __launch_bounds__(FB_WARPS * 32) k_build_fast(int ncell, const float4* __restrict__ qs,
                                                             const int* __restrict__ cell_start, const int* __restrict__ stencil,
                                                             Box bx, int ncx, int ncy, int ncz, float r2list, int cap,
                                                             PairFilter F, uint32_t* __restrict__ rows,
                                                             int* __restrict__ row_len, int* __restrict__ flags) {
    __shared__ float4 s_loc[FB_WARPS][FB_CAP];       // local x,y,z ; w = sorted index t (int bits)
    __shared__ uint32_t s_img[FB_WARPS][FB_CAP];
    __shared__ int s_pre[FB_WARPS][28];              // candidate-index prefix over the 27 stencil cells
    __shared__ int s_cs[FB_WARPS][27];               // cell_start of the stencil cells
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * FB_WARPS + w;
    if (c >= ncell) return;
    const int a0 = cell_start[c], na = cell_start[c + 1] - a0;
    if (na == 0) return;
    if (flags[6] | flags[7]) {
        for (int a = lane; a < na; a += 32) row_len[a0 + a] = 0;
        return;
    }
    int kself = 0;
    {   // stencil prefix (warp scan over the 27 counts) and the slot of the cell itself
        int cc = lane < 27 ? stencil[c * 27 + lane] : -1;
        int cs = lane < 27 ? cell_start[cc] : 0;
        int cnt = lane < 27 ? cell_start[cc + 1] - cs : 0;
        int x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane < 27) { s_pre[w][lane + 1] = x; s_cs[w][lane] = cs; }
        if (lane == 0) s_pre[w][0] = 0;
        kself = __ffs(__ballot_sync(0xffffffffu, cc == c)) - 1;
    }
    __syncwarp();
    const int total = s_pre[w][27];
    const int cx = c % ncx, cy = (c / ncx) % ncy, cz = c / (ncx * ncy);
    const float ox = (float)cx / (float)ncx, oy = (float)cy / (float)ncy, oz = (float)cz / (float)ncz;
    const bool filt = (F.sel_a != nullptr) || (F.n_ex > 0);
    const uint32_t lt = (1u << lane) - 1u;
    const bool staged = total <= FB_CAP;

    // ---- stage (or, for over-full stencils, only check image uniformity) ---------------------------
    uint32_t img0 = 0;
    bool uniform = true;
    {
        int kk = 0;
        for (int a = lane; a < total; a += 32) {
            while (a >= s_pre[w][kk + 1]) ++kk;
            int t = s_cs[w][kk] + (a - s_pre[w][kk]);
            float4 qj = qs[t];
            float lx, ly, lz;
            int Ix, Iy, Iz;
            local_coord(qj.x, bx.L[0], bx.invL[0], ox, lx, Ix);
            local_coord(qj.y, bx.L[1], bx.invL[1], oy, ly, Iy);
            local_coord(qj.z, bx.L[2], bx.invL[2], oz, lz, Iz);
            uint32_t im = pack_img(Ix, Iy, Iz);
            if (staged) {
                s_loc[w][a] = make_float4(lx, ly, lz, __int_as_float(t));
                s_img[w][a] = im;
            }
            if (a == lane) img0 = __shfl_sync(__activemask(), im, 0);
            uniform = uniform && (im == img0);
        }
        img0 = __shfl_sync(0xffffffffu, img0, 0);
        uniform = __all_sync(0xffffffffu, uniform);
    }
    __syncwarp();
    const uint32_t ZERO_CODE = (1u | (1u << 2) | (1u << 4)) << MDG_IDX_BITS;

    // ---- scan: one atom of the cell at a time, all lanes on its candidates ---------------------------
    for (int i = 0; i < na; ++i) {
        const int s = a0 + i;
        float cix, ciy, ciz;
        uint32_t imi;
        int idi = 0;
        if (staged) {
            float4 ci = s_loc[w][s_pre[w][kself] + i];
            cix = ci.x; ciy = ci.y; ciz = ci.z;
            imi = s_img[w][s_pre[w][kself] + i];
        } else {
            float4 qi = qs[s];
            int Ix, Iy, Iz;
            local_coord(qi.x, bx.L[0], bx.invL[0], ox, cix, Ix);
            local_coord(qi.y, bx.L[1], bx.invL[1], oy, ciy, Iy);
            local_coord(qi.z, bx.L[2], bx.invL[2], oz, ciz, Iz);
            imi = pack_img(Ix, Iy, Iz);
        }
        if (filt) idi = __float_as_int(qs[s].w);
        uint32_t* row = rows + (size_t)s * cap;
        int cnt = 0;
        int kk = 0;
        for (int base = 0; base < total; base += 32) {
            const int a = base + lane;
            bool acc = false;
            uint32_t entry = 0;
            if (a < total) {
                float lx, ly, lz;
                int t;
                uint32_t imj;
                if (staged) {
                    float4 lj = s_loc[w][a];
                    lx = lj.x; ly = lj.y; lz = lj.z;
                    t = __float_as_int(lj.w);
                } else {
                    while (a >= s_pre[w][kk + 1]) ++kk;
                    t = s_cs[w][kk] + (a - s_pre[w][kk]);
                    float4 qj = qs[t];
                    int Ix, Iy, Iz;
                    local_coord(qj.x, bx.L[0], bx.invL[0], ox, lx, Ix);
                    local_coord(qj.y, bx.L[1], bx.invL[1], oy, ly, Iy);
                    local_coord(qj.z, bx.L[2], bx.invL[2], oz, lz, Iz);
                    imj = pack_img(Ix, Iy, Iz);
                }
                float dx = lx - cix, dy = ly - ciy, dz = lz - ciz;
                float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                acc = (d2 < r2list) && (t != s);
                entry = (uint32_t)t | ZERO_CODE;
                if (acc && (!uniform || filt)) {           // rare: boundary cells / filtered lists
                    if (staged) imj = s_img[w][a];
                    bool ok;
                    entry = (uint32_t)t | img_code(imj, imi, ok);
                    acc = ok;
                    if (acc && filt) acc = pair_allowed(F, idi, __float_as_int(qs[t].w));
                }
            }
            uint32_t m = __ballot_sync(0xffffffffu, acc);
            int pos = cnt + __popc(m & lt);
            if (acc && pos < cap) row[pos] = entry;
            cnt += __popc(m);
        }
        if (lane == 0) {
            if (cnt > cap) { atomicMax(&flags[2], cnt); flags[0] = 1; cnt = cap; }
            row_len[s] = cnt;
        }
    }
}
