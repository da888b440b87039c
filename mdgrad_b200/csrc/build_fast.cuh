// build_fast.cuh - fast builder for the engine's Verlet-SKIN list (included by nbr.cu).
//
// Membership at the LIST radius does not have to be bit-exact: the force kernel re-tests every
// entry against rc^2 with the reference arithmetic, and the skin criterion (no atom moved more than
// skin/2 since the build) has orders of magnitude more slack than the ~1e-5 coordinate error of the
// cell-local frame used here.  So candidates are screened with a 7-instruction FMA distance test.
//
// One warp per cell, two phases per batch of <= 1024 stencil candidates:
//   phase 1  lane = CANDIDATE (coalesced float4 loads of the 27 stencil cells, 98% lane use):
//            loop over the cell's atoms (broadcast LDS), branch-free test, one ballot per
//            (atom, 32-candidate chunk) -> bitmask in shared memory.
//   phase 2  lane = ATOM: walk the set bits in ascending candidate order and append the row entries
//            (ascending neighbor index, as the force kernel and the export expect).
// The image code of an accepted pair is the integer image difference of the two atoms
// (x = (frac + I) L in the cell's frame): off = -(I_j - I_i), equal to the reference's
// -(red > 0.5) + (red < -0.5) for every pair that can be inside the cutoff; pairs with |m| >= 2 are
// the ones the reference's single +-1 correction loses (SURVEY 7 "unwrapped positions"): dropped.
// Requires >= 5 cells per axis (stencil extent < half a box), else the exact builder is used.
#pragma once

#ifndef FB_WARPS
#define FB_WARPS 4                // cells (warps) per CTA; build variant fbw2 (8 would exceed the 48 KB static shared-memory limit)
#endif
#ifndef FB_DYNAMIC
#define FB_DYNAMIC 1            // warps take cells from a global counter (0: cell = f(blockIdx, warp), build variant fbst)
#endif
#ifndef FB_PAIR
#define FB_PAIR 1               // two candidates per lane in phase 1 (0: one, build variant fbp0): 246 -> 228 us per rebuild
#endif
#define FB_BATCH 736             // candidates per batch (23 chunks of 32; a 27-cell stencil holds ~650 at liquid density)
#define FB_CHUNKS (FB_BATCH / 32)

__device__ __forceinline__ void local_coord(float x, float L, float invL, float origin, float& l, int& I) {
    float f = x * invL;
    float nf = floorf(f);
    float u = (f - nf) - origin;        // fractional position relative to the cell origin
    float r = rintf(u);                 // nearest periodic image of the cell frame
    l = (u - r) * L;
    I = (int)nf + (int)r;
}

// Row storage order: inside every aligned block of 16 entries, logical entry k is stored at slot
// 4*(k%4) + (k/4)%4, so that the four lanes that stream a row in k_force_rows (one uint4 = 4 slots each) gather
// four CONSECUTIVE neighbors per load instruction (same 128-byte line): 57.8 -> 53.4 us on the 256k-atom box.
// (Only the engine's force kernel reads these rows; it is order-agnostic inside a block.)
__device__ __forceinline__ int fb_slot(int k) { return (k & ~15) | ((k & 3) << 2) | ((k >> 2) & 3); }

// ---- experimental build variant MDG_BUILD_INT8_SCREEN (mdgrad_b200/build.py VARIANTS["i8"]; NOT the default) -------------
// Phase 1 screens candidates on cell-local coordinates quantised to 8 bits per axis: one VABSDIFF4.U8 + one IDP.4A.U8.U8
// give the squared distance in quantisation units (both single SASS instructions on sm_100a) instead of 3 FADD + FMUL +
// 2 FFMA.  Local coordinates lie in [-c, 2c) per axis (the cell and its +-1 neighbours), so with ONE scale S = 255 / (3 c_max)
// for all axes q = floor((l + c) S) fits a byte and |dq_k| < |dl_k| S + 1, hence
//      sum dq^2 < (S d + sqrt(3))^2   for every pair with true distance d,
// and the threshold T = (S r_list + sqrt(3) + 0.05)^2 can never lose a pair inside r_list (0.05: fp32 slack of the quantiser).
// The price: pairs up to ~2 sqrt(3)/S (0.12 sigma at the benchmark geometry) beyond r_list also pass - ~6% longer rows, which the
// force kernel's exact re-test discards.  Whether the cheaper phase 1 pays for the longer rows is a GPU measurement (round 2).
__device__ __forceinline__ uint32_t fb_quant(float lx, float ly, float lz, float cx, float cy, float cz, float S) {
    int qx = min(max((int)floorf((lx + cx) * S), 0), 255);
    int qy = min(max((int)floorf((ly + cy) * S), 0), 255);
    int qz = min(max((int)floorf((lz + cz) * S), 0), 255);
    return (uint32_t)qx | ((uint32_t)qy << 8) | ((uint32_t)qz << 16);
}

// Predicated store / opaque pointer: without them nvcc branches around every row store and re-derives the row base (s * cap,
// 64-bit) per entry - 20 instructions per entry in phase 2 instead of 10.
__device__ __forceinline__ void fb_store(uint32_t* dst, uint32_t v, bool ok) {
#ifdef MDG_EMU
    if (ok) *dst = v;
#else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.global.u32 [%0], %1;\n\t}" ::"l"(dst), "r"(v), "r"((int)ok) : "memory");
#endif
}
__device__ __forceinline__ void fb_opaque(uint32_t*& p) {
#ifndef MDG_EMU
    asm volatile("" : "+l"(p));
#endif
}

__device__ __forceinline__ uint32_t pack_img(int Ix, int Iy, int Iz) {
    return (uint32_t)((Ix + 512) & 1023) | ((uint32_t)((Iy + 512) & 1023) << 10) | ((uint32_t)((Iz + 512) & 1023) << 20);
}

// Phase 2 of k_build_fast for one lane (= one atom): walk the set bits of its mask words in ascending candidate order and
// append the row entries.  UNIFORM (warp-uniform): every atom of the cell and every candidate share one periodic image, all
// entries carry `uni_code`; otherwise the image code comes from s_dim and pairs outside +-1 image (or filtered ones) are dropped.
template <bool UNIFORM>
__device__ __forceinline__ int fb_walk(uint32_t nz, const uint32_t* mrow, const unsigned short* s_tk, const unsigned short* s_dim,
                                       const int* s_cs, uint32_t* rowp, int cnt, int cap, uint32_t uni_code, int dIx, int dIy,
                                       int dIz, bool filt, const PairFilter& F, int idi, const float4* __restrict__ qs) {
    if (nz == 0u) return cnt;
    const int ch0 = __ffs(nz) - 1;
    nz &= nz - 1;
    uint32_t m = mrow[ch0];
    const unsigned short* tkp = s_tk + (ch0 << 5);
    const unsigned short* dmp = s_dim + (ch0 << 5);
    do {
        const int b = __ffs(m) - 1;            // candidate b of the current word
        m &= m - 1;
        const unsigned tk = tkp[b];
        unsigned dc = 0;
        if (!UNIFORM) dc = dmp[b];
        if (m == 0u && nz != 0u) {             // next non-empty word
            const int ch = __ffs(nz) - 1;
            nz &= nz - 1;
            m = mrow[ch];
            tkp = s_tk + (ch << 5);
            dmp = s_dim + (ch << 5);
        }
        const uint32_t ref = (uint32_t)(s_cs[tk >> 11] + (int)(tk & 2047u));
        if (UNIFORM) {
            fb_store(rowp + fb_slot(cnt), ref | uni_code, cnt < cap);
            ++cnt;
        } else {
            // dc == 0xFFFF: image far outside the window, never a listed pair (see s_dim)
            const int mx = (int)(dc & 31u) - dIx, my = (int)((dc >> 5) & 31u) - dIy, mz = (int)(dc >> 10) - dIz;
            bool keep = dc != 0xFFFFu && (unsigned)(mx + 1) <= 2u && (unsigned)(my + 1) <= 2u && (unsigned)(mz + 1) <= 2u;
            if (filt && keep) keep = pair_allowed(F, idi, __float_as_int(qs[ref].w));
            if (keep) {
                const uint32_t code = (uint32_t)((1 - mx) | ((1 - my) << 2) | ((1 - mz) << 4)) << MDG_IDX_BITS;
                fb_store(rowp + fb_slot(cnt), ref | code, cnt < cap);
                ++cnt;
            }
        }
    } while (m != 0u);
    return cnt;
}

#ifdef FB_MINBLOCKS
#define FB_BOUNDS __launch_bounds__(FB_WARPS * 32, FB_MINBLOCKS)
#else
#define FB_BOUNDS __launch_bounds__(FB_WARPS * 32)
#endif
#ifndef FB_TRIP
#define FB_TRIP 2                 // atoms per trip of the paired phase-1 loop (build variant fbp4: 4)
#endif
__global__ void FB_BOUNDS k_build_fast(int cell0, int ncell, const float4* __restrict__ qs,
                                                             const int* __restrict__ cell_start, const int* __restrict__ stencil,
                                                             Box bx, int ncx, int ncy, int ncz, float r2list, int cap,
                                                             PairFilter F, uint32_t* __restrict__ rows,
                                                             int* __restrict__ row_len, int* __restrict__ flags,
                                                             int* __restrict__ work) {
    __shared__ uint32_t s_mask[FB_WARPS][32][FB_CHUNKS + 1];   // [atom][chunk], padded: conflict-free for lane = atom
    // Per-candidate state is kept small - 3 bytes instead of 8 - because shared memory per warp is what bounds the occupancy of
    // this kernel (ncu r02: 27% of the warp slots at 9.75 KB per warp, issue rate 54% with 2.8 "wait" + 2.3 "short scoreboard"
    // stalls per issue: too few warps to cover its dependent ALU / LDS chains):
    //   s_tk  = (stencil slot << 11) | index inside that cell      -> sorted index = s_cs[slot] + index
    //   s_dim = image of the candidate RELATIVE to the pass's reference image I0: 5 bits per axis holding dI + 16, 0xFFFF = outside
    //           [-16, 15] (a pair is only ever listed for |I_j - I_i| <= 1, so this loses nothing unless two atoms of ONE cell
    //           are themselves more than 14 box lengths apart in raw coordinates)
    __shared__ unsigned short s_dim[FB_WARPS][FB_BATCH];
    __shared__ unsigned short s_tk[FB_WARPS][FB_BATCH];
    __shared__ float4 s_ctr[FB_WARPS][32];                     // local coords of the cell's atoms, w = packed image
#ifdef MDG_BUILD_INT8_SCREEN
    __shared__ uint32_t s_cq[FB_WARPS][32];                    // the same, quantised (fb_quant)
    const float i8_cx = bx.L[0] / (float)ncx, i8_cy = bx.L[1] / (float)ncy, i8_cz = bx.L[2] / (float)ncz;
    const float i8_S = 255.0f / (3.0f * fmaxf(i8_cx, fmaxf(i8_cy, i8_cz)));
    const float i8_t = i8_S * sqrtf(r2list) + 1.7820508f;      // sqrt(3) + 0.05
    const uint32_t i8_T = (uint32_t)ceilf(i8_t * i8_t);
#endif
    __shared__ int s_pre[FB_WARPS][28];                        // candidate-index prefix over the 27 stencil cells
    __shared__ int s_cs[FB_WARPS][27];                         // cell_start of the stencil cells
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Cells are handed out by a global counter: a warp that finishes a cheap cell takes the next one, and CTAs that start late
    // find the counter exhausted.  The static form (cell = f(blockIdx, warp)) ran 2.25 waves of CTAs whose four warps end at
    // different times: 40 % of the warp slots active against 50 % theoretical in the r02 capture.
#if FB_DYNAMIC
  for (;;) {
    int c = 0;
    if (lane == 0) c = cell0 + atomicAdd(work, 1);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= ncell) return;
#else
  for (int once = 0; once < 1; ++once) {
    (void)work;
    const int c = cell0 + blockIdx.x * FB_WARPS + w;          // cells [cell0, ncell) (ncell = end of this rank's range)
    if (c >= ncell) return;
#endif
    const int a0 = cell_start[c], na = cell_start[c + 1] - a0;
    if (na == 0) continue;
    if (flags[6] | flags[7]) {
        for (int a = lane; a < na; a += 32) row_len[a0 + a] = 0;
        continue;
    }
    // stencil prefix (warp scan over 27 counts) and the stencil slot of the cell itself
    int kself_slot = 0;
    {
        int cc = lane < 27 ? stencil[c * 27 + lane] : 0;
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
        kself_slot = __ffs(__ballot_sync(0xffffffffu, lane < 27 && cc == c)) - 1;
        if (__any_sync(0xffffffffu, cnt >= 2048)) {        // (s_tk holds 11 bits of in-cell index: such a cell is a collapsed system)
            if (lane == 0) flags[7] = 2048;
            for (int a = lane; a < na; a += 32) row_len[a0 + a] = 0;
            continue;
        }
    }
    __syncwarp();
    const int total = s_pre[w][27];
    const int cx = c % ncx, cy = (c / ncx) % ncy, cz = c / (ncx * ncy);
    const float ox = (float)cx / (float)ncx, oy = (float)cy / (float)ncy, oz = (float)cz / (float)ncz;
    const bool filt = (F.sel_a != nullptr) || (F.n_ex > 0);

    for (int pass = 0; pass < na; pass += 32) {
        const int np = min(32, na - pass);                     // atoms of this pass
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
            s_ctr[w][lane] = make_float4(lx, ly, lz, 0.f);
#ifdef MDG_BUILD_INT8_SCREEN
            s_cq[w][lane] = fb_quant(lx, ly, lz, i8_cx, i8_cy, i8_cz, i8_S);
#endif
            imc = pack_img(Iix, Iiy, Iiz);
        }
        // image of the pass: if every atom of the cell and every candidate share it, all codes are "no shift"
        const uint32_t im0 = __shfl_sync(0xffffffffu, imc, 0);
        const bool ctr_uniform = __all_sync(0xffffffffu, !act || imc == im0);
        const int I0x = __shfl_sync(0xffffffffu, Iix, 0), I0y = __shfl_sync(0xffffffffu, Iiy, 0), I0z = __shfl_sync(0xffffffffu, Iiz, 0);
        uint32_t* row = rows + (size_t)(act ? s : a0) * cap;
        int cnt = 0;
        // a row is PURE when its single batch is uniform: bare indices, flagged in row_len (force kernel skips the
        // index mask and the image-code test)
        const bool pure_ok = (total <= FB_BATCH);
        bool row_pure = false;
        for (int B = 0; B < total; B += FB_BATCH) {
            const int nb = min(FB_BATCH, total - B);
            const int nch = (nb + 31) >> 5;
            __syncwarp();
            // ---------------- phase 1: lane = candidate -------------------------------------------
#if FB_PAIR && !defined(MDG_BUILD_INT8_SCREEN)
            // ---- build variant FB_PAIR: TWO candidates per lane (chunks ch and ch + 1), so that every LDS.128 of an atom's
            // coordinates feeds two distance tests (half the shared-memory loads and half as many exposed load latencies).
            bool cand_uniform = true;
            int kk_run = 0;
            auto fetch = [&](int chx, int& a, bool& valid, int& kk, float4& q) {
                a = B + (chx << 5) + lane;
                valid = (chx < nch) && (a < B + nb);
                if (valid) {
                    while (a >= s_pre[w][kk_run + 1]) ++kk_run;
                    q = qs[s_cs[w][kk_run] + (a - s_pre[w][kk_run])];
                }
                kk = kk_run;
            };
            auto stage = [&](int a, bool valid, int kk, const float4& qj, float& lx, float& ly, float& lz) {
                lx = 1e30f; ly = 1e30f; lz = 1e30f;
                if (valid) {
                    int Ix, Iy, Iz;
                    local_coord(qj.x, bx.L[0], bx.invL[0], ox, lx, Ix);
                    local_coord(qj.y, bx.L[1], bx.invL[1], oy, ly, Iy);
                    local_coord(qj.z, bx.L[2], bx.invL[2], oz, lz, Iz);
                    uint32_t imj = pack_img(Ix, Iy, Iz);
                    const int ex = Ix - I0x + 16, ey = Iy - I0y + 16, ez = Iz - I0z + 16;
                    s_dim[w][a - B] = ((unsigned)ex | (unsigned)ey | (unsigned)ez) > 31u ? (unsigned short)0xFFFF
                                                                                     : (unsigned short)(ex | (ey << 5) | (ez << 10));
                    s_tk[w][a - B] = (unsigned short)((kk << 11) | (a - s_pre[w][kk]));
                    cand_uniform = cand_uniform && (imj == im0);
                }
            };
            int aA_n, aB_n, kA_n, kB_n;
            bool vA_n, vB_n;
            float4 qA_n = make_float4(0.f, 0.f, 0.f, 0.f), qB_n = qA_n;
            fetch(0, aA_n, vA_n, kA_n, qA_n);
            fetch(1, aB_n, vB_n, kB_n, qB_n);
            for (int ch = 0; ch < nch; ch += 2) {
                const int aA = aA_n, aB = aB_n, kA = kA_n, kB = kB_n;
                const bool vA = vA_n, vB = vB_n;
                const float4 qA = qA_n, qB = qB_n;
                fetch(ch + 2, aA_n, vA_n, kA_n, qA_n);
                fetch(ch + 3, aB_n, vB_n, kB_n, qB_n);
                float ax, ay, az, bx_, by_, bz_;
                stage(aA, vA, kA, qA, ax, ay, az);
                stage(aB, vB, kB, qB, bx_, by_, bz_);
                int i = 0;
#if FB_TRIP == 4
                for (; i + 4 <= np; i += 4) {
                    const float4 c0 = s_ctr[w][i], c1 = s_ctr[w][i + 1], c2 = s_ctr[w][i + 2], c3 = s_ctr[w][i + 3];
                    uint32_t mA[4], mB[4];
#define FB_T(k, cc)                                                                                                         \
    {                                                                                                                       \
        const float x = ax - cc.x, y = ay - cc.y, z = az - cc.z, u = bx_ - cc.x, v = by_ - cc.y, t = bz_ - cc.z;             \
        mA[k] = __ballot_sync(0xffffffffu, fmaf(z, z, fmaf(y, y, x * x)) < r2list);                                         \
        mB[k] = __ballot_sync(0xffffffffu, fmaf(t, t, fmaf(v, v, u * u)) < r2list);                                         \
    }
                    FB_T(0, c0) FB_T(1, c1) FB_T(2, c2) FB_T(3, c3)
#undef FB_T
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) { s_mask[w][i + k][ch] = mA[k]; s_mask[w][i + k][ch + 1] = mB[k]; }
                    }
                }
#endif
                for (; i + 2 <= np; i += 2) {           // (the self pair passes here and is dropped in phase 2)
                    const float4 c0 = s_ctr[w][i], c1 = s_ctr[w][i + 1];
                    const float x0 = ax - c0.x, y0 = ay - c0.y, z0 = az - c0.z, x1 = ax - c1.x, y1 = ay - c1.y, z1 = az - c1.z;
                    const float u0 = bx_ - c0.x, v0 = by_ - c0.y, w0 = bz_ - c0.z, u1 = bx_ - c1.x, v1 = by_ - c1.y, w1 = bz_ - c1.z;
                    const float dA0 = fmaf(z0, z0, fmaf(y0, y0, x0 * x0)), dA1 = fmaf(z1, z1, fmaf(y1, y1, x1 * x1));
                    const float dB0 = fmaf(w0, w0, fmaf(v0, v0, u0 * u0)), dB1 = fmaf(w1, w1, fmaf(v1, v1, u1 * u1));
                    const uint32_t mA0 = __ballot_sync(0xffffffffu, dA0 < r2list), mA1 = __ballot_sync(0xffffffffu, dA1 < r2list);
                    const uint32_t mB0 = __ballot_sync(0xffffffffu, dB0 < r2list), mB1 = __ballot_sync(0xffffffffu, dB1 < r2list);
                    if (lane == 0) {
                        s_mask[w][i][ch] = mA0; s_mask[w][i][ch + 1] = mB0; s_mask[w][i + 1][ch] = mA1; s_mask[w][i + 1][ch + 1] = mB1;
                    }
                }
                for (; i < np; ++i) {
                    const float4 ci = s_ctr[w][i];
                    const float x0 = ax - ci.x, y0 = ay - ci.y, z0 = az - ci.z, u0 = bx_ - ci.x, v0 = by_ - ci.y, w0 = bz_ - ci.z;
                    const uint32_t mA = __ballot_sync(0xffffffffu, fmaf(z0, z0, fmaf(y0, y0, x0 * x0)) < r2list);
                    const uint32_t mB = __ballot_sync(0xffffffffu, fmaf(w0, w0, fmaf(v0, v0, u0 * u0)) < r2list);
                    if (lane == 0) { s_mask[w][i][ch] = mA; s_mask[w][i][ch + 1] = mB; }
                }
            }
#else
            bool cand_uniform = true;
            // The candidate of the NEXT chunk is fetched while this chunk is screened (the global load was an exposed long-scoreboard
            // stall once per chunk: 5.6 % of the samples in the r02 capture).
            int a_n = B + lane, kk_n = 0, t_n = 0;
            bool valid_n = a_n < B + nb;
            float4 q_n = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid_n) {
                while (a_n >= s_pre[w][kk_n + 1]) ++kk_n;
                t_n = s_cs[w][kk_n] + (a_n - s_pre[w][kk_n]);
                q_n = qs[t_n];
            }
            for (int ch = 0; ch < nch; ++ch) {
                const int a = a_n, kk = kk_n;
                const bool valid = valid_n;
                const float4 qj = q_n;
                if (ch + 1 < nch) {
                    a_n = B + ((ch + 1) << 5) + lane;
                    valid_n = a_n < B + nb;
                    if (valid_n) {
                        while (a_n >= s_pre[w][kk_n + 1]) ++kk_n;
                        t_n = s_cs[w][kk_n] + (a_n - s_pre[w][kk_n]);
                        q_n = qs[t_n];
                    }
                }
                float lx = 1e30f, ly = 1e30f, lz = 1e30f;
                if (valid) {
                    int Ix, Iy, Iz;
                    local_coord(qj.x, bx.L[0], bx.invL[0], ox, lx, Ix);
                    local_coord(qj.y, bx.L[1], bx.invL[1], oy, ly, Iy);
                    local_coord(qj.z, bx.L[2], bx.invL[2], oz, lz, Iz);
                    uint32_t imj = pack_img(Ix, Iy, Iz);
                    const int ex = Ix - I0x + 16, ey = Iy - I0y + 16, ez = Iz - I0z + 16;
                    s_dim[w][a - B] = ((unsigned)ex | (unsigned)ey | (unsigned)ez) > 31u ? (unsigned short)0xFFFF
                                                                                     : (unsigned short)(ex | (ey << 5) | (ez << 10));
                    s_tk[w][a - B] = (unsigned short)((kk << 11) | (a - s_pre[w][kk]));
                    cand_uniform = cand_uniform && (imj == im0);
                }
#ifdef MDG_BUILD_INT8_SCREEN
                const uint32_t cq = valid ? fb_quant(lx, ly, lz, i8_cx, i8_cy, i8_cz, i8_S) : 0u;
                const uint32_t thr = valid ? i8_T : 0u;        // lanes past the end of the stream never pass
#pragma unroll 4
                for (int i = 0; i < np; ++i) {          // (the self pair passes here and is dropped in phase 2)
                    uint32_t d = __vabsdiffu4(cq, s_cq[w][i]);
                    uint32_t d2 = __dp4a(d, d, 0u);
                    uint32_t m = __ballot_sync(0xffffffffu, d2 < thr);
                    if (lane == 0) s_mask[w][i][ch] = m;
                }
#else
                // Four atoms per trip, their coordinates loaded BEFORE the first use: as one-atom iterations every LDS.128 was
                // followed by its dependent FADD (short-scoreboard stall per atom: 22 % of the samples in the r02 capture).
                int i = 0;
                for (; i + 4 <= np; i += 4) {           // (the self pair passes here and is dropped in phase 2)
                    const float4 c0 = s_ctr[w][i], c1 = s_ctr[w][i + 1], c2 = s_ctr[w][i + 2], c3 = s_ctr[w][i + 3];
                    const float x0 = lx - c0.x, y0 = ly - c0.y, z0 = lz - c0.z;
                    const float x1 = lx - c1.x, y1 = ly - c1.y, z1 = lz - c1.z;
                    const float x2 = lx - c2.x, y2 = ly - c2.y, z2 = lz - c2.z;
                    const float x3 = lx - c3.x, y3 = ly - c3.y, z3 = lz - c3.z;
                    const float d0 = fmaf(z0, z0, fmaf(y0, y0, x0 * x0)), d1 = fmaf(z1, z1, fmaf(y1, y1, x1 * x1));
                    const float d2 = fmaf(z2, z2, fmaf(y2, y2, x2 * x2)), d3 = fmaf(z3, z3, fmaf(y3, y3, x3 * x3));
                    const uint32_t m0 = __ballot_sync(0xffffffffu, d0 < r2list), m1 = __ballot_sync(0xffffffffu, d1 < r2list);
                    const uint32_t m2 = __ballot_sync(0xffffffffu, d2 < r2list), m3 = __ballot_sync(0xffffffffu, d3 < r2list);
                    if (lane == 0) {
                        s_mask[w][i][ch] = m0; s_mask[w][i + 1][ch] = m1; s_mask[w][i + 2][ch] = m2; s_mask[w][i + 3][ch] = m3;
                    }
                }
                for (; i < np; ++i) {
                    const float4 ci = s_ctr[w][i];
                    const float dx = lx - ci.x, dy = ly - ci.y, dz = lz - ci.z;
                    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const uint32_t m = __ballot_sync(0xffffffffu, d2 < r2list);
                    if (lane == 0) s_mask[w][i][ch] = m;
                }
#endif
            }
#endif
            const bool uniform = ctr_uniform && __all_sync(0xffffffffu, cand_uniform) && !filt;
            row_pure = pure_ok && uniform;
            const uint32_t uni_code = row_pure ? 0u : ((1u | (1u << 2) | (1u << 4)) << MDG_IDX_BITS);
            __syncwarp();
            // ---------------- phase 2: lane = atom ------------------------------------------------
            // One flattened loop per lane over ALL its set bits of the batch: lanes drift apart across chunks, but every
            // iteration of every active lane emits one entry.  Empty words are skipped through a bitmask of the non-empty ones
            // (corner cells of the stencil are mostly empty) and the self pair is cleared up front, so the loop body has no
            // "nothing to emit" iteration (ncu r02: 58 warp instructions per iteration, one in six of them an empty one).
            if (act) {
                uint32_t* const mrow = &s_mask[w][lane][0];
                {
                    const int as = s_pre[w][kself_slot] + pass + lane - B;          // this atom in the candidate stream
                    if ((unsigned)as < (unsigned)nb) mrow[as >> 5] &= ~(1u << (as & 31));
                }
                uint32_t nz = 0;
                for (int ch = 0; ch < nch; ++ch) nz |= (mrow[ch] != 0u ? 1u : 0u) << ch;
                const int dIx = Iix - I0x + 16, dIy = Iiy - I0y + 16, dIz = Iiz - I0z + 16;
                uint32_t* rowp = row;
                fb_opaque(rowp);                               // (keeps the row base in a register pair: no per-entry s * cap)
                if (uniform)
                    cnt = fb_walk<true>(nz, mrow, s_tk[w], s_dim[w], s_cs[w], rowp, cnt, cap, uni_code, dIx, dIy, dIz, false, F, idi, qs);
                else
                    cnt = fb_walk<false>(nz, mrow, s_tk[w], s_dim[w], s_cs[w], rowp, cnt, cap, 0u, dIx, dIy, dIz, filt, F, idi, qs);
            }
        }
        if (act) {
            if (cnt > cap) { atomicMax(&flags[2], cnt); flags[0] = 1; cnt = cap; }
            {   // length (+ PURE flag) and padding to a whole 32-entry block with self entries (see mdg_pad_row)
                row_len[s] = cnt | (row_pure ? MDG_ROW_PURE : 0);
                const uint32_t pad_code = row_pure ? 0u : ((1u | (1u << 2) | (1u << 4)) << MDG_IDX_BITS);
                const uint32_t self_ref = (uint32_t)s;
                int end = (cnt + 31) & ~31;
                if (end > cap) end = cap;
                for (int k = cnt; k < end; ++k) row[fb_slot(k)] = self_ref | pad_code;
            }
        }
        __syncwarp();
    }
  }
}
