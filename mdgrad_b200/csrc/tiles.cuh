// tiles.cuh - geometry of the engine's TILE list (build_tiles.cuh builds it, force_tiles.cuh consumes it).
//
// The engine's Verlet-skin list in its second form (round 2).  The first form (build_fast.cuh / k_force_rows) stores one row
// of 4-byte GLOBAL indices per atom and gathers neighbor positions through L1: ncu on the 256k-atom box shows that kernel
// bound by L1 wavefronts (83% L1 throughput, ~10 distinct 128-byte lines per 32-lane gather) and by its own row stream
// (104 MB of DRAM reads per launch = 16 us at the HBM peak).  The tile form removes both:
//
//   * the cell grid is cut into BLOCKS of <= 4 consecutive cells of one x-row; the force kernel runs one CTA per block and
//     stages the block's whole 3 x 3 x (w + 2) cell stencil - 9 runs of cells that are contiguous in the cell-sorted position
//     array, <= 18 contiguous pieces with the periodic wrap - into shared memory with TMA bulk copies (cp.async.bulk, coalesced
//     16-byte elements, completion on an mbarrier);
//   * row entries are 16-bit LOCAL indices into that staged stream (pre-multiplied by 16 = the byte offset of the float4), so
//     the row stream is half the bytes and a neighbor position is one LDS.128;
//   * the rows of the 8 atoms a warp works on are stored INTERLEAVED chunk by chunk (8 rows x 16 entries = 256 contiguous
//     bytes), so one warp-wide 8-byte load per lane is two full 128-byte lines instead of eight partial ones;
//   * entries that need a periodic image shift live in a separate segment of the row (32-bit: index + image code), so the
//     common loop has no code test at all, also in boundary blocks.
//
// Membership at the list radius is the builder's approximate (FMA) test; the force kernel re-tests every entry with the
// reference arithmetic (topology.py:59-68) exactly as k_force_rows does, so the contributing pair set is the reference's.
#pragma once
#include "common.cuh"

#define MDG_TILE_GROUP 8            // rows per warp
#define MDG_TILE_LANES 4            // lanes per row
#define MDG_TILE_CHUNK 16           // 16-bit slots per row per chunk (4 lanes x 4 slots = one 8-byte load per lane)
#define MDG_TILE_GCHUNK (MDG_TILE_GROUP * MDG_TILE_CHUNK)   // 16-bit slots per group per chunk (256 bytes)
#define MDG_TILE_MAXW 4             // cells per block along x
#define MDG_TILE_MAXST (9 * (MDG_TILE_MAXW + 2))            // stencil cells of a block
#define MDG_TILE_MAXSCAP 4095       // staged atoms: (index << 4) must fit 16 bits
#define MDG_TILE_DESC 64            // ints per block descriptor: [0] first row, [1] rows, [2] home offset in the stream, [3] staged atoms,
                                    // [4] pieces, then (source index, stream offset, count) per contiguous piece from [8]


__host__ __device__ __forceinline__ int tile_bx0(const TileGeom& G, int bi) { return bi * G.wbase + (bi < G.wrem ? bi : G.wrem); }

// t / kw and t % kw for t < 64, kw in [3, 6] without the integer-division sequence
__device__ __forceinline__ int tile_div_kw(int t, int kw) { return (t * (65536 / kw + 1)) >> 16; }

// stencil cell (r, k) of block (bi, cy, cz): r = (dz + 1) * 3 + (dy + 1), k = 0 .. w + 1  ->  linear cell id
__device__ __forceinline__ int tile_stencil_cell(const TileGeom& G, int bx0, int cy, int cz, int r, int k) {
    int x = bx0 - 1 + k;
    x = x < 0 ? x + G.ncx : (x >= G.ncx ? x - G.ncx : x);
    int y = cy + (r % 3) - 1;
    y = y < 0 ? y + G.ncy : (y >= G.ncy ? y - G.ncy : y);
    int z = cz + (r / 3) - 1;
    z = z < 0 ? z + G.ncz : (z >= G.ncz ? z - G.ncz : z);
    return (z * G.ncy + y) * G.ncx + x;
}

// first group of block b (upper bound construction: every block before b adds at most one partially filled group)
__device__ __forceinline__ int tile_group0(const TileGeom& G, int b, int a0) { return (a0 >> 3) + (b - G.b_base) - G.g_base; }

// Exclusive prefix over the block's stencil cell counts, computed by warp 0 into s_off[0 .. nst] (s_off[nst] = total).
// s_cs / s_cn receive cell_start / count of every stencil cell.  Callers __syncthreads() afterwards.
__device__ __forceinline__ void tile_stencil_prefix(const TileGeom& G, int bx0, int w, int cy, int cz,
                                                    const int* __restrict__ cell_start, int* s_cs, int* s_cn, int* s_off) {
    const int lane = threadIdx.x & 31;
    if ((threadIdx.x >> 5) != 0) return;
    const int kw = w + 2, nst = 9 * kw;
    int carry = 0;
    for (int base = 0; base < nst; base += 32) {
        const int t = base + lane;
        int cs = 0, cn = 0;
        if (t < nst) {
            const int r = tile_div_kw(t, kw);
            const int cc = tile_stencil_cell(G, bx0, cy, cz, r, t - r * kw);
            cs = cell_start[cc];
            cn = cell_start[cc + 1] - cs;
            s_cs[t] = cs;
            s_cn[t] = cn;
        }
        int x = cn;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (t < nst) s_off[t] = carry + x - cn;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) s_off[nst] = carry;
}

// sorted (global) index of the staged atom with stream index `loc` (rare paths only: pair filters)
__device__ __forceinline__ int tile_global_of(const int* s_cs, const int* s_off, int nst, int loc) {
    int t = 0;
    while (t + 1 < nst && loc >= s_off[t + 1]) ++t;
    return s_cs[t] + (loc - s_off[t]);
}
