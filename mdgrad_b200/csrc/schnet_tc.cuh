// schnet_tc.cuh - the SchNet dense layers on the 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32.
//
//     C (M x N) = epilogue( A (M x K, row-major fp32) * Bt^T ),   Bt = (N x K) row-major = torch Linear weight (out x in)
//
// EXPERIMENTAL - opt-in with MDG_SCHNET_TC=1, default OFF: written after the round's GPU budget was spent, it has NOT run on
// a B200 yet.  Its LOGIC (staging layout, descriptors, K loop, guards, epilogue mapping) runs under the CPU emulation through a
// functional model of the tensor-core primitives (below) and equals the SIMT path there; what only hardware can tell - the
// descriptor semantics as the model reads them from the CUTLASS headers, the fences, the timing - is the first task of round 2
// (tests/test_schnet.py::test_tc_gemm_*, tools/tc_check.py), before it becomes the default for the configs[4] layer sizes.
// The default path stays the SIMT kernel k_sn_gemm of schnet.cu; both share the epilogue codes.
//
// Precision: the reference computes these layers in fp32 and the parity bar is 1e-5, so one TF32 product (10-bit
// mantissa) is not enough.  3xTF32: x = hi + lo with hi = tf32(x), lo = tf32(x - hi);  A B ~ Ah Bh + Ah Bl + Al Bh, three
// kind::tf32 MMAs into the same fp32 TMEM accumulator (the dropped Al Bl term is ~2^-22 relative).
//
// Structure (one CTA = one 128 x NT tile, 128 threads, deliberately simple - single smem stage, no TMA):
//   per 32-wide k-block: all threads stage the fp32 rows of A and Bt from global memory, split them into hi / lo and store
//   them in the canonical K-major no-swizzle UMMA layout (8 x 16-byte core matrices; CuTe ((8,n),2):((1,SBO),LBO) in 16-byte
//   units: LBO = 128 B between the two 16-byte K chunks of one MMA, SBO = 1024 B between 8-row groups);
//   fence.proxy.async + barrier; ONE thread issues 4 x 3 tcgen05.mma (M128, N = NT, K8) and a tcgen05.commit onto an
//   mbarrier; everybody waits on it before the buffers are overwritten.  Epilogue: each warp reads its 32 TMEM lanes with
//   tcgen05.ld.32x32b.x16, applies bias / ssp / sigmoid-gate / residual and writes 64-byte row segments.
// Descriptor bit fields: cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor) of the vendored CUTLASS headers.
#pragma once

#define TC_M 128
#define TC_KB 32
#define TC_LBO 128u
#define TC_SBO 1024u

// ---------------------------------------------------------------------------------------------
// Primitives.  GPU build: thin wrappers around the tcgen05 / mbarrier PTX.  MDG_EMU build (tests/cuemu): a functional model
// that INTERPRETS the same 64-bit shared-memory descriptors and the 32-bit instruction descriptor (start address, LBO, SBO,
// M, N per cute/arch/mma_sm100_desc.hpp) on the emulated shared memory and keeps the accumulator in an emulated TMEM, so that
// the kernel below - staging layout, descriptor arithmetic, K loop / accumulate flags, tile guards, TMEM lane / column ->
// row / column mapping of the epilogue - runs unchanged on the CPU against the SIMT kernel (tests/test_emu_schnet.py).  The
// model says nothing about the hardware's timing or about the PTX syntax (ptxas checks the latter at build time).
// ---------------------------------------------------------------------------------------------
#ifndef MDG_EMU
#define TC_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
struct TcBarrier { uint64_t w; };
__device__ __forceinline__ uint32_t tc_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t tc_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tc_mbar_init(TcBarrier* b) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_addr(b)), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// warp-collective: allocate `cols` TMEM columns, the base address is written to *slot (shared memory)
__device__ __forceinline__ void tc_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_addr(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t tmem, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the tensor core (async proxy)
__device__ __forceinline__ void tc_fence_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
// completion of everything issued so far -> one arrival on the mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(TcBarrier* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_addr(b)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(TcBarrier* b, uint32_t parity) {
    const uint32_t bar = tc_smem_addr(b);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// warp-collective: this thread's TMEM lane (32 * warp + lane), 16 consecutive columns from `col`
__device__ __forceinline__ void tc_ld16(uint32_t tmem, int warp, int col, uint32_t* r) {
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)col;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
#else   // ------------------------------------------------------------------------------------ functional model
#define TC_DYN_SMEM(name) unsigned char* name = (unsigned char*)cuemu::dyn_smem()
struct TcBarrier { uint32_t phase; };
static float    g_tc_tmem[128][512];              // one block at a time in the emulator
static uint32_t g_tc_next_col;
static inline uint32_t tc_smem_addr(const void* p) { return (uint32_t)((const unsigned char*)p - (const unsigned char*)cuemu::dyn_smem()); }
static inline uint32_t tc_tf32(float x) {          // cvt.rna.tf32.f32: nearest, ties away from zero, 10 explicit mantissa bits
    uint32_t b;
    memcpy(&b, &x, 4);
    return (b + 0x1000u) & 0xFFFFE000u;
}
static inline void tc_mbar_init(TcBarrier* b) { b->phase = 0; }
static inline void tc_alloc(uint32_t* slot, uint32_t cols) {
    if (cuemu::lane_id() == 0) {
        if (cols < 32 || (cols & (cols - 1)) || cols > 512) { fprintf(stderr, "tc model: tcgen05.alloc of %u columns (power of two in [32, 512] required)\n", cols); abort(); }
        g_tc_next_col = 0;
        *slot = g_tc_next_col;                     // lane 0, column 0
        g_tc_next_col += cols;
        for (int l = 0; l < 128; ++l)
            for (uint32_t cidx = 0; cidx < cols; ++cidx) g_tc_tmem[l][cidx] = __int_as_float(0x7fc00000);   // TMEM is not zeroed
    }
}
static inline void tc_dealloc(uint32_t, uint32_t) {}
static inline void tc_fence_before() {}
static inline void tc_fence_after() {}
static inline void tc_fence_smem() {}
static inline float tc_elem(uint64_t desc, int row, int k) {      // element (row, k) of a K-major no-swizzle operand, k in [0, 8)
    const uint32_t start = (uint32_t)(desc & 0x3FFFu) << 4, lbo = (uint32_t)((desc >> 16) & 0x3FFFu) << 4,
                   sbo = (uint32_t)((desc >> 32) & 0x3FFFu) << 4;
    if (((desc >> 46) & 3u) != 1u || (desc >> 61) != 0u) { fprintf(stderr, "tc model: descriptor version / layout bits not as expected\n"); abort(); }
    const uint32_t byte = start + (uint32_t)(row >> 3) * sbo + (uint32_t)(row & 7) * 16u + (uint32_t)(k >> 2) * lbo + (uint32_t)(k & 3) * 4u;
    float v;
    memcpy(&v, (const unsigned char*)cuemu::dyn_smem() + byte, 4);
    return v;
}
static inline void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    const int N = (int)((idesc >> 17) & 0x3Fu) << 3, M = (int)((idesc >> 24) & 0x1Fu) << 4;
    if (M != 128 || N < 16 || N > 256 || (N & 15) || ((idesc >> 4) & 3u) != 1u || ((idesc >> 7) & 7u) != 2u || ((idesc >> 10) & 7u) != 2u ||
        ((idesc >> 15) & 3u) != 0u) {
        fprintf(stderr, "tc model: instruction descriptor %08x is not F32 += TF32 x TF32, K-major, M128, N%%16\n", idesc);
        abort();
    }
    const int col0 = (int)(tmem_d & 0xFFFFu);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float acc = accumulate ? g_tc_tmem[m][col0 + n] : 0.f;
            for (int k = 0; k < 8; ++k) acc += tc_elem(da, m, k) * tc_elem(db, n, k);
            g_tc_tmem[m][col0 + n] = acc;
        }
}
static inline void tc_commit(TcBarrier* b) { b->phase ^= 1u; }     // the model's MMAs complete at issue
static inline void tc_mbar_wait(TcBarrier* b, uint32_t parity) {
    while (b->phase == parity) cuemu::fiber_yield();               // phase bit == parity: that phase has not completed yet
}
static inline void tc_ld16(uint32_t tmem, int warp, int col, uint32_t* r) {
    const int lane = 32 * warp + cuemu::lane_id(), col0 = (int)(tmem & 0xFFFFu) + col;
    for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(g_tc_tmem[lane][col0 + i]);
}
#endif

__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);               // start address,        bits [0,14)
    d |= (uint64_t)((TC_LBO >> 4) & 0x3FFFu) << 16;        // leading byte offset,  bits [16,30)
    d |= (uint64_t)((TC_SBO >> 4) & 0x3FFFu) << 32;        // stride byte offset,   bits [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version 1 (sm_100)
    return d;                                              // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (bits 61-63 = 0)
}

// rows [row0, row0 + ROWS) x k [k0, k0 + 32) of a row-major fp32 matrix -> hi / lo tiles in the canonical layout
template <int ROWS>
__device__ __forceinline__ void tc_stage(const float* __restrict__ G, int ld, int row0, int nrows, int k0, int K,
                                         unsigned char* s_hi, unsigned char* s_lo) {
    for (int r = threadIdx.x; r < ROWS; r += blockDim.x) {
        const int row = row0 + r;
        const uint32_t base = (uint32_t)(r >> 3) * TC_SBO + (uint32_t)(r & 7) * 16u;
#pragma unroll
        for (int c = 0; c < TC_KB / 4; ++c) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int k = k0 + 4 * c;
            if (row < nrows && k < K) v = *reinterpret_cast<const float4*>(G + (size_t)row * ld + k);   // K % 4 == 0
            uint4 hi, lo;
            hi.x = tc_tf32(v.x); hi.y = tc_tf32(v.y); hi.z = tc_tf32(v.z); hi.w = tc_tf32(v.w);
            lo.x = tc_tf32(v.x - __uint_as_float(hi.x)); lo.y = tc_tf32(v.y - __uint_as_float(hi.y));
            lo.z = tc_tf32(v.z - __uint_as_float(hi.z)); lo.w = tc_tf32(v.w - __uint_as_float(hi.w));
            *reinterpret_cast<uint4*>(s_hi + base + (uint32_t)c * TC_LBO) = hi;
            *reinterpret_cast<uint4*>(s_lo + base + (uint32_t)c * TC_LBO) = lo;
        }
    }
}

template <int NT, int EPI>
__global__ void __launch_bounds__(128) k_sn_gemm_tc(int M, int N, int K, const float* __restrict__ A, const float* __restrict__ Bt,
                                                    const float* __restrict__ bias, float* __restrict__ aux, float* __restrict__ C) {
    TC_DYN_SMEM(tc_smem);
    __shared__ __align__(8) TcBarrier s_bar;
    __shared__ uint32_t s_tmem;
    unsigned char* sA_hi = tc_smem;
    unsigned char* sA_lo = sA_hi + TC_M * 128;
    unsigned char* sB_hi = sA_lo + TC_M * 128;
    unsigned char* sB_lo = sB_hi + NT * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_M, n0 = blockIdx.x * NT;

    if (threadIdx.x == 0) tc_mbar_init(&s_bar);
    if (warp == 0) tc_alloc(&s_tmem, (uint32_t)NT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = NT, M = 128 (cute UMMA::InstrDescriptor)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    uint32_t parity = 0;
    const int nkb = (K + TC_KB - 1) / TC_KB;
    for (int kb = 0; kb < nkb; ++kb) {
        tc_stage<TC_M>(A, K, m0, M, kb * TC_KB, K, sA_hi, sA_lo);
        tc_stage<NT>(Bt, K, n0, N, kb * TC_KB, K, sB_hi, sB_lo);
        tc_fence_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < TC_KB / 8; ++ks) {
                const uint32_t off = (uint32_t)ks * 2u * TC_LBO;         // one MMA consumes two 16-byte K chunks
                const uint64_t dah = tc_desc(tc_smem_addr(sA_hi) + off), dal = tc_desc(tc_smem_addr(sA_lo) + off);
                const uint64_t dbh = tc_desc(tc_smem_addr(sB_hi) + off), dbl = tc_desc(tc_smem_addr(sB_lo) + off);
                tc_mma(tmem, dah, dbh, idesc, (kb | ks) ? 1u : 0u);
                tc_mma(tmem, dah, dbl, idesc, 1u);
                tc_mma(tmem, dal, dbh, idesc, 1u);
            }
            tc_commit(&s_bar);
        }
        tc_mbar_wait(&s_bar, parity);     // the MMAs have read the staged tiles (and, after the last block, written D)
        parity ^= 1u;
        tc_fence_after();
    }

    // epilogue: warp w owns TMEM lanes [32 w, 32 w + 32) = rows m0 + 32 w + lane
    const int m = m0 + 32 * warp + lane;
    for (int c0 = 0; c0 < NT; c0 += 16) {
        uint32_t r[16];
        tc_ld16(tmem, warp, c0, r);
        if (m < M) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int n = n0 + c0 + i;
                if (n >= N) continue;
                const size_t o = (size_t)m * N + n;
                const float v = __uint_as_float(r[i]);
                if (EPI == SN_EPI_STORE) C[o] = v;
                else if (EPI == SN_EPI_BIAS) C[o] = v + bias[n];
                else if (EPI == SN_EPI_BIAS_SSP) { float p = v + bias[n]; aux[o] = p; C[o] = sn_ssp(p); }
                else if (EPI == SN_EPI_BIAS_ADD) C[o] += v + bias[n];
                else if (EPI == SN_EPI_MUL_SIG) C[o] = v * sn_sigmoid(aux[o]);
                else C[o] += v;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem, (uint32_t)NT);
}

// =====================================================================================================================
// Second version (round 2, after the first hardware runs of the kernel above - profiles/r02_schnet.md):
//   * staging is COALESCED: 8 consecutive threads read the 128 contiguous bytes of one row's 32-wide k-block (the first
//     version gave every thread a whole row: 32 different 128-byte lines per warp request, the L1 wavefronts of that made a
//     k-block ~5x longer than its 12 MMAs);
//   * two shared-memory stages: the threads stage k-block kb + 1 while the tensor core works on kb (one mbarrier per stage,
//     armed by tcgen05.commit, waited before the stage is overwritten);
//   * the accumulation over K is CHUNKED: the tensor core accumulates at most TC_KCH k-blocks (128 k) into one half of a
//     double-buffered TMEM accumulator, the warps add finished chunks into fp32 registers with round-to-nearest adds while the
//     next chunk runs.  Measured reason: with one long TMEM accumulation the K = 512 layers of configs[4] miss the 1e-5 parity
//     bar against the reference (the tensor core's fp32 accumulator does not round to nearest: the error grows with the number
//     of accumulations), K = 128 passes;
//   * the epilogue reads / writes rows in float4.
// Same operand layout, descriptors and 3xTF32 split as the first version (validated on hardware).
// =====================================================================================================================
#define TC_KCH 4            // k-blocks per TMEM accumulation chunk
#define TC_STAGES 2

// Staging of one 32-wide k-block in two halves: tc_load2 requests the rows' float4 (registers), tc_store2 splits them into
// hi / lo TF32 and stores them in the canonical layout.  Thread mapping (128 threads, warp w, lane l): a warp instruction
// covers row group g = 4 * it' + w (8 rows), rows r = 8 g + (l & 7), float4 column c = 4 * half + (l >> 3):
//   * shared memory: the 8 lanes of a quarter-warp write the 8 rows of a core matrix = 8 consecutive 16-byte slots ->
//     conflict-free (the first coalesced mapping put a row's 8 float4 at stride LBO = 128 B: 8-way bank conflicts, 3084
//     instead of 384 store wavefronts per k-block - ncu, profiles/r02_schnet.md);
//   * global memory: a warp reads 64 contiguous bytes of each of 8 rows; the other half of every 128-byte line is read by the
//     next instruction of the same thread (L1 hit).
template <int ROWS>
__device__ __forceinline__ void tc_load2(const float* __restrict__ G, int ld, int row0, int nrows, int k0, int K, float4* v) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int it = 0; it < ROWS / 16; ++it) {
        const int g = (it >> 1) * 4 + w, half = it & 1;
        const int r = 8 * g + (lane & 7), c = 4 * half + (lane >> 3);
        const int row = row0 + r, k = k0 + 4 * c;
        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < nrows && k < K) v[it] = __ldg(reinterpret_cast<const float4*>(G + (size_t)row * ld + k));   // K % 4 == 0
    }
}
template <int ROWS>
__device__ __forceinline__ void tc_store2(const float4* v, unsigned char* s_hi, unsigned char* s_lo) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int it = 0; it < ROWS / 16; ++it) {
        const int g = (it >> 1) * 4 + w, half = it & 1;
        const int c = 4 * half + (lane >> 3);
        const uint32_t base = (uint32_t)g * TC_SBO + (uint32_t)(lane & 7) * 16u + (uint32_t)c * TC_LBO;
        uint4 hi, lo;
        hi.x = tc_tf32(v[it].x); hi.y = tc_tf32(v[it].y); hi.z = tc_tf32(v[it].z); hi.w = tc_tf32(v[it].w);
        lo.x = tc_tf32(v[it].x - __uint_as_float(hi.x)); lo.y = tc_tf32(v[it].y - __uint_as_float(hi.y));
        lo.z = tc_tf32(v[it].z - __uint_as_float(hi.z)); lo.w = tc_tf32(v[it].w - __uint_as_float(hi.w));
        *reinterpret_cast<uint4*>(s_hi + base) = hi;
        *reinterpret_cast<uint4*>(s_lo + base) = lo;
    }
}

template <int NT, int EPI>
__global__ void __launch_bounds__(128) k_sn_gemm_tc2(int M, int N, int K, const float* __restrict__ A, const float* __restrict__ Bt,
                                                     const float* __restrict__ bias, float* __restrict__ aux, float* __restrict__ C) {
    TC_DYN_SMEM(tc_smem);
    __shared__ __align__(8) TcBarrier s_done[TC_STAGES];      // the MMAs that read stage s have completed
    __shared__ __align__(8) TcBarrier s_chunk[2];             // every MMA of the chunk accumulated in TMEM half h has completed
    __shared__ uint32_t s_tmem;
    constexpr uint32_t STAGE = (uint32_t)(2 * TC_M + 2 * NT) * 128u;
    const int warp = threadIdx.x >> 5;
    const int m0 = blockIdx.y * TC_M, n0 = blockIdx.x * NT;

    if (threadIdx.x == 0) { tc_mbar_init(&s_done[0]); tc_mbar_init(&s_done[1]); tc_mbar_init(&s_chunk[0]); tc_mbar_init(&s_chunk[1]); }
    if (warp == 0) tc_alloc(&s_tmem, (uint32_t)(2 * NT));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

    float racc[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) racc[i] = 0.f;
    uint32_t done_ph = 0, chunk_ph = 0;                       // bit k = parity of the next wait on barrier k
    const int nkb = (K + TC_KB - 1) / TC_KB;
    const int nchunks = (nkb + TC_KCH - 1) / TC_KCH;
    int drained = 0;

    auto drain = [&](int c) {                                 // add the finished chunk c (TMEM half c & 1) into the registers
        const int h = c & 1;
        tc_mbar_wait(&s_chunk[h], (chunk_ph >> h) & 1u);
        chunk_ph ^= 1u << h;
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < NT; c0 += 16) {
            uint32_t r[16];
            tc_ld16(tmem + (uint32_t)(h * NT), warp, c0, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) racc[c0 + i] += __uint_as_float(r[i]);
        }
        tc_fence_before();
    };

    float4 va[TC_M / 16], vb[NT / 16];                       // the k-block being staged (requested one iteration ahead)
    tc_load2<TC_M>(A, K, m0, M, 0, K, va);
    tc_load2<NT>(Bt, K, n0, N, 0, K, vb);
    for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        if (kb >= TC_STAGES) {                                // the MMAs of k-block kb - 2 have read this stage
            tc_mbar_wait(&s_done[s], (done_ph >> s) & 1u);
            done_ph ^= 1u << s;
        }
        unsigned char* sA_hi = tc_smem + (size_t)s * STAGE;
        unsigned char* sA_lo = sA_hi + TC_M * 128;
        unsigned char* sB_hi = sA_lo + TC_M * 128;
        unsigned char* sB_lo = sB_hi + NT * 128;
        tc_store2<TC_M>(va, sA_hi, sA_lo);
        tc_store2<NT>(vb, sB_hi, sB_lo);
        if (kb + 1 < nkb) {                                   // next k-block's rows in flight behind this block's barrier + MMAs
            tc_load2<TC_M>(A, K, m0, M, (kb + 1) * TC_KB, K, va);
            tc_load2<NT>(Bt, K, n0, N, (kb + 1) * TC_KB, K, vb);
        }
        tc_fence_smem();
        // a chunk that ended two k-blocks ago has (almost certainly) completed: fold it in while the tensor core is busy
        if (kb % TC_KCH == 1 && kb / TC_KCH >= 1 && drained < kb / TC_KCH) { drain(drained); ++drained; }
        __syncthreads();
        if (threadIdx.x == 0) {
            tc_fence_after();
            const int chunk = kb / TC_KCH, h = chunk & 1;
            const uint32_t d = tmem + (uint32_t)(h * NT);
#pragma unroll
            for (int ks = 0; ks < TC_KB / 8; ++ks) {
                const uint32_t off = (uint32_t)ks * 2u * TC_LBO;         // one MMA consumes two 16-byte K chunks
                const uint64_t dah = tc_desc(tc_smem_addr(sA_hi) + off), dal = tc_desc(tc_smem_addr(sA_lo) + off);
                const uint64_t dbh = tc_desc(tc_smem_addr(sB_hi) + off), dbl = tc_desc(tc_smem_addr(sB_lo) + off);
                tc_mma(d, dah, dbh, idesc, ((kb % TC_KCH) | ks) ? 1u : 0u);
                tc_mma(d, dah, dbl, idesc, 1u);
                tc_mma(d, dal, dbh, idesc, 1u);
            }
            tc_commit(&s_done[s]);
            if (kb % TC_KCH == TC_KCH - 1 || kb == nkb - 1) tc_commit(&s_chunk[h]);
        }
    }
    for (; drained < nchunks; ++drained) drain(drained);

    // epilogue: warp w owns TMEM lanes [32 w, 32 w + 32) = rows m0 + 32 w + lane; this thread holds its row's NT sums
    const int m = m0 + 32 * warp + (threadIdx.x & 31);
    if (m < M) {
#pragma unroll
        for (int c0 = 0; c0 < NT; c0 += 4) {
            const int n = n0 + c0;
            if (n >= N) break;
            const size_t o = (size_t)m * N + n;
            if (n + 3 < N && ((N & 3) == 0)) {
                float4 v = make_float4(racc[c0], racc[c0 + 1], racc[c0 + 2], racc[c0 + 3]);
                if (EPI == SN_EPI_BIAS || EPI == SN_EPI_BIAS_SSP || EPI == SN_EPI_BIAS_ADD) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
                    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                }
                if (EPI == SN_EPI_BIAS_SSP) {
                    *reinterpret_cast<float4*>(aux + o) = v;
                    v = make_float4(sn_ssp(v.x), sn_ssp(v.y), sn_ssp(v.z), sn_ssp(v.w));
                } else if (EPI == SN_EPI_BIAS_ADD || EPI == SN_EPI_ADD) {
                    const float4 c = *reinterpret_cast<const float4*>(C + o);
                    v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
                } else if (EPI == SN_EPI_MUL_SIG) {
                    const float4 a = *reinterpret_cast<const float4*>(aux + o);
                    v.x *= sn_sigmoid(a.x); v.y *= sn_sigmoid(a.y); v.z *= sn_sigmoid(a.z); v.w *= sn_sigmoid(a.w);
                }
                *reinterpret_cast<float4*>(C + o) = v;
            } else {
                for (int i = 0; i < 4 && n + i < N; ++i) {
                    const float v = racc[c0 + i];
                    const size_t oi = o + i;
                    if (EPI == SN_EPI_STORE) C[oi] = v;
                    else if (EPI == SN_EPI_BIAS) C[oi] = v + bias[n + i];
                    else if (EPI == SN_EPI_BIAS_SSP) { float p = v + bias[n + i]; aux[oi] = p; C[oi] = sn_ssp(p); }
                    else if (EPI == SN_EPI_BIAS_ADD) C[oi] += v + bias[n + i];
                    else if (EPI == SN_EPI_MUL_SIG) C[oi] = v * sn_sigmoid(aux[oi]);
                    else C[oi] += v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem, (uint32_t)(2 * NT));
}

// Bt: (N x K) row-major.  Returns MDG_E_STATE when the shape is not covered (caller falls back to the SIMT kernel).
template <int EPI>
static int sn_gemm_tc(int M, int N, int K, const float* A, const float* Bt, const float* bias, float* aux, float* C, cudaStream_t st) {
    if (M <= 0 || N <= 0) return MDG_OK;
    if ((K & 3) || (((uintptr_t)A | (uintptr_t)Bt) & 15)) return MDG_E_STATE;
    // 128-wide tiles only when they still give >= 2 CTAs per SM: the kernel is single-staged (stage -> MMA -> wait), so it is
    // co-resident CTAs that overlap one tile's staging with another's MMAs; at configs[4] (M 4096, N 256 / 512) the 64-wide
    // tile gives 128 - 256 CTAs of 48 KB instead of 64 - 128 of 64 KB
    const bool wide = N > 64 && (int64_t)((N + 127) / 128) * ((M + TC_M - 1) / TC_M) >= 2 * 148;
    if (getenv("MDG_SCHNET_TC_V1") == nullptr) {            // second version: coalesced 2-stage pipeline, chunked accumulation
        const size_t smem = (size_t)TC_STAGES * (2 * TC_M + 2 * 64) * 128;
        MDG_CUDA(cudaFuncSetAttribute(k_sn_gemm_tc2<64, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((N + 63) / 64, (M + TC_M - 1) / TC_M);
        k_sn_gemm_tc2<64, EPI><<<grid, 128, smem, st>>>(M, N, K, A, Bt, bias, aux, C);
        MDG_KERNEL_CHECK();
        return MDG_OK;
    }
    if (wide) {
        const size_t smem = (size_t)(2 * TC_M + 2 * 128) * 128;
        MDG_CUDA(cudaFuncSetAttribute(k_sn_gemm_tc<128, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((N + 127) / 128, (M + TC_M - 1) / TC_M);
        k_sn_gemm_tc<128, EPI><<<grid, 128, smem, st>>>(M, N, K, A, Bt, bias, aux, C);
    } else {
        const size_t smem = (size_t)(2 * TC_M + 2 * 64) * 128;
        MDG_CUDA(cudaFuncSetAttribute(k_sn_gemm_tc<64, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((N + 63) / 64, (M + TC_M - 1) / TC_M);
        k_sn_gemm_tc<64, EPI><<<grid, 128, smem, st>>>(M, N, K, A, Bt, bias, aux, C);
    }
    MDG_KERNEL_CHECK();
    return MDG_OK;
}
