// schnet.cu - native SchNet energy + forces (K5 in SURVEY.md 2c, rows a12-a14 of section 8).
//
// One call evaluates the reference network  (nff/nn/models/schnet.py:113-171 SchNet.convolve/forward,
// nff/nn/modules.py:514-575 SchNetConv, nff/nn/graphconv.py:32-53 MessagePassingModule, nff/nn/layers.py:14-31
// gaussian smearing, :86-134 Dense, nff/nn/activations.py:5-11 shifted_softplus, nff/nn/modules.py:761-809 readout,
// nff/nn/graphop.py:9-30 atom sum)  AND its analytic first derivative w.r.t. the positions (the autograd force of
// torchmd/md.py:227-228), as a fixed program of kernels with no autograd tape:
//
//   forward   d_e = |x_i - x_j - off_e|                                     k_pair_dis (force.cu)
//             per layer:  W_e  = Dense(ssp(Dense(gauss(d_e))))   (E x F)    k_sn_edge_fwd   (fused; weights broadcast
//                         h    = Dense(r)                        (N x F)    k_sn_gemm        from L1, activations from smem)
//                         agg  = sum_{e at k} h[other] * W_e     (N x F)    k_cfconv_agg    (graph.cu, atomics-free CSR)
//                         r   += Dense(ssp(Dense(agg)))          (N x A)    k_sn_gemm x2
//             readout     E    = sum_n Dense(ssp(Dense(r_n)))               k_sn_gemm + k_sn_readout
//   backward  the same chain transposed; the edge part is ONE fused kernel per layer that never materialises dE/dW_e:
//             gW_e = h_i*g_j + h_j*g_i  ->  (x We2) * ssp'  ->  (x We1)  ->  . dgauss/dd  ->  gd_e += ...   k_sn_edge_bwd
//             F_k  = - sum_{e at k} +-gd_e * rvec_e / d_e   (CSR, deterministic, no atomics)               k_sn_edge_force
//
// fp32 throughout (reference dtype), plain FMA arithmetic: parity with the reference is a tolerance (1e-5), not bits.
// The dense node-side layers use a shared-memory tiled SIMT GEMM here; they are the part the tcgen05 path replaces
// (DESIGN.md section 7).  Weight gradients are not produced: training goes through the autograd route of the mirror.
#include <stdlib.h>
#include "common.cuh"

int mdg_i_graph_build(mdg_ctx* c, const int64_t* d_nbr, int64_t n_edges, int n, cudaStream_t st, const int* d_n_edges = nullptr);
int mdg_i_cfconv_agg(mdg_ctx* c, const float* d_h, const float* d_W, int n, int F, float* d_out, cudaStream_t st);

#define SN_GMAX 64          // max gaussians (29 / 33 in the reference configs)
#define SN_TE 32            // edges per CTA in the fused edge-filter forward
#define SN_TEB 16           // edges per CTA in the fused edge backward

__device__ __forceinline__ float sn_ssp(float x) {          // softplus(x) - ln 2, torch threshold 20
    float sp = x > 20.0f ? x : log1pf(expf(x));
    return sp - 0.69314718055994531f;
}
__device__ __forceinline__ float sn_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }   // d ssp / dx

// ---------------------------------------------------------------------------------------------
// small elementwise kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_sn_embed(int n, int A, const int64_t* __restrict__ z, const float* __restrict__ embed, float* __restrict__ r) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * A) return;
    int i = (int)(idx / A), a = (int)(idx - (int64_t)i * A);
    r[idx] = embed[z[i] * A + a];
}

// W (out x in, torch Linear layout) -> WT (in x out)
__global__ void k_sn_transpose(int rows, int cols, const float* __restrict__ W, float* __restrict__ WT) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    int r = idx / cols, c = idx - r * cols;
    WT[(size_t)c * rows + r] = W[idx];
}

// ---------------------------------------------------------------------------------------------
// fused edge filter, forward:  W[e][f] = be2[f] + sum_k ssp(be1[k] + sum_k' g[e][k'] We1[k][k']) We2[f][k]
// One CTA = SN_TE edges.  Activations live TRANSPOSED in shared memory ([k][edge]) so that one LDS.128 feeds four
// FMAs with the same weight; weights come pre-transposed ([in][out]) so that the lanes of a warp (consecutive output
// index) read one coalesced line per k.
// ---------------------------------------------------------------------------------------------
#ifdef SN_EDGE_MINBLOCKS          // build variant sne3: 3 CTAs per SM (77 registers, no spills) for the fused filter generator
#define SN_EDGE_BOUNDS __launch_bounds__(256, SN_EDGE_MINBLOCKS)
#else                            // default: 100 registers, 2 CTAs per SM
#define SN_EDGE_BOUNDS __launch_bounds__(256)
#endif
__global__ void SN_EDGE_BOUNDS k_sn_edge_fwd(int64_t E, const int* __restrict__ dE, int G, int F, const float* __restrict__ dis,
                                                     const float* __restrict__ mu, const float* __restrict__ width,
                                                     const float* __restrict__ We1T, const float* __restrict__ be1,
                                                     const float* __restrict__ We2T, const float* __restrict__ be2,
                                                     float* __restrict__ preT1, float* __restrict__ W) {
    __shared__ __align__(16) float s_g[SN_GMAX][SN_TE];
    __shared__ __align__(16) float s_a[SN_GMAX][SN_TE];
    if (dE) E = min(E, (int64_t)*dE);          // asynchronous engine steps: E is the launch bound, the count is on the device
    const int64_t e0 = (int64_t)blockIdx.x * SN_TE;
    if (e0 >= E) return;                       // (block-uniform)
    const int t = threadIdx.x;
    for (int idx = t; idx < G * SN_TE; idx += blockDim.x) {
        int k = idx / SN_TE, e = idx - k * SN_TE;
        float v = 0.f;
        if (e0 + e < E) {
            float w = width[k];
            float diff = dis[e0 + e] - mu[k];
            v = expf((-0.5f / (w * w)) * (diff * diff));
        }
        s_g[k][e] = v;
    }
    __syncthreads();
    // phase 1: thread = (k, group of 8 edges)
    {
        const int k = t & 63, eg = t >> 6;         // 4 groups x 8 edges = SN_TE
        if (k < G) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            for (int kk = 0; kk < G; ++kk) {
                float w = We1T[kk * G + k];
                float4 a = *reinterpret_cast<const float4*>(&s_g[kk][eg * 8]);
                float4 b = *reinterpret_cast<const float4*>(&s_g[kk][eg * 8 + 4]);
                acc[0] += a.x * w; acc[1] += a.y * w; acc[2] += a.z * w; acc[3] += a.w * w;
                acc[4] += b.x * w; acc[5] += b.y * w; acc[6] += b.z * w; acc[7] += b.w * w;
            }
            float b1 = be1[k];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int e = eg * 8 + j;
                float pre = acc[j] + b1;
                if (e0 + e < E) preT1[(e0 + e) * G + k] = pre;
                s_a[k][e] = sn_ssp(pre);
            }
        }
    }
    __syncthreads();
    // phase 2: thread = output filter f (strided by the block size), all SN_TE edges in registers
    for (int f = t; f < F; f += blockDim.x) {
        float acc[SN_TE];
#pragma unroll
        for (int j = 0; j < SN_TE; ++j) acc[j] = 0.f;
        for (int k = 0; k < G; ++k) {
            float w = We2T[(size_t)k * F + f];
#pragma unroll
            for (int j4 = 0; j4 < SN_TE / 4; ++j4) {
                float4 a = *reinterpret_cast<const float4*>(&s_a[k][j4 * 4]);
                acc[j4 * 4 + 0] += a.x * w; acc[j4 * 4 + 1] += a.y * w; acc[j4 * 4 + 2] += a.z * w; acc[j4 * 4 + 3] += a.w * w;
            }
        }
        float b2 = be2[f];
#pragma unroll
        for (int j = 0; j < SN_TE; ++j)
            if (e0 + j < E) W[(e0 + j) * F + f] = acc[j] + b2;
    }
}

// ---------------------------------------------------------------------------------------------
// fused edge filter, backward (per layer): accumulates dE/dd_e into gd[e].
//   gW[e][f]  = h[i][f] g[j][f] + h[j][f] g[i][f]                (g = dE/dagg; never written to HBM)
//   gT[e][k]  = sigmoid(preT1[e][k]) * sum_f gW[e][f] We2[f][k]
//   gg[e][k'] = sum_k gT[e][k] We1[k][k']
//   gd[e]    += sum_k' gg[e][k'] * gauss_k'(d) * 2 coeff_k' (d - mu_k')
// dynamic shared memory: F x SN_TEB floats (gW transposed).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sn_edge_bwd(int64_t E, const int* __restrict__ dE, int G, int F, const int64_t* __restrict__ nbr,
                                                     const float* __restrict__ h, const float* __restrict__ g,
                                                     const float* __restrict__ dis, const float* __restrict__ mu,
                                                     const float* __restrict__ width, const float* __restrict__ We1,
                                                     const float* __restrict__ We2, const float* __restrict__ preT1,
                                                     float* __restrict__ gd) {
    extern __shared__ float s_gw[];                               // [F][SN_TEB]
    __shared__ __align__(16) float s_gt[SN_GMAX][SN_TEB];         // [k][edge]
    __shared__ float s_out[SN_TEB];
    __shared__ int s_i[SN_TEB], s_j[SN_TEB];
    if (dE) E = min(E, (int64_t)*dE);
    const int64_t e0 = (int64_t)blockIdx.x * SN_TEB;
    if (e0 >= E) return;                       // (block-uniform)
    const int t = threadIdx.x;
    if (t < SN_TEB) {
        bool ok = e0 + t < E;
        s_i[t] = ok ? (int)nbr[2 * (e0 + t)] : -1;
        s_j[t] = ok ? (int)nbr[2 * (e0 + t) + 1] : -1;
        s_out[t] = 0.f;
    }
    __syncthreads();
    for (int f = t; f < F; f += blockDim.x) {
#pragma unroll
        for (int e = 0; e < SN_TEB; ++e) {
            int i = s_i[e], j = s_j[e];
            float v = 0.f;
            if (i >= 0) v = h[(size_t)i * F + f] * g[(size_t)j * F + f] + h[(size_t)j * F + f] * g[(size_t)i * F + f];
            s_gw[f * SN_TEB + e] = v;
        }
    }
    __syncthreads();
    const int k = t & 63, eg = t >> 6;             // thread = (k, group of 4 edges)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (k < G) {
        for (int f = 0; f < F; ++f) {
            float w = We2[(size_t)f * G + k];
            float4 a = *reinterpret_cast<const float4*>(&s_gw[f * SN_TEB + eg * 4]);
            acc[0] += a.x * w; acc[1] += a.y * w; acc[2] += a.z * w; acc[3] += a.w * w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int e = eg * 4 + j;
            float pre = (e0 + e < E) ? preT1[(e0 + e) * G + k] : 0.f;
            s_gt[k][e] = acc[j] * sn_sigmoid(pre);
        }
    }
    __syncthreads();
    float part[4] = {0.f, 0.f, 0.f, 0.f};
    if (k < G) {                                    // here k plays k' (input gaussian index)
        float gg[4] = {0.f, 0.f, 0.f, 0.f};
        for (int kk = 0; kk < G; ++kk) {
            float w = We1[kk * G + k];
            float4 a = *reinterpret_cast<const float4*>(&s_gt[kk][eg * 4]);
            gg[0] += a.x * w; gg[1] += a.y * w; gg[2] += a.z * w; gg[3] += a.w * w;
        }
        float wk = width[k], m = mu[k];
        float coeff = -0.5f / (wk * wk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int e = eg * 4 + j;
            if (e0 + e < E) {
                float diff = dis[e0 + e] - m;
                float gauss = expf(coeff * (diff * diff));
                part[j] = gg[j] * gauss * (2.0f * coeff * diff);
            }
        }
    }
    // reduce over k' (the 64 threads sharing eg = two warps) and accumulate per edge
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v = part[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((t & 31) == 0) atomicAdd(&s_out[eg * 4 + j], v);
    }
    __syncthreads();
    if (t < SN_TEB && e0 + t < E) gd[e0 + t] += s_out[t];
}

// F_k = - sum over edges e incident to k of  s * gd[e] * rvec_e / d_e,  s = +1 if k is the first atom of e, else -1;
// rvec_e = x_i - x_j - off_e * scale  (the reference's raw-offset quirk = scale (1,1,1))
// One WARP per atom: lanes stride the atom's incident edges, fixed-tree shuffle reduction (deterministic).  (A thread per atom
// walked ~83 dependent gathers serially on the water box: 154 us.)
__global__ void __launch_bounds__(128) k_sn_edge_force(int n, const int* __restrict__ off, const int* __restrict__ inc_edge,
                                const int64_t* __restrict__ nbr, const float* __restrict__ offsets, float sx, float sy, float sz,
                                const float* __restrict__ xyz, const float* __restrict__ dis, const float* __restrict__ gd,
                                float* __restrict__ force) {
    const int lane = threadIdx.x & 31;
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n) return;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    for (int p = off[k] + lane; p < off[k + 1]; p += 32) {
        int e = inc_edge[p];
        int64_t i = nbr[2 * (int64_t)e], j = nbr[2 * (int64_t)e + 1];
        float rx = (xyz[3 * i] - xyz[3 * j]) - offsets[3 * (int64_t)e] * sx;
        float ry = (xyz[3 * i + 1] - xyz[3 * j + 1]) - offsets[3 * (int64_t)e + 1] * sy;
        float rz = (xyz[3 * i + 2] - xyz[3 * j + 2]) - offsets[3 * (int64_t)e + 2] * sz;
        float d = dis[e];
        float w = d > 0.f ? gd[e] / d : 0.f;
        if (i != k) w = -w;
        fx -= w * rx; fy -= w * ry; fz -= w * rz;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, o);
        fy += __shfl_xor_sync(0xffffffffu, fy, o);
        fz += __shfl_xor_sync(0xffffffffu, fz, o);
    }
    if (lane == 0) { force[3 * k] = fx; force[3 * k + 1] = fy; force[3 * k + 2] = fz; }
}

// ---------------------------------------------------------------------------------------------
// dense layers: C (M x N) = epilogue( A (M x K, row-major) * B ),  B(k,n) = TRANSB ? Wt[n*ldb + k] : Wt[k*ldb + n]
// 64 x 64 x 16 tiles, 256 threads, 4 x 4 outputs per thread.
// ---------------------------------------------------------------------------------------------
#define SN_EPI_STORE 0      // C = acc
#define SN_EPI_BIAS 1       // C = acc + bias[n]
#define SN_EPI_BIAS_SSP 2   // aux = acc + bias[n] ; C = ssp(aux)
#define SN_EPI_BIAS_ADD 3   // C += acc + bias[n]
#define SN_EPI_MUL_SIG 4    // C = acc * sigmoid(aux)
#define SN_EPI_ADD 5        // C += acc
#define GT 64
#define GK 16

template <bool TRANSB, int EPI>
__global__ void __launch_bounds__(256) k_sn_gemm(int M, int N, int K, const float* __restrict__ A, const float* __restrict__ B,
                                                 int ldb, const float* __restrict__ bias, float* __restrict__ aux,
                                                 float* __restrict__ C) {
    __shared__ __align__(16) float As[GK][GT + 4];
    __shared__ __align__(16) float Bs[GK][GT + 4];
    const int t = threadIdx.x;
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    const int tm = (t >> 4) * 4, tn = (t & 15) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += GK) {
        // A tile: 64 rows x 16 k -> As[k][m]; thread loads 4 elements
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int idx = t + q * 256;               // 0..1023
            int m = idx >> 4, k = idx & 15;
            float v = 0.f;
            if (m0 + m < M && k0 + k < K) v = A[(size_t)(m0 + m) * K + k0 + k];
            As[k][m] = v;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int idx = t + q * 256;
            float v = 0.f;
            if (TRANSB) {
                int n = idx >> 4, k = idx & 15;
                if (n0 + n < N && k0 + k < K) v = B[(size_t)(n0 + n) * ldb + k0 + k];
                Bs[k][n] = v;
            } else {
                int k = idx >> 6, n = idx & 63;
                if (n0 + n < N && k0 + k < K) v = B[(size_t)(k0 + k) * ldb + n0 + n];
                Bs[k][n] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&As[k][tm]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + tm + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tn + j;
            if (n >= N) continue;
            size_t o = (size_t)m * N + n;
            float v = acc[i][j];
            if (EPI == SN_EPI_STORE) C[o] = v;
            else if (EPI == SN_EPI_BIAS) C[o] = v + bias[n];
            else if (EPI == SN_EPI_BIAS_SSP) { float p = v + bias[n]; aux[o] = p; C[o] = sn_ssp(p); }
            else if (EPI == SN_EPI_BIAS_ADD) C[o] += v + bias[n];
            else if (EPI == SN_EPI_MUL_SIG) C[o] = v * sn_sigmoid(aux[o]);
            else C[o] += v;
        }
    }
}

#include "schnet_tc.cuh"

// Dense-layer route.  MDG_SCHNET_TC=1 forces the tcgen05 kernel (schnet_tc.cuh), =0 forces the SIMT kernel; unset = AUTO:
// tensor cores for the wide / tall layers (M >= 1024 rows and K, N >= 128: the configs[4] shapes, 4096 x 512 x 256), SIMT for
// small systems, where one 128-row tile per CTA leaves the GPU empty (measured on the 192-atom water box: 481 vs 574 steps/s).
// First hardware run of the tensor-core path: round 2 (profiles/r02_schnet.md) - equal to the reference at 1e-5 on the
// configured-width fixtures (tests/test_schnet.py::test_native_schnet_configured_widths_vs_reference_fixture).
static int sn_tc_mode() {          // 1 = always, 0 = never, 2 = auto  (read per call: tests switch inside one process)
    const char* e = getenv("MDG_SCHNET_TC");
    return !e ? 2 : (e[0] == '1' ? 1 : (e[0] == '0' ? 0 : 2));
}
static bool sn_tc_enabled(int M, int N, int K) {
    const int mode = sn_tc_mode();
    if (mode != 2) return mode == 1;
#ifdef MDG_EMU
    return false;                  // (the functional tensor-core model is exercised explicitly by test_emu_tc_*)
#else
    return M >= 1024 && N >= 128 && K >= 128;
#endif
}

template <bool TRANSB, int EPI>
static int sn_gemm(mdg_ctx* c, int M, int N, int K, const float* A, const float* B, int ldb, const float* bias, float* aux, float* C,
                   cudaStream_t st) {
    if (M <= 0 || N <= 0) return MDG_OK;
    if (sn_tc_enabled(M, N, K)) {
        const float* Bt = B;                     // the tensor-core kernel wants B as (N x K) row-major = K-major
        if (!TRANSB) {                           // backward layers multiply by W (K x N): transpose the (small) weight first
            MDG_TRY(c->sn_wt.reserve(sizeof(float) * (size_t)K * N));
            k_sn_transpose<<<(K * N + 255) / 256, 256, 0, st>>>(K, N, B, c->sn_wt.as<float>());
            Bt = c->sn_wt.as<float>();
        }
        if ((TRANSB ? ldb == K : ldb == N)) {
            int r = sn_gemm_tc<EPI>(M, N, K, A, Bt, bias, aux, C, st);
            if (r != MDG_E_STATE) return r;
        }
    }
    dim3 grid((N + GT - 1) / GT, (M + GT - 1) / GT);
    k_sn_gemm<TRANSB, EPI><<<grid, 256, 0, st>>>(M, N, K, A, B, ldb, bias, aux, C);
    MDG_KERNEL_CHECK();
    return MDG_OK;
}

// readout tail: e_n = br2 + sum_j y[n][j] Wr2[j];  block partial sums in double;  gy[n][j] = Wr2[j] * sigmoid(preY[n][j])
__global__ void __launch_bounds__(256) k_sn_readout(int n, int R, const float* __restrict__ y, const float* __restrict__ preY,
                                                    const float* __restrict__ Wr2, const float* __restrict__ br2,
                                                    float* __restrict__ gy, double* __restrict__ part) {
    __shared__ double sm[8];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc = 0;
    for (int i = blockIdx.x * 8 + w; i < n; i += gridDim.x * 8) {
        float s = 0.f;
        for (int j = lane; j < R; j += 32) {
            float wj = Wr2[j];
            s += y[(size_t)i * R + j] * wj;
            gy[(size_t)i * R + j] = wj * sn_sigmoid(preY[(size_t)i * R + j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) acc += (double)(s + br2[0]);
    }
    if (lane == 0) sm[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tsum = 0;
        for (int k = 0; k < 8; ++k) tsum += sm[k];
        part[blockIdx.x] = tsum;
    }
}

__global__ void k_sn_energy_final(int np, const double* __restrict__ part, float* __restrict__ energy) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double tsum = 0;
        for (int k = 0; k < np; ++k) tsum += part[k];
        *energy = (float)tsum;
    }
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
static inline size_t al(size_t x) { return (x + 63) & ~(size_t)63; }   // floats, 256-byte granules

int mdg_i_pair_dis_fwd(const float* d_xyz, const int64_t* d_nbr, const float* d_offsets, int64_t n_pairs, const int* d_n_pairs,
                       const float* h_cell3, float* d_dis, cudaStream_t st);
int mdg_i_schnet_energy_force(mdg_ctx* c, const mdg_schnet_model* m, const int64_t* d_z, const float* d_xyz, int n,
                              const int64_t* d_nbr, const float* d_offsets, int64_t E, const int* d_E, const float* h_off_scale3,
                              float* d_energy, float* d_force, void* stream);

extern "C" int mdg_schnet_energy_force(mdg_ctx* c, const mdg_schnet_model* m, const int64_t* d_z, const float* d_xyz, int n,
                                       const int64_t* d_nbr, const float* d_offsets, int64_t E, const float* h_off_scale3,
                                       float* d_energy, float* d_force, void* stream) {
    return mdg_i_schnet_energy_force(c, m, d_z, d_xyz, n, d_nbr, d_offsets, E, nullptr, h_off_scale3, d_energy, d_force, stream);
}

// d_E != nullptr (asynchronous engine steps, engine.cu): E is the CAPACITY of the edge buffers / the launch bound, the
// actual edge count is read on the device (flags[4] of the list build); nothing here depends on it on the host.
int mdg_i_schnet_energy_force(mdg_ctx* c, const mdg_schnet_model* m, const int64_t* d_z, const float* d_xyz, int n,
                              const int64_t* d_nbr, const float* d_offsets, int64_t E, const int* d_E, const float* h_off_scale3,
                              float* d_energy, float* d_force, void* stream) {
    if (!c || !m || !d_energy || (n > 0 && (!d_z || !d_xyz)) || (E > 0 && (!d_nbr || !d_offsets)) || !h_off_scale3) {
        mdg_set_error("mdg_schnet_energy_force: null argument");
        return MDG_E_BADARG;
    }
    const int A = m->n_atom_basis, F = m->n_filters, G = m->n_gaussians, L = m->n_convolutions, R = m->n_readout;
    if (A <= 0 || F <= 0 || G <= 0 || G > SN_GMAX || L <= 0 || L > MDG_SCHNET_MAX_LAYERS || R <= 0 || n < 0 || E < 0) {
        mdg_set_error("mdg_schnet_energy_force: unsupported sizes A=%d F=%d G=%d (max %d) L=%d R=%d", A, F, G, SN_GMAX, L, R);
        return MDG_E_BADARG;
    }
    MDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { MDG_CUDA(cudaMemsetAsync(d_energy, 0, sizeof(float), st)); return MDG_OK; }
    MDG_TRY(mdg_i_graph_build(c, d_nbr, E, n, st, d_E));

    // workspace carve-up (floats)
    const size_t sE = al((size_t)E), sEG = al((size_t)E * G), sEF = al((size_t)E * F);
    const size_t sNA = al((size_t)n * A), sNF = al((size_t)n * F), sNR = al((size_t)n * R);
    const size_t sWt = 0;       // transposed filter weights live in their own cached buffer (sn_wcache)
    size_t total = 2 * sE + (size_t)L * (sEG + sEF + sNF + sNA + sWt) + 4 * sNA + 3 * sNF + 3 * sNR + 1024;
    MDG_TRY(c->sn_ws.reserve(total * sizeof(float)));
    float* p = c->sn_ws.as<float>();
    auto take = [&](size_t cnt) { float* q = p; p += cnt; return q; };
    float* dis = take(sE);
    float* gd = take(sE);
    float *preT1[MDG_SCHNET_MAX_LAYERS], *W[MDG_SCHNET_MAX_LAYERS], *h[MDG_SCHNET_MAX_LAYERS], *preU1[MDG_SCHNET_MAX_LAYERS];
    float *We1T[MDG_SCHNET_MAX_LAYERS], *We2T[MDG_SCHNET_MAX_LAYERS];
    for (int l = 0; l < L; ++l) {
        preT1[l] = take(sEG); W[l] = take(sEF); h[l] = take(sNF); preU1[l] = take(sNA);
    }
    // transposed filter weights: rebuilt only when the model's weights_tag (or a weight pointer) changed
    {
        const size_t per = al((size_t)G * G) + al((size_t)G * F);
        MDG_TRY(c->sn_wcache.reserve(sizeof(float) * per * (size_t)L));
        bool fresh = m->weights_tag != 0 && m->weights_tag == c->sn_wtag;
        for (int l = 0; l < L; ++l) {
            We1T[l] = c->sn_wcache.as<float>() + per * (size_t)l;
            We2T[l] = We1T[l] + al((size_t)G * G);
            fresh = fresh && c->sn_wkey[2 * l] == m->layers[l].We1 && c->sn_wkey[2 * l + 1] == m->layers[l].We2;
        }
        if (!fresh) {
            for (int l = 0; l < L; ++l) {
                k_sn_transpose<<<(G * G + 255) / 256, 256, 0, (cudaStream_t)stream>>>(G, G, m->layers[l].We1, We1T[l]);
                k_sn_transpose<<<(F * G + 255) / 256, 256, 0, (cudaStream_t)stream>>>(F, G, m->layers[l].We2, We2T[l]);
                c->sn_wkey[2 * l] = m->layers[l].We1;
                c->sn_wkey[2 * l + 1] = m->layers[l].We2;
            }
            c->sn_wtag = m->weights_tag;
            c->stat_launches += 2 * L;
        }
    }
    float* r = take(sNA);
    float* u1 = take(sNA);
    float* gr = take(sNA);
    float* gu = take(sNA);
    float* agg = take(sNF);
    float* gagg = take(sNF);
    float* gh = take(sNF);
    float* y = take(sNR);
    float* preY = take(sNR);
    float* gy = take(sNR);
    double* epart = (double*)take(512);

    const int T = 256;
    // ---- forward ---------------------------------------------------------------------------------------------
    if (E > 0) MDG_TRY(mdg_i_pair_dis_fwd(d_xyz, d_nbr, d_offsets, E, d_E, h_off_scale3, dis, st));
    k_sn_embed<<<(unsigned)(((int64_t)n * A + T - 1) / T), T, 0, st>>>(n, A, d_z, m->embed, r);
    const unsigned eb = (unsigned)((E + SN_TE - 1) / SN_TE);
    for (int l = 0; l < L; ++l) {
        const mdg_schnet_layer& Y = m->layers[l];
        if (E > 0)
            k_sn_edge_fwd<<<eb, 256, 0, st>>>(E, d_E, G, F, dis, Y.mu, Y.width, We1T[l], Y.be1, We2T[l], Y.be2, preT1[l], W[l]);
        MDG_TRY((sn_gemm<true, SN_EPI_BIAS>(c, n, F, A, r, Y.Wn, A, Y.bn, nullptr, h[l], st)));
        MDG_TRY(mdg_i_cfconv_agg(c, h[l], W[l], n, F, agg, st));
        MDG_TRY((sn_gemm<true, SN_EPI_BIAS_SSP>(c, n, A, F, agg, Y.Wu1, F, Y.bu1, preU1[l], u1, st)));
        MDG_TRY((sn_gemm<true, SN_EPI_BIAS_ADD>(c, n, A, A, u1, Y.Wu2, A, Y.bu2, nullptr, r, st)));
    }
    MDG_TRY((sn_gemm<true, SN_EPI_BIAS_SSP>(c, n, R, A, r, m->Wr1, A, m->br1, preY, y, st)));
    int rb = (n + 7) / 8;
    if (rb > 256) rb = 256;
    k_sn_readout<<<rb, 256, 0, st>>>(n, R, y, preY, m->Wr2, m->br2, gy, epart);
    k_sn_energy_final<<<1, 32, 0, st>>>(rb, epart, d_energy);
    c->stat_launches += 4 + 5 * L;
    MDG_KERNEL_CHECK();
    if (!d_force) return MDG_OK;

    // ---- backward (dE/dxyz) ----------------------------------------------------------------------------------
    MDG_TRY((sn_gemm<false, SN_EPI_STORE>(c, n, A, R, gy, m->Wr1, A, nullptr, nullptr, gr, st)));
    if (E > 0) MDG_CUDA(cudaMemsetAsync(gd, 0, sizeof(float) * (size_t)E, st));
    const unsigned ebb = (unsigned)((E + SN_TEB - 1) / SN_TEB);
    const size_t bwd_smem = sizeof(float) * (size_t)F * SN_TEB;
    if (bwd_smem > 40 * 1024) {
        if (bwd_smem > 200 * 1024) { mdg_set_error("mdg_schnet_energy_force: n_filters=%d too large for the fused edge backward", F); return MDG_E_BADARG; }
        MDG_CUDA(cudaFuncSetAttribute(k_sn_edge_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem));
    }
    for (int l = L - 1; l >= 0; --l) {
        const mdg_schnet_layer& Y = m->layers[l];
        MDG_TRY((sn_gemm<false, SN_EPI_MUL_SIG>(c, n, A, A, gr, Y.Wu2, A, nullptr, preU1[l], gu, st)));
        MDG_TRY((sn_gemm<false, SN_EPI_STORE>(c, n, F, A, gu, Y.Wu1, F, nullptr, nullptr, gagg, st)));
        if (E > 0) {
            k_sn_edge_bwd<<<ebb, 256, bwd_smem, st>>>(E, d_E, G, F, d_nbr, h[l], gagg, dis, Y.mu, Y.width, Y.We1,
                                                                             Y.We2, preT1[l], gd);
        }
        if (l > 0) {     // the embedding below layer 0 does not depend on the positions
            MDG_TRY(mdg_i_cfconv_agg(c, gagg, W[l], n, F, gh, st));
            MDG_TRY((sn_gemm<false, SN_EPI_ADD>(c, n, A, F, gh, Y.Wn, A, nullptr, nullptr, gr, st)));
        }
    }
    k_sn_edge_force<<<(n + 3) / 4, 128, 0, st>>>(n, c->g_off.as<int>(), c->g_edge.as<int>(), d_nbr, d_offsets, h_off_scale3[0],
                                                     h_off_scale3[1], h_off_scale3[2], d_xyz, dis, gd, d_force);
    c->stat_launches += 3 + 5 * L;
    MDG_KERNEL_CHECK();
    return MDG_OK;
}
