// adjoint.cu - analytic Hessian-vector / parameter-Jacobian-vector products of the listed-pair force
// (SURVEY.md section 8, row f1 / a17).
//
// The adjoint solver (reference torchmd/sovlers.py:211-293, OdeintAdjointMethod.backward, and the len(y)==8 branch of
// NHverlet_update :129-164) needs, at every reverse step, the vector-Jacobian products of the equation of motion
// (torchmd/md.py:210-240) with the adjoint state: the reference gets them by differentiating the autograd force a
// SECOND time (double backward through compute_dis / u(r), torch.autograd.grad(..., create_graph=True)).
// For the power-law pair family the second derivative is closed-form, so one row-streaming kernel - the same data
// flow as the force kernel - gives both products for a given adjoint vector a (N x 3):
//
//     hv      = (dF/dq)^T a = -H a,         (H a)_i = sum_j [ c (r.(a_i - a_j)) r - g (a_i - a_j) ],   r = x_j - x_i + off L
//     dtheta  = (dF/dtheta)^T a = - sum_pairs (dg/dtheta) (r.(a_i - a_j))
// with, for u = 4 eps (s^p - B s^q), s = sigma / |r|  (LennardJones p,q = 12,6; LennardJones69 9,6; LJFamily rep,attr;
// ExcludedVolume p = power, B = 0;  reference torchmd/potentials.py:61-73,317-352):
//     g = -u'/r        = 4 eps (p s^p - B q s^q) / r^2
//     c = (u''-u'/r)/r^2 = 4 eps (p (p+2) s^p - B q (q+2) s^q) / r^4
//     dg/dsigma        = 4 eps (p^2 s^p - B q^2 s^q) / (sigma r^2),        dg/deps = g / eps
// Buck (u = A e^{-B r} - C r^-6) and ModifiedMorse have their own closed forms in hvp_eval below.
#include "common.cuh"

struct PowLaw {
    int   kind;                     // MDG_POT_*: power-law family, Buck or ModifiedMorse
    float sigma, eps, p, q, B;      // power law: u = 4 eps (s^p - B s^q)
    int   ip, iq;                   // integer exponents (>= 0) or -1: use powf
    float a0, a1, a2, aux;          // Buck: A, B, C ; ModifiedMorse: a, phi, -, 1 / (1 + A0)
    float scale[3];                 // post-loop scales of the three parameter sums
};

static int make_powlaw(int kind, const float* h_params, int n_params, PowLaw* out) {
    PowLaw P;
    memset(&P, 0, sizeof(P));
    float v[MDG_MAX_POT_PARAMS] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < n_params && k < MDG_MAX_POT_PARAMS; ++k) v[k] = h_params[k];
    P.kind = kind;
    P.sigma = v[0];
    P.eps = v[1];
    P.B = 1.f;
    P.scale[0] = v[0] != 0.f ? 1.0f / v[0] : 0.f;       // d/dsigma carries 1/sigma, d/deps carries 1/eps
    P.scale[1] = v[1] != 0.f ? 1.0f / v[1] : 0.f;
    P.scale[2] = 0.f;
    if (kind == MDG_POT_LJ) { P.p = 12.f; P.q = 6.f; }
    else if (kind == MDG_POT_LJ69) { P.p = 9.f; P.q = 6.f; }
    else if (kind == MDG_POT_LJFAM) { P.p = v[2]; P.q = v[3]; }
    else if (kind == MDG_POT_EXV) { P.p = v[2]; P.q = 0.f; P.B = 0.f; }
    else if (kind == MDG_POT_BUCK) { P.a0 = v[0]; P.a1 = v[1]; P.a2 = v[2]; P.scale[0] = P.scale[1] = P.scale[2] = 1.f; }
    else if (kind == MDG_POT_MORSE) {
        P.a0 = v[0]; P.a1 = v[1];
        double a = v[0], phi = v[1];
        double A0 = (phi >= 0) ? 0.0 : (exp(2 * a / phi) - 2 * exp(a / phi));      // potentials.py:82-85
        P.aux = (float)(1.0 / (1.0 + A0));
        P.scale[0] = P.scale[1] = P.scale[2] = 0.f;                                // a, phi are plain floats: no parameters
    }
    else return MDG_E_BADARG;
    auto as_int = [](float x) { return (x == floorf(x) && x >= 0.f && x < 64.f) ? (int)x : -1; };
    P.ip = as_int(P.p);
    P.iq = as_int(P.q);
    *out = P;
    return MDG_OK;
}

__device__ __forceinline__ float powlaw_pow(float s, float e, int ie) { return ie >= 0 ? mdg_ipow(s, ie) : powf(s, e); }

// g = -u'/r, c = (u'' - u'/r)/r^2 and the UNSCALED parameter derivatives of g (times PowLaw::scale after the loop)
__device__ __forceinline__ void hvp_eval(const PowLaw& P, float d2, float& g, float& c, float* dg) {
    const float r2i = 1.0f / d2;
    if (P.kind == MDG_POT_BUCK) {
        // u = A e^{-B r} - C r^-6 (potentials.py:354-365):  u' = -A B e + 6 C r^-7,  u'' = A B^2 e - 42 C r^-8
        const float r = sqrtf(d2), ri = 1.0f / r;
        const float ex = expf(-P.a1 * r);
        const float r8i = r2i * r2i * r2i * r2i;
        g = P.a0 * P.a1 * ex * ri - 6.0f * P.a2 * r8i;
        c = (P.a0 * P.a1 * P.a1 * ex + P.a0 * P.a1 * ex * ri - 48.0f * P.a2 * r8i) * r2i;
        dg[0] = P.a1 * ex * ri;                               // dg/dA
        dg[1] = P.a0 * ex * (1.0f - P.a1 * r) * ri;           // dg/dB
        dg[2] = -6.0f * r8i;                                  // dg/dC
    } else if (P.kind == MDG_POT_MORSE) {
        // u = (e^{2x} - 2 e^x - A0) / (1 + A0), x = a (1 - r^phi) / phi (potentials.py:75-93)
        const float r = sqrtf(d2), ri = 1.0f / r;
        const float rphi = powf(r, P.a1);
        const float x = P.a0 * (1.0f - rphi) / P.a1;
        const float e1 = expf(x), e2 = e1 * e1;
        const float x1 = -P.a0 * rphi * ri;                   // dx/dr
        const float x2 = -P.a0 * (P.a1 - 1.0f) * rphi * r2i;  // d2x/dr2
        const float u1 = (2.0f * e2 - 2.0f * e1) * x1 * P.aux;
        const float u2 = ((4.0f * e2 - 2.0f * e1) * x1 * x1 + (2.0f * e2 - 2.0f * e1) * x2) * P.aux;
        g = -u1 * ri;
        c = (u2 - u1 * ri) * r2i;
        dg[0] = dg[1] = dg[2] = 0.f;
    } else {
        const float e4 = 4.0f * P.eps;
        const float sr = P.sigma * sqrtf(r2i);
        const float sp = powlaw_pow(sr, P.p, P.ip);
        const float sq = P.B != 0.f ? P.B * powlaw_pow(sr, P.q, P.iq) : 0.f;
        const float tp = P.p * sp, tq = P.q * sq;
        g = e4 * (tp - tq) * r2i;
        c = e4 * ((P.p + 2.0f) * tp - (P.q + 2.0f) * tq) * r2i * r2i;
        dg[0] = e4 * (P.p * tp - P.q * tq) * r2i;             // * 1/sigma
        dg[1] = g;                                            // * 1/eps
        dg[2] = 0.f;
    }
}

__global__ void k_hvp_gather(int n, const float* __restrict__ xyz, const float* __restrict__ avec, const int* __restrict__ perm,
                             float4* __restrict__ qs, float4* __restrict__ as4) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int i = perm[s];
    qs[s] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], __int_as_float(i));
    as4[s] = make_float4(avec[3 * i], avec[3 * i + 1], avec[3 * i + 2], 0.f);
}

#define HVP_GROUP 4
__global__ void __launch_bounds__(256) k_pair_hvp(int n, const float4* __restrict__ qs, const float4* __restrict__ as4,
                                                  const uint32_t* __restrict__ rows, const int* __restrict__ row_len, int cap,
                                                  Box bx, PowLaw P, const int* __restrict__ perm, float* __restrict__ hv,
                                                  double* __restrict__ dp_partials) {
    const int lane = threadIdx.x % HVP_GROUP;
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) / HVP_GROUP;
    float hx = 0.f, hy = 0.f, hz = 0.f, dpar[3] = {0.f, 0.f, 0.f};
    if (s < n) {
        const float4 qi = qs[s], ai = as4[s];
        const uint32_t* row = rows + (size_t)s * cap;
        const int m = row_len[s] & MDG_ROW_LEN_MASK;
        for (int k = lane; k < m; k += HVP_GROUP) {
            const uint32_t e = row[k];
            const uint32_t j = e & MDG_IDX_MASK, code = e >> MDG_IDX_BITS;
            const float4 qj = qs[j], aj = as4[j];
            float dx = (qj.x - qi.x) + mdg_code_shift(code & 3u, bx.L[0]);
            float dy = (qj.y - qi.y) + mdg_code_shift((code >> 2) & 3u, bx.L[1]);
            float dz = (qj.z - qi.z) + mdg_code_shift((code >> 4) & 3u, bx.L[2]);
            float d2 = dx * dx + dy * dy + dz * dz;
            if (d2 == 0.0f) continue;                       // padding (self) entries and coincident atoms, like the list
            float g, c, dg[3];
            hvp_eval(P, d2, g, c, dg);
            float ax = ai.x - aj.x, ay = ai.y - aj.y, az = ai.z - aj.z;
            float rda = dx * ax + dy * ay + dz * az;
            hx += c * rda * dx - g * ax;
            hy += c * rda * dy - g * ay;
            hz += c * rda * dz - g * az;
            dpar[0] += dg[0] * rda;
            dpar[1] += dg[1] * rda;
            dpar[2] += dg[2] * rda;
        }
    }
#pragma unroll
    for (int o = HVP_GROUP / 2; o > 0; o >>= 1) {
        hx += __shfl_xor_sync(0xffffffffu, hx, o);
        hy += __shfl_xor_sync(0xffffffffu, hy, o);
        hz += __shfl_xor_sync(0xffffffffu, hz, o);
    }
    if (s < n && lane == 0) {
        int i = perm[s];
        hv[3 * i] = -hx; hv[3 * i + 1] = -hy; hv[3 * i + 2] = -hz;    // (dF/dq)^T a = -H a
    }
    // parameter products: every undirected pair is visited twice -> factor 1/2; sign: dF/dtheta = -(dg/dtheta) r ...
    __shared__ double sm[8][3];
    double v0 = (double)dpar[0], v1 = (double)dpar[1], v2 = (double)dpar[2];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
    }
    if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5][0] = v0; sm[threadIdx.x >> 5][1] = v1; sm[threadIdx.x >> 5][2] = v2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w][threadIdx.x];
        dp_partials[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) k_hvp_finalize(int nblocks, const double* __restrict__ part, float s0, float s1, float s2,
                                                      float* __restrict__ dtheta) {
    __shared__ double sm[8];
    for (int what = 0; what < 3; ++what) {
        double v = 0;
        for (int i = threadIdx.x; i < nblocks; i += blockDim.x) v += part[(size_t)i * 3 + what];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0;
            for (int w = 0; w < 8; ++w) t += sm[w];
            dtheta[what] = (float)(-0.5 * t * (double)(what == 0 ? s0 : (what == 1 ? s1 : s2)));
        }
    }
    if (threadIdx.x == 0) dtheta[3] = 0.f;
}

extern "C" int mdg_pair_hvp(mdg_ctx* c, int kind, const float* h_params, int n_params, const float* d_xyz, int n,
                            const float* d_avec, float* d_hv, float* d_dtheta, void* stream) {
    if (!c || !h_params || (n > 0 && (!d_xyz || !d_avec || !d_hv))) { mdg_set_error("mdg_pair_hvp: null argument"); return MDG_E_BADARG; }
    PowLaw P;
    if (make_powlaw(kind, h_params, n_params, &P) != MDG_OK) {
        mdg_set_error("mdg_pair_hvp: unknown potential kind %d", kind);
        return MDG_E_BADARG;
    }
    if (!c->built || n != c->n) { mdg_set_error("mdg_pair_hvp: no list built for n=%d", n); return MDG_E_STATE; }
    MDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        if (d_dtheta) MDG_CUDA(cudaMemsetAsync(d_dtheta, 0, sizeof(float) * MDG_MAX_POT_PARAMS, st));
        return MDG_OK;
    }
    const int T = 256;
    const int nb = (n + T - 1) / T;
    const int hb = (int)(((int64_t)n * HVP_GROUP + T - 1) / T);
    MDG_TRY(c->f4b.reserve(sizeof(float4) * (size_t)n));                    // sorted adjoint vector
    MDG_TRY(c->partials.reserve(sizeof(double) * (size_t)hb * 3 + 64));
    float4* qs = c->qs_ptr;
    float4* as4 = c->f4b.as<float4>();
    k_hvp_gather<<<nb, T, 0, st>>>(n, d_xyz, d_avec, c->perm.as<int>(), qs, as4);
    k_pair_hvp<<<hb, T, 0, st>>>(n, qs, as4, c->rows.as<uint32_t>(), c->row_len.as<int>(), c->cap, c->box, P, c->perm.as<int>(),
                                 d_hv, c->partials.as<double>());
    if (d_dtheta)
        k_hvp_finalize<<<1, 256, 0, st>>>(hb, c->partials.as<double>(), P.scale[0], P.scale[1], P.scale[2], d_dtheta);
    c->stat_launches += 3;
    MDG_KERNEL_CHECK();
    return MDG_OK;
}
