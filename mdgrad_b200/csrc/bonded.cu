// bonded.cu - bonded terms of a Stack force field: BondPotentials / AnglePotentials
// (reference torchmd/interface.py:406-454 / :456-510) as one native energy + force + dE/dparam program.
//
// Reference arithmetic (fp32):
//   bond  (i, j):     v = x_i - x_j;  off = -(v >= L/2) + (v < -L/2)  (get_offsets, topology.py:74-80);  v += off * L;
//                     b = (vx^2 + vy^2) + vz^2   -- the SQUARED length, the reference never takes the root (:447);
//                     E = 0.5 k sum_t (b_t - ro)^2
//   angle (a, c, e):  v1 = x_a - x_c, v2 = x_e - x_c (same image rule);  cos = v1.v2 / sqrt(|v1|^2 |v2|^2);
//                     theta = acos(cos);  E = 0.5 k sum_t (theta_t - theta0)^2
// Data flow (no atomics, deterministic): k_bonded_terms - one thread per term - writes the term's position gradients
// into a slot table (bond: 1 slot, angle: 2 slots; the remaining atom of a term gets minus the sum) and block partial
// sums of the energies / parameter derivatives in fp64;  k_bonded_gather - one thread per atom - walks the atom's
// reference list (a CSR over the STATIC topology, built once by the caller) and sums its slots in term order;
// k_bonded_reduce folds the block partials.  Algorithmic bytes: 16 (B + 2A) written + read, 24 B + 36 A of int64
// topology, 12 N forces: a latency-bound helper next to the pair kernels (a 64-bead chain has 63 bonds).
#include "common.cuh"

#define BD_T 128

struct BdArgs {
    const int64_t* bond_top;
    const int64_t* angle_top;
    int   nb, na;
    float kb, r0, ka, th0;
    float L[3];
};

// get_offsets (topology.py:74-80): v >= L/2 -> -1, v < -L/2 -> +1; then v + off * L (two rounded ops, as the reference)
__device__ __forceinline__ float bd_image(float v, float L) {
    float off = 0.f;
    if (v >= 0.5f * L) off = -1.f;
    else if (v < -0.5f * L) off = 1.f;
    return __fadd_rn(v, __fmul_rn(off, L));
}

__device__ __forceinline__ void bd_vec(const float* __restrict__ xyz, int64_t i, int64_t j, const float* L, float& x, float& y, float& z) {
    x = bd_image(__fsub_rn(xyz[3 * i], xyz[3 * j]), L[0]);
    y = bd_image(__fsub_rn(xyz[3 * i + 1], xyz[3 * j + 1]), L[1]);
    z = bd_image(__fsub_rn(xyz[3 * i + 2], xyz[3 * j + 2]), L[2]);
}

// partial sums per block: [0] E_bond, [1] E_angle, [2] dE/dk_b, [3] dE/dr0, [4] dE/dk_a, [5] dE/dtheta0
__global__ void __launch_bounds__(BD_T) k_bonded_terms(BdArgs A, int n_atoms, const float* __restrict__ xyz, float4* __restrict__ slots,
                                                       double* __restrict__ partials) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (t < A.nb) {
        const int64_t i = A.bond_top[2 * t], j = A.bond_top[2 * t + 1];
        if (i < 0 || j < 0 || i >= n_atoms || j >= n_atoms) {      // a term naming a non-existent atom contributes nothing
            slots[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            float x, y, z;
            bd_vec(xyz, i, j, A.L, x, y, z);
            const float b = mdg_d2_exact(x, y, z);
            const float d = b - A.r0;
            const float c = 2.0f * A.kb * d;              // dE/dv = k (b - ro) * 2 v
            slots[t] = make_float4(c * x, c * y, c * z, 0.f);
            acc[0] = 0.5 * (double)A.kb * (double)d * (double)d;
            acc[2] = 0.5 * (double)d * (double)d;
            acc[3] = -(double)A.kb * (double)d;
        }
    } else if (t < A.nb + A.na) {
        const int a = t - A.nb;
        const int64_t i0 = A.angle_top[3 * a], i1 = A.angle_top[3 * a + 1], i2 = A.angle_top[3 * a + 2];
        float4* s = slots + A.nb + 2 * a;
        if (i0 < 0 || i1 < 0 || i2 < 0 || i0 >= n_atoms || i1 >= n_atoms || i2 >= n_atoms) {
            s[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            s[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            float x1, y1, z1, x2, y2, z2;
            bd_vec(xyz, i0, i1, A.L, x1, y1, z1);
            bd_vec(xyz, i2, i1, A.L, x2, y2, z2);
            const float dot = x1 * x2 + y1 * y2 + z1 * z2;
            const float n1 = mdg_d2_exact(x1, y1, z1), n2 = mdg_d2_exact(x2, y2, z2);
            const float nrm = sqrtf(n1 * n2);
            const float cs = dot / nrm;
            const float th = acosf(cs);
            const float d = th - A.th0;
            // dE/dcos = k (theta - theta0) * (-1 / sqrt(1 - cos^2));  dcos/dv1 = v2 / nrm - cos v1 / |v1|^2  (same for v2)
            const float w = -A.ka * d / sqrtf(1.0f - cs * cs);
            const float a1 = cs / n1, a2 = cs / n2, inv = 1.0f / nrm;
            s[0] = make_float4(w * (x2 * inv - a1 * x1), w * (y2 * inv - a1 * y1), w * (z2 * inv - a1 * z1), 0.f);
            s[1] = make_float4(w * (x1 * inv - a2 * x2), w * (y1 * inv - a2 * y2), w * (z1 * inv - a2 * z2), 0.f);
            acc[1] = 0.5 * (double)A.ka * (double)d * (double)d;
            acc[4] = 0.5 * (double)d * (double)d;
            acc[5] = -(double)A.ka * (double)d;
        }
    }
    __shared__ double sm[BD_T / 32][6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        double v = acc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double s = 0;
        for (int w = 0; w < BD_T / 32; ++w) s += sm[w][threadIdx.x];
        partials[(size_t)blockIdx.x * 6 + threadIdx.x] = s;
    }
}

// refs[r] = slot * 4 + role;  role 0: + slot, role 1: - slot (second atom of a bond), role 2: - (slot + next slot)
// (centre atom of an angle).  F = - sum.
__global__ void k_bonded_gather(int n, const int* __restrict__ ref_start, const int* __restrict__ refs,
                                const float4* __restrict__ slots, float* __restrict__ force) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int r = ref_start[i]; r < ref_start[i + 1]; ++r) {
        const int ref = refs[r];
        const int slot = ref >> 2, role = ref & 3;
        float4 g = slots[slot];
        if (role == 2) {
            const float4 h = slots[slot + 1];
            g.x += h.x; g.y += h.y; g.z += h.z;
        }
        if (role == 0) { gx += g.x; gy += g.y; gz += g.z; }
        else { gx -= g.x; gy -= g.y; gz -= g.z; }
    }
    force[3 * i] = -gx; force[3 * i + 1] = -gy; force[3 * i + 2] = -gz;
}

__global__ void k_bonded_reduce(const double* __restrict__ partials, int nblocks, float* __restrict__ energy2,
                                float* __restrict__ dparams4) {
    __shared__ double sm[6][32];
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;      // 6 warps, one per quantity
    double s = 0;
    for (int b = lane; b < nblocks; b += 32) s += partials[(size_t)b * 6 + q];
    sm[q][lane] = s;
    __syncwarp();
    if (lane == 0) {
        double t = 0;
        for (int l = 0; l < 32; ++l) t += sm[q][l];
        if (q < 2) { if (energy2) energy2[q] = (float)t; }
        else if (dparams4) dparams4[q - 2] = (float)t;
    }
}

extern "C" int mdg_bonded_force(mdg_ctx* c, const mdg_bonded_terms* h, const float* d_xyz, int n, const float* h_cell3,
                                float* d_energy2, float* d_force, float* d_dparams4, void* stream) {
    if (!c || !h || !h_cell3) { mdg_set_error("mdg_bonded_force: null argument"); return MDG_E_BADARG; }
    if (n < 0 || h->n_bonds < 0 || h->n_angles < 0) { mdg_set_error("mdg_bonded_force: negative size"); return MDG_E_BADARG; }
    if ((h->n_bonds > 0 && !h->d_bond_top) || (h->n_angles > 0 && !h->d_angle_top) || (n > 0 && !d_xyz)) {
        mdg_set_error("mdg_bonded_force: null topology / positions");
        return MDG_E_BADARG;
    }
    if (d_force && n > 0 && (!h->d_ref_start || (h->n_bonds + h->n_angles > 0 && !h->d_refs))) {
        mdg_set_error("mdg_bonded_force: forces need the atom -> term reference list (d_ref_start, d_refs)");
        return MDG_E_BADARG;
    }
    if ((int64_t)h->n_bonds + 2 * (int64_t)h->n_angles >= (1 << 29)) { mdg_set_error("mdg_bonded_force: too many terms"); return MDG_E_BADARG; }
    MDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = h->n_bonds + h->n_angles;
    if (nt == 0) {
        if (d_energy2) MDG_CUDA(cudaMemsetAsync(d_energy2, 0, sizeof(float) * 2, st));
        if (d_dparams4) MDG_CUDA(cudaMemsetAsync(d_dparams4, 0, sizeof(float) * 4, st));
        if (d_force && n > 0) MDG_CUDA(cudaMemsetAsync(d_force, 0, sizeof(float) * 3 * (size_t)n, st));
        return MDG_OK;
    }
    const int nblk = (nt + BD_T - 1) / BD_T;
    MDG_TRY(c->bd_slots.reserve(sizeof(float4) * ((size_t)h->n_bonds + 2 * (size_t)h->n_angles)));
    MDG_TRY(c->bd_part.reserve(sizeof(double) * 6 * (size_t)nblk));
    BdArgs A;
    A.bond_top = h->d_bond_top; A.angle_top = h->d_angle_top;
    A.nb = h->n_bonds; A.na = h->n_angles;
    A.kb = h->k_bond; A.r0 = h->r0; A.ka = h->k_angle; A.th0 = h->theta0;
    for (int k = 0; k < 3; ++k) A.L[k] = h_cell3[k];
    k_bonded_terms<<<nblk, BD_T, 0, st>>>(A, n, d_xyz, c->bd_slots.as<float4>(), c->bd_part.as<double>());
    if (d_force && n > 0)
        k_bonded_gather<<<(n + 255) / 256, 256, 0, st>>>(n, h->d_ref_start, h->d_refs, c->bd_slots.as<float4>(), d_force);
    if (d_energy2 || d_dparams4) k_bonded_reduce<<<1, 192, 0, st>>>(c->bd_part.as<double>(), nblk, d_energy2, d_dparams4);
    c->stat_launches += 1 + (d_force && n > 0 ? 1 : 0) + ((d_energy2 || d_dparams4) ? 1 : 0);
    MDG_KERNEL_CHECK();
    return MDG_OK;
}
