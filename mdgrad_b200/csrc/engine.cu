// engine.cu - fused Velocity-Verlet / Nose-Hoover-chain integrator kernels (K4 in SURVEY.md 2c)
// and the device-resident MD epoch driver (single GPU, or one slab of a multi-GPU run - dist.cu).
//
// Replaces the hot loop of Simulations.simulate (reference torchmd/md.py:73-96) ->
// FixedGridODESolver.integrate (torchmd/tinydiffeq.py:56-76) -> NHverlet_update / verlet_update
// (torchmd/sovlers.py:110-127 / :25-40) -> NoseHooverChain.forward / NVE.forward
// (torchmd/md.py:210-240 / :131-148).
//
// This translation unit is compiled with -fmad=false: the integrator algebra is memory-bound,
// and keeping every product/sum individually rounded reproduces the reference's separate
// elementwise ATen ops (only the global kinetic-energy reduction order differs).
#include <stdlib.h>
#include <vector>
#include "common.cuh"
#include "dist.cuh"

#define PROF_MAX_EVENTS 16384

// ---------------------------------------------------------------------------------------------
// Development aid (MDG_TIMELINE=<path prefix>): CUDA events at the phase boundaries of a window of steps, on the streams the
// phases run on, written as "<label> <ms since the first mark>" to <prefix><rank>.txt after the epoch.  Events do not serialise
// the streams, so this is the pipelined timeline (nsys is not available on the GPU boxes).  Off unless the variable is set.
// ---------------------------------------------------------------------------------------------
struct Timeline {
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    std::vector<int> step;
    const char* prefix = nullptr;
    int g0 = 20, g1 = 36;
    bool init = false;
};
static Timeline g_tl;
static inline bool tl_on(int g) {
    if (!g_tl.init) {
        g_tl.init = true;
        g_tl.prefix = getenv("MDG_TIMELINE");
        const char* w = getenv("MDG_TIMELINE_STEPS");
        if (w) { int a = 0, b = 0; if (sscanf(w, "%d:%d", &a, &b) == 2 && b > a) { g_tl.g0 = a; g_tl.g1 = b; } }
    }
    return g_tl.prefix != nullptr && g >= g_tl.g0 && g < g_tl.g1;
}
static inline void tl_mark(int g, const char* label, cudaStream_t st) {
    if (!tl_on(g)) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    g_tl.marks.emplace_back(label, e);
    g_tl.step.push_back(g);
}
static void tl_flush(int rank) {
    if (g_tl.marks.empty()) return;
    char path[512];
    snprintf(path, sizeof(path), "%s%d.txt", g_tl.prefix, rank);
    FILE* f = fopen(path, "a");
    for (size_t i = 0; i < g_tl.marks.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_tl.marks[0].second, g_tl.marks[i].second);
        if (f) fprintf(f, "%d %s %.3f\n", g_tl.step[i], g_tl.marks[i].first, 1e3 * ms);
    }
    if (f) { fprintf(f, "--\n"); fclose(f); }
    for (auto& m : g_tl.marks) cudaEventDestroy(m.second);
    g_tl.marks.clear();
    g_tl.step.clear();
}

int mdg_i_force_blocks(mdg_ctx* c);

#define INT_THREADS 256
#define INT_MAX_BLOCKS 592   // 148 SMs x 4 resident CTAs; grid-stride beyond that

struct IntArgs {
    int    s0, s1;                 // sorted-atom range integrated by this context
    int    integrator;
    int    M;                      // chains
    float  Q[MDG_MAX_CHAINS];
    float  T;                      // fl32(T)
    float  target;                 // fl32(T * ndof * 0.5) computed in double like the python expression
    float  half_skin2;             // (skin/2)^2
};

// scalar state on device: ping-pong bath momenta
struct Scalars {
    float pv[2][MDG_MAX_CHAINS];
    float ph[MDG_MAX_CHAINS];
};

__device__ __forceinline__ double block_sum_double(double v, double* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    return t;  // valid in thread 0
}

// fixed-order sum of per-block partials, identical in every block -> deterministic broadcast
__device__ __forceinline__ float sum_partials(const double* __restrict__ part, int np, double* sm, float* bc) {
    double v = 0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) v += part[i];
    double t = block_sum_double(v, sm);
    if (threadIdx.x == 0) *bc = (float)t;
    __syncthreads();
    return *bc;
}

// NHC bath derivative (reference torchmd/md.py:234-236)
__device__ __forceinline__ void nhc_dpv(const IntArgs& A, float ke, const float* pv, float* dpv) {
    int M = A.M;
    dpv[0] = 2.0f * (ke - A.target) - pv[0] * pv[1] / A.Q[1];
    for (int k = 1; k < M - 1; ++k)
        dpv[k] = (pv[k - 1] * pv[k - 1] / A.Q[k - 1] - A.T) - pv[k + 1] * pv[k] / A.Q[k + 1];
    dpv[M - 1] = pv[M - 2] * pv[M - 2] / A.Q[M - 2] - A.T;
}

// kinetic energy partials of the initial velocities: ke = 0.5 * sum(p^2/m), p = v*m (md.py:221-223)
__global__ void __launch_bounds__(INT_THREADS) k_ke_init(IntArgs A, const float4* __restrict__ v4,
                                                         double* __restrict__ ke_part) {
    __shared__ double sm[INT_THREADS / 32];
    double acc = 0;
    for (int s = A.s0 + blockIdx.x * blockDim.x + threadIdx.x; s < A.s1; s += gridDim.x * blockDim.x) {
        float4 v = v4[s];
        float m = v.w;
        float px = v.x * m, py = v.y * m, pz = v.z * m;
        acc += (double)(px * px / m) + (double)(py * py / m) + (double)(pz * pz / m);
    }
    double t = block_sum_double(acc, sm);
    if (threadIdx.x == 0) ke_part[blockIdx.x] = 0.5 * t;
}

// multi-GPU: collapse two local partial arrays into dke[0..1] (then one 2-double all-reduce)
__global__ void __launch_bounds__(INT_THREADS) k_ke_pack(const double* __restrict__ a, int na, const double* __restrict__ b, int nb,
                                                         double* __restrict__ dke) {
    __shared__ double sm[INT_THREADS / 32];
    double va = 0, vb = 0;
    for (int i = threadIdx.x; i < na; i += blockDim.x) va += a[i];
    for (int i = threadIdx.x; i < nb; i += blockDim.x) vb += b[i];
    double ta = block_sum_double(va, sm);
    double tb = block_sum_double(vb, sm);
    if (threadIdx.x == 0) { dke[0] = ta; dke[1] = tb; }
}

// ---------------------------------------------------------------------------------------------------------------------
// peer-to-peer step path (dist.cuh): flags and kinetic energies live in DistSync blocks that the peers write over NVLink
// ---------------------------------------------------------------------------------------------------------------------
struct DistArgs {
    DistSync* mine;            // nullptr: single GPU, or the NCCL path (kinetic energies arrive in ke_part arrays)
    DistSync* below;           // the neighbours' blocks as mapped here (acknowledgements)
    DistSync* above;
    int world, seq;
    int pos_seq;               // pull path: > 0 -> once ALL blocks have stored their positions, raise pos_flag = pos_seq at both neighbours
    int wait_pull;             // pull path: > 0 -> both neighbours must have copied my boundary layers of that step before q is overwritten
};

// Prologue of the B kernels on the peer-to-peer path: acknowledge the ghosts of this step (the forces that read them are
// complete - stream order), wait for every rank's kinetic energies and sum them in rank order (identical on all ranks).
__device__ __forceinline__ void dist_ack(const DistArgs& D) {
    if (D.mine && blockIdx.x == 0 && threadIdx.x == 0) {
        vstore_i(&D.below->ack_flag[1], D.seq);      // I am the neighbour ABOVE of the rank below me
        vstore_i(&D.above->ack_flag[0], D.seq);
        D.mine->dbg[8] = mdg_globaltimer_ns();
    }
}

// Pull path (default): nothing is stored into a neighbour's arrays on the step path.  A rank raises pos_flag at its neighbours when
// its integrator kernel has stored the new positions (grid-wide ticket; only LOCAL stores to fence), the neighbours copy its
// boundary layer with NVLink loads (k_dist_pull) and answer with pull_flag, which this rank's next integrator kernel awaits
// before it overwrites the positions.  r02 stamps of the push form: 11 us from kernel start to flag (system-scope fences behind
// remote stores: 3-8 us, ticket + second fence + flag: 3 us) plus the stream hop - the ghosts arrived 16-20 us into the step.
__device__ __forceinline__ void dist_wait_pulled(const DistArgs& D) {
    if (D.mine && D.wait_pull > 0) {
        if (threadIdx.x == 0) {
            spin_until_ge(&D.mine->pull_flag[0], D.wait_pull, &D.mine->pad[1]);
            spin_until_ge(&D.mine->pull_flag[1], D.wait_pull, &D.mine->pad[1]);
        }
        __syncthreads();
    }
}
__device__ __forceinline__ void dist_raise_pos(const DistArgs& D) {
    if (D.mine && D.pos_seq > 0) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(&D.mine->ba_ticket, 1) == (int)gridDim.x - 1) {
            D.mine->ba_ticket = 0;
            __threadfence_system();
            vstore_i(&D.below->pos_flag[1], D.pos_seq);      // I am the neighbour ABOVE of the rank below me
            vstore_i(&D.above->pos_flag[0], D.pos_seq);
            D.mine->dbg[9] = mdg_globaltimer_ns();
        }
    }
}

__device__ __forceinline__ void dist_gather_ke(const DistArgs& D, double* sm, float* ke0, float* ke1) {
    const int par = D.seq & 1;
    if ((int)threadIdx.x < D.world) spin_until_ge(&D.mine->ke_flag[par][threadIdx.x], D.seq, &D.mine->pad[1]);
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int r = 0; r < D.world; ++r) {
            a += *(const volatile double*)&D.mine->ke[par][r][0];
            b += *(const volatile double*)&D.mine->ke[par][r][1];
        }
        sm[0] = a;
        sm[1] = b;
    }
    __syncthreads();
    *ke0 = (float)sm[0];
    *ke1 = (float)sm[1];
    __syncthreads();
}

// One launch per step on the communication stream: (1) wait until both neighbours have consumed the ghosts of the previous
// step, (2) store my bottom / top layer of positions into the ghost ranges of the rank below / above (same global sorted
// indices on both sides), (3) block 0: reduce my kinetic-energy partials and store them into every rank's table, flag after a
// system-scope fence, (4) the last block to finish raises the halo flags on the two neighbours.
__global__ void __launch_bounds__(256) k_dist_push(const float4* __restrict__ q, int lo0, int lo1, int hi0, int hi1,
                                                   float4* __restrict__ q_below, float4* __restrict__ q_above,
                                                   const double* __restrict__ ke_v_part, const double* __restrict__ ke_h_part,
                                                   int n_part, int nhc, PeerTab T, int me, int world, int below, int above, int seq,
                                                   int halo_blocks) {
    // blocks [0, halo_blocks): ghost layers; block halo_blocks (present when nhc): the kinetic energies.  The two paths end in
    // their own system-scope fence + flags, side by side (one after the other they cost ~8 us more per step: r02 timeline).
    __shared__ double sm[256 / 32];
    __shared__ int s_last;
    DistSync* mine = T.s[me];
    if (blockIdx.x == 0 && threadIdx.x == 0) { vstore_i(&mine->push_started, seq); mine->dbg[0] = mdg_globaltimer_ns(); }
    if ((int)blockIdx.x >= halo_blocks) {
        if (!nhc) return;
        double va = 0, vb = 0;
        for (int i = threadIdx.x; i < n_part; i += blockDim.x) { va += ke_v_part[i]; vb += ke_h_part[i]; }
        __shared__ double s_t[2];
        double ta = block_sum_double(va, sm);
        double tb = block_sum_double(vb, sm);        // (totals are valid in thread 0 only)
        if (threadIdx.x == 0) { s_t[0] = ta; s_t[1] = tb; }
        __syncthreads();
        ta = s_t[0]; tb = s_t[1];
        const int par = seq & 1;
        if ((int)threadIdx.x < world) {
            const int r = threadIdx.x;
            *(volatile double*)&T.s[r]->ke[par][me][0] = ta;
            *(volatile double*)&T.s[r]->ke[par][me][1] = tb;
            __threadfence_system();
            vstore_i(&T.s[r]->ke_flag[par][me], seq);
            if (r == 0) mine->dbg[5] = mdg_globaltimer_ns();
        }
        return;
    }
    if (threadIdx.x == 0) {
        spin_until_ge(&mine->ack_flag[0], seq - 1, &mine->pad[1]);
        spin_until_ge(&mine->ack_flag[1], seq - 1, &mine->pad[1]);
        if (blockIdx.x == 0) mine->dbg[1] = mdg_globaltimer_ns();
    }
    __syncthreads();
    const int nlo = lo1 - lo0, nhi = hi1 - hi0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlo + nhi; i += halo_blocks * blockDim.x) {
        if (i < nlo) q_below[lo0 + i] = q[lo0 + i];
        else q_above[hi0 + (i - nlo)] = q[hi0 + (i - nlo)];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) mine->dbg[2] = mdg_globaltimer_ns();
    __threadfence_system();
    if (blockIdx.x == 0 && threadIdx.x == 0) mine->dbg[3] = mdg_globaltimer_ns();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&mine->ticket, 1) == halo_blocks - 1) ? 1 : 0;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        mine->ticket = 0;
        __threadfence_system();
        vstore_i(&T.s[below]->halo_flag[1], seq);    // my bottom layer is the ghost layer ABOVE the rank below me
        vstore_i(&T.s[above]->halo_flag[0], seq);
        mine->dbg[4] = mdg_globaltimer_ns();
    }
}
// Gate on the main stream in front of the interior rows: returns once this step's push kernel is RUNNING (it then holds its few
// CTA slots).  Without it the interior force - which has no dependency to resolve - takes every CTA slot first and the push only
// starts when the first wave retires, ~15 us into the step: the ghosts land after the interior rows are done.
__global__ void k_dist_wait_started(DistSync* mine, int seq) {
    if (threadIdx.x == 0) spin_until_ge(&mine->push_started, seq, &mine->pad[1]);
}
__global__ void k_dist_ack(DistSync* below, DistSync* above, int seq) {
    if (threadIdx.x == 0) {
        vstore_i(&below->ack_flag[1], seq);
        vstore_i(&above->ack_flag[0], seq);
    }
}

// Pull path: copy the ghost layers (the top layer of the rank below, the bottom layer of the rank above - same global index
// ranges there) out of the neighbours' position arrays once they are final, then tell the neighbours.
__global__ void __launch_bounds__(256) k_dist_pull(float4* __restrict__ q, const float4* q_below, const float4* q_above, int gl0, int gl1,
                                                   int gu0, int gu1, DistSync* mine, DistSync* below, DistSync* above, int seq) {
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) mine->dbg[6] = mdg_globaltimer_ns();
        spin_until_ge(&mine->pos_flag[0], seq, &mine->pad[1]);
        spin_until_ge(&mine->pos_flag[1], seq, &mine->pad[1]);
        if (blockIdx.x == 0) mine->dbg[7] = mdg_globaltimer_ns();
    }
    __syncthreads();
    const int nl = gl1 - gl0, nu = gu1 - gu0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nl + nu; i += gridDim.x * blockDim.x) {
        if (i < nl) q[gl0 + i] = mdg_ld_peer(q_below + gl0 + i);
        else q[gu0 + (i - nl)] = mdg_ld_peer(q_above + gu0 + (i - nl));
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&mine->pull_ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        mine->pull_ticket = 0;
        __threadfence_system();
        vstore_i(&below->pull_flag[1], seq);         // I am the neighbour ABOVE of the rank below me
        vstore_i(&above->pull_flag[0], seq);
        mine->dbg[10] = mdg_globaltimer_ns();
    }
}

// Rebuild on the peer-to-peer path: my bottom two layers of (q, v, vh) go to the rank below, my top two to the rank above, into
// the SAME global index ranges of their arrays (what state_exchange does with 12 NCCL send / recv: ~50-60 us per rebuild in the r02
// timeline, launch and handshake latency rather than bytes).  Same write-after-read rule as the ghost push: the neighbours read
// these ranges last in the forces of the previous step, acknowledged by their B kernel.
__global__ void __launch_bounds__(256) k_dist_push_state(const float4* __restrict__ q, const float4* __restrict__ v,
                                                         const float4* __restrict__ vh, int lo0, int lo1, int hi0, int hi1,
                                                         float4* __restrict__ q_below, float4* __restrict__ v_below,
                                                         float4* __restrict__ vh_below, float4* __restrict__ q_above,
                                                         float4* __restrict__ v_above, float4* __restrict__ vh_above, DistSync* mine,
                                                         DistSync* below, DistSync* above, int seq) {
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        spin_until_ge(&mine->ack_flag[0], seq - 1, &mine->pad[1]);
        spin_until_ge(&mine->ack_flag[1], seq - 1, &mine->pad[1]);
    }
    __syncthreads();
    const int nlo = lo1 - lo0, nhi = hi1 - hi0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlo + nhi; i += gridDim.x * blockDim.x) {
        if (i < nlo) {
            const int k = lo0 + i;
            q_below[k] = q[k]; v_below[k] = v[k]; vh_below[k] = vh[k];
        } else {
            const int k = hi0 + (i - nlo);
            q_above[k] = q[k]; v_above[k] = v[k]; vh_above[k] = vh[k];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&mine->ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        mine->ticket = 0;
        __threadfence_system();
        vstore_i(&below->state_flag[1], seq);        // I am the neighbour ABOVE of the rank below me
        vstore_i(&above->state_flag[0], seq);
    }
}
__global__ void k_dist_wait_state(DistSync* mine, int seq) {
    if (threadIdx.x == 0) {
        spin_until_ge(&mine->state_flag[0], seq, &mine->pad[1]);
        spin_until_ge(&mine->state_flag[1], seq, &mine->pad[1]);
    }
}
__global__ void k_dist_wait(DistSync* mine, int seq) {
    if (threadIdx.x == 0) {
        mine->dbg[6] = mdg_globaltimer_ns();
        spin_until_ge(&mine->halo_flag[0], seq, &mine->pad[1]);
        spin_until_ge(&mine->halo_flag[1], seq, &mine->pad[1]);
        mine->dbg[7] = mdg_globaltimer_ns();
    }
}

// step part A (sovlers.py:111-118): a0 from (v, f, pv); vh = 1/2 a0 dt; q += (v + vh) dt;
// accumulates ke(v + vh) and checks the skin criterion against the positions of the last build.
__global__ void __launch_bounds__(INT_THREADS) k_step_a(IntArgs A, float dt, int pv_sel, const Scalars* __restrict__ sc,
                                                        const float4* __restrict__ v4, float4* __restrict__ vh4,
                                                        float4* __restrict__ q4, const float4* __restrict__ f4,
                                                        const float4* __restrict__ qref, int check_skin,
                                                        double* __restrict__ ke_half_part, int* __restrict__ flags, DistArgs D) {
    __shared__ double sm[INT_THREADS / 32];
    float pv0 = 0.f, Q0 = 1.f;
    dist_wait_pulled(D);
    if (A.integrator == MDG_INT_NHC) { pv0 = sc->pv[pv_sel][0]; Q0 = A.Q[0]; }
    double acc = 0;
    bool viol = false;
    for (int s = A.s0 + blockIdx.x * blockDim.x + threadIdx.x; s < A.s1; s += gridDim.x * blockDim.x) {
        float4 v = v4[s];
        float4 f = f4[s];
        float4 q = q4[s];
        float m = v.w;
        float ax, ay, az;
        if (A.integrator == MDG_INT_NHC) {
            float px = v.x * m, py = v.y * m, pz = v.z * m;
            ax = (f.x - pv0 * px / Q0) / m;
            ay = (f.y - pv0 * py / Q0) / m;
            az = (f.z - pv0 * pz / Q0) / m;
        } else {
            ax = f.x; ay = f.y; az = f.z;          // NVE: dv/dt = f, no mass division (md.py:146)
        }
        float hx = 0.5f * ax * dt, hy = 0.5f * ay * dt, hz = 0.5f * az * dt;
        float ux = v.x + hx, uy = v.y + hy, uz = v.z + hz;
        q.x = q.x + ux * dt;
        q.y = q.y + uy * dt;
        q.z = q.z + uz * dt;
        vh4[s] = make_float4(hx, hy, hz, 0.f);
        q4[s] = q;
        if (A.integrator == MDG_INT_NHC) {
            float px = ux * m, py = uy * m, pz = uz * m;
            acc += (double)(px * px / m) + (double)(py * py / m) + (double)(pz * pz / m);
        }
        if (check_skin) {
            float4 r = qref[s];
            float dx = q.x - r.x, dy = q.y - r.y, dz = q.z - r.z;
            viol |= (dx * dx + dy * dy + dz * dz) > A.half_skin2;
        }
    }
    if (viol) flags[5] = 1;
    if (A.integrator == MDG_INT_NHC) {
        double t = block_sum_double(acc, sm);
        if (threadIdx.x == 0) ke_half_part[blockIdx.x] = 0.5 * t;
    }
    dist_raise_pos(D);
}

// step part B (sovlers.py:120-127 + tinydiffeq.py:69): a1 from (v + vh, f_new, pv + ph);
// v += vh + 1/2 a1 dt; pv += ph + 1/2 dpv1 dt; writes the trajectory frame (original order) and
// accumulates ke(v_new) for the next step.
__global__ void __launch_bounds__(INT_THREADS) k_step_b(IntArgs A, float dt, int pv_sel, Scalars* __restrict__ sc,
                                                        float4* __restrict__ v4, const float4* __restrict__ vh4,
                                                        const float4* __restrict__ q4, const float4* __restrict__ f4,
                                                        const double* __restrict__ ke_part, const double* __restrict__ ke_half_part,
                                                        int n_part, double* __restrict__ ke_next_part,
                                                        float* __restrict__ traj_v, float* __restrict__ traj_q,
                                                        float* __restrict__ traj_pv_row, DistArgs D) {
    __shared__ double sm[INT_THREADS / 32];
    __shared__ float bc[2];
    __shared__ float s_pvh0;
    float pvh0 = 0.f, Q0 = 1.f;
    dist_ack(D);
    if (A.integrator == MDG_INT_NHC) {
        float ke0, ke1;
        if (D.mine) dist_gather_ke(D, sm, &ke0, &ke1);
        else {
            ke0 = sum_partials(ke_part, n_part, sm, &bc[0]);
            ke1 = sum_partials(ke_half_part, n_part, sm, &bc[1]);
        }
        if (threadIdx.x == 0) {
            float pv[MDG_MAX_CHAINS], ph[MDG_MAX_CHAINS], pvh[MDG_MAX_CHAINS], d0[MDG_MAX_CHAINS], d1[MDG_MAX_CHAINS];
            for (int k = 0; k < A.M; ++k) pv[k] = sc->pv[pv_sel][k];
            nhc_dpv(A, ke0, pv, d0);
            for (int k = 0; k < A.M; ++k) { ph[k] = 0.5f * d0[k] * dt; pvh[k] = pv[k] + ph[k]; }
            nhc_dpv(A, ke1, pvh, d1);
            s_pvh0 = pvh[0];
            if (blockIdx.x == 0) {
                for (int k = 0; k < A.M; ++k) {
                    float pn = pv[k] + (ph[k] + 0.5f * d1[k] * dt);
                    sc->pv[pv_sel ^ 1][k] = pn;
                    if (traj_pv_row) traj_pv_row[k] = pn;
                }
            }
        }
        __syncthreads();
        pvh0 = s_pvh0;
        Q0 = A.Q[0];
    }
    double acc = 0;
    for (int s = A.s0 + blockIdx.x * blockDim.x + threadIdx.x; s < A.s1; s += gridDim.x * blockDim.x) {
        float4 v = v4[s];
        float4 h = vh4[s];
        float4 f = f4[s];
        float m = v.w;
        float ax, ay, az;
        if (A.integrator == MDG_INT_NHC) {
            float px = (v.x + h.x) * m, py = (v.y + h.y) * m, pz = (v.z + h.z) * m;
            ax = (f.x - pvh0 * px / Q0) / m;
            ay = (f.y - pvh0 * py / Q0) / m;
            az = (f.z - pvh0 * pz / Q0) / m;
        } else {
            ax = f.x; ay = f.y; az = f.z;
        }
        v.x = v.x + (h.x + 0.5f * ax * dt);
        v.y = v.y + (h.y + 0.5f * ay * dt);
        v.z = v.z + (h.z + 0.5f * az * dt);
        v4[s] = v;
        if (A.integrator == MDG_INT_NHC) {
            float px = v.x * m, py = v.y * m, pz = v.z * m;
            acc += (double)(px * px / m) + (double)(py * py / m) + (double)(pz * pz / m);
        }
        if (traj_v) {
            float4 q = q4[s];
            int id = __float_as_int(q.w);
            traj_v[3 * (size_t)id] = v.x; traj_v[3 * (size_t)id + 1] = v.y; traj_v[3 * (size_t)id + 2] = v.z;
            traj_q[3 * (size_t)id] = q.x; traj_q[3 * (size_t)id + 1] = q.y; traj_q[3 * (size_t)id + 2] = q.z;
        }
    }
    if (A.integrator == MDG_INT_NHC) {
        double t = block_sum_double(acc, sm);
        if (threadIdx.x == 0) ke_next_part[blockIdx.x] = 0.5 * t;
    }
}

// Fused B(n) + A(n+1): the second half-kick of step n and the first half-kick + drift of step n+1 read the same
// force, velocities and positions, so one pass does both (saves one read of v, f, q and one launch per step).
// Bath scalars of step n+1 come from the same redundant per-block computation as in k_step_b.
__global__ void __launch_bounds__(INT_THREADS) k_step_ba(IntArgs A, float dt, float dt_next, int pv_sel, Scalars* __restrict__ sc,
                                                         float4* __restrict__ v4, float4* __restrict__ vh4, float4* __restrict__ q4,
                                                         const float4* __restrict__ f4, const double* __restrict__ ke_part,
                                                         const double* __restrict__ ke_half_part, int n_part,
                                                         double* __restrict__ ke_next_part, double* __restrict__ ke_half_next_part,
                                                         float* __restrict__ traj_v, float* __restrict__ traj_q,
                                                         float* __restrict__ traj_pv_row, const float4* __restrict__ qref,
                                                         int check_skin_next, int* __restrict__ flags, DistArgs D) {
    __shared__ double sm[INT_THREADS / 32];
    __shared__ float bc[2];
    __shared__ float s_pvh0, s_pvn0;
    float pvh0 = 0.f, pvn0 = 0.f, Q0 = 1.f;
    dist_ack(D);
    dist_wait_pulled(D);
    if (A.integrator == MDG_INT_NHC) {
        float ke0, ke1;
        if (D.mine) dist_gather_ke(D, sm, &ke0, &ke1);
        else {
            ke0 = sum_partials(ke_part, n_part, sm, &bc[0]);
            ke1 = sum_partials(ke_half_part, n_part, sm, &bc[1]);
        }
        if (threadIdx.x == 0) {
            float pv[MDG_MAX_CHAINS], ph[MDG_MAX_CHAINS], pvh[MDG_MAX_CHAINS], d0[MDG_MAX_CHAINS], d1[MDG_MAX_CHAINS];
            for (int k = 0; k < A.M; ++k) pv[k] = sc->pv[pv_sel][k];
            nhc_dpv(A, ke0, pv, d0);
            for (int k = 0; k < A.M; ++k) { ph[k] = 0.5f * d0[k] * dt; pvh[k] = pv[k] + ph[k]; }
            nhc_dpv(A, ke1, pvh, d1);
            s_pvh0 = pvh[0];
            s_pvn0 = pv[0] + (ph[0] + 0.5f * d1[0] * dt);
            if (blockIdx.x == 0) {
                for (int k = 0; k < A.M; ++k) {
                    float pn = pv[k] + (ph[k] + 0.5f * d1[k] * dt);
                    sc->pv[pv_sel ^ 1][k] = pn;
                    if (traj_pv_row) traj_pv_row[k] = pn;
                }
            }
        }
        __syncthreads();
        pvh0 = s_pvh0;
        pvn0 = s_pvn0;
        Q0 = A.Q[0];
    }
    double acc_v = 0, acc_h = 0;
    bool viol = false;
    for (int s = A.s0 + blockIdx.x * blockDim.x + threadIdx.x; s < A.s1; s += gridDim.x * blockDim.x) {
        float4 v = v4[s];
        float4 h = vh4[s];
        float4 f = f4[s];
        float4 q = q4[s];
        float m = v.w;
        float ax, ay, az;
        // ---- B(n) ----
        if (A.integrator == MDG_INT_NHC) {
            float px = (v.x + h.x) * m, py = (v.y + h.y) * m, pz = (v.z + h.z) * m;
            ax = (f.x - pvh0 * px / Q0) / m;
            ay = (f.y - pvh0 * py / Q0) / m;
            az = (f.z - pvh0 * pz / Q0) / m;
        } else {
            ax = f.x; ay = f.y; az = f.z;
        }
        v.x = v.x + (h.x + 0.5f * ax * dt);
        v.y = v.y + (h.y + 0.5f * ay * dt);
        v.z = v.z + (h.z + 0.5f * az * dt);
        v4[s] = v;
        if (traj_v) {
            int id = __float_as_int(q.w);
            traj_v[3 * (size_t)id] = v.x; traj_v[3 * (size_t)id + 1] = v.y; traj_v[3 * (size_t)id + 2] = v.z;
            traj_q[3 * (size_t)id] = q.x; traj_q[3 * (size_t)id + 1] = q.y; traj_q[3 * (size_t)id + 2] = q.z;
        }
        // ---- A(n+1) ----
        float px = v.x * m, py = v.y * m, pz = v.z * m;
        if (A.integrator == MDG_INT_NHC) {
            acc_v += (double)(px * px / m) + (double)(py * py / m) + (double)(pz * pz / m);
            ax = (f.x - pvn0 * px / Q0) / m;
            ay = (f.y - pvn0 * py / Q0) / m;
            az = (f.z - pvn0 * pz / Q0) / m;
        } else {
            ax = f.x; ay = f.y; az = f.z;
        }
        float hx = 0.5f * ax * dt_next, hy = 0.5f * ay * dt_next, hz = 0.5f * az * dt_next;
        float ux = v.x + hx, uy = v.y + hy, uz = v.z + hz;
        q.x = q.x + ux * dt_next;
        q.y = q.y + uy * dt_next;
        q.z = q.z + uz * dt_next;
        vh4[s] = make_float4(hx, hy, hz, 0.f);
        q4[s] = q;
        if (A.integrator == MDG_INT_NHC) {
            float qx = ux * m, qy = uy * m, qz = uz * m;
            acc_h += (double)(qx * qx / m) + (double)(qy * qy / m) + (double)(qz * qz / m);
        }
        if (check_skin_next) {
            float4 r = qref[s];
            float dx = q.x - r.x, dy = q.y - r.y, dz = q.z - r.z;
            viol |= (dx * dx + dy * dy + dz * dz) > A.half_skin2;
        }
    }
    if (viol) flags[5] = 1;
    if (A.integrator == MDG_INT_NHC) {
        double t = block_sum_double(acc_v, sm);
        if (threadIdx.x == 0) ke_next_part[blockIdx.x] = 0.5 * t;
        double t2 = block_sum_double(acc_h, sm);
        if (threadIdx.x == 0) ke_half_next_part[blockIdx.x] = 0.5 * t2;
    }
    dist_raise_pos(D);
}

// state (re)ordering ---------------------------------------------------------------------------
__global__ void k_init_v(int s0, int n, const int* __restrict__ perm, const float* __restrict__ v0,
                         const float* __restrict__ mass, float4* __restrict__ v4) {
    int s = s0 + blockIdx.x * blockDim.x + threadIdx.x;       // own range [s0, n): perm is only defined there in slab mode
    if (s >= n) return;
    int i = perm[s];
    v4[s] = make_float4(v0[3 * i], v0[3 * i + 1], v0[3 * i + 2], mass[i]);
}

__global__ void k_permute2(int s0, int n, const int* __restrict__ perm, const float4* __restrict__ a_in, float4* __restrict__ a_out,
                           const float4* __restrict__ b_in, float4* __restrict__ b_out) {
    int s = s0 + blockIdx.x * blockDim.x + threadIdx.x;      // new sorted range [s0, n)
    if (s >= n) return;
    int i = perm[s];
    a_out[s] = a_in[i];
    b_out[s] = b_in[i];
}

// multi-GPU frame 0: each rank writes the atoms it owns (frames are summed across ranks by the caller)
__global__ void k_frame_own(int s0, int s1, const float4* __restrict__ v4, const float4* __restrict__ q4,
                            float* __restrict__ traj_v, float* __restrict__ traj_q) {
    int s = s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= s1) return;
    float4 q = q4[s], v = v4[s];
    int id = __float_as_int(q.w);
    traj_v[3 * (size_t)id] = v.x; traj_v[3 * (size_t)id + 1] = v.y; traj_v[3 * (size_t)id + 2] = v.z;
    traj_q[3 * (size_t)id] = q.x; traj_q[3 * (size_t)id + 1] = q.y; traj_q[3 * (size_t)id + 2] = q.z;
}

__global__ void __launch_bounds__(256) k_energy_sum(int s0, int s1, const float4* __restrict__ fs, double* __restrict__ part) {
    __shared__ double sm[8];
    double v = 0;
    for (int s = s0 + blockIdx.x * blockDim.x + threadIdx.x; s < s1; s += gridDim.x * blockDim.x) v += (double)fs[s].w;
    double t = block_sum_double(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU communication steps (NCCL on the engine's stream)
// ---------------------------------------------------------------------------------------------
static int owner_of_layer(int z, int ncz, int world) {
    int plan[4];
    for (int r = 0; r < world; ++r) {
        mdg_slab_plan(ncz, world, r, plan);
        if (z >= plan[0] && z < plan[1]) return r;
    }
    return 0;
}

// ghost positions: my bottom layer -> rank below, my top layer -> rank above; receive their counterparts
static int halo_exchange_nogroup(mdg_ctx* c, float4* q, cudaStream_t st) {
    NcclApi* N = mdg_nccl();
    const int ncz = c->n_layers - 1, R = c->dist_world;
    const int zlo = c->slab_zlo, zhi = c->slab_zhi;
    const int zl = (zlo - 1 + ncz) % ncz, zu = zhi % ncz;             // ghost layers
    const int below = owner_of_layer(zl, ncz, R), above = owner_of_layer(zu, ncz, R);
    const int* L = c->h_layers;
    // sends (own boundary layers)
    MDG_TRY(mdg_nccl_check(N->Send(q + L[zlo], (size_t)(L[zlo + 1] - L[zlo]) * 4, MDG_NCCL_FLOAT32, below, c->dist_comm, st), "Send"));
    MDG_TRY(mdg_nccl_check(N->Send(q + L[zhi - 1], (size_t)(L[zhi] - L[zhi - 1]) * 4, MDG_NCCL_FLOAT32, above, c->dist_comm, st), "Send"));
    // receives (ghost layers); message order per peer: the peer's "up" send matches my "below" recv
    MDG_TRY(mdg_nccl_check(N->Recv(q + L[zu], (size_t)(L[zu + 1] - L[zu]) * 4, MDG_NCCL_FLOAT32, above, c->dist_comm, st), "Recv"));
    MDG_TRY(mdg_nccl_check(N->Recv(q + L[zl], (size_t)(L[zl + 1] - L[zl]) * 4, MDG_NCCL_FLOAT32, below, c->dist_comm, st), "Recv"));
    return MDG_OK;
}

static int halo_exchange(mdg_ctx* c, float4* q, cudaStream_t st) {
    NcclApi* N = mdg_nccl();
    MDG_TRY(mdg_nccl_check(N->GroupStart(), "GroupStart"));
    MDG_TRY(halo_exchange_nogroup(c, q, st));
    MDG_TRY(mdg_nccl_check(N->GroupEnd(), "GroupEnd"));
    return MDG_OK;
}

// Rebuild exchange (old slab plan): each rank sends its bottom / top TWO layers of (q, v, vh) to the rank below /
// above and receives theirs in place (same global index ranges on both sides).  Afterwards a rank holds valid state
// on the old layers [zlo-2, zhi+2) - a superset of everything that can be in its new slab + ghost layers.
static int state_exchange(mdg_ctx* c, float4* q, float4* v, float4* vh, cudaStream_t st) {
    NcclApi* N = mdg_nccl();
    const int ncz = c->n_layers - 1, R = c->dist_world;
    const int zlo = c->slab_zlo, zhi = c->slab_zhi;
    if (zhi - zlo < 2) { mdg_set_error("distributed run: every rank needs >= 2 cell layers (ncz=%d, world=%d)", ncz, R); return MDG_E_BADARG; }
    const int below = owner_of_layer((zlo - 1 + ncz) % ncz, ncz, R), above = owner_of_layer(zhi % ncz, ncz, R);
    const int* L = c->h_layers;
    const int rl0 = (zlo - 2 + ncz) % ncz, ru0 = zhi % ncz;      // first of the two layers received from below / above
    float4* arr[3] = {q, v, vh};
    MDG_TRY(mdg_nccl_check(N->GroupStart(), "GroupStart"));
    for (int a = 0; a < 3; ++a) {
        float4* x = arr[a];
        MDG_TRY(mdg_nccl_check(N->Send(x + L[zlo], (size_t)(L[zlo + 2] - L[zlo]) * 4, MDG_NCCL_FLOAT32, below, c->dist_comm, st), "Send"));
        MDG_TRY(mdg_nccl_check(N->Send(x + L[zhi - 2], (size_t)(L[zhi] - L[zhi - 2]) * 4, MDG_NCCL_FLOAT32, above, c->dist_comm, st), "Send"));
        MDG_TRY(mdg_nccl_check(N->Recv(x + L[ru0], (size_t)(L[ru0 + 2] - L[ru0]) * 4, MDG_NCCL_FLOAT32, above, c->dist_comm, st), "Recv"));
        MDG_TRY(mdg_nccl_check(N->Recv(x + L[rl0], (size_t)(L[rl0 + 2] - L[rl0]) * 4, MDG_NCCL_FLOAT32, below, c->dist_comm, st), "Recv"));
    }
    MDG_TRY(mdg_nccl_check(N->GroupEnd(), "GroupEnd"));
    return MDG_OK;
}

// ---------------------------------------------------------------------------------------------
// epoch driver
// ---------------------------------------------------------------------------------------------
static int run_once(mdg_ctx* c, const mdg_md_params* p, int n, const float* d_mass, const float* d_v0,
                    const float* d_q0, const float* h_pv0, const float* h_tgrid, int n_grid, float* d_traj_v,
                    float* d_traj_q, float* h_traj_pv, float* h_last_energy, int rebuild_every, cudaStream_t st) {
    const int nhc = p->integrator == MDG_INT_NHC;
    const int M = nhc ? p->n_chains : 0;
    const int stride = p->traj_stride < 1 ? 1 : p->traj_stride;
    const double rlist = p->cutoff + (double)p->skin;
    const bool retest = p->skin > 0.f;
    const bool dist = c->dist_world > 1;
    NcclApi* N = mdg_nccl();
    if (dist && !retest) { mdg_set_error("multi-GPU runs need a Verlet skin (skin > 0)"); return MDG_E_BADARG; }
    if (dist && !c->comm_stream) {
        // highest priority: the NCCL kernels must get SM slots while the (long) interior force kernel is running
        int prio_lo = 0, prio_hi = 0;
        MDG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        MDG_CUDA(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, prio_hi));
        MDG_CUDA(cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
        MDG_CUDA(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
        MDG_CUDA(cudaEventCreateWithFlags(&c->ev_ke, cudaEventDisableTiming));
        MDG_CUDA(cudaStreamCreateWithPriority(&c->bnd_stream, cudaStreamNonBlocking, prio_hi));
        MDG_CUDA(cudaEventCreateWithFlags(&c->ev_bnd, cudaEventDisableTiming));
    }
    const char* bz = getenv("MDG_DIST_BND_STREAM");
    const char* pz = getenv("MDG_DIST_PUSH_SIDE");
    const bool push_side = !(pz && pz[0] == '0');     // peer-to-peer push on the communication stream (0: main stream, ahead of the forces)
    const char* plz = getenv("MDG_DIST_PULL");
    const char* pbz = getenv("MDG_DIST_PUSH_BLOCKS");
    const int push_blocks_env = pbz ? atoi(pbz) : 0;
    const char* gz = getenv("MDG_DIST_GATE");
    const bool gate = !(gz && gz[0] == '0');          // interior rows wait until the push kernel has started
    const bool bnd_side = !(bz && bz[0] == '0');     // boundary layers on their own stream, concurrent with the interior rows
    // Ghost layers: PULLED by the consumer (NVLink loads once the neighbour's positions are final) for slabs below ~190 000 atoms per
    // rank, PUSHED by the producer above - measured on 2 x B200: 131 072 atoms per GPU 83.5 (pull) vs 86.3 us per step (push),
    // 256 000 per GPU 122.3 vs 116.1 (the longer interior launch hides the push; the pull's spinning CTAs and the ticket in the
    // integrator kernel then only cost).  n and the world size are the same on every rank, so all ranks take the same path.
    // MDG_DIST_PULL=1 / 0 forces one.
    const bool pull_auto = dist && (n / (c->dist_world > 0 ? c->dist_world : 1)) < 190000;
    const bool pull = (plz ? plz[0] != '0' : pull_auto) && bnd_side;
    IntArgs A;
    memset(&A, 0, sizeof(A));
    A.integrator = p->integrator;
    A.M = M;
    for (int k = 0; k < M; ++k) A.Q[k] = p->Q[k];
    A.T = (float)p->T;
    A.target = (float)(p->T * (double)p->ndof * 0.5);
    A.half_skin2 = 0.25f * p->skin * p->skin;
    PotParams P = mdg_make_pot(p->pot_kind, p->pot_params, MDG_MAX_POT_PARAMS);

    const int T = 256;
    const int n_frames = (n_grid - 1) / stride + 1;

    MDG_TRY(c->v4.reserve(2 * sizeof(float4) * (size_t)n));
    MDG_TRY(c->vh4.reserve(2 * sizeof(float4) * (size_t)n));
    MDG_TRY(c->qref.reserve(sizeof(float4) * (size_t)n));
    MDG_TRY(c->fs.reserve(sizeof(float4) * (size_t)n));
    MDG_TRY(c->pvbuf.reserve(sizeof(Scalars) + sizeof(float) * (size_t)n_frames * MDG_MAX_CHAINS));
    MDG_TRY(c->kebuf.reserve(sizeof(double) * (5 * INT_MAX_BLOCKS + 8)));
    float4* vbuf[2] = {c->v4.as<float4>(), c->v4.as<float4>() + n};
    float4* hbuf[2] = {c->vh4.as<float4>(), c->vh4.as<float4>() + n};
    int vsel = 0;
    Scalars* sc = c->pvbuf.as<Scalars>();
    float* d_traj_pv = (float*)(sc + 1);
    double* kb = c->kebuf.as<double>();
    double* e_part = kb + 4 * INT_MAX_BLOCKS;
    double* dke = kb + 5 * INT_MAX_BLOCKS;          // [0]=ke(v), [1]=ke(v+vh), global (multi-GPU)
    // kinetic-energy partial arrays: [cur_v, cur_half] are read by B(n), [nxt_v, nxt_half] written by B(n) / A(n+1)
    double* ke_v_cur = kb, *ke_h_cur = kb + INT_MAX_BLOCKS, *ke_v_nxt = kb + 2 * INT_MAX_BLOCKS, *ke_h_nxt = kb + 3 * INT_MAX_BLOCKS;
    const char* fz = getenv("MDG_FUSED_STEP");
    const bool fused = !(fz && fz[0] == '0');

    // scalars + flags
    Scalars hs;
    memset(&hs, 0, sizeof(hs));
    for (int k = 0; k < M; ++k) hs.pv[0][k] = h_pv0 ? h_pv0[k] : 0.f;
    MDG_CUDA(cudaMemcpyAsync(sc, &hs, sizeof(Scalars), cudaMemcpyHostToDevice, st));
    MDG_TRY(c->flags.reserve(sizeof(int) * 8));
    MDG_CUDA(cudaMemsetAsync(c->flags.p, 0, sizeof(int) * 8, st));
    c->flags_sticky = true;      // the rebuilds of this epoch accumulate into the flags; they are read once at its end

    // initial sort + list at q0, state into sorted order (every rank holds the full inputs)
    c->sel_a = c->eng_sel_a; c->sel_b = c->eng_sel_b; c->ex_keys = c->eng_ex_keys; c->n_ex = c->eng_n_ex;
    c->rows_wanted = true;
    c->fast_build = retest;      // skin list: every entry is re-tested exactly by the force kernel
    c->slab = dist;
    MDG_TRY(mdg_i_build_list(c, d_q0, nullptr, n, p->cell, rlist, p->cutoff, st));
    if (dist) MDG_TRY(mdg_i_dist_p2p_setup(c, st));      // (no-op once the mappings exist; collective when it is not)
    float4* q = c->qs_ptr;
    if (c->own_s1 > c->own_s0)
        k_init_v<<<(c->own_s1 - c->own_s0 + T - 1) / T, T, 0, st>>>(c->own_s0, c->own_s1, c->perm.as<int>(), d_v0, d_mass, vbuf[vsel]);
    MDG_CUDA(cudaMemcpyAsync(c->qref.p, q, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    MDG_CUDA(cudaMemsetAsync(hbuf[vsel], 0, sizeof(float4) * (size_t)n, st));
    c->stat_launches += 1;
    A.s0 = c->own_s0; A.s1 = c->own_s1;
    int nown = A.s1 - A.s0;
    int ib = (nown + T - 1) / T;
    ib = ib < 1 ? 1 : (ib > INT_MAX_BLOCKS ? INT_MAX_BLOCKS : ib);
    // the per-atom energy is only read after the LAST force evaluation of the epoch (h_last_energy)
    const bool energy_free = getenv("MDG_FORCE_ENERGY_ALWAYS") == nullptr;
    c->force_energy = !(energy_free && n_grid > 1);
    MDG_TRY(mdg_i_force_sorted(c, P, q, c->fs.as<float4>(), retest, false, nullptr, st));
    c->force_energy = true;
    if (dist && c->dist_p2p) {
        const int W = c->dist_world, me = c->dist_rank;
        ++c->dist_seq;                     // the epoch-start pseudo-step: the first real step waits for THIS acknowledgement
        k_dist_ack<<<1, 32, 0, st>>>((DistSync*)c->peer_sync[(me - 1 + W) % W], (DistSync*)c->peer_sync[(me + 1) % W], c->dist_seq);
        c->stat_launches++;
    }
    if (nhc) { k_ke_init<<<ib, INT_THREADS, 0, st>>>(A, vbuf[vsel], ke_v_cur); c->stat_launches++; }
    if (!dist) {   // frame 0 = the initial state, verbatim
        MDG_CUDA(cudaMemcpyAsync(d_traj_v, d_v0, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToDevice, st));
        MDG_CUDA(cudaMemcpyAsync(d_traj_q, d_q0, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    } else {       // owned atoms only; frames are assembled by summing over ranks
        MDG_CUDA(cudaMemsetAsync(d_traj_v, 0, sizeof(float) * 3 * (size_t)n * n_frames, st));
        MDG_CUDA(cudaMemsetAsync(d_traj_q, 0, sizeof(float) * 3 * (size_t)n * n_frames, st));
        if (nown > 0) k_frame_own<<<(nown + T - 1) / T, T, 0, st>>>(A.s0, A.s1, vbuf[vsel], q, d_traj_v, d_traj_q);
    }
    if (M) MDG_CUDA(cudaMemcpyAsync(d_traj_pv, sc->pv[0], sizeof(float) * M, cudaMemcpyDeviceToDevice, st));

    // Loop structure.  Unfused: A(g) ; [rebuild | halo] ; force ; B(g).   Fused (default): A(0) once, then per step
    // [rebuild | halo] ; force ; BA(g) = B(g) + A(g+1) in one pass (plain B for the last step).
    int pv_sel = 0;
    int ib_prev = ib;                       // block count of the launch(es) that wrote the *_cur partial arrays
    const int nsteps = n_grid - 1;
    bool a_done = false;                    // A(g) already executed by the previous fused kernel
    bool prev_split_pull = false;           // pull path: the previous step pulled its ghosts on the boundary stream (split step)
    int last_pull_seq = 0;                  // pull path: sequence number of the last step whose ghosts were pulled (0: none pending)
    for (int g = 0; g < nsteps; ++g) {
        float dt = h_tgrid[g + 1] - h_tgrid[g];          // fp32 subtraction, like t1 - t0 in tinydiffeq.py:67-68
        bool do_rebuild = ((g + 1) % rebuild_every) == 0;
        if (!a_done) {
            DistArgs DAa{};
            if (dist && c->dist_p2p && pull) {      // pull path: A overwrites positions the neighbours copy, and produces the next ones
                const int W = c->dist_world, me = c->dist_rank;
                DAa.mine = (DistSync*)c->dsync.p;
                DAa.below = (DistSync*)c->peer_sync[(me - 1 + W) % W];
                DAa.above = (DistSync*)c->peer_sync[(me + 1) % W];
                DAa.world = W;
                DAa.pos_seq = c->dist_seq + 1;
                DAa.wait_pull = last_pull_seq;
            }
            k_step_a<<<ib, INT_THREADS, 0, st>>>(A, dt, pv_sel, sc, vbuf[vsel], hbuf[vsel], q, c->fs.as<float4>(),
                                                 c->qref.as<float4>(), (retest && !do_rebuild) ? 1 : 0, ke_h_cur,
                                                 c->flags.as<int>(), DAa);
            last_pull_seq = 0;
            c->stat_launches++;
        }
        tl_mark(g, "step_begin", st);
        if (do_rebuild) {
            if (dist && c->dist_p2p && c->peer_v[0]) {
                // peer-to-peer: one push kernel + one wait kernel (dist.cu maps the neighbours' v / vh arrays as well)
                const int* Ly = c->h_layers;
                const int zlo = c->slab_zlo, zhi = c->slab_zhi;
                const int W = c->dist_world, me = c->dist_rank, below = (me - 1 + W) % W, above = (me + 1) % W;
                if (zhi - zlo < 2) { mdg_set_error("distributed run: every rank needs >= 2 cell layers"); return MDG_E_BADARG; }
                const int sel = (q == c->qs_buf[0].as<float4>()) ? 0 : 1;
                const size_t voff = (size_t)(vbuf[vsel] - c->v4.as<float4>()), hoff = (size_t)(hbuf[vsel] - c->vh4.as<float4>());
                const int nst = (Ly[zlo + 2] - Ly[zlo]) + (Ly[zhi] - Ly[zhi - 2]);
                int pb = (nst + 1023) / 1024;
                pb = pb < 1 ? 1 : (pb > 96 ? 96 : pb);
                const int rseq = c->dist_seq + 1;                  // the sequence number this step's push will carry
                k_dist_push_state<<<pb, 256, 0, st>>>(q, vbuf[vsel], hbuf[vsel], Ly[zlo], Ly[zlo + 2], Ly[zhi - 2], Ly[zhi],
                                                     (float4*)c->peer_qs[0][sel], (float4*)c->peer_v[0] + voff, (float4*)c->peer_vh[0] + hoff,
                                                     (float4*)c->peer_qs[1][sel], (float4*)c->peer_v[1] + voff, (float4*)c->peer_vh[1] + hoff,
                                                     (DistSync*)c->dsync.p, (DistSync*)c->peer_sync[below], (DistSync*)c->peer_sync[above], rseq);
                k_dist_wait_state<<<1, 32, 0, st>>>((DistSync*)c->dsync.p, rseq);
                c->stat_launches += 2;
            } else if (dist) MDG_TRY(state_exchange(c, q, vbuf[vsel], hbuf[vsel], st));
            tl_mark(g, "rb_xchg_end", st);
            c->slab_local = dist;
            MDG_TRY(mdg_i_build_list(c, nullptr, q, n, p->cell, rlist, p->cutoff, st));
            c->slab_local = false;
            tl_mark(g, "rb_build_end", st);
            q = c->qs_ptr;
            A.s0 = c->own_s0; A.s1 = c->own_s1;
            nown = A.s1 - A.s0;
            if (c->path == 0) {    // v, vh follow their atoms into the new order (own range only: ghosts carry positions only)
                if (nown > 0)
                    k_permute2<<<(nown + T - 1) / T, T, 0, st>>>(A.s0, A.s1, c->perm.as<int>(), vbuf[vsel], vbuf[vsel ^ 1],
                                                               hbuf[vsel], hbuf[vsel ^ 1]);
                c->stat_launches++;
                vsel ^= 1;
            }
            ib = (nown + T - 1) / T;
            ib = ib < 1 ? 1 : (ib > INT_MAX_BLOCKS ? INT_MAX_BLOCKS : ib);
            if (retest && nown > 0)
                MDG_CUDA(cudaMemcpyAsync(c->qref.as<float4>() + A.s0, q + A.s0, sizeof(float4) * (size_t)nown, cudaMemcpyDeviceToDevice, st));
        }
        // ---- communication + forces ------------------------------------------------------------------------
        //  single GPU            : force on all rows.
        //  multi GPU, rebuild    : state_exchange above already refreshed everything -> KE all-reduce, force on all rows.
        //  multi GPU, plain step : side stream = ghost-layer halo, then the 2-double KE all-reduce; main stream =
        //                          forces of the INTERIOR layers (no ghost needed) meanwhile, then the two boundary
        //                          layers once the ghosts arrived, then B once the kinetic energies arrived.
        const double* ke_a = ke_v_cur;
        const double* ke_b = ke_h_cur;
        int n_part = ib_prev;
        DistArgs DA{nullptr, nullptr, nullptr, 1, 0};
        bool split_pull_now = false;
        c->force_energy = !(energy_free && g + 1 < nsteps);
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (c->prof_enable) {
            std::vector<cudaEvent_t>* pool = (std::vector<cudaEvent_t>*)c->prof_events;
            if (!pool) { pool = new std::vector<cudaEvent_t>(); c->prof_events = pool; }
            if (c->prof_used + 2 <= PROF_MAX_EVENTS) {
                while ((int)pool->size() < c->prof_used + 2) {
                    cudaEvent_t e;
                    MDG_CUDA(cudaEventCreate(&e));
                    pool->push_back(e);
                }
                ev0 = (*pool)[c->prof_used];
                ev1 = (*pool)[c->prof_used + 1];
                c->prof_used += 2;
            }
        }
        if (!dist) {
            if (ev0) MDG_CUDA(cudaEventRecord(ev0, st));
            MDG_TRY(mdg_i_force_sorted(c, P, q, c->fs.as<float4>(), retest, false, nullptr, st));
            if (ev1) MDG_CUDA(cudaEventRecord(ev1, st));
        } else {
            cudaStream_t cs = c->comm_stream;
            MDG_CUDA(cudaEventRecord(c->ev_a, st));                    // positions + KE partials of this step are ready
            if (!(c->dist_p2p && !push_side)) MDG_CUDA(cudaStreamWaitEvent(cs, c->ev_a, 0));
            const int* Ly = c->h_layers;
            const int zlo = c->slab_zlo, zhi = c->slab_zhi;
            const int ncz = c->n_layers - 1;
            const bool p2p = c->dist_p2p && ((zlo - 1 + ncz) % ncz != zhi % ncz);     // (distinct ghost layers below / above)
            bool split = !do_rebuild && (zhi - zlo) >= 3;
            if (p2p) {
                // ---- peer-to-peer: one push kernel (NVLink stores + flags), no NCCL kernel on the step path ----------------
                const int seq = ++c->dist_seq;
                const int W = c->dist_world, me = c->dist_rank, below = (me - 1 + W) % W, above = (me + 1) % W;
                const int sel = (q == c->qs_buf[0].as<float4>()) ? 0 : 1;
                PeerTab PT;
                for (int r = 0; r < MDG_DIST_MAXW; ++r) PT.s[r] = (DistSync*)c->peer_sync[r < W ? r : me];
                const int nh = (Ly[zlo + 1] - Ly[zlo]) + (Ly[zhi] - Ly[zhi - 1]);
                int hb = do_rebuild ? 0 : (nh + 511) / 512;          // blocks that copy ghost layers
                hb = (do_rebuild || pull) ? 0 : (hb < 1 ? 1 : (hb > 48 ? 48 : hb));
                if (push_blocks_env > 0 && hb > 0) hb = push_blocks_env;
                // The push runs on the communication stream beside the interior rows; MDG_DIST_PUSH_SIDE=0 puts it on the main
                // stream ahead of them (measured: 8 us per step slower - the push is ~15 us of fence / flag latency, not bytes).
                cudaStream_t ps = push_side ? cs : st;
                if (hb + (nhc ? 1 : 0) > 0) {
                    k_dist_push<<<hb + (nhc ? 1 : 0), 256, 0, ps>>>(q, Ly[zlo], Ly[zlo + 1], Ly[zhi - 1], Ly[zhi], (float4*)c->peer_qs[0][sel],
                                                                   (float4*)c->peer_qs[1][sel], ke_v_cur, ke_h_cur, ib_prev, nhc, PT, me, W, below,
                                                                   above, seq, hb);
                    if (push_side && split && gate && !pull && (Ly[zhi - 1] - Ly[zlo + 1]) < 160000) { k_dist_wait_started<<<1, 32, 0, st>>>((DistSync*)c->dsync.p, seq); c->stat_launches++; }
                }
                if (push_side) MDG_CUDA(cudaEventRecord(c->ev_push, cs));
                tl_mark(g, "push_end", ps);
                c->stat_launches++;
                DA.mine = (DistSync*)c->dsync.p;
                DA.below = (DistSync*)c->peer_sync[below];
                DA.above = (DistSync*)c->peer_sync[above];
                DA.world = W;
                DA.seq = seq;
                if (ev0) MDG_CUDA(cudaEventRecord(ev0, st));
                if (split) {
                    // The two boundary layers need the ghosts, the interior layers do not.  Boundary side stream: wait kernel
                    // (spins on the halo flags) + one launch over the bottom and top layer; main stream: the interior rows.  The
                    // boundary CTAs fill in as the interior's last wave drains, so a slab step costs about one full force launch
                    // (serial wait + small boundary launch behind the interior rows: +10 us per step in the r02 measurements).
                    const int nxy = c->nc[0] * c->nc[1];
                    cudaStream_t bs = bnd_side ? c->bnd_stream : st;
                    // Pull path: the copy kernel does NOT wait for this rank's own integrator kernel - it is enqueued behind the
                    // previous step's boundary rows (the last readers of the ghost ranges, same stream) and spins until the
                    // neighbours' positions are final, so the ghosts are here ~one NVLink round trip after the neighbours' B + A
                    // kernels end instead of a stream hop (5-13 us in the r02 stamps) + kernel later.  After a rebuild step (or at
                    // the start of an epoch) the last reader was the full force launch on the main stream: wait for ev_a then.
                    const bool early_pull = pull && prev_split_pull;
                    if (bnd_side && !early_pull) MDG_CUDA(cudaStreamWaitEvent(bs, c->ev_a, 0));
                    if (!bnd_side) MDG_TRY(mdg_i_force_range(c, P, q, c->fs.as<float4>(), retest, Ly[zlo + 1], Ly[zhi - 1], (zlo + 1) * nxy,
                                                             (zhi - 1) * nxy, st));                                       // interior
                    if (pull) {
                        const int zl = (zlo - 1 + ncz) % ncz, zu = zhi % ncz;
                        const int ng = (Ly[zl + 1] - Ly[zl]) + (Ly[zu + 1] - Ly[zu]);
                        int pk = (ng + 255) / 256;                       // one NVLink load per thread: one round trip
                        pk = pk < 1 ? 1 : (pk > 96 ? 96 : pk);
                        k_dist_pull<<<pk, 256, 0, bs>>>(q, (const float4*)c->peer_qs[0][sel], (const float4*)c->peer_qs[1][sel], Ly[zl], Ly[zl + 1],
                                                       Ly[zu], Ly[zu + 1], (DistSync*)c->dsync.p, (DistSync*)c->peer_sync[below],
                                                       (DistSync*)c->peer_sync[above], seq);
                        last_pull_seq = seq;
                        if (early_pull) MDG_CUDA(cudaStreamWaitEvent(bs, c->ev_a, 0));      // the boundary rows need MY positions too
                        split_pull_now = true;
                    } else
                        k_dist_wait<<<1, 32, 0, bs>>>((DistSync*)c->dsync.p, seq);                                         // ghosts landed
                    tl_mark(g, "wait_end", bs);
                    if (!c->tiles) {   // bottom + top layer in one launch
                        MDG_TRY(mdg_i_force_range2(c, P, q, c->fs.as<float4>(), retest, Ly[zlo], Ly[zlo + 1], Ly[zhi - 1], Ly[zhi], bs));
                    } else {
                        MDG_TRY(mdg_i_force_range(c, P, q, c->fs.as<float4>(), retest, Ly[zlo], Ly[zlo + 1], zlo * nxy,
                                                  (zlo + 1) * nxy, bs));                                                  // bottom layer
                        MDG_TRY(mdg_i_force_range(c, P, q, c->fs.as<float4>(), retest, Ly[zhi - 1], Ly[zhi], (zhi - 1) * nxy,
                                                  zhi * nxy, bs));                                                        // top layer
                    }
                    tl_mark(g, "bnd_force_end", bs);
                    if (bnd_side) {
                        MDG_CUDA(cudaEventRecord(c->ev_bnd, bs));
                        MDG_TRY(mdg_i_force_range(c, P, q, c->fs.as<float4>(), retest, Ly[zlo + 1], Ly[zhi - 1], (zlo + 1) * nxy,
                                                  (zhi - 1) * nxy, st));                                                  // interior
                        tl_mark(g, "int_force_end", st);
                        MDG_CUDA(cudaStreamWaitEvent(st, c->ev_bnd, 0));
                    }
                    c->stat_launches++;
                } else {
                    if (!do_rebuild && pull) {
                        const int zl = (zlo - 1 + ncz) % ncz, zu = zhi % ncz;
                        const int ng = (Ly[zl + 1] - Ly[zl]) + (Ly[zu + 1] - Ly[zu]);
                        int pk = (ng + 511) / 512;
                        pk = pk < 1 ? 1 : (pk > 48 ? 48 : pk);
                        k_dist_pull<<<pk, 256, 0, st>>>(q, (const float4*)c->peer_qs[0][sel], (const float4*)c->peer_qs[1][sel], Ly[zl], Ly[zl + 1],
                                                       Ly[zu], Ly[zu + 1], (DistSync*)c->dsync.p, (DistSync*)c->peer_sync[below],
                                                       (DistSync*)c->peer_sync[above], seq);
                        last_pull_seq = seq;
                        c->stat_launches++;
                    } else if (!do_rebuild) { k_dist_wait<<<1, 32, 0, st>>>((DistSync*)c->dsync.p, seq); c->stat_launches++; }
                    MDG_TRY(mdg_i_force_sorted(c, P, q, c->fs.as<float4>(), retest, false, nullptr, st));
                }
                if (ev1) MDG_CUDA(cudaEventRecord(ev1, st));
                if (push_side) MDG_CUDA(cudaStreamWaitEvent(st, c->ev_push, 0));      // B overwrites q: my own stores to the neighbours must have read it
            } else {
            if (!do_rebuild) {
                MDG_TRY(halo_exchange(c, q, cs));
                MDG_CUDA(cudaEventRecord(c->ev_halo, cs));
            }
            if (nhc) {
                k_ke_pack<<<1, INT_THREADS, 0, cs>>>(ke_v_cur, ib_prev, ke_h_cur, ib_prev, dke);
                MDG_TRY(mdg_nccl_check(N->AllReduce(dke, dke, 2, MDG_NCCL_FLOAT64, MDG_NCCL_SUM, c->dist_comm, cs), "AllReduce"));
                MDG_CUDA(cudaEventRecord(c->ev_ke, cs));
                ke_a = dke; ke_b = dke + 1;
                n_part = 1;
                c->stat_launches++;
            }
            if (ev0) MDG_CUDA(cudaEventRecord(ev0, st));
            if (split) {
                const int nxy = c->nc[0] * c->nc[1];
                cudaStream_t bs = bnd_side ? c->bnd_stream : st;
                if (!bnd_side) MDG_TRY(mdg_i_force_range(c, P, q, c->fs.as<float4>(), retest, Ly[zlo + 1], Ly[zhi - 1], (zlo + 1) * nxy,
                                                         (zhi - 1) * nxy, st));                                       // interior
                MDG_CUDA(cudaStreamWaitEvent(bs, c->ev_halo, 0));
                if (!c->tiles) {       // bottom + top layer in one launch
                    MDG_TRY(mdg_i_force_range2(c, P, q, c->fs.as<float4>(), retest, Ly[zlo], Ly[zlo + 1], Ly[zhi - 1], Ly[zhi], bs));
                } else {
                    MDG_TRY(mdg_i_force_range(c, P, q, c->fs.as<float4>(), retest, Ly[zlo], Ly[zlo + 1], zlo * nxy,
                                              (zlo + 1) * nxy, bs));                                                  // bottom layer
                    MDG_TRY(mdg_i_force_range(c, P, q, c->fs.as<float4>(), retest, Ly[zhi - 1], Ly[zhi], (zhi - 1) * nxy,
                                              zhi * nxy, bs));                                                        // top layer
                }
                if (bnd_side) {        // (see the peer-to-peer branch: boundary rows beside the interior rows)
                    MDG_CUDA(cudaEventRecord(c->ev_bnd, bs));
                    MDG_TRY(mdg_i_force_range(c, P, q, c->fs.as<float4>(), retest, Ly[zlo + 1], Ly[zhi - 1], (zlo + 1) * nxy,
                                              (zhi - 1) * nxy, st));                                                  // interior
                    MDG_CUDA(cudaStreamWaitEvent(st, c->ev_bnd, 0));
                }
            } else {
                if (!do_rebuild) MDG_CUDA(cudaStreamWaitEvent(st, c->ev_halo, 0));
                MDG_TRY(mdg_i_force_sorted(c, P, q, c->fs.as<float4>(), retest, false, nullptr, st));
            }
            if (ev1) MDG_CUDA(cudaEventRecord(ev1, st));
            if (nhc) MDG_CUDA(cudaStreamWaitEvent(st, c->ev_ke, 0));
            }
        }
        tl_mark(g, "force_end", st);
        c->force_energy = true;
        int gp = g + 1;
        bool keep = (gp % stride) == 0;
        size_t fr = (size_t)(gp / stride);
        float* tv = keep ? d_traj_v + fr * 3 * (size_t)n : nullptr;
        float* tq = keep ? d_traj_q + fr * 3 * (size_t)n : nullptr;
        float* tp = (keep && M) ? d_traj_pv + fr * M : nullptr;
        if (fused && g + 1 < nsteps) {
            float dt_next = h_tgrid[g + 2] - h_tgrid[g + 1];
            bool next_rebuild = ((g + 2) % rebuild_every) == 0;
            if (DA.mine && pull) {             // B + A: overwrites the positions the neighbours pulled in this step, produces the next ones
                DA.pos_seq = DA.seq + 1;
                DA.wait_pull = last_pull_seq;
                last_pull_seq = 0;
            }
            k_step_ba<<<ib, INT_THREADS, 0, st>>>(A, dt, dt_next, pv_sel, sc, vbuf[vsel], hbuf[vsel], q, c->fs.as<float4>(),
                                                  ke_a, ke_b, n_part, ke_v_nxt, ke_h_nxt, tv, tq, tp, c->qref.as<float4>(),
                                                  (retest && !next_rebuild) ? 1 : 0, c->flags.as<int>(), DA);
            a_done = true;
            { double* t1 = ke_v_cur; ke_v_cur = ke_v_nxt; ke_v_nxt = t1; }
            { double* t2 = ke_h_cur; ke_h_cur = ke_h_nxt; ke_h_nxt = t2; }
        } else {
            k_step_b<<<ib, INT_THREADS, 0, st>>>(A, dt, pv_sel, sc, vbuf[vsel], hbuf[vsel], q, c->fs.as<float4>(),
                                                 ke_a, ke_b, n_part, ke_v_nxt, tv, tq, tp, DA);
            a_done = false;
            { double* t1 = ke_v_cur; ke_v_cur = ke_v_nxt; ke_v_nxt = t1; }
        }
        c->stat_launches++;
        tl_mark(g, "step_end", st);
        prev_split_pull = split_pull_now;
        pv_sel ^= 1;
        ib_prev = ib;
    }
    if (h_last_energy) {
        k_energy_sum<<<ib, 256, 0, st>>>(A.s0, A.s1, c->fs.as<float4>(), e_part);
        k_ke_pack<<<1, INT_THREADS, 0, st>>>(e_part, ib, e_part, 0, dke + 6);
        if (dist) MDG_TRY(mdg_nccl_check(N->AllReduce(dke + 6, dke + 6, 1, MDG_NCCL_FLOAT64, MDG_NCCL_SUM, c->dist_comm, st), "AllReduce"));
        c->stat_launches += 2;
    }
    int h_p2p_timeout = 0;
    if (dist && c->dist_p2p)
        MDG_CUDA(cudaMemcpyAsync(&h_p2p_timeout, c->dsync.as<char>() + offsetof(DistSync, pad) + sizeof(int), sizeof(int),
                                 cudaMemcpyDeviceToHost, st));
    if (dist)   // all ranks must take the same retry decision
        MDG_TRY(mdg_nccl_check(N->AllReduce(c->flags.p, c->flags.p, 8, MDG_NCCL_INT32, MDG_NCCL_MAX, c->dist_comm, st), "AllReduce"));
    MDG_KERNEL_CHECK();
    // read-backs (SYNC)
    double h_e = 0;
    if (h_last_energy) MDG_CUDA(cudaMemcpyAsync(&h_e, dke + 6, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (M && h_traj_pv)
        MDG_CUDA(cudaMemcpyAsync(h_traj_pv, d_traj_pv, sizeof(float) * (size_t)n_frames * M, cudaMemcpyDeviceToHost, st));
    MDG_CUDA(cudaMemcpyAsync(c->h_pinned, c->flags.p, sizeof(int) * 8, cudaMemcpyDeviceToHost, st));
    MDG_CUDA(cudaStreamSynchronize(st));
    if (h_last_energy) *h_last_energy = (float)h_e;
    if (g_tl.prefix && dist && c->dist_p2p) {      // stamps of the LAST step's push / wait kernels (ns, relative to the push start)
        unsigned long long h[16];
        if (cudaMemcpy(h, c->dsync.as<char>() + offsetof(DistSync, dbg), sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
            char path[512];
            snprintf(path, sizeof(path), "%s%d.txt", g_tl.prefix, c->dist_rank);
            FILE* f = fopen(path, "a");
            if (f) {
                fprintf(f, "# push: start 0 ack_seen %lld copied %lld fenced %lld halo_flag %lld ke_flag %lld | wait / pull kernel: start %lld seen %lld pulled %lld | B prologue(ack sent) %lld pos raised %lld\n",
                        (long long)(h[1] - h[0]), (long long)(h[2] - h[0]), (long long)(h[3] - h[0]), (long long)(h[4] - h[0]),
                        (long long)(h[5] - h[0]), (long long)(h[6] - h[0]), (long long)(h[7] - h[0]), (long long)(h[10] - h[0]), (long long)(h[8] - h[0]),
                        (long long)(h[9] - h[0]));
                fclose(f);
            }
        }
    }
    tl_flush(c->dist_world > 1 ? c->dist_rank : 0);
    if (h_p2p_timeout) {
        mdg_set_error("distributed step: a peer-to-peer wait timed out (a neighbouring rank stopped making progress)");
        return MDG_E_NCCL;
    }
    if (c->prof_enable && c->prof_events) {
        std::vector<cudaEvent_t>* pool = (std::vector<cudaEvent_t>*)c->prof_events;
        c->prof_force_ms = 0.0;
        c->prof_force_launches = 0;
        for (int i = 0; i + 1 < c->prof_used; i += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, (*pool)[i], (*pool)[i + 1]) == cudaSuccess) {
                c->prof_force_ms += ms;
                c->prof_force_launches++;
            }
        }
    }
    if (c->h_pinned[6] || c->h_pinned[7]) {
        mdg_set_error("mdg_md_run: non-finite coordinates or collapsed cell (occupancy %d) - the dynamics diverged",
                      c->h_pinned[7]);
        return MDG_E_NUMERIC;
    }
    if (c->h_pinned[0]) return MDG_E_CAPACITY;
    if (c->h_pinned[5]) return MDG_E_SKIN;
    return MDG_OK;
}

extern "C" int mdg_md_run(mdg_ctx* c, const mdg_md_params* p, int n, const float* d_mass, const float* d_v0,
                          const float* d_q0, const float* h_pv0, const float* h_tgrid, int n_grid, float* d_traj_v,
                          float* d_traj_q, float* h_traj_pv, float* h_last_energy, void* stream) {
    if (!c || !p) { mdg_set_error("null ctx/params"); return MDG_E_BADARG; }
    if (n <= 0 || n_grid < 1) { mdg_set_error("mdg_md_run: n=%d n_grid=%d", n, n_grid); return MDG_E_BADARG; }
    if (p->integrator == MDG_INT_NHC && (p->n_chains < 2 || p->n_chains > MDG_MAX_CHAINS)) {
        mdg_set_error("NHC needs 2 <= n_chains <= %d (got %d)", MDG_MAX_CHAINS, p->n_chains);
        return MDG_E_BADARG;
    }
    if (p->integrator != MDG_INT_NHC && p->integrator != MDG_INT_NVE) { mdg_set_error("bad integrator"); return MDG_E_BADARG; }
    MDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    int K = p->rebuild_every < 1 ? 1 : p->rebuild_every;
    if (!(p->skin > 0.f)) K = 1;
    int status = MDG_E_CAPACITY;
    for (int attempt = 0; attempt < 12; ++attempt) {
        c->stat_launches = 0;
        c->stat_rebuilds = 0;
        c->prof_used = 0;
        int s = run_once(c, p, n, d_mass, d_v0, d_q0, h_pv0, h_tgrid, n_grid, d_traj_v, d_traj_q, h_traj_pv,
                         h_last_energy, K, st);
        c->force_energy = true;     // (an error return inside the loop must not leak the force-only mode)
        c->flags_sticky = false;
        if (s == MDG_E_CAPACITY) {
            if (c->h_pinned[1] > 0) {          // tile list: a block's stencil did not fit the staged-atom capacity
                c->tile_scap_min = c->h_pinned[1] + c->h_pinned[1] / 8 + 32;
                if (c->h_pinned[2] == 0) continue;
            }
            int need = c->h_pinned[2];
            int cap = ((need + need / 8 + 31) / 32) * 32;
            if (cap <= c->cap) cap = c->cap + 32;
            c->cap = cap;
            continue;
        }
        if (s == MDG_E_SKIN) {
            if (K == 1) { mdg_set_error("skin violated with rebuild_every=1"); status = MDG_E_SKIN; break; }
            K = K / 2 < 1 ? 1 : K / 2;
            continue;
        }
        c->stat_maxrow = K;   // report the rebuild interval that was finally used
        status = s;
        break;
    }
    c->slab = false;
    if (status == MDG_E_CAPACITY) mdg_set_error("mdg_md_run: could not satisfy capacity/skin constraints");
    return status;
}

// ---------------------------------------------------------------------------------------------
// GNN epoch: SchNet (+ pair priors) force provider under the same integrator kernels.
// State stays in ORIGINAL atom order (q4.w = id = index): the lists are rebuilt from scratch at every step anyway
// (reference semantics, topology_update_freq = 1) and each member sorts privately inside its own context.
// ---------------------------------------------------------------------------------------------
int mdg_i_export_count(mdg_ctx* c, cudaStream_t st, int64_t* h_npairs);
int mdg_i_export_fill(mdg_ctx* c, int64_t* d_nbr, float* d_offsets, float* d_dis, cudaStream_t st);

__global__ void k_gnn_init(int n, const float* __restrict__ v0, const float* __restrict__ q0, const float* __restrict__ mass,
                           float4* __restrict__ v4, float4* __restrict__ q4, float4* __restrict__ vh4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    v4[i] = make_float4(v0[3 * i], v0[3 * i + 1], v0[3 * i + 2], mass[i]);
    q4[i] = make_float4(q0[3 * i], q0[3 * i + 1], q0[3 * i + 2], __int_as_float(i));
    vh4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void k_gnn_q_to_xyz(int n, const float4* __restrict__ q4, float* __restrict__ xyz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 q = q4[i];
    xyz[3 * i] = q.x; xyz[3 * i + 1] = q.y; xyz[3 * i + 2] = q.z;
}

// f4 = f_gnn + sum of the prior forces (fp3 holds n_priors consecutive N x 3 blocks), summed in member order like
// Stack.forward (interface.py:397-403) followed by one autograd pass
__global__ void k_gnn_sum_forces(int n, const float* __restrict__ f3, const float* __restrict__ fp3, int n_priors,
                                 float4* __restrict__ f4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float fx = f3[3 * i], fy = f3[3 * i + 1], fz = f3[3 * i + 2];
    for (int k = 0; k < n_priors; ++k) {
        const float* f = fp3 + (size_t)k * 3 * n;
        fx += f[3 * i]; fy += f[3 * i + 1]; fz += f[3 * i + 2];
    }
    f4[i] = make_float4(fx, fy, fz, 0.f);
}

int mdg_i_nbr_build_async(mdg_ctx* c, const float* d_xyz, const float4* d_q4, int n, const float* h_cell3, double cutoff, const uint8_t* d_sel_a,
                          const uint8_t* d_sel_b, const int64_t* d_ex_keys, int n_ex, bool want_export, int64_t cap_pairs,
                          int64_t* d_nbr, float* d_offsets, cudaStream_t st);
int mdg_i_schnet_energy_force(mdg_ctx* c, const mdg_schnet_model* m, const int64_t* d_z, const float* d_xyz, int n,
                              const int64_t* d_nbr, const float* d_offsets, int64_t E, const int* d_E, const float* h_off_scale3,
                              float* d_energy, float* d_force, void* stream);

// One force evaluation of the GNN / Stack model.
//   cap_pairs <  0 (synchronous): every list build reads its pair count back (one SYNC per member), buffers are sized
//                  exactly; *pairs_out = the GNN list's pair count.
//   cap_pairs >= 0 (asynchronous): NOTHING is read back.  The GNN list is exported into buffers of cap_pairs pairs (sized by
//                  the caller from an earlier, synchronous evaluation), its count stays in flags[4] and is consumed on the
//                  device; overflows are latched in flags[3] of the member's context (mdg_i_nbr_build_async).
static int gnn_force(mdg_ctx* c, const mdg_gnn_md_params* p, const mdg_schnet_model* model, const int64_t* d_z, int n,
                     const float4* q4, float4* f4, float* d_e_gnn, int64_t* launches, int64_t cap_pairs, int64_t* pairs_out,
                     cudaStream_t st) {
    const int T = 256, nb = (n + T - 1) / T;
    const bool async = cap_pairs >= 0;
    *launches += c->stat_launches;               // (the list builds restart the context's counter)
    float* xyz = c->gnn_xyz.as<float>();
    float* f3 = c->gnn_f3.as<float>();
    float* fp3 = c->gnn_fp3.as<float>();
    k_gnn_q_to_xyz<<<nb, T, 0, st>>>(n, q4, xyz);
    if (model) {
        // GNN list at the current positions (exact membership, reference layout)
        int64_t P = 0;
        if (!async) {
            MDG_TRY(mdg_nbr_build(c, xyz, n, p->cell, p->cutoff, nullptr, nullptr, p->d_ex_keys, p->n_ex, (void*)st, &P));
            MDG_TRY(c->gnn_nbr.reserve(sizeof(int64_t) * 2 * (size_t)(P + 1)));
            MDG_TRY(c->gnn_off.reserve(sizeof(float) * 3 * (size_t)(P + 1)));
            MDG_TRY(mdg_i_export_fill(c, c->gnn_nbr.as<int64_t>(), c->gnn_off.as<float>(), nullptr, st));
            MDG_TRY(mdg_schnet_energy_force(c, model, d_z, xyz, n, c->gnn_nbr.as<int64_t>(), c->gnn_off.as<float>(), P, p->off_scale,
                                            d_e_gnn, f3, (void*)st));
            if (pairs_out) *pairs_out = P;
        } else {
            MDG_TRY(mdg_i_nbr_build_async(c, xyz, q4, n, p->cell, p->cutoff, nullptr, nullptr, p->d_ex_keys, p->n_ex, true, cap_pairs,
                                          c->gnn_nbr.as<int64_t>(), c->gnn_off.as<float>(), st));
            MDG_TRY(mdg_i_schnet_energy_force(c, model, d_z, xyz, n, c->gnn_nbr.as<int64_t>(), c->gnn_off.as<float>(), cap_pairs,
                                              c->flags.as<int>() + 4, p->off_scale, d_e_gnn, f3, (void*)st));
        }
    } else {    // Stack of analytic pair members only
        MDG_CUDA(cudaMemsetAsync(f3, 0, sizeof(float) * 3 * (size_t)n, st));
        MDG_CUDA(cudaMemsetAsync(d_e_gnn, 0, sizeof(float), st));
    }
    for (int k = 0; k < p->n_priors; ++k) {
        const mdg_prior_spec& R = p->priors[k];
        if (!async) {
            int64_t Pk = 0;
            MDG_TRY(mdg_nbr_build(R.ctx, xyz, n, p->cell, R.cutoff, R.d_sel_a, R.d_sel_b, R.d_ex_keys, R.n_ex, (void*)st, &Pk));
        } else {    // the force kernel streams the member's rows: no export, no count
            MDG_TRY(mdg_i_nbr_build_async(R.ctx, xyz, q4, n, p->cell, R.cutoff, R.d_sel_a, R.d_sel_b, R.d_ex_keys, R.n_ex, false, 0,
                                          nullptr, nullptr, st));
        }
        MDG_TRY(mdg_pair_force(R.ctx, R.kind, R.params, R.n_params, xyz, n, nullptr, fp3 + (size_t)k * 3 * n, nullptr, (void*)st));
    }
    int n_extra = p->n_priors;
    if (p->bonded.n_bonds + p->bonded.n_angles > 0) {      // bonded members of the Stack (bonded.cu)
        MDG_TRY(mdg_bonded_force(c, &p->bonded, xyz, n, p->cell, nullptr, fp3 + (size_t)n_extra * 3 * n, nullptr, (void*)st));
        n_extra++;
    }
    k_gnn_sum_forces<<<nb, T, 0, st>>>(n, f3, fp3, n_extra, f4);
    c->stat_launches += 2;
    for (int k = 0; k < p->n_priors; ++k) *launches += p->priors[k].ctx->stat_launches;
    c->stat_rebuilds++;
    MDG_KERNEL_CHECK();
    return MDG_OK;
}

static int gnn_run_once(mdg_ctx* c, const mdg_gnn_md_params* p, const mdg_schnet_model* model, const int64_t* d_z, int n,
                        const float* d_mass, const float* d_v0, const float* d_q0, const float* h_pv0, const float* h_tgrid,
                        int n_grid, float* d_traj_v, float* d_traj_q, float* h_traj_pv, float* h_last_energy, bool async,
                        int* latched, cudaStream_t st);

extern "C" int mdg_md_run_gnn(mdg_ctx* c, const mdg_gnn_md_params* p, const mdg_schnet_model* model, const int64_t* d_z, int n,
                              const float* d_mass, const float* d_v0, const float* d_q0, const float* h_pv0,
                              const float* h_tgrid, int n_grid, float* d_traj_v, float* d_traj_q, float* h_traj_pv,
                              float* h_last_energy, void* stream) {
    if (!c || !p || (model && !d_z)) { mdg_set_error("mdg_md_run_gnn: null argument"); return MDG_E_BADARG; }
    if (!model && p->n_priors < 1 && p->bonded.n_bonds + p->bonded.n_angles < 1) {
        mdg_set_error("mdg_md_run_gnn: no SchNet model, no pair member and no bonded term");
        return MDG_E_BADARG;
    }
    if (n <= 0 || n_grid < 1) { mdg_set_error("mdg_md_run_gnn: n=%d n_grid=%d", n, n_grid); return MDG_E_BADARG; }
    if (p->integrator != MDG_INT_NHC && p->integrator != MDG_INT_NVE) { mdg_set_error("bad integrator"); return MDG_E_BADARG; }
    if (p->integrator == MDG_INT_NHC && (p->n_chains < 2 || p->n_chains > MDG_MAX_CHAINS)) {
        mdg_set_error("NHC needs 2 <= n_chains <= %d (got %d)", MDG_MAX_CHAINS, p->n_chains);
        return MDG_E_BADARG;
    }
    if (p->n_priors < 0 || p->n_priors > MDG_MAX_PRIORS) { mdg_set_error("mdg_md_run_gnn: n_priors=%d", p->n_priors); return MDG_E_BADARG; }
    for (int k = 0; k < p->n_priors; ++k)
        if (!p->priors[k].ctx || p->priors[k].ctx == c) { mdg_set_error("mdg_md_run_gnn: prior %d needs its own context", k); return MDG_E_BADARG; }
    if (c->dist_world > 1) { mdg_set_error("mdg_md_run_gnn: single GPU only"); return MDG_E_BADARG; }
    MDG_CUDA(cudaSetDevice(c->device));
    // Steps after the first run ASYNCHRONOUSLY (no pair-count read-back, see gnn_force); if a capacity sized from the first
    // evaluation turns out too small during the epoch, the latched flags say so at the end and the epoch is repeated on
    // the synchronous path (inputs are never modified).  MDG_GNN_SYNC=1 forces the synchronous path.
    const bool try_async = getenv("MDG_GNN_SYNC") == nullptr && n_grid > 2;
    const int64_t replays0 = c->stat_graph_replays;
    if (try_async) {
        int latched = 0;
        MDG_TRY(gnn_run_once(c, p, model, d_z, n, d_mass, d_v0, d_q0, h_pv0, h_tgrid, n_grid, d_traj_v, d_traj_q, h_traj_pv,
                             h_last_energy, true, &latched, (cudaStream_t)stream));
        // stats slot 3: 1 = the epoch completed on the asynchronous path, 2 = with its force evaluations replayed as a graph
        if (!latched) { c->stat_maxrow = c->stat_graph_replays > replays0 ? 2 : 1; return MDG_OK; }
        c->stat_async_retries++;
    }
    c->stat_maxrow = 0;
    return gnn_run_once(c, p, model, d_z, n, d_mass, d_v0, d_q0, h_pv0, h_tgrid, n_grid, d_traj_v, d_traj_q, h_traj_pv, h_last_energy,
                        false, nullptr, (cudaStream_t)stream);
}

static int gnn_run_once(mdg_ctx* c, const mdg_gnn_md_params* p, const mdg_schnet_model* model, const int64_t* d_z, int n,
                        const float* d_mass, const float* d_v0, const float* d_q0, const float* h_pv0, const float* h_tgrid,
                        int n_grid, float* d_traj_v, float* d_traj_q, float* h_traj_pv, float* h_last_energy, bool async,
                        int* latched, cudaStream_t st_user) {
    // CUDA-graph replay (default since its first hardware runs in round 2 - bit-identical to the plain asynchronous epoch,
    // tests/test_zzz_gpu_last.py, and 2540 vs 2196 steps/s on the 64-water SchNet box; MDG_GNN_GRAPH=0 turns it off): on the
    // asynchronous path a force evaluation is a FIXED launch sequence
    // with fixed arguments (same buffers, capacities, device-side counts), so it is captured once (at the second step, after the
    // first one has grown every buffer) and replayed as a CUDA graph.  Stream capture is not allowed on the legacy default
    // stream - which is what PyTorch's current stream usually is - so such an epoch runs on a private stream, ordered after the
    // caller's stream by an event and synchronised before the call returns.
    struct GraphGuard {
        cudaGraphExec_t exec = nullptr;
        ~GraphGuard() { if (exec) cudaGraphExecDestroy(exec); }
    } gg;
    const char* ge = getenv("MDG_GNN_GRAPH");
    const bool use_graph = async && !(ge && ge[0] == '0') && n_grid > 3;
    bool graph_failed = false;
    cudaStream_t st = st_user;
    if (use_graph) {
        if (!c->gnn_stream) {
            MDG_CUDA(cudaStreamCreateWithFlags(&c->gnn_stream, cudaStreamNonBlocking));
            MDG_CUDA(cudaEventCreateWithFlags(&c->ev_gnn, cudaEventDisableTiming));
        }
        MDG_CUDA(cudaEventRecord(c->ev_gnn, st_user));
        MDG_CUDA(cudaStreamWaitEvent(c->gnn_stream, c->ev_gnn, 0));
        st = c->gnn_stream;
    }
    const int nhc = p->integrator == MDG_INT_NHC;
    const int M = nhc ? p->n_chains : 0;
    const int stride = p->traj_stride < 1 ? 1 : p->traj_stride;
    const int n_frames = (n_grid - 1) / stride + 1;
    const int T = 256, nb = (n + T - 1) / T;
    c->stat_launches = 0;
    c->stat_rebuilds = 0;

    IntArgs A;
    memset(&A, 0, sizeof(A));
    A.s0 = 0; A.s1 = n;
    A.integrator = p->integrator;
    A.M = M;
    for (int k = 0; k < M; ++k) A.Q[k] = p->Q[k];
    A.T = (float)p->T;
    A.target = (float)(p->T * (double)p->ndof * 0.5);
    A.half_skin2 = 0.f;

    MDG_TRY(c->v4.reserve(sizeof(float4) * (size_t)n));
    MDG_TRY(c->vh4.reserve(sizeof(float4) * (size_t)n));
    MDG_TRY(c->f4b.reserve(sizeof(float4) * (size_t)n));
    MDG_TRY(c->gnn_xyz.reserve(sizeof(float) * 3 * (size_t)n));
    MDG_TRY(c->gnn_f3.reserve(sizeof(float) * 3 * (size_t)n));
    MDG_TRY(c->gnn_fp3.reserve(sizeof(float) * 3 * (size_t)n * (size_t)(p->n_priors + 1)));
    MDG_TRY(c->pvbuf.reserve(sizeof(Scalars) + sizeof(float) * (size_t)n_frames * MDG_MAX_CHAINS + 64));
    MDG_TRY(c->kebuf.reserve(sizeof(double) * (5 * INT_MAX_BLOCKS + 8)));
    MDG_TRY(c->flags.reserve(sizeof(int) * 8));
    float4 *v4 = c->v4.as<float4>(), *vh4 = c->vh4.as<float4>(), *f4 = c->f4b.as<float4>();
    // positions: `qref` (free here - no skin list).  The list builds of this context use q4b / qs_buf / rows / ..., the
    // SchNet program sn_ws, so v4 / vh4 / f4b / qref are untouched by the force evaluation.
    MDG_TRY(c->qref.reserve(sizeof(float4) * (size_t)n));
    float4* q4 = c->qref.as<float4>();
    Scalars* sc = c->pvbuf.as<Scalars>();
    float* d_traj_pv = (float*)(sc + 1);
    float* d_e_gnn = d_traj_pv + (size_t)n_frames * MDG_MAX_CHAINS;
    double* kb = c->kebuf.as<double>();
    double *ke_v_cur = kb, *ke_h_cur = kb + INT_MAX_BLOCKS, *ke_v_nxt = kb + 2 * INT_MAX_BLOCKS, *ke_h_nxt = kb + 3 * INT_MAX_BLOCKS;

    Scalars hs;
    memset(&hs, 0, sizeof(hs));
    for (int k = 0; k < M; ++k) hs.pv[0][k] = h_pv0 ? h_pv0[k] : 0.f;
    MDG_CUDA(cudaMemcpyAsync(sc, &hs, sizeof(Scalars), cudaMemcpyHostToDevice, st));
    MDG_CUDA(cudaMemsetAsync(c->flags.p, 0, sizeof(int) * 8, st));
    for (int k = 0; k < p->n_priors; ++k)      // the members' overflow latches
        MDG_CUDA(cudaMemsetAsync(p->priors[k].ctx->flags.as<int>() + 3, 0, sizeof(int), st));
    k_gnn_init<<<nb, T, 0, st>>>(n, d_v0, d_q0, d_mass, v4, q4, vh4);
    int ib = nb < 1 ? 1 : (nb > INT_MAX_BLOCKS ? INT_MAX_BLOCKS : nb);
    int64_t launches = 0;
    // The first evaluation of an epoch is synchronous - its pair count sizes the edge buffers of the asynchronous steps -
    // unless the previous epoch of the same system completed asynchronously: then its capacity is reused and this epoch has no
    // read-back at all (a capacity that no longer fits is latched like any other overflow and the epoch repeated).
    int64_t P0 = 0, cap_pairs = -1;
    const char* mg = getenv("MDG_GNN_MARGIN");                 // (tests force the overflow / retry path with a tiny margin)
    if (async && !mg && c->gnn_last_cap > 0 && c->gnn_last_n == n) {
        cap_pairs = c->gnn_last_cap;
        if (model) {
            MDG_TRY(c->gnn_nbr.reserve(sizeof(int64_t) * 2 * (size_t)(cap_pairs + 1)));
            MDG_TRY(c->gnn_off.reserve(sizeof(float) * 3 * (size_t)(cap_pairs + 1)));
        }
        MDG_TRY(gnn_force(c, p, model, d_z, n, q4, f4, d_e_gnn, &launches, cap_pairs, nullptr, st));
    } else {
        MDG_TRY(gnn_force(c, p, model, d_z, n, q4, f4, d_e_gnn, &launches, -1, &P0, st));
        if (async) {
            cap_pairs = P0 + (mg ? (int64_t)atoll(mg) : P0 / 4 + 256);
            if (model) {
                MDG_TRY(c->gnn_nbr.reserve(sizeof(int64_t) * 2 * (size_t)(cap_pairs + 1)));
                MDG_TRY(c->gnn_off.reserve(sizeof(float) * 3 * (size_t)(cap_pairs + 1)));
            }
        }
    }
    c->gnn_last_cap = 0;                                       // set again below when this epoch completes asynchronously
    if (nhc) k_ke_init<<<ib, INT_THREADS, 0, st>>>(A, v4, ke_v_cur);
    MDG_CUDA(cudaMemcpyAsync(d_traj_v, d_v0, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    MDG_CUDA(cudaMemcpyAsync(d_traj_q, d_q0, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    if (M) MDG_CUDA(cudaMemcpyAsync(d_traj_pv, sc->pv[0], sizeof(float) * M, cudaMemcpyDeviceToDevice, st));
    c->stat_launches += 2;

    int pv_sel = 0;
    const int nsteps = n_grid - 1;
    for (int g = 0; g < nsteps; ++g) {
        float dt = h_tgrid[g + 1] - h_tgrid[g];
        if (g == 0) {
            k_step_a<<<ib, INT_THREADS, 0, st>>>(A, dt, pv_sel, sc, v4, vh4, q4, f4, nullptr, 0, ke_h_cur, c->flags.as<int>(), DistArgs{});
            c->stat_launches++;
        }
        if (use_graph && g >= 1 && !gg.exec && !graph_failed) {       // capture the evaluation of the second step
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                int64_t dummy = 0;
                const int r = gnn_force(c, p, model, d_z, n, q4, f4, d_e_gnn, &dummy, cap_pairs, nullptr, st);
                const cudaError_t e = cudaStreamEndCapture(st, &graph);        // (always: never leave the stream capturing)
                if (r != MDG_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&gg.exec, graph, 0) != cudaSuccess) {
                    graph_failed = true;
                    gg.exec = nullptr;
                }
                if (graph) cudaGraphDestroy(graph);
            } else {
                graph_failed = true;
            }
            if (graph_failed) (void)cudaGetLastError();                        // plain launches from here on
        }
        if (gg.exec) {
            MDG_CUDA(cudaGraphLaunch(gg.exec, st));
            c->stat_graph_replays++;
            launches++;
        } else {
            MDG_TRY(gnn_force(c, p, model, d_z, n, q4, f4, d_e_gnn, &launches, cap_pairs, nullptr, st));
        }
        int gp = g + 1;
        bool keep = (gp % stride) == 0;
        size_t fr = (size_t)(gp / stride);
        float* tv = keep ? d_traj_v + fr * 3 * (size_t)n : nullptr;
        float* tq = keep ? d_traj_q + fr * 3 * (size_t)n : nullptr;
        float* tp = (keep && M) ? d_traj_pv + fr * M : nullptr;
        if (g + 1 < nsteps) {
            float dt_next = h_tgrid[g + 2] - h_tgrid[g + 1];
            k_step_ba<<<ib, INT_THREADS, 0, st>>>(A, dt, dt_next, pv_sel, sc, v4, vh4, q4, f4, ke_v_cur, ke_h_cur, ib, ke_v_nxt,
                                                  ke_h_nxt, tv, tq, tp, nullptr, 0, c->flags.as<int>(), DistArgs{nullptr, nullptr, nullptr, 1, 0});
            { double* t1 = ke_v_cur; ke_v_cur = ke_v_nxt; ke_v_nxt = t1; }
            { double* t2 = ke_h_cur; ke_h_cur = ke_h_nxt; ke_h_nxt = t2; }
        } else {
            k_step_b<<<ib, INT_THREADS, 0, st>>>(A, dt, pv_sel, sc, v4, vh4, q4, f4, ke_v_cur, ke_h_cur, ib, ke_v_nxt, tv, tq, tp,
                                                 DistArgs{nullptr, nullptr, nullptr, 1, 0});
            { double* t1 = ke_v_cur; ke_v_cur = ke_v_nxt; ke_v_nxt = t1; }
        }
        c->stat_launches++;
        pv_sel ^= 1;
    }
    MDG_KERNEL_CHECK();
    float h_e = 0.f;
    if (h_last_energy) MDG_CUDA(cudaMemcpyAsync(&h_e, d_e_gnn, sizeof(float), cudaMemcpyDeviceToHost, st));
    if (M && h_traj_pv)
        MDG_CUDA(cudaMemcpyAsync(h_traj_pv, d_traj_pv, sizeof(float) * (size_t)n_frames * M, cudaMemcpyDeviceToHost, st));
    if (async) {    // the overflow latches of all members -> pinned slots 8.., read after the one synchronisation of the epoch
        MDG_CUDA(cudaMemcpyAsync(c->h_pinned + 8, c->flags.as<int>() + 3, sizeof(int), cudaMemcpyDeviceToHost, st));
        for (int k = 0; k < p->n_priors; ++k)
            MDG_CUDA(cudaMemcpyAsync(c->h_pinned + 9 + k, p->priors[k].ctx->flags.as<int>() + 3, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    MDG_CUDA(cudaStreamSynchronize(st));
    if (async) {
        int any = c->h_pinned[8];
        for (int k = 0; k < p->n_priors; ++k) any |= c->h_pinned[9 + k];
        *latched = any;
        if (any & 2) { mdg_set_error("mdg_md_run_gnn: non-finite coordinates or collapsed cell during the epoch"); return MDG_E_NUMERIC; }
        if (!any && !mg) { c->gnn_last_cap = cap_pairs; c->gnn_last_n = n; }
    }
    if (h_last_energy) *h_last_energy = h_e;      // SchNet energy of the last evaluation (priors not included)
    c->stat_launches += launches;
    return MDG_OK;
}

void mdg_i_release_profile(mdg_ctx* c) {
    std::vector<cudaEvent_t>* pool = (std::vector<cudaEvent_t>*)c->prof_events;
    if (!pool) return;
    for (cudaEvent_t e : *pool) cudaEventDestroy(e);
    delete pool;
    c->prof_events = nullptr;
}

// Per-kernel timing of the force launches inside mdg_md_run (CUDA events on the launch stream
// around every force kernel).  out[0] = summed force-kernel ms of the last run, out[1] = launches.
extern "C" int mdg_set_profile(mdg_ctx* c, int enable) {
    if (!c) { mdg_set_error("null ctx"); return MDG_E_BADARG; }
    c->prof_enable = enable;
    return MDG_OK;
}

extern "C" int mdg_get_profile(mdg_ctx* c, double* h_out2) {
    if (!c || !h_out2) { mdg_set_error("null argument"); return MDG_E_BADARG; }
    h_out2[0] = c->prof_force_ms;
    h_out2[1] = (double)c->prof_force_launches;
    return MDG_OK;
}
