/*
 * oracle_c.c - plain-C restatement of the reference's hot path.  TEST INFRASTRUCTURE / CPU
 * BASELINE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs; never by the product path (mdgrad_b200/*).
 *
 * Follows the reference algorithm (torchmd/mdgrad @ cea2332e), not the GPU design:
 *   - orc_nbr_list        all-pairs O(N^2) minimum-image list, torchmd/topology.py:30-73
 *                         (d = x_j - x_i :35; strict +-0.5 image test on d * (1/L) :59-62;
 *                          d += off*L :64; d2 = (dx^2+dy^2)+dz^2; d2 < fl32(cutoff^2) && d2 != 0
 *                          over the upper triangle :66-68), emitted in (i, j) row-major order.
 *   - orc_pair_rows       per evaluation: rebuild the all-pairs list rows and accumulate
 *                         E = sum u(r), F = -dE/dx  (torchmd/interface.py:263-300, md.py:227-228,
 *                         potentials.py:317-327) for the atoms [i0, i1) - the reference rebuilds
 *                         the list at every force evaluation (topology_update_freq = 1).
 *   - orc_nhc_md          NH-Verlet / NHC epoch on top of it (torchmd/sovlers.py:110-127,
 *                         torchmd/md.py:210-240), one force evaluation per step.
 * Unlike the torch oracle (oracle/oracle_torch.py) it needs O(N) memory, so it reaches the
 * 256k-atom benchmark configuration; it is validated against the torch oracle / the reference at
 * small sizes in tests/test_oracle_golden.py.  Compile with -ffp-contract=off (no FMA) so the
 * fp32 membership arithmetic rounds exactly like the reference's separate ATen ops.
 * Parity pinning: see oracle/oracle_torch.py header.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline float axis_min_image(float xi, float xj, float L, float invL, float* off) {
    float d = xj - xi;
    float red = d * invL;
    float o = 0.0f;
    if (red > 0.5f) o = -1.0f;
    else if (red < -0.5f) o = 1.0f;
    *off = o;
    return d + o * L;
}

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Returns the number of pairs P (i<j). If nbr != NULL writes at most cap pairs:
 * nbr[2p], nbr[2p+1]; off[3p..]; dis[p] (may be NULL). Order: ascending i, then ascending j. */
int64_t orc_nbr_list(const float* xyz, int n, const float* cell3, double cutoff, int64_t* nbr, float* off,
                     float* dis, int64_t cap) {
    float L[3], invL[3];
    for (int k = 0; k < 3; ++k) { L[k] = cell3[k]; invL[k] = 1.0f / cell3[k]; }
    const float rc2 = (float)(cutoff * cutoff);
    int64_t* cnt = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n; ++i) {
        int64_t c = 0;
        for (int j = i + 1; j < n; ++j) {
            float ox, oy, oz;
            float dx = axis_min_image(xyz[3 * i], xyz[3 * j], L[0], invL[0], &ox);
            float dy = axis_min_image(xyz[3 * i + 1], xyz[3 * j + 1], L[1], invL[1], &oy);
            float dz = axis_min_image(xyz[3 * i + 2], xyz[3 * j + 2], L[2], invL[2], &oz);
            float d2 = (dx * dx + dy * dy) + dz * dz;
            c += (d2 < rc2) && (d2 != 0.0f);
        }
        cnt[i + 1] = c;
    }
    for (int i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
    int64_t P = cnt[n];
    if (nbr) {
#pragma omp parallel for schedule(dynamic, 16)
        for (int i = 0; i < n; ++i) {
            int64_t p = cnt[i];
            for (int j = i + 1; j < n; ++j) {
                float ox, oy, oz;
                float dx = axis_min_image(xyz[3 * i], xyz[3 * j], L[0], invL[0], &ox);
                float dy = axis_min_image(xyz[3 * i + 1], xyz[3 * j + 1], L[1], invL[1], &oy);
                float dz = axis_min_image(xyz[3 * i + 2], xyz[3 * j + 2], L[2], invL[2], &oz);
                float d2 = (dx * dx + dy * dy) + dz * dz;
                if ((d2 < rc2) && (d2 != 0.0f)) {
                    if (p < cap) {
                        nbr[2 * p] = i; nbr[2 * p + 1] = j;
                        off[3 * p] = ox; off[3 * p + 1] = oy; off[3 * p + 2] = oz;
                        if (dis) dis[p] = sqrtf(d2);
                    }
                    ++p;
                }
            }
        }
    }
    free(cnt);
    return P;
}

/* Reference rows of SELECTED atoms at any box size (tests/fullsize_checks.py): for every i = sel[r] the pairs (i, j), j > i,
 * that torchmd/topology.py:59-68 lists (upper triangle, d = x_j - x_i, strict +-0.5 image test, d2 < fl32(cutoff^2) && d2 != 0),
 * in ascending j like torch.nonzero.  out_j / out_off hold `cap` slots per selected row; out_cnt[r] is the TRUE count (may
 * exceed cap: the caller re-runs with a larger cap).  O(nsel * N) work, O(nsel * cap) memory. */
void orc_nbr_rows_upper(const float* xyz, int n, const float* cell3, double cutoff, const int64_t* sel, int nsel,
                        int64_t* out_j, float* out_off, int* out_cnt, int cap) {
    float L[3], invL[3];
    for (int k = 0; k < 3; ++k) { L[k] = cell3[k]; invL[k] = 1.0f / cell3[k]; }
    const float rc2 = (float)(cutoff * cutoff);
#pragma omp parallel for schedule(dynamic, 8)
    for (int r = 0; r < nsel; ++r) {
        const int i = (int)sel[r];
        int c = 0;
        for (int j = i + 1; j < n; ++j) {
            float ox, oy, oz;
            float dx = axis_min_image(xyz[3 * i], xyz[3 * j], L[0], invL[0], &ox);
            float dy = axis_min_image(xyz[3 * i + 1], xyz[3 * j + 1], L[1], invL[1], &oy);
            float dz = axis_min_image(xyz[3 * i + 2], xyz[3 * j + 2], L[2], invL[2], &oz);
            float d2 = (dx * dx + dy * dy) + dz * dz;
            if ((d2 < rc2) && (d2 != 0.0f)) {
                if (c < cap) {
                    out_j[(size_t)r * cap + c] = j;
                    out_off[((size_t)r * cap + c) * 3] = ox;
                    out_off[((size_t)r * cap + c) * 3 + 1] = oy;
                    out_off[((size_t)r * cap + c) * 3 + 2] = oz;
                }
                ++c;
            }
        }
        out_cnt[r] = c;
    }
}

/* LJ u(r) = 4 eps ((s/r)^12 - (s/r)^6): e and g = -u'(r)/r from d2 (double precision algebra). */
static inline void lj_eval(double d2, double sigma, double eps, double* e, double* g) {
    double r2i = 1.0 / d2, s2 = sigma * sigma * r2i, s6 = s2 * s2 * s2, s12 = s6 * s6;
    *e = 4.0 * eps * (s12 - s6);
    *g = 24.0 * eps * (2.0 * s12 - s6) * r2i;
}

/* One force evaluation of the reference for the atoms [i0, i1): all-pairs membership (every j != i,
 * exact fp32 test) then energy/force accumulation. f: (i1-i0) x 3 floats; returns sum_i e_i with
 * e_i = 1/2 sum_j u(r_ij). */
double orc_pair_rows(const float* xyz, int n, const float* cell3, double cutoff, double sigma, double eps,
                     int i0, int i1, float* f) {
    float L[3], invL[3];
    for (int k = 0; k < 3; ++k) { L[k] = cell3[k]; invL[k] = 1.0f / cell3[k]; }
    const float rc2 = (float)(cutoff * cutoff);
    double etot = 0.0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : etot)
    for (int i = i0; i < i1; ++i) {
        double fx = 0, fy = 0, fz = 0, ei = 0;
        const float xi = xyz[3 * i], yi = xyz[3 * i + 1], zi = xyz[3 * i + 2];
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            float ox, oy, oz;
            float dx = axis_min_image(xi, xyz[3 * j], L[0], invL[0], &ox);
            float dy = axis_min_image(yi, xyz[3 * j + 1], L[1], invL[1], &oy);
            float dz = axis_min_image(zi, xyz[3 * j + 2], L[2], invL[2], &oz);
            float d2 = (dx * dx + dy * dy) + dz * dz;
            if ((d2 < rc2) && (d2 != 0.0f)) {
                double e, g;
                lj_eval((double)d2, sigma, eps, &e, &g);
                fx -= g * dx; fy -= g * dy; fz -= g * dz;
                ei += 0.5 * e;
            }
        }
        f[3 * (i - i0)] = (float)fx; f[3 * (i - i0) + 1] = (float)fy; f[3 * (i - i0) + 2] = (float)fz;
        etot += ei;
    }
    return etot;
}

/* NHC derivative pieces (torchmd/md.py:221-240) in fp32 like the reference. */
static void nhc_dpv(int M, const float* Q, float T, float target, float ke, const float* pv, float* d) {
    d[0] = 2.0f * (ke - target) - pv[0] * pv[1] / Q[1];
    for (int k = 1; k < M - 1; ++k) d[k] = (pv[k - 1] * pv[k - 1] / Q[k - 1] - T) - pv[k + 1] * pv[k] / Q[k + 1];
    d[M - 1] = pv[M - 2] * pv[M - 2] / Q[M - 2] - T;
}

static float kinetic(const float* v, const float* m, int n) {
    double acc = 0;
#pragma omp parallel for reduction(+ : acc)
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) { float p = v[3 * i + k] * m[i]; acc += (double)(p * p / m[i]); }
    return (float)(0.5 * acc);
}

/* nsteps NH-Verlet steps in place on (v, q, pv); dt[s] per step. `rows` > 0 evaluates the force only
 * for `rows` atoms per evaluation (bounded-sample timing mode; the trajectory is then NOT physical);
 * rows <= 0 = all atoms.  Returns the last potential energy. */
double orc_nhc_md(float* v, float* q, float* pv, const float* mass, int n, const float* cell3, double cutoff,
                  double sigma, double eps, int M, const float* Q, double T, int ndof, const float* dt, int nsteps,
                  int rows) {
    float* f = (float*)calloc((size_t)3 * n, sizeof(float));
    float* vh = (float*)calloc((size_t)3 * n, sizeof(float));
    float target = (float)(T * ndof * 0.5), Tf = (float)T;
    int r1 = (rows > 0 && rows < n) ? rows : n;
    double epot = orc_pair_rows(q, n, cell3, cutoff, sigma, eps, 0, r1, f);
    float d0[32], d1[32], ph[32], pvh[32];
    for (int s = 0; s < nsteps; ++s) {
        float h = dt[s];
        float ke0 = kinetic(v, mass, n);
        float pv0 = pv[0], Q0 = Q[0];
#pragma omp parallel for
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) {
                float m = mass[i], p = v[3 * i + k] * m;
                float a = (f[3 * i + k] - pv0 * p / Q0) / m;
                float hh = 0.5f * a * h;
                vh[3 * i + k] = hh;
                q[3 * i + k] = q[3 * i + k] + (v[3 * i + k] + hh) * h;
            }
        nhc_dpv(M, Q, Tf, target, ke0, pv, d0);
        for (int k = 0; k < M; ++k) { ph[k] = 0.5f * d0[k] * h; pvh[k] = pv[k] + ph[k]; }
        epot = orc_pair_rows(q, n, cell3, cutoff, sigma, eps, 0, r1, f);
        double acc = 0;
#pragma omp parallel for reduction(+ : acc)
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) { float p = (v[3 * i + k] + vh[3 * i + k]) * mass[i]; acc += (double)(p * p / mass[i]); }
        float ke1 = (float)(0.5 * acc);
        nhc_dpv(M, Q, Tf, target, ke1, pvh, d1);
        float pvh0 = pvh[0];
#pragma omp parallel for
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) {
                float m = mass[i], p = (v[3 * i + k] + vh[3 * i + k]) * m;
                float a = (f[3 * i + k] - pvh0 * p / Q0) / m;
                v[3 * i + k] = v[3 * i + k] + (vh[3 * i + k] + 0.5f * a * h);
            }
        for (int k = 0; k < M; ++k) pv[k] = pv[k] + (ph[k] + 0.5f * d1[k] * h);
    }
    free(f);
    free(vh);
    return epot;
}
