/*
 * mdgrad_b200.h - C ABI of libmdgrad_b200.so: the sm_100a (B200) implementation of
 * torchmd/mdgrad's MD hot path (periodic neighbor list, pair distance / force, fused
 * Velocity-Verlet / Nose-Hoover-chain step, RDF).
 *
 * The reference (torchmd/mdgrad @ cea2332e) is pure Python/PyTorch and has no FFI; the
 * boundary it exposes for this path is its Python API.  Every entry point below therefore
 * cites the reference Python function whose work it replaces (file:line relative to the
 * reference tree); the binding a reference maintainer would add is the ctypes stub shown
 * in INTEGRATION.md (and shipped as mdgrad_b200/_lib.py).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types.
 *  - every function returns an int status: MDG_OK or a negative MDG_E_* code;
 *    mdg_last_error() returns a thread-local message for the last failure. Nothing throws.
 *  - pointers named d_* are DEVICE pointers (e.g. tensor.data_ptr()); h_* are HOST pointers.
 *  - all work is enqueued on the cudaStream_t passed as `stream` (void*; pass PyTorch's
 *    current stream). Calls are asynchronous except where "SYNC" is stated (a count
 *    read-back that sizes a caller allocation).
 *  - a context (mdg_ctx) owns all scratch / internal lists for one device and one box;
 *    it is thread-compatible (one thread at a time), not re-entrant.
 *  - fp32 on device throughout, like the reference (torch.Tensor default dtype).
 *  - cells are orthorhombic: h_cell3 = (Lx, Ly, Lz). (The reference accepts a general 3x3
 *    cell, topology.py:55-59; every config of the path is cubic.  The Python layer raises
 *    for non-diagonal cells instead of silently falling back.)
 */
#ifndef MDGRAD_B200_H
#define MDGRAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDG_OK            0
#define MDG_E_BADARG     -1
#define MDG_E_CUDA       -2
#define MDG_E_CAPACITY   -3   /* a fixed-capacity internal table overflowed; grow and retry */
#define MDG_E_STATE      -4   /* call order violated (e.g. export before build) */
#define MDG_E_SKIN       -5   /* an atom moved more than skin/2 between list rebuilds */
#define MDG_E_NCCL       -6
#define MDG_E_NUMERIC    -7   /* non-finite coordinates (diverged dynamics) or a collapsed cell */

/* pair potential kinds: reference torchmd/potentials.py */
#define MDG_POT_LJ        0   /* LennardJones   :317-327  params (sigma, epsilon)            */
#define MDG_POT_LJFAM     1   /* LJFamily       :61-73    params (sigma, epsilon, rep, attr) */
#define MDG_POT_LJ69      2   /* LennardJones69 :329-339  params (sigma, epsilon)            */
#define MDG_POT_EXV       3   /* ExcludedVolume :341-352  params (sigma, epsilon, power)     */
#define MDG_POT_BUCK      4   /* Buck           :354-365  params (A, B, C)                   */
#define MDG_POT_MORSE     5   /* ModifiedMorse  :75-93    params (a, phi)                    */
#define MDG_MAX_POT_PARAMS 4

/* integrators: reference torchmd/md.py + torchmd/sovlers.py */
#define MDG_INT_NVE       0   /* NVE.forward md.py:131-148 + verlet_update sovlers.py:25-40   */
#define MDG_INT_NHC       1   /* NoseHooverChain.forward md.py:210-240 + NHverlet_update
                                 sovlers.py:110-127                                          */
#define MDG_MAX_CHAINS    16

typedef struct mdg_ctx mdg_ctx;

int         mdg_version(void);
const char* mdg_last_error(void);

/* Create / destroy a context on CUDA device `device`. */
int mdg_create(int device, mdg_ctx** out);
int mdg_destroy(mdg_ctx* ctx);

/* ------------------------------------------------------------------------------------------
 * K1  neighbor list - replaces generate_nbr_list(xyz, cutoff, cell, index_tuple, ex_pairs,
 *     get_dis)  torchmd/topology.py:30-73 (and PairPotentials._reset_topology,
 *     torchmd/interface.py:263-282).
 *
 * mdg_nbr_build: bins the N atoms of d_xyz (N x 3 fp32, any order, need not be wrapped) into
 * cells by a device-wide counting sort and finds, for every atom, all neighbors with
 * d2 < fl32(cutoff^2) and d2 != 0 using the reference's exact fp32 arithmetic
 * (d = x_j - x_i; image = -(d/L > 0.5) + (d/L < -0.5); d += image*L; d2 = (dx^2+dy^2)+dz^2,
 * no FMA contraction; topology.py:35,59-67).  Systems with fewer than 3 cells on an axis (or
 * N <= small) use an all-pairs tiled search with identical single-image semantics.
 *   d_sel_a / d_sel_b : optional per-atom 0/1 flags = membership in index_tuple[0] / [1]
 *                       (generate_pair_index, topology.py:15-27); both NULL = all pairs.
 *   d_ex_keys, n_ex   : optional exclusions (ex_pairs, topology.py:44-53) as SORTED unique
 *                       int64 keys min(i,j)*N + max(i,j).
 * SYNC: returns in *h_npairs the number P of undirected pairs (i<j) so the caller can size
 * the export buffers.
 *
 * mdg_nbr_export: writes the list in the reference's exact layout and ORDER (row-major
 * (i, j) with i<j, torch.nonzero order, topology.py:68): d_nbr P x 2 int64, d_offsets P x 3
 * fp32 in {-1,0,1} (topology.py:60-62,73), d_dis P fp32 = sqrt(d2) or NULL (topology.py:71).
 * ------------------------------------------------------------------------------------------ */
int mdg_nbr_build(mdg_ctx* ctx, const float* d_xyz, int n, const float* h_cell3, double cutoff,
                  const uint8_t* d_sel_a, const uint8_t* d_sel_b,
                  const int64_t* d_ex_keys, int n_ex,
                  void* stream, int64_t* h_npairs);
int mdg_nbr_export(mdg_ctx* ctx, int64_t* d_nbr, float* d_offsets, float* d_dis, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2+K3  listed-pair energy and forces - replaces PairPotentials.forward
 *     (torchmd/interface.py:284-300: compute_dis topology.py:5-12 -> u(r).sum()) together
 *     with the autograd force F = -dE/dxyz (torchmd/md.py:227-228, nff/utils/scatter.py:5-21).
 *
 * Evaluates E = sum over the pairs of the list held by ctx (last mdg_nbr_build) with the
 * CURRENT d_xyz and the stored image offsets - no cutoff re-test, exactly as the reference
 * evaluates a (possibly stale) stored list.  Outputs (any may be NULL):
 *   d_energy  : 1 fp32 (total energy)
 *   d_force   : N x 3 fp32, F = -dE/dxyz
 *   d_dparams : MDG_MAX_POT_PARAMS fp32, dE/dparam for the differentiable parameters of the
 *               kind (sigma, epsilon | A, B, C)
 *   d_dis     : not provided here - use mdg_pair_dis_*.
 * ------------------------------------------------------------------------------------------ */
int mdg_pair_force(mdg_ctx* ctx, int kind, const float* h_params, int n_params,
                   const float* d_xyz, int n,
                   float* d_energy, float* d_force, float* d_dparams, void* stream);

/* ------------------------------------------------------------------------------------------
 * Adjoint support: analytic second-order products of the listed-pair force - replaces the
 * DOUBLE backward through compute_dis / u(r) that OdeintAdjointMethod.backward (torchmd/sovlers.py:211-293, with the
 * len(y)==8 branch of NHverlet_update :129-164) performs via torch.autograd.grad(..., create_graph=True) on
 * NoseHooverChain.forward / NVE.forward (torchmd/md.py:210-240 / :131-148).
 * Over the list held by ctx (last mdg_nbr_build, stored image offsets, no re-test - as mdg_pair_force) and a given
 * adjoint vector d_avec (N x 3):
 *   d_hv     (N x 3)  = (dF/dxyz)^T a   ( = -Hessian(E) a )
 *   d_dtheta (MDG_MAX_POT_PARAMS, may be NULL) = (dF/dparam)^T a  for the kind's parameters, in mdg_pair_force order
 * All analytic kinds: the power-law family (MDG_POT_LJ, _LJFAM, _LJ69, _EXV: sigma, epsilon), MDG_POT_BUCK (A, B, C) and
 * MDG_POT_MORSE (no differentiable parameters).  Learned u(r) keeps the autograd route in the Python layer.
 * ------------------------------------------------------------------------------------------ */
int mdg_pair_hvp(mdg_ctx* ctx, int kind, const float* h_params, int n_params,
                 const float* d_xyz, int n, const float* d_avec,
                 float* d_hv, float* d_dtheta, void* stream);

/* Generic listed-pair distance op for learned u(r) (pairMLP etc.): replaces compute_dis
 * (torchmd/topology.py:5-12) and its autograd backward.  The list is given explicitly in the
 * reference layout (d_nbr P x 2 int64, d_offsets P x 3 fp32).
 *   fwd: d_dis[p] = | x_i - x_j - offsets_p * cell |
 *   bwd: d_grad_xyz (N x 3, zero-initialised by the callee) += scatter of d_grad_dis[p] * unit vector */
int mdg_pair_dis_fwd(const float* d_xyz, int n, const int64_t* d_nbr, const float* d_offsets,
                     int64_t n_pairs, const float* h_cell3, float* d_dis, void* stream);
int mdg_pair_dis_bwd(const float* d_xyz, int n, const int64_t* d_nbr, const float* d_offsets,
                     int64_t n_pairs, const float* h_cell3, const float* d_dis,
                     const float* d_grad_dis, float* d_grad_xyz, void* stream);

/* ------------------------------------------------------------------------------------------
 * K6  RDF - replaces rdf.forward (torchmd/observable.py:62-76) + generate_vol_bins (:10-21):
 * Gaussian-smeared pair-distance histogram over all pairs with d < end + 0.5 of ONE frame
 * (call once per frame and sum counts for a batch): d_count[k] += sum_p exp(-0.5/w^2 (d_p - mu_k)^2),
 * mu = linspace(start, end, nbins), w = `width` (<=0: mu_1 - mu_0).  Normalisation
 * (count/sum, / (vol_bins/V)) is nbins-sized host-side algebra in the Python layer.
 * d_count must be zero-initialised by the caller (accumulates).
 * ------------------------------------------------------------------------------------------ */
int mdg_rdf_accumulate(mdg_ctx* ctx, const float* d_xyz, int n, const float* h_cell3,
                       double start, double end, int nbins, double width,
                       const uint8_t* d_sel_a, const uint8_t* d_sel_b,
                       float* d_count, void* stream);

/* ------------------------------------------------------------------------------------------
 * Velocity autocorrelation - replaces vacf.forward (torchmd/observable.py:153-163):
 * d_vel = (n_frames, n_atoms, dim) fp32; d_out[0] = mean(v * v), d_out[t] = mean(v[t:] * v[:-t]) for t < t_range
 * (un-normalised means over all elements, exactly what the reference returns).  1 <= t_range <= n_frames.
 * ------------------------------------------------------------------------------------------ */
int mdg_vacf(mdg_ctx* ctx, const float* d_vel, int n_frames, int n_atoms, int dim, int t_range, float* d_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K4 + driver  fused MD epoch - replaces the hot loop of Simulations.simulate
 *     (torchmd/md.py:73-96) -> odeint (torchmd/sovlers.py:171-193) ->
 *     FixedGridODESolver.integrate (torchmd/tinydiffeq.py:56-76) -> NHverlet_update /
 *     verlet_update (sovlers.py:110-127 / :25-40) -> NoseHooverChain.forward / NVE.forward
 *     (md.py:210-240 / :131-148) with a PairPotentials model whose list is rebuilt at every
 *     evaluation (topology_update_freq = 1).
 *
 * Integrates n_grid-1 steps over the fp32 time grid h_tgrid (n_grid points; per-step dt is
 * t[i+1]-t[i] in fp32 exactly as tinydiffeq.py:67-68) starting from (d_v0, d_q0, h_pv0) and
 * writes the stacked trajectory like the reference: d_traj_v / d_traj_q  (n_grid x N x 3,
 * frame 0 = initial state), h_traj_pv (n_grid x n_chains; NULL for NVE).  One force
 * evaluation per step (bitwise-equivalent to the reference's two).
 * The pair set contributing at every evaluation equals a freshly built reference list: a
 * Verlet list with `skin` is re-tested against fl32(cutoff^2) with the exact arithmetic of
 * mdg_nbr_build; it is rebuilt every `rebuild_every` steps and a device-side check raises
 * MDG_E_SKIN (state untouched) if any atom moved > skin/2 in between.
 * traj_stride > 1 keeps only every traj_stride-th grid point (extension; reference = 1).
 * SYNC at the end (h_traj_pv, error flags).
 * ------------------------------------------------------------------------------------------ */
typedef struct mdg_md_params {
    int    integrator;            /* MDG_INT_NVE | MDG_INT_NHC                                  */
    int    pot_kind;              /* MDG_POT_*                                                  */
    float  pot_params[MDG_MAX_POT_PARAMS];
    double cutoff;                /* as passed to PairPotentials(cutoff=)                       */
    float  cell[3];
    int    n_chains;              /* NHC: number of bath variables M                            */
    float  Q[MDG_MAX_CHAINS];     /* NHC bath masses (md.py:191-193), fp32 as the reference     */
    double T;                     /* target temperature in energy units (md.py:186), python float */
    int    ndof;                  /* N * dim (md.py:187)                                        */
    float  skin;                  /* Verlet skin (0 = rebuild every step, no re-test needed)    */
    int    rebuild_every;         /* steps between list rebuilds (>=1)                          */
    int    traj_stride;           /* 1 = every grid point (reference behaviour)                 */
} mdg_md_params;

/* Pair filter applied by mdg_md_run's internal list builds: the index_tuple / ex_pairs arguments
 * of PairPotentials (torchmd/interface.py:228, topology.py:15-27,44-53) in the same encoding as
 * mdg_nbr_build.  Pointers must stay valid until replaced; all NULL/0 = no filter. */
int mdg_set_pair_filter(mdg_ctx* ctx, const uint8_t* d_sel_a, const uint8_t* d_sel_b,
                        const int64_t* d_ex_keys, int n_ex);

int mdg_md_run(mdg_ctx* ctx, const mdg_md_params* p, int n, const float* d_mass,
               const float* d_v0, const float* d_q0, const float* h_pv0,
               const float* h_tgrid, int n_grid,
               float* d_traj_v, float* d_traj_q, float* h_traj_pv,
               float* h_last_energy, void* stream);

/* Statistics of the last mdg_md_run / mdg_nbr_build on this context (for bench / tests):
 * out[0]=kernels launched, out[1]=list rebuilds, out[2]=directed list entries (last build),
 * out[3]=max row length, out[4]=ncell_x, out[5]=ncell_y, out[6]=ncell_z, out[7]=path
 * (0 = cell list, 1 = all-pairs, 2 = cell list in the engine's tile form). */
int mdg_get_stats(mdg_ctx* ctx, int64_t* h_out8);

/* ------------------------------------------------------------------------------------------
 * K5  SchNet continuous-filter convolution, aggregation part - replaces the gather-multiply-
 *     scatter of the reference message passing: SchNetConv.message nff/nn/modules.py:568-572
 *     (h[a0]*W, h[a1]*W) + MessagePassingModule.forward/aggregate nff/nn/graphconv.py:43-53
 *     (two scatter_add, nff/utils/scatter.py:24-45):
 *         out[k] = sum over edges e incident to node k of h[other(e,k)] * W[e]     (N x F)
 * mdg_graph_build  : node -> incident-edge CSR from the reference-layout list d_nbr (E x 2 int64,
 *                    must stay valid until the next build); deterministic order, no atomics in
 *                    the reduction.
 * mdg_cfconv_agg   : the forward above; ALSO the gradient w.r.t. h (apply it to the upstream grad).
 * mdg_cfconv_edge_grad : gW[e] = h[a0]*g[a1] + h[a1]*g[a0]                        (E x F)
 * ------------------------------------------------------------------------------------------ */
int mdg_graph_build(mdg_ctx* ctx, const int64_t* d_nbr, int64_t n_edges, int n, void* stream);
int mdg_cfconv_agg(mdg_ctx* ctx, const float* d_h, const float* d_W, int n, int n_filters, float* d_out, void* stream);
int mdg_cfconv_edge_grad(mdg_ctx* ctx, const float* d_h, const float* d_g, int n, int n_filters, float* d_gW, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5'  SchNet energy + forces in one call - replaces SchNet.forward (nff/nn/models/schnet.py:113-171:
 *     convolve -> atomwisereadout -> batch_and_sum) over a given neighbor list together with the autograd
 *     force F = -dE/dxyz (torchmd/md.py:227-228) for a GNNPotentials model (torchmd/interface.py:125-136).
 *
 * The model is described by DEVICE pointers to its parameters in the reference's `state_dict` layout (torch
 * Linear weight = out x in, row-major):
 *   embed  : atom_embed.weight (100 x A)
 *   layer l: mu/width = convolutions.l.moduledict.message_edge_filter.0.{offsets,width} (G)
 *            We1,be1  = ...message_edge_filter.1.{weight,bias}   (G x G, G)
 *            We2,be2  = ...message_edge_filter.3.{weight,bias}   (F x G, F)
 *            Wn,bn    = ...message_node_filter.{weight,bias}     (F x A, F)
 *            Wu1,bu1  = ...update_function.0.{weight,bias}       (A x F, A)
 *            Wu2,bu2  = ...update_function.2.{weight,bias}       (A x A, A)
 *   readout: Wr1,br1  = atomwisereadout.readout.energy.linear0   (R x A, R),  Wr2,br2 = ...linear2 (1 x R, 1)
 * Inputs: d_z (N int64 atomic numbers), d_xyz (N x 3), the list in the reference layout (d_nbr E x 2 int64 with
 * i < j, d_offsets E x 3 fp32) and h_off_scale3: edge vector = x_i - x_j - offsets * h_off_scale3 - (1,1,1)
 * reproduces the reference's raw-offset quirk (schnet.py:140-142), the cell lengths give true PBC.
 * Outputs: d_energy (1 fp32), d_force (N x 3 fp32, may be NULL = energy only).  Asynchronous on `stream`.
 * ------------------------------------------------------------------------------------------ */
#define MDG_SCHNET_MAX_LAYERS 8
typedef struct mdg_schnet_layer {
    const float *mu, *width;
    const float *We1, *be1, *We2, *be2;
    const float *Wn, *bn;
    const float *Wu1, *bu1, *Wu2, *bu2;
} mdg_schnet_layer;

typedef struct mdg_schnet_model {
    int n_atom_basis, n_filters, n_gaussians, n_convolutions, n_readout;   /* A, F, G, L, R */
    const float* embed;
    mdg_schnet_layer layers[MDG_SCHNET_MAX_LAYERS];
    const float *Wr1, *br1, *Wr2, *br2;
    uint64_t weights_tag;   /* 0 = unknown: derived weight layouts are rebuilt at every call.  Non-zero: the caller promises to
                               change the tag whenever any parameter VALUE changed, so the library may cache the transposed
                               filter weights it derives (the Python layer hashes the tensors' data pointers and version counters) */
} mdg_schnet_model;

int mdg_schnet_energy_force(mdg_ctx* ctx, const mdg_schnet_model* h_model, const int64_t* d_z,
                            const float* d_xyz, int n, const int64_t* d_nbr, const float* d_offsets,
                            int64_t n_edges, const float* h_off_scale3,
                            float* d_energy, float* d_force, void* stream);

/* ------------------------------------------------------------------------------------------
 * K4 + K5  fused MD epoch with a SchNet force field (+ analytic pair priors) - the epoch loop of
 *     Simulations.simulate (torchmd/md.py:73-96 -> tinydiffeq.py:56-76 -> sovlers.py:110-127 / :25-40 ->
 *     md.py:210-240 / :131-148) for a model = GNNPotentials (torchmd/interface.py:86-136) or a Stack
 *     (:364-403) of one GNNPotentials and PairPotentials priors, topology_update_freq = 1: every step rebuilds
 *     each member's exact neighbor list at the current positions (mdg_nbr_build semantics, each member with its
 *     own cutoff / species selection / exclusions), evaluates SchNet energy+forces (mdg_schnet_energy_force) and
 *     the priors (mdg_pair_force), and applies the fused integrator kernels - all from one host call, no Python
 *     and no autograd inside the loop.  Same trajectory outputs as mdg_md_run.  SYNC: the FIRST evaluation of an
 *     epoch reads its pair counts back (they size the edge buffers, +25%) unless the previous epoch of the same
 *     context and atom count completed asynchronously (its capacity is reused); every later step is enqueued without any
 *     read-back - the pair count is consumed on the device, a capacity that turns out too small is latched on the
 *     device, read once at the end of the epoch, and the epoch is then repeated with a read-back per list build
 *     (inputs are never modified; MDG_GNN_SYNC=1 forces that path).  mdg_get_stats slot 3 = 1 if the epoch
 *     completed on the asynchronous path.
 *     ctx owns the GNN list and the SchNet workspace; every prior brings the context that owns its list.
 *     h_model == NULL: a Stack of analytic PairPotentials only (e.g. the three species-pair members of
 *     scripts/fit_2_comp.py:182), every member evaluated on its own exact per-step list.
 * ------------------------------------------------------------------------------------------ */
#define MDG_MAX_PRIORS 4

/* ------------------------------------------------------------------------------------------
 * Bonded terms: BondPotentials.forward (torchmd/interface.py:436-451) and AnglePotentials.forward
 *     (:489-510) with their autograd forces (md.py:227-228) - the `prior` member of demo/fold.py:131.
 *     bond (i, j):     v = x_i - x_j + get_offsets(v) * L (topology.py:74-80),  b = |v|^2 (SQUARED, as the reference),
 *                      E = 0.5 k sum (b - ro)^2
 *     angle (a, c, e): E = 0.5 k sum (acos(v1.v2 / sqrt(|v1|^2 |v2|^2)) - theta0)^2,  v1 = x_a - x_c, v2 = x_e - x_c
 *     d_energy2 = (E_bond, E_angle); d_force (n,3) = -dE/dx of both (overwritten); d_dparams4 = dE/d(k_bond, ro,
 *     k_angle, theta0); each may be NULL.  Forces need the atom -> term reference list of the (static) topology:
 *     d_ref_start (n+1) and d_refs, refs of atom i = d_refs[d_ref_start[i] .. d_ref_start[i+1]), each
 *     ref = slot * 4 + role with slot = t for bond t, n_bonds + 2 t for angle t; role 0 = first atom of a bond /
 *     first atom of an angle (and, with slot + 1, its third atom), 1 = second atom of a bond, 2 = centre atom of an
 *     angle (mdgrad_b200/interface.py builds it once per topology).  Terms naming atoms outside [0, n) contribute
 *     nothing.  Deterministic (no atomics).  Asynchronous.
 * ------------------------------------------------------------------------------------------ */
typedef struct mdg_bonded_terms {
    const int64_t* d_bond_top;                    /* (n_bonds, 2) int64 or NULL                  */
    int            n_bonds;
    float          k_bond, r0;
    const int64_t* d_angle_top;                   /* (n_angles, 3) int64 or NULL                 */
    int            n_angles;
    float          k_angle, theta0;
    const int32_t* d_ref_start;                   /* (n + 1)                                      */
    const int32_t* d_refs;
} mdg_bonded_terms;

int mdg_bonded_force(mdg_ctx* ctx, const mdg_bonded_terms* h_terms, const float* d_xyz, int n, const float* h_cell3,
                     float* d_energy2, float* d_force, float* d_dparams4, void* stream);

typedef struct mdg_prior_spec {
    mdg_ctx*       ctx;
    int            kind;                          /* MDG_POT_*                                   */
    float          params[MDG_MAX_POT_PARAMS];
    int            n_params;
    double         cutoff;
    const uint8_t* d_sel_a;                       /* index_tuple flags or NULL (as mdg_nbr_build) */
    const uint8_t* d_sel_b;
    const int64_t* d_ex_keys;
    int            n_ex;
} mdg_prior_spec;

typedef struct mdg_gnn_md_params {
    int     integrator;                           /* MDG_INT_NVE | MDG_INT_NHC                   */
    int     n_chains;
    float   Q[MDG_MAX_CHAINS];
    double  T;
    int     ndof;
    float   cell[3];
    double  cutoff;                               /* GNNPotentials(cutoff=)                      */
    float   off_scale[3];                         /* see mdg_schnet_energy_force                 */
    const int64_t* d_ex_keys;                     /* GNNPotentials(ex_pairs=) as sorted keys     */
    int     n_ex;
    int     n_priors;
    mdg_prior_spec priors[MDG_MAX_PRIORS];
    int     traj_stride;
    mdg_bonded_terms bonded;                      /* BondPotentials / AnglePotentials members of the Stack
                                                     (n_bonds = n_angles = 0: none)               */
} mdg_gnn_md_params;

int mdg_md_run_gnn(mdg_ctx* ctx, const mdg_gnn_md_params* p, const mdg_schnet_model* h_model,
                   const int64_t* d_z, int n, const float* d_mass, const float* d_v0, const float* d_q0,
                   const float* h_pv0, const float* h_tgrid, int n_grid,
                   float* d_traj_v, float* d_traj_q, float* h_traj_pv, float* h_last_energy, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU (one process per GPU).  The reference has no distributed code; this is the
 * spatial decomposition of SURVEY 8e: slabs of whole z-layers of cells in the global cell-sorted
 * index space, per-step ghost-position halo (ncclSend/ncclRecv of two contiguous ranges), one
 * 2-double KE all-reduce per step, state all-gather + identical re-sort at every list rebuild.
 *   mdg_slab_plan      : host-only; layers [out4[0], out4[1]) and the lower/upper neighbour ranks.
 *   mdg_dist_unique_id : rank 0 creates the NCCL unique id (128 bytes); broadcast it out of band
 *                        (e.g. torch.distributed) and call mdg_dist_init on every rank.
 *   nccl_lib_path      : the libnccl.so.2 the process already uses (torch's), NULL = default search.
 * After mdg_dist_init(world > 1) mdg_md_run integrates only the context's own slab: every rank
 * passes the same full (N) inputs; trajectory frames hold the owned atoms only (zero elsewhere;
 * sum across ranks to assemble) ; bath trajectory and energy are replicated.
 * ------------------------------------------------------------------------------------------ */
int mdg_slab_plan(int ncz, int world, int rank, int* h_out4);
int mdg_dist_unique_id(const char* nccl_lib_path, char* h_out128);
int mdg_dist_init(mdg_ctx* ctx, const char* nccl_lib_path, const char* h_id128, int rank, int world);
int mdg_dist_finalize(mdg_ctx* ctx);

/* Measurement aid (bench.py roofline): when enabled, mdg_md_run brackets every pair-force kernel
 * launch with CUDA events on the launch stream; mdg_get_profile returns h_out2[0] = summed
 * force-kernel milliseconds and h_out2[1] = number of force launches of the last run. */
int mdg_set_profile(mdg_ctx* ctx, int enable);
int mdg_get_profile(mdg_ctx* ctx, double* h_out2);

#ifdef __cplusplus
}
#endif
#endif /* MDGRAD_B200_H */
