"""CPU oracle for the MD hot path - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, in plain torch-CPU fp32 tensor algebra, of the algorithm the
reference (torchmd/mdgrad @ cea2332e, pure Python/PyTorch) runs for the path in
SURVEY.md section 8: all-pairs minimum-image neighbor list, listed-pair
distances, analytic pair energies with autograd forces, the Nose-Hoover-chain /
NVE equations of motion, the (NH-)velocity-Verlet step, the epoch driver and the
Gaussian-smeared RDF.  Each function cites the reference file:line it follows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module, and only as the checker / CPU baseline.
The product path (mdgrad_b200/*.py + libmdgrad_b200.so) never does.

Parity pinning: validated in this container against the reference itself
imported in-process (oracle/ref_import.py) - see tests/test_oracle_vs_reference.py
(runs only where /root/reference exists) - and against the committed fixtures in
tests/golden/ produced by oracle/make_golden.py from the reference.  The
reference's only own known answer for the path (torchmd/topology.py:126-147:
FCC 3x3x3, a=1.679, rc=2.5 -> 5832 directed pairs) is checked in
tests/test_oracle_golden.py.

The O(N^2) temporaries are evaluated in row blocks so the oracle reaches ~64k
atoms; the arithmetic per element (and therefore every bit of the result) is the
same as the unblocked reference expression.
"""
import math

import numpy as np
import torch

F32 = torch.float32


# ----------------------------------------------------------------------------
# A1  neighbor list  (reference torchmd/topology.py:30-73)
# ----------------------------------------------------------------------------
def pair_selection_mask(n, index_tuple):
    """Dense species-selection mask (reference torchmd/topology.py:15-27): 1 where
    (i in A and j in B) or (i in B and j in A)."""
    m = torch.zeros(n, n)
    if index_tuple is not None:
        a = torch.as_tensor(np.asarray(index_tuple[0]), dtype=torch.long)
        b = torch.as_tensor(np.asarray(index_tuple[1]), dtype=torch.long)
        ia = a.repeat_interleave(len(b))
        ib = b.repeat(len(a))
        m[ia, ib] = 1
        m[ib, ia] = 1
    return m


def neighbor_list(xyz, cutoff, cell, index_tuple=None, ex_pairs=None, get_dis=False,
                  block=1024):
    """All-pairs minimum-image list for one frame.

    xyz (N,3) fp32, cell (3,3) or (3,) fp32.  Returns nbr (P,2) int64 with i<j in
    row-major (i, then j) order, offsets (P,3) fp32 in {-1,0,1}, and dis (P,) if
    get_dis.  Follows torchmd/topology.py:35 (d = x_j - x_i), :37-53 (masks zero the
    displacement), :55-56 (1-D cell -> diag), :59-62 (strict +-0.5 image test on
    d @ inv(cell)), :64, :66-68 (upper triangle, d2 < cutoff**2, d2 != 0, nonzero).
    """
    xyz = xyz.detach().to(F32)
    n = xyz.shape[-2]
    cell = cell.detach().to(F32)
    if cell.dim() == 1:
        cell = torch.diag(cell)
    inv = cell.inverse()
    half = torch.tensor([0.5, 0.5, 0.5])
    sel = pair_selection_mask(n, index_tuple) if index_tuple is not None else None
    exm = None
    if ex_pairs is not None:
        ex_pairs = torch.as_tensor(ex_pairs, dtype=torch.long)
        exm = torch.ones(n, n)
        exm[ex_pairs[:, 0], ex_pairs[:, 1]] = 0
        exm[ex_pairs[:, 1], ex_pairs[:, 0]] = 0
    rc2 = cutoff ** 2  # python double; compared against fp32 tensor -> rounded to fp32
    nbrs, offs, diss = [], [], []
    for i0 in range(0, n, block):
        i1 = min(n, i0 + block)
        d = xyz[None, :, :] - xyz[i0:i1, None, :]            # (b,N,3)  x_j - x_i
        if sel is not None:
            d = d * sel[i0:i1, :, None]
        if exm is not None:
            d = d * exm[i0:i1, :, None]
        red = d.matmul(inv)
        off = -(red > half).to(F32) + (red < -half).to(F32)
        d = d + off.matmul(cell)
        d2 = d.pow(2).sum(-1)                                # (b,N)
        jj = torch.arange(n)[None, :]
        ii = torch.arange(i0, i1)[:, None]
        upper = jj >= ii                                      # triu incl. diagonal (d2==0 there)
        d2u = torch.where(upper, d2, torch.zeros_like(d2))
        keep = (d2u < rc2) & (d2u != 0)
        idx = torch.nonzero(keep, as_tuple=False)
        gi = idx[:, 0] + i0
        nbrs.append(torch.stack([gi, idx[:, 1]], 1))
        offs.append(off[idx[:, 0], idx[:, 1], :])
        if get_dis:
            diss.append(d2u[keep].sqrt())
    nbr = torch.cat(nbrs) if nbrs else torch.zeros(0, 2, dtype=torch.long)
    off = torch.cat(offs) if offs else torch.zeros(0, 3)
    if get_dis:
        return nbr, torch.cat(diss) if diss else torch.zeros(0), off
    return nbr, off


# ----------------------------------------------------------------------------
# A2  listed-pair distance  (reference torchmd/topology.py:5-12)
# ----------------------------------------------------------------------------
def pair_distance(xyz, nbr, offsets, cell):
    """|x_i - x_j - offsets @ cell| for the listed pairs, shape (P,1)."""
    if cell.dim() == 1:
        cell = torch.diag(cell)
    return (xyz[nbr[:, 0]] - xyz[nbr[:, 1]] - offsets.matmul(cell)).pow(2).sum(1).sqrt()[:, None]


# ----------------------------------------------------------------------------
# A2/a3  analytic pair energies u(r)  (reference torchmd/potentials.py)
# ----------------------------------------------------------------------------
def u_pair(r, kind, p):
    """kind/params follow the reference classes:
    'lj'   LennardJones      potentials.py:317-327   p=(sigma, epsilon)
    'ljfam' LJFamily          potentials.py:61-73    p=(sigma, epsilon, rep_pow, attr_pow)
    'lj69' LennardJones69    potentials.py:329-339   p=(sigma, epsilon)
    'exv'  ExcludedVolume    potentials.py:341-352   p=(sigma, epsilon, power)
    'buck' Buck              potentials.py:354-365   p=(A, B, C)
    'morse' ModifiedMorse    potentials.py:75-93     p=(a, phi)
    """
    if kind == "lj":
        s, e = p[0], p[1]
        return 4 * e * ((s / r) ** 12 - (s / r) ** 6)
    if kind == "ljfam":
        s, e, rp, ap = p
        return 4 * e * ((s / r) ** rp - (s / r) ** ap)
    if kind == "lj69":
        s, e = p[0], p[1]
        return 4 * e * ((s / r) ** 9 - (s / r) ** 6)
    if kind == "exv":
        s, e, pw = p
        return 4 * e * ((s / r) ** pw)
    if kind == "buck":
        A, B, C = p
        return A * torch.exp(-B * r) - C / r ** 6
    if kind == "morse":
        a, phi = p
        A0 = 0.0 if phi >= 0 else math.exp(2 * a / phi) - 2 * math.exp(a / phi)
        ex = a * (1 - r ** phi) / phi
        return (torch.exp(2 * ex) - 2 * torch.exp(ex) - A0) / (1 + A0)
    raise ValueError(kind)


def pair_energy_forces(xyz, nbr, offsets, cell, kind, params, need_param_grads=False):
    """E = sum u(r_ij) (reference torchmd/interface.py:298-300) and F = -dE/dxyz by
    reverse-mode autograd (reference torchmd/md.py:227-228, nff/utils/scatter.py:5-21).
    Returns (E 0-d, F (N,3)[, dE/dparams list])."""
    q = xyz.detach().clone().requires_grad_(True)
    ps = [torch.tensor([float(v)], requires_grad=need_param_grads) for v in params]
    if kind in ("ljfam", "exv", "morse"):
        # integer/float exponents are plain python numbers in the reference
        if kind == "ljfam":
            pp = (ps[0], ps[1], params[2], params[3])
        elif kind == "exv":
            pp = (ps[0], ps[1], params[2])
        else:
            pp = (float(params[0]), float(params[1]))
    else:
        pp = tuple(ps)
    r = pair_distance(q, nbr, offsets, cell)
    e = u_pair(r, kind, pp).sum()
    n_tensor_params = {"lj": 2, "ljfam": 2, "lj69": 2, "exv": 2, "buck": 3, "morse": 0}[kind]
    wrt = [q] + (ps[:n_tensor_params] if need_param_grads else [])
    grads = torch.autograd.grad(e, wrt, allow_unused=True)
    out = (e.detach(), -grads[0])
    if need_param_grads:
        out = out + (list(grads[1:]),)
    return out


# ----------------------------------------------------------------------------
# A3  equations of motion  (reference torchmd/md.py:131-148, 179-240)
# ----------------------------------------------------------------------------
def nhc_bath_masses(Q, n_atoms, num_chains):
    """Q = [Q, Q/N, ..., Q/N] (reference torchmd/md.py:191-193), built in numpy fp64 then
    narrowed to fp32 exactly as torch.Tensor(np.array) does."""
    q = np.array([Q, *[Q / n_atoms] * (num_chains - 1)])
    return torch.Tensor(q)


def nhc_derivative(v, f, pv, mass, Qb, T, ndof):
    """(dv/dt, dq/dt, dpv/dt) of reference NoseHooverChain.forward, torchmd/md.py:221-240,
    with the force f supplied by the caller."""
    m = mass[:, None]
    p = v * m
    ke = 0.5 * (p.pow(2) / m).sum()
    coupled = (pv[0] * p.reshape(-1) / Qb[0]).reshape(-1, 3)
    dpdt = f - coupled
    d0 = 2 * (ke - T * ndof * 0.5) - pv[0] * pv[1] / Qb[1]
    dmid = (pv[:-2].pow(2) / Qb[:-2] - T) - pv[2:] * pv[1:-1] / Qb[2:]
    dlast = pv[-2].pow(2) / Qb[-2] - T
    return dpdt / m, v, torch.cat((d0[None], dmid, dlast[None]))


def nve_derivative(v, f):
    """reference NVE.forward, torchmd/md.py:131-148: dv/dt = f (no mass division), dq/dt = v."""
    return f, v


def time_grid(dt, frequency):
    """fp32 grid t_i = fl32(dt*i) (reference torchmd/md.py:81)."""
    return torch.Tensor([dt * i for i in range(frequency)])


class PairSystemOracle:
    """Bundles what one force evaluation of the reference does for a PairPotentials model
    with topology_update_freq=1: rebuild list at q (md.py:200-204 -> interface.py:263-282),
    then E, F (interface.py:284-300, md.py:227-228)."""

    def __init__(self, cell, cutoff, kind, params, index_tuple=None, ex_pairs=None):
        self.cell = torch.as_tensor(cell, dtype=F32)
        self.cutoff, self.kind, self.params = cutoff, kind, params
        self.index_tuple, self.ex_pairs = index_tuple, ex_pairs
        self.n_eval = 0

    def force(self, q):
        nbr, off = neighbor_list(q, self.cutoff, self.cell, self.index_tuple, self.ex_pairs)
        e, f = pair_energy_forces(q, nbr, off, self.cell, self.kind, self.params)
        self.n_eval += 1
        self.last_energy = e
        return f


def nh_verlet_step_2eval(force_fn, v, q, pv, dt, mass, Qb, T, ndof):
    """One forward NH-Verlet step exactly as the reference does it, with TWO force
    evaluations (reference torchmd/sovlers.py:110-127 + tinydiffeq.py:67-70)."""
    a0, _, dp0 = nhc_derivative(v, force_fn(q), pv, mass, Qb, T, ndof)
    vh = 1 / 2 * a0 * dt
    ph = 1 / 2 * dp0 * dt
    dq = (v + vh) * dt
    a1, _, dp1 = nhc_derivative(v + vh, force_fn(q + dq), pv + ph, mass, Qb, T, ndof)
    dv = vh + 1 / 2 * a1 * dt
    dpv = ph + 1 / 2 * dp1 * dt
    return v + dv, q + dq, pv + dpv


def nh_verlet_trajectory(force_fn, v0, q0, pv0, t, mass, Qb, T, ndof, reuse_force=True):
    """Integrate over the fp32 grid t (frequency points -> frequency-1 steps) and stack every
    grid point, as reference FixedGridODESolver.integrate (tinydiffeq.py:56-76).
    reuse_force=True is the one-evaluation-per-step form (the second
    evaluation of step n and the first of step n+1 are at the same q); it is bitwise
    identical to the two-evaluation form."""
    vs, qs, ps = [v0], [q0], [pv0]
    v, q, pv = v0, q0, pv0
    f = force_fn(q) if reuse_force else None
    for i in range(len(t) - 1):
        dt = t[i + 1] - t[i]
        if not reuse_force:
            v, q, pv = nh_verlet_step_2eval(force_fn, v, q, pv, dt, mass, Qb, T, ndof)
        else:
            a0, _, dp0 = nhc_derivative(v, f, pv, mass, Qb, T, ndof)
            vh = 1 / 2 * a0 * dt
            ph = 1 / 2 * dp0 * dt
            dq = (v + vh) * dt
            f = force_fn(q + dq)
            a1, _, dp1 = nhc_derivative(v + vh, f, pv + ph, mass, Qb, T, ndof)
            v, q, pv = v + (vh + 1 / 2 * a1 * dt), q + dq, pv + (ph + 1 / 2 * dp1 * dt)
        vs.append(v), qs.append(q), ps.append(pv)
    return torch.stack(vs), torch.stack(qs), torch.stack(ps)


def verlet_trajectory(force_fn, v0, q0, t):
    """NVE velocity Verlet, reference torchmd/sovlers.py:25-40 with NVE.forward (md.py:131-148):
    a = f (no mass division)."""
    vs, qs = [v0], [q0]
    v, q = v0, q0
    f = force_fn(q)
    for i in range(len(t) - 1):
        dt = t[i + 1] - t[i]
        vh = 0.5 * f * dt
        dq = (v + vh) * dt
        f = force_fn(q + dq)
        v, q = v + (vh + 0.5 * f * dt), q + dq
        vs.append(v), qs.append(q)
    return torch.stack(vs), torch.stack(qs)


# ----------------------------------------------------------------------------
# A6  radial distribution function  (reference torchmd/observable.py:10-21, 33-76)
# ----------------------------------------------------------------------------
def rdf(xyz, cell_diag, nbins, r_range, index_tuple=None, width=None):
    """Gaussian-smeared g(r) of one frame (N,3).  bins = linspace(start,end,nbins+1),
    vol_bins = 4pi/3 (b1^3-b0^3), V = 4pi/3 end^3, centres linspace(start,end,nbins),
    list cutoff end+0.5 (observable.py:59).  Returns (count, bins, g)."""
    start, end = r_range
    bins = torch.linspace(start, end, nbins + 1)
    vbins = torch.Tensor(4 * np.pi / 3 * (bins[1:] ** 3 - bins[:-1] ** 3))
    V = (4 / 3) * np.pi * end ** 3
    mu = torch.linspace(start, bins[-1], nbins)
    w = (mu[1] - mu[0]) * torch.ones_like(mu) if width is None else width * torch.ones_like(mu)
    nbr, dis, _ = neighbor_list(xyz, end + 5e-1, torch.as_tensor(cell_diag, dtype=F32),
                                index_tuple=index_tuple, get_dis=True)
    coeff = -0.5 / torch.pow(w, 2)
    count = torch.zeros(nbins)
    for c0 in range(0, dis.shape[0], 1 << 18):
        d = dis[c0:c0 + (1 << 18), None] - mu
        count = count + torch.exp(coeff * torch.pow(d, 2)).sum(0)
    count = count / count.sum()
    return count, bins, count / (vbins / V)


# ----------------------------------------------------------------------------
# A5  SchNet energy  (reference nff/nn/models/schnet.py:113-171, nff/nn/modules.py:550-575,
#     nff/nn/graphconv.py:43-53, nff/nn/layers.py:14-31, nff/nn/activations.py:5-11)
# ----------------------------------------------------------------------------
def ssp(x):
    return torch.nn.functional.softplus(x) - math.log(2.0)


def schnet_energy(sd, z, xyz, nbr, offsets, cell=None, pbc_mode="reference"):
    """Flat edge/node program of the reference SchNet forward given its `state_dict` sd.
    pbc_mode='reference' subtracts the raw integer offsets (schnet.py:140-142 quirk,
    SURVEY 3c); 'correct' subtracts offsets @ cell."""
    n_conv = len({k.split(".")[1] for k in sd if k.startswith("convolutions.")})
    a0, a1 = nbr[:, 0], nbr[:, 1]
    shift = offsets if pbc_mode == "reference" else offsets.matmul(cell)
    e = (xyz[a0] - xyz[a1] - shift).pow(2).sum(1).sqrt()[:, None]
    r = sd["atom_embed.weight"][z]
    n = r.shape[0]
    for l in range(n_conv):
        pre = "convolutions.%d.moduledict." % l
        mu, w = sd[pre + "message_edge_filter.0.offsets"], sd[pre + "message_edge_filter.0.width"]
        g = torch.exp(-0.5 / torch.pow(w, 2) * torch.pow(e - mu, 2))
        W = torch.nn.functional.linear(g, sd[pre + "message_edge_filter.1.weight"],
                                       sd[pre + "message_edge_filter.1.bias"])
        W = torch.nn.functional.linear(ssp(W), sd[pre + "message_edge_filter.3.weight"],
                                       sd[pre + "message_edge_filter.3.bias"])
        h = torch.nn.functional.linear(r, sd[pre + "message_node_filter.weight"],
                                       sd[pre + "message_node_filter.bias"])
        agg = torch.zeros(n, W.shape[1]).index_add(0, a1, h[a0] * W)
        agg = agg.index_add(0, a0, h[a1] * W)
        u = torch.nn.functional.linear(agg, sd[pre + "update_function.0.weight"],
                                       sd[pre + "update_function.0.bias"])
        u = torch.nn.functional.linear(ssp(u), sd[pre + "update_function.2.weight"],
                                       sd[pre + "update_function.2.bias"])
        r = r + u
    ro = "atomwisereadout.readout.energy."
    y = torch.nn.functional.linear(r, sd[ro + "linear0.weight"], sd[ro + "linear0.bias"])
    y = torch.nn.functional.linear(ssp(y), sd[ro + "linear2.weight"], sd[ro + "linear2.bias"])
    return y.sum()


# ----------------------------------------------------------------------------
# synthetic systems (SURVEY 8d generators; numpy default_rng(seed))
# ----------------------------------------------------------------------------
# ----------------------------------------------------------------------------
# bonded / electrostatic Stack members  (reference torchmd/interface.py:406-510, :303-361)
# ----------------------------------------------------------------------------
def image_offsets(vecs, cell3):
    """get_offsets (reference torchmd/topology.py:74-80): -(v >= L/2) + (v < -L/2)"""
    return -vecs.ge(0.5 * cell3).to(F32) + vecs.lt(-0.5 * cell3).to(F32)


def bond_energy(xyz, top, cell3, k, ro):
    """BondPotentials.forward (interface.py:444-451): 0.5 k sum (|v|^2 - ro)^2 - the SQUARED length is compared with ro"""
    v = xyz[top[:, 0]] - xyz[top[:, 1]]
    v = v + image_offsets(v, cell3) * cell3
    return 0.5 * k * (v.pow(2).sum(-1) - ro).pow(2).sum(-1)


def angle_energy(xyz, top, cell3, k, theta0):
    """AnglePotentials.forward (interface.py:497-510)"""
    v1 = xyz[top[:, 0]] - xyz[top[:, 1]]
    v2 = xyz[top[:, 2]] - xyz[top[:, 1]]
    v1 = v1 + image_offsets(v1, cell3) * cell3
    v2 = v2 + image_offsets(v2, cell3) * cell3
    cos = (v1 * v2).sum(-1) / (v1.pow(2).sum(-1) * v2.pow(2).sum(-1)).sqrt()
    return 0.5 * k * (torch.acos(cos) - theta0).pow(2).sum(-1)


def coulomb_energy(xyz, charges, cell3, cutoff, conversion, ex_pairs=None):
    """Electrostatics.forward (interface.py:347-360) INCLUDING its overwritten first charge: U = -conv sum q_j^2 / r over the
    minimum-image list (rebuilt at the call)"""
    nbr, dis, _ = neighbor_list(xyz, cutoff, cell3, ex_pairs=ex_pairs, get_dis=True)
    # differentiable distances over the (constant) list
    off = neighbor_list(xyz.detach(), cutoff, cell3, ex_pairs=ex_pairs)[1]
    r = pair_distance(xyz, nbr, off, cell3).squeeze(-1)
    qj = charges[nbr[:, 1]]
    return (-conversion * (qj * qj / r)).sum()


def fcc_positions(ncell, a):
    """FCC lattice, unit cells in (i outer, j, k inner) order, basis innermost - the ASE
    FaceCenteredCubic ordering."""
    basis = np.array([(0, 0, 0), (0.5, 0.5, 0), (0.5, 0, 0.5), (0, 0.5, 0.5)], dtype=np.float64)
    g = np.stack(np.meshgrid(np.arange(ncell), np.arange(ncell), np.arange(ncell),
                             indexing="ij"), -1).reshape(-1, 1, 3)
    return ((g + basis[None]) * a).reshape(-1, 3)


def lj_system(ncell, rho=0.845, jitter=0.05, T=1.0, mass=1.008, seed=1, a=None):
    """Positions (fp64), velocities (fp64), box length of an FCC LJ box."""
    a = (4.0 / rho) ** (1.0 / 3.0) if a is None else a
    pos = fcc_positions(ncell, a)
    L = ncell * a
    if jitter:
        pos = pos + np.random.default_rng(seed).normal(0.0, jitter * a, pos.shape)
    vel = np.random.default_rng(seed + 1).standard_normal(pos.shape) * math.sqrt(T / mass)
    return pos, vel, L
