"""Bonded and electrostatic members of a `Stack` - mirror of reference torchmd/interface.py:
BondPotentials :406-454, AnglePotentials :456-510, Electrostatics :303-361 (the `prior` of
demo/fold.py:130-131).

Same constructor signatures, attributes (.cell .k .ro / .thetao .top .device, .charges .conversion .cutoff
.index_tuple .ex_pairs) and `forward(xyz) -> energy`.  Bond / angle energies, forces and dE/d(k, ro | thetao) come
from one native program (`mdg_bonded_force`, csrc/bonded.cu: term kernel + atomics-free per-atom gather over a CSR of
the static topology); a pure-torch restatement serves double backward (adjoint reverse sweep), as for the pair terms.
Electrostatics runs on the native neighbor list + distance op with the charge algebra in torch.
"""
import numpy as np
import torch

from . import _lib
from .topology import _exclusion_keys, _selection_flags, cell_lengths, compute_dis, compute_dis_torch, get_offsets


def _device_of(system):
    d = system.device
    return torch.device("cuda:%d" % d if isinstance(d, int) else d)


def term_refs(n_atoms, bond_top=None, angle_top=None):
    """atom -> term reference list of a static bonded topology (layout: include/mdgrad_b200.h, mdg_bonded_terms):
    returns (ref_start int32 (n_atoms+1), refs int32).  refs of one atom are ordered by slot (= term order)."""
    nb = 0 if bond_top is None else int(bond_top.shape[0])
    atoms, refs = [], []
    if nb:
        b = np.asarray(bond_top, dtype=np.int64).reshape(-1, 2)
        t = np.arange(nb, dtype=np.int64)
        atoms += [b[:, 0], b[:, 1]]
        refs += [t * 4 + 0, t * 4 + 1]
    if angle_top is not None and angle_top.shape[0]:
        a = np.asarray(angle_top, dtype=np.int64).reshape(-1, 3)
        s = nb + 2 * np.arange(a.shape[0], dtype=np.int64)
        atoms += [a[:, 0], a[:, 1], a[:, 2]]
        refs += [s * 4 + 0, s * 4 + 2, (s + 1) * 4 + 0]
    if not atoms:
        return np.zeros(n_atoms + 1, dtype=np.int32), np.zeros(0, dtype=np.int32)
    atoms, refs = np.concatenate(atoms), np.concatenate(refs)
    if atoms.min() < 0 or atoms.max() >= n_atoms:
        raise IndexError("bonded topology names atom %d of %d" % (int(atoms.max() if atoms.max() >= n_atoms else atoms.min()), n_atoms))
    order = np.lexsort((refs, atoms))
    start = np.zeros(n_atoms + 1, dtype=np.int64)
    np.cumsum(np.bincount(atoms, minlength=n_atoms), out=start[1:])
    return start.astype(np.int32), refs[order].astype(np.int32)


class _BondedEnergy(torch.autograd.Function):
    """E_bond + E_angle of the owner's terms; backward = the forces / parameter derivatives of the same launch."""

    @staticmethod
    def forward(ctx, owner, xyz, *ptensors):
        need_f = xyz.requires_grad
        need_p = any(isinstance(p, torch.Tensor) and p.requires_grad for p in ptensors)
        e2, f, dp = owner._ctx.bonded_force(owner._terms(), xyz, owner._L, want_force=need_f, want_dparams=need_p)
        ctx.save_for_backward(f if f is not None else torch.empty(0), dp if dp is not None else torch.empty(0))
        ctx.flags = (need_f, need_p, owner._param_slots, [tuple(p.shape) if isinstance(p, torch.Tensor) else None for p in ptensors])
        return e2[owner._energy_slot]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        f, dp = ctx.saved_tensors
        need_f, need_p, slots, shapes = ctx.flags
        gx = (-g) * f if need_f else None
        gps = tuple((g * dp[s]).reshape(shp) if (need_p and shp is not None) else None for s, shp in zip(slots, shapes))
        return (None, gx) + gps


class _Bonded(torch.nn.Module):
    _energy_slot = 0
    _param_slots = (0, 1)

    def _init_common(self, system, top, width):
        self.device = system.device
        self._dev = _device_of(system)
        cell = torch.Tensor(np.asarray(system.get_cell()))
        self._L = cell_lengths(cell)
        self.cell = cell.diag().to(self._dev)            # reference: "transform into a diagonal"
        top = torch.as_tensor(top).to(torch.int64).reshape(-1, width)
        self._n_atoms = len(system)
        self._refs = term_refs(self._n_atoms, **{("bond_top" if width == 2 else "angle_top"): top.cpu().numpy()})
        self.top = top.to(self._dev)
        self._ctx = None
        self._dev_refs = None
        self.second_order = False

    def _reset_topology(self, xyz):
        """static topology (reference BondPotentials._reset_topology :433-434)"""
        return None

    def _scalars(self):
        raise NotImplementedError

    def _terms(self):
        if self._ctx is None:
            self._ctx = _lib.Context(self._dev)
        if self._dev_refs is None:
            self._dev_refs = (torch.from_numpy(self._refs[0]).to(self._dev), torch.from_numpy(self._refs[1]).to(self._dev),
                              self.top.contiguous())
        t = _lib.BondedTerms()
        rs, rf, top = self._dev_refs
        k, x0 = self._scalars()
        if self._energy_slot == 0:
            t.d_bond_top, t.n_bonds, t.k_bond, t.r0 = top.data_ptr(), top.shape[0], k, x0
        else:
            t.d_angle_top, t.n_angles, t.k_angle, t.theta0 = top.data_ptr(), top.shape[0], k, x0
        t.d_ref_start, t.d_refs = rs.data_ptr(), (rf.data_ptr() if rf.numel() else 0)
        return t

    def native_ready(self):
        return not self.second_order

    def native_force(self, xyz):
        if self._ctx is None:
            self._ctx = _lib.Context(self._dev)
        return self._ctx.bonded_force(self._terms(), xyz, self._L, want_force=True)[1]

    def forward(self, xyz):
        if _lib.on_device(xyz) and not (self.second_order and torch.is_grad_enabled()):
            if self._ctx is None:
                self._ctx = _lib.Context(self._dev)
            return _BondedEnergy.apply(self, xyz, *self._ptensors())
        if not _lib.on_device(xyz):
            _lib.require_cuda(xyz, "xyz")
        return self.forward_torch(xyz)


def _val(v):
    return float(v.detach().reshape(-1)[0]) if isinstance(v, torch.Tensor) else float(v)


class BondPotentials(_Bonded):
    """BondPotentials(system, top, k, ro)   (reference torchmd/interface.py:417-431).  E = 0.5 k sum (|v|^2 - ro)^2 -
    the reference compares the SQUARED bond length with `ro` (:447-449); kept."""
    _energy_slot = 0
    _param_slots = (0, 1)

    def __init__(self, system, top, k, ro):
        super().__init__()
        self._init_common(system, top, 2)
        self.k = k
        self.ro = ro

    def _scalars(self):
        return _val(self.k), _val(self.ro)

    def _ptensors(self):
        return (self.k, self.ro)

    def forward_torch(self, xyz):
        """the reference's op sequence (:444-451), differentiable to any order"""
        bond_vec = xyz[self.top[:, 0]] - xyz[self.top[:, 1]]
        offsets = get_offsets(bond_vec, self.cell, xyz.device)
        bond_vec = bond_vec + offsets * self.cell
        bond = bond_vec.pow(2).sum(-1)
        return 0.5 * self.k * (bond - self.ro).pow(2).sum(-1)


class AnglePotentials(_Bonded):
    """AnglePotentials(system, top, k, thetao)   (reference torchmd/interface.py:468-487)."""
    _energy_slot = 1
    _param_slots = (2, 3)

    def __init__(self, system, top, k, thetao):
        super().__init__()
        self._init_common(system, top, 3)
        self.k = k
        self.thetao = thetao

    def _scalars(self):
        return _val(self.k), _val(self.thetao)

    def _ptensors(self):
        return (self.k, self.thetao)

    def forward_torch(self, xyz):
        """the reference's op sequence (:497-510)"""
        v1 = xyz[self.top[:, 0]] - xyz[self.top[:, 1]]
        v2 = xyz[self.top[:, 2]] - xyz[self.top[:, 1]]
        v1 = v1 + get_offsets(v1, self.cell, xyz.device) * self.cell
        v2 = v2 + get_offsets(v2, self.cell, xyz.device) * self.cell
        angle_dot = (v1 * v2).sum(-1)
        norm = (v1.pow(2).sum(-1) * v2.pow(2).sum(-1)).sqrt()
        angle = torch.acos(angle_dot / norm)
        return 0.5 * self.k * (angle - self.thetao).pow(2).sum(-1)


class Electrostatics(torch.nn.Module):
    """Electrostatics(charges, cell, device=0, cutoff=2.5, index_tuple=None, ex_pairs=None)   (reference
    torchmd/interface.py:317-360): bare truncated Coulomb sum over the minimum-image list, rebuilt at every call.

    `charge_product="reference"` (default) reproduces the reference's arithmetic, in which the first charge is
    overwritten by the second (`q1 = charges[nbr[:,0]]; q1 = charges[nbr[:,1]]`, :355-356) and the sign is negative:
    U = -conv * sum q_j^2 / r.  `charge_product="physical"` evaluates +conv * sum q_i q_j / r."""

    def __init__(self, charges, cell, device=0, cutoff=2.5, index_tuple=None, ex_pairs=None, charge_product="reference"):
        super().__init__()
        from .system import HAVE_ASE
        if HAVE_ASE:
            from ase import units
        else:
            from ._ase_compat import units
        if charge_product not in ("reference", "physical"):
            raise ValueError("charge_product must be 'reference' or 'physical'")
        dev = torch.device("cuda:%d" % device if isinstance(device, int) else device)
        self._dev = dev
        self.charges = charges.to(dev)
        k_e = 8.987551787e9
        EV_TO_J = 1.60210e-19
        self.conversion = k_e * units.C ** -2 * (1 / EV_TO_J) * (units.m)
        self.cell = torch.Tensor(np.asarray(cell)).to(dev)
        self._L = cell_lengths(self.cell)
        self.device = device
        self.cutoff = cutoff
        self.index_tuple = index_tuple
        self.ex_pairs = ex_pairs
        self.charge_product = charge_product
        n = int(self.charges.shape[0])
        self._sel = _selection_flags(n, index_tuple, dev)
        self._exk = _exclusion_keys(n, ex_pairs, dev)
        self._ctx = None
        self.second_order = False

    def _reset_topology(self, xyz):
        return None

    def forward(self, x):
        if not _lib.on_device(x):
            _lib.require_cuda(x, "xyz")
        if self._ctx is None:
            self._ctx = _lib.Context(self._dev)
        nbr, off = self._ctx.nbr_list(x, self._L, self.cutoff, self._sel[0], self._sel[1], self._exk)
        cell = torch.diag(self.cell) if self.cell.dim() == 1 else self.cell
        dis_fn = compute_dis_torch if (self.second_order and torch.is_grad_enabled()) else compute_dis
        pair_dis = dis_fn(x, nbr, off, cell).squeeze(-1)
        qj = self.charges.reshape(-1)[nbr[:, 1]]
        if self.charge_product == "reference":
            U = -self.conversion * (qj * qj / pair_dis)
        else:
            U = self.conversion * (self.charges.reshape(-1)[nbr[:, 0]] * qj / pair_dis)
        return U.sum()
