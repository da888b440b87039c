"""Interaction wrappers - mirror of reference torchmd/interface.py: GeneralInteraction :33-57,
PairPotentials :217-300, TPairPotentials :139-215, Stack :364-403 (GNNPotentials :86-136 lives in
gnn.py; BondPotentials :406-454, AnglePotentials :456-510, Electrostatics :303-361 in bonded.py).

Same constructor signatures, attributes (.nbr_list .offsets .cell .cutoff .model .index_tuple
.ex_pairs .nbr_list_device) and `forward(xyz) -> energy`, `_reset_topology(xyz)`.  The list is
built by the cell-list kernels and, for the analytic potential family, energy + forces + dE/dparam
come from one list-streaming kernel (no autograd tape through u(r)).
"""
import numpy as np
import torch
from torch.nn import ModuleDict

from . import _lib
from .topology import (_exclusion_keys, _selection_flags, cell_lengths, compute_dis, compute_dis_torch)


class GeneralInteraction(torch.nn.Module):
    """Base: holds the system, its (3,3) fp32 cell as a leaf with requires_grad (reference :47-57)."""

    def __init__(self, system):
        super().__init__()
        self.system = system
        self.cell = torch.Tensor(np.asarray(system.get_cell())).to(system.device)
        self.cell.requires_grad = True
        self.device = system.device


class SpecificInteraction(torch.nn.Module):
    """Base for interactions over a fixed topology (reference :59-83); nothing in the reference derives from it."""

    def __init__(self, system, topology):
        super().__init__()
        self.system = system
        self.cell = torch.Tensor(np.asarray(system.get_cell())).to(system.device)
        self.cell.requires_grad = True
        self.device = system.device
        self.topology = topology


class topology:
    """Placeholder kept for import compatibility: the reference's class (:513-545) is an unfinished stub whose constructor cannot
    be called (`def __init__():`)."""

    def __init__(self, top=None):
        self.top = top

    def _mutate(self, xyz, boundary):
        pass

    def _get_topology(self):
        return self.top

    def _stack(self):
        pass


class _PairEnergy(torch.autograd.Function):
    """E(xyz, params) over the context's stored list; backward = the forces / dE/dparam the same
    kernel launch produced (first order only - see PairPotentials.second_order)."""

    @staticmethod
    def forward(ctx, owner, xyz, *ptensors):
        kind, values, _ = owner.model.native_spec()
        need_f = xyz.requires_grad
        need_p = any(p.requires_grad for p in ptensors)
        e, f, dp = owner._ctx.pair_force(kind, values, xyz, want_force=need_f, want_dparams=need_p)
        ctx.save_for_backward(f if f is not None else torch.empty(0), dp if dp is not None else torch.empty(0))
        ctx.flags = (need_f, need_p, len(ptensors))
        return e

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        f, dp = ctx.saved_tensors
        need_f, need_p, npar = ctx.flags
        gx = (-g) * f if need_f else None
        gps = tuple((g * dp[k]).reshape(1) if need_p else None for k in range(npar))
        return (None, gx) + gps


class PairPotentials(GeneralInteraction):
    """PairPotentials(system, pair_model, cutoff=2.5, index_tuple=None, ex_pairs=None,
    nbr_list_device=None)   (reference torchmd/interface.py:228)."""

    def __init__(self, system, pair_model, cutoff=2.5, index_tuple=None, ex_pairs=None, nbr_list_device=None):
        super().__init__(system)
        self.nbr_list_device = system.device if nbr_list_device is None else nbr_list_device
        self.model = pair_model
        self.cutoff = cutoff
        self.index_tuple = index_tuple
        self.ex_pairs = ex_pairs
        self.second_order = False      # set by the adjoint solver when a double-backward graph is needed
        dev = torch.device(system.device if not isinstance(system.device, int) else "cuda:%d" % system.device)
        self._dev = dev
        self._ctx = _lib.Context(dev)           # owns this potential's list
        self._L = cell_lengths(self.cell)
        n = len(system)
        self._sel = _selection_flags(n, index_tuple, dev)
        self._exk = _exclusion_keys(n, ex_pairs, dev)
        self._nbr_dev = None
        self._reset_topology(torch.Tensor(system.get_positions()).to(dev))

    # the reference keeps nbr_list on the CPU (interface.py:259); materialise that copy lazily
    @property
    def nbr_list(self):
        self._refresh_if_stale()
        if self._nbr_cpu is None:
            self._nbr_cpu = self._nbr_dev.to("cpu")
        return self._nbr_cpu

    @property
    def offsets(self):
        self._refresh_if_stale()
        return self._offsets

    @offsets.setter
    def offsets(self, value):
        self._offsets = value

    # After an epoch on the fused engine (mdg_md_run keeps its own skin list) the Python-visible topology must be the list of
    # the reference's last evaluation, i.e. the list at the final positions.  Building + exporting it costs a list build per
    # epoch that nothing reads in a plain MD loop, so it is deferred until somebody looks (nbr_list / offsets / forward).
    def _mark_stale(self, xyz):
        self._stale_xyz = xyz.detach().clone()

    def _refresh_if_stale(self):
        xyz = getattr(self, "_stale_xyz", None)
        if xyz is not None:
            self._stale_xyz = None
            self._reset_topology(xyz)

    def native_kind(self):
        return self.model.native_spec() if hasattr(self.model, "native_spec") else None

    def _reset_topology(self, xyz):
        """Rebuild the list at xyz; returns (nbr_list, pair_dis, offsets) (reference :263-282)."""
        nbr, off, dis = self._ctx.nbr_list(xyz, self._L, self.cutoff, self._sel[0], self._sel[1], self._exk, get_dis=True)
        self._stale_xyz = None
        self._nbr_dev, self._nbr_cpu = nbr, None
        self._offsets = off
        return nbr, dis, off

    # -- native force route (no autograd tape) ---------------------------------------------------------------------
    def native_ready(self):
        return type(self) is PairPotentials and self.native_kind() is not None and not self.second_order

    def native_force(self, xyz):
        """-dE/dxyz over the stored list straight from the force kernel (mdg_pair_force)."""
        self._refresh_if_stale()
        kind, values, _ = self.native_kind()
        return self._ctx.pair_force(kind, values, xyz, want_force=True)[1]

    def forward(self, xyz):
        """sum_pairs u(|x_i - x_j - offsets@cell|) over the STORED list (reference :284-300)."""
        self._refresh_if_stale()
        spec = self.native_kind()
        if spec is not None and not (self.second_order and torch.is_grad_enabled()):
            return _PairEnergy.apply(self, xyz, *spec[2])
        dis_fn = compute_dis_torch if self.second_order else compute_dis
        pair_dis = dis_fn(xyz, self._nbr_dev, self.offsets, self.cell)
        return self.model(pair_dis).sum()


class TPairPotentials(PairPotentials):
    """Temperature-conditioned pair model u(r, kB*T) (reference torchmd/interface.py:139-215)."""

    def __init__(self, system, pair_model, T, cutoff=2.5, index_tuple=None, ex_pairs=None, nbr_list_device=None):
        super().__init__(system, pair_model, cutoff, index_tuple, ex_pairs, nbr_list_device)
        self.T = T

    def native_kind(self):
        return None

    def forward(self, xyz):
        from .system import HAVE_ASE
        if HAVE_ASE:
            from ase import units
        else:
            from ._ase_compat import units
        dis_fn = compute_dis_torch if self.second_order else compute_dis
        pair_dis = dis_fn(xyz, self._nbr_dev, self.offsets, self.cell)
        return self.model(pair_dis, units.kB * self.T).sum()


class Stack(torch.nn.Module):
    """Sum of member energies; `_reset_topology` fans out (reference torchmd/interface.py:364-403)."""

    def __init__(self, model_dict, mode="sum"):
        super().__init__()
        self.models = ModuleDict(model_dict)

    def _reset_topology(self, x):
        for key in self.models.keys():
            self.models[key]._reset_topology(x)

    def native_ready(self):
        return all(hasattr(m, "native_ready") and m.native_ready() for m in self.models.values())

    def native_force(self, x):
        f = None
        for m in self.models.values():
            fm = m.native_force(x)
            f = fm if f is None else f + fm
        return f

    def forward(self, x):
        result = None
        for key in self.models.keys():
            term = self.models[key](x).sum().reshape(-1)
            result = term if result is None else result + term
        return result


def __getattr__(name):
    if name == "GNNPotentials":          # lives in gnn.py (imports this module)
        from .gnn import GNNPotentials
        return GNNPotentials
    if name in ("BondPotentials", "AnglePotentials", "Electrostatics"):      # bonded.py
        from . import bonded
        return getattr(bonded, name)
    raise AttributeError(name)
