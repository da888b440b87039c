"""Bulk thermodynamic observables - mirror of reference torchmd/thermo.py (BulkObservable :9-14, Temperature :57-66,
Pressure :16-54).

`Temperature` is the reference's algebra.  The reference's `Pressure.forward` cannot run (it uses undefined names `x` and
`pair`, thermo.py:36,41 - SURVEY 8f); the class here keeps the constructor signature `Pressure(system, model)` and the call
`pressure(q, v)` and computes what that code sets out to compute, the virial pressure of a pair model

    P = N T / V  -  1/(3 V) * sum_pairs r_ij u'(r_ij),        T = 2 KE / N_dof  (energy units)

over the native neighbor list of `model` (PairPotentials) rebuilt at q, with u'(r) from autograd through the pair energy
function, so it works for analytic and learned u(r) alike and stays differentiable w.r.t. the model parameters.
"""
import torch

from .system import check_system
from .topology import compute_dis


class BulkObservable(torch.nn.Module):
    def __init__(self, system):
        super().__init__()
        check_system(system)
        self.device = system.device
        self.system = system


class Temperature(BulkObservable):
    """T = KE / (N_dof / 2) in energy units (reference thermo.py:57-66)"""

    def __init__(self, system):
        super().__init__(system)
        self.mass = torch.Tensor(system.get_masses()).to(system.device)

    def forward(self, v):
        N_dof = self.mass.shape[0] * self.system.dim
        p = v * self.mass[:, None]
        ke = 0.5 * (p.pow(2) / self.mass[:, None]).sum()
        return ke / (N_dof * 0.5)


class Pressure(BulkObservable):
    """virial pressure of a PairPotentials model (see the module docstring)"""

    def __init__(self, system, model):
        super().__init__(system)
        self.model = model
        self.mass = torch.Tensor(system.get_masses()).to(system.device)

    def forward(self, q, v):
        nbr, _, offsets = self.model._reset_topology(q.detach())
        with torch.enable_grad():
            dis = compute_dis(q.detach(), nbr, offsets, self.model.cell.detach()).detach().requires_grad_(True)
            u = self.model.model(dis).sum()
            dudr, = torch.autograd.grad(u, dis, create_graph=torch.is_grad_enabled() and any(
                p.requires_grad for p in self.model.model.parameters()))
        N_dof = self.mass.shape[0] * self.system.dim
        p = v * self.mass[:, None]
        ke = 0.5 * (p.pow(2) / self.mass[:, None]).sum()
        temperature = ke / (N_dof * 0.5)
        volume = self.system.get_volume()
        p_ideal = self.system.get_number_of_atoms() * temperature / volume
        p_virial = (dis * dudr).sum() / (self.system.dim * volume)
        return p_ideal - p_virial
