"""Observables - mirror of reference torchmd/observable.py: generate_vol_bins :10-21, Observable
:24-31, rdf :33-76, vacf :153-163.

`rdf(system, nbins, r_range, index_tuple=None, width=None)(xyz) -> (count, bins, g)`; the pair
search + Gaussian-smeared histogram run in one cell-list traversal kernel (mdg_rdf_accumulate)
instead of a second O(N^2) neighbor list and a (P, nbins) tensor.  The kernel path is not
differentiable; when `xyz.requires_grad` the smearing is evaluated with torch ops over the native
neighbor list so that losses still reach the trajectory.
"""
import numpy as np
import torch

from . import _lib
from .system import check_system
from .potentials import GaussianSmearing
from .topology import (_selection_flags, cell_lengths, compute_dis_torch, context_for, generate_angle_list,
                       generate_nbr_list, get_offsets)


def generate_vol_bins(start, end, nbins, dim):
    """bins = linspace(start, end, nbins+1); shell volumes; V = 4/3 pi end^3 (reference :10-21)."""
    bins = torch.linspace(start, end, nbins + 1)
    if dim == 3:
        Vbins = 4 * np.pi / 3 * (bins[1:] ** 3 - bins[:-1] ** 3)
        V = (4 / 3) * np.pi * (end) ** 3
    elif dim == 2:
        Vbins = np.pi * (bins[1:] ** 2 - bins[:-1] ** 2)
        V = np.pi * (end) ** 2
    return V, torch.Tensor(Vbins), bins


class Observable(torch.nn.Module):
    def __init__(self, system):
        super().__init__()
        check_system(system)
        self.device = system.device
        self.volume = system.get_volume()
        self.cell = torch.Tensor(np.asarray(system.get_cell())).diag().to(self.device)
        self.natoms = len(system)


class rdf(Observable):
    def __init__(self, system, nbins, r_range, index_tuple=None, width=None):
        super().__init__(system)
        start, end = r_range[0], r_range[1]
        V, vol_bins, bins = generate_vol_bins(start, end, nbins, dim=system.dim)
        self.V = V
        self.vol_bins = vol_bins.to(self.device)
        self.r_axis = np.linspace(start, end, nbins)
        self.bins = bins
        self.start, self.end, self.width = float(start), float(bins[-1]), width
        self.nbins = nbins
        self.cutoff_boundary = end + 5e-1
        self.index_tuple = index_tuple
        # Gaussian centres / widths exactly as nff GaussianSmearing builds them (layers.py:55-59)
        self.mu = torch.linspace(start, bins[-1], nbins).to(self.device)
        w = (self.mu[1] - self.mu[0]) if width is None else width
        self.w = float(w)
        self._L = cell_lengths(self.cell)
        self._sel = None

    def forward(self, xyz):
        _lib.require_cuda(xyz, "xyz")
        frames = xyz.reshape(-1, xyz.shape[-2], 3)
        n = frames.shape[1]
        if xyz.requires_grad and torch.is_grad_enabled():
            count = 0
            for f in range(frames.shape[0]):
                nbr, off = generate_nbr_list(frames[f], self.cutoff_boundary, self.cell,
                                             index_tuple=self.index_tuple, _ctx_key="rdf")
                d = compute_dis_torch(frames[f], nbr, off, self.cell)
                count = count + torch.exp(-0.5 / self.w ** 2 * (d - self.mu) ** 2).sum(0)
        else:
            if self._sel is None:
                self._sel = _selection_flags(n, self.index_tuple, xyz.device)
            ctx = context_for(xyz.device, "rdf")
            count = torch.zeros(self.nbins, dtype=torch.float32, device=xyz.device)
            for f in range(frames.shape[0]):
                ctx.rdf_accumulate(frames[f], self._L, self.start, self.end, self.nbins,
                                   self.width, count, self._sel[0], self._sel[1])
        count = count / count.sum()
        g = count / (self.vol_bins / self.V)
        return count, self.bins, g


def compute_angle(xyz, angle_list, cell, N):
    """cos of the angle a-b-c at the centre b with minimum-image bond vectors (reference observable.py:166-179)"""
    device = xyz.device
    xyz = xyz.reshape(-1, N, 3)
    bond_vec1 = xyz[angle_list[:, 0], angle_list[:, 1]] - xyz[angle_list[:, 0], angle_list[:, 2]]
    bond_vec2 = xyz[angle_list[:, 0], angle_list[:, 3]] - xyz[angle_list[:, 0], angle_list[:, 2]]
    bond_vec1 = bond_vec1 + get_offsets(bond_vec1, cell, device) * cell
    bond_vec2 = bond_vec2 + get_offsets(bond_vec2, cell, device) * cell
    angle_dot = (bond_vec1 * bond_vec2).sum(-1)
    norm = (bond_vec1.pow(2).sum(-1) * bond_vec2.pow(2).sum(-1)).sqrt()
    return angle_dot / norm


class Angles(Observable):
    """cos(angle) of every bonded triple within `cutoff` (reference observable.py:78-110); the neighbor list comes
    from the native cell-list kernels, the triples are enumerated on the device (topology.generate_angle_list)."""

    def __init__(self, system, nbins, angle_range, cutoff=3.0, index_tuple=None, width=None):
        super().__init__(system)
        start, end = angle_range[0], angle_range[1]
        self.bins = torch.linspace(start, end, nbins + 1).to(self.device)
        self.smear = GaussianSmearing(start=start, stop=self.bins[-1], n_gaussians=nbins, width=width,
                                      trainable=False).to(self.device)
        self.width = (self.smear.width[0]).item()
        self.cutoff = cutoff
        self.index_tuple = index_tuple

    def _cos_angles(self, xyz):
        xyz = xyz.reshape(-1, self.natoms, 3)
        nbr_list, _ = generate_nbr_list(xyz, self.cutoff, self.cell, index_tuple=self.index_tuple, get_dis=False,
                                        _ctx_key="angles")
        angle_list = generate_angle_list(nbr_list)
        return compute_angle(xyz, angle_list, self.cell, N=self.natoms)

    def forward(self, xyz):
        return self._cos_angles(xyz)


class angle_distribution(Angles):
    """Gaussian-smeared, normalised histogram of the bond angles (reference observable.py:112-151):
    returns (bins, count, angles)."""

    def forward(self, xyz):
        angles = self._cos_angles(xyz).acos()
        count = self.smear(angles.reshape(-1).squeeze()[..., None]).sum(0)
        count = count / count.sum()
        return self.bins, count, angles


class vacf(Observable):
    """velocity autocorrelation (reference observable.py:153-163)"""

    def __init__(self, system, t_range):
        super().__init__(system)
        self.t_window = [i for i in range(1, t_range, 1)]

    def forward(self, vel):
        if (_lib.on_device(vel) and vel.dim() == 3 and len(self.t_window) + 1 <= vel.shape[0]
                and not (torch.is_grad_enabled() and vel.requires_grad)):
            # nobody differentiates it: one lag-product reduction kernel (fp64 accumulation) instead of t_range sliced products
            return context_for(vel.device, "vacf").vacf(vel, len(self.t_window) + 1)
        vacf = [(vel * vel).mean()[None]]
        vacf += [(vel[t:] * vel[:-t]).mean()[None] for t in self.t_window]
        return torch.stack(vacf).reshape(-1)          # un-normalised, exactly as the reference


def compute_dihe(xyz, dihes):
    """cos(phi) of the listed dihedrals for every frame (reference observable.py:181-197), without the reference's
    (frames, N, N, 3) difference tensor: only the four bond vectors of each dihedral are formed.  No periodic images (as the
    reference)."""
    assert len(xyz.shape) == 3
    dihes = torch.as_tensor(dihes).to(xyz.device)
    p0, p1, p2, p3 = (xyz[:, dihes[:, k]] for k in range(4))
    vec1, vec2 = p1 - p0, p1 - p2          # D[:, d1, d0] = x[d1] - x[d0],  D[:, d1, d2] = x[d1] - x[d2]
    vec3, vec4 = p2 - p1, p2 - p3
    cross1 = torch.cross(vec1, vec2, dim=-1)
    cross2 = torch.cross(vec3, vec4, dim=-1)
    norm = (cross1.pow(2).sum(-1) * cross2.pow(2).sum(-1)).sqrt()
    return 1.0 * ((cross1 * cross2).sum(-1) / norm)
