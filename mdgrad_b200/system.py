"""State container - mirror of reference torchmd/system.py:16-70 (`System(ase.Atoms)`).

Derives from the real `ase.Atoms` when ASE is importable, otherwise from the ASE-3.20
restatement in `_ase_compat` (ASE is absent from this image).  Positions / momenta / cell stay
numpy fp64 on the host exactly like the reference; device tensors are created by the
integrators (`get_inital_states`).
"""
import numpy as np
import torch

try:  # pragma: no cover - depends on the environment
    from ase import Atoms as _Atoms
    if getattr(__import__("ase"), "_mdgrad_standin", False):
        raise ImportError
    HAVE_ASE = True
except ImportError:
    from ._ase_compat import Atoms as _Atoms
    HAVE_ASE = False


def check_system(obj):
    """reference torchmd/system.py:11-14"""
    if obj.__class__ is not System:
        raise TypeError("input should be a torchmd.system.System")


class System(_Atoms):
    """System(atoms_or_symbols, ..., device=, dim=3, props={})  (reference system.py:27-37)."""

    def __init__(self, *args, device, dim=3, props={}, **kwargs):
        super().__init__(*args, **kwargs)
        self.props = props
        self.device = device
        self.dim = dim

    def get_nxyz(self):
        """(N,4) [Z, x, y, z]  (reference system.py:39-52)"""
        return np.concatenate([self.get_atomic_numbers().reshape(-1, 1),
                               self.get_positions().reshape(-1, 3)], axis=1)

    def get_cell_len(self):
        return np.diag(np.asarray(self.get_cell()))

    def get_batch(self):
        """reference system.py:56-62"""
        return {"nxyz": torch.Tensor(self.get_nxyz()),
                "num_atoms": torch.LongTensor([len(self)]),
                "energy": 0.0}

    def get_number_of_atoms(self):
        return len(self)

    def set_temperature(self, T):
        """Maxwell-Boltzmann momenta at T (energy units); zero the last column if dim < 3
        (reference system.py:64-70)."""
        if HAVE_ASE:
            from ase.md.velocitydistribution import MaxwellBoltzmannDistribution
        else:
            from ._ase_compat import MaxwellBoltzmannDistribution
        MaxwellBoltzmannDistribution(self, T)
        if self.dim < 3:
            vel = self.get_velocities()
            vel[:, -1] = 0.0
            self.set_velocities(vel)
