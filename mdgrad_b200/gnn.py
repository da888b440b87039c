"""GNNPotentials - mirror of reference torchmd/interface.py:86-136.

`GNNPotentials(system, gnn, cutoff, ex_pairs=None)`: holds the model input dict
{nxyz, num_atoms, energy, nbr_list, offsets}; `_reset_topology(xyz)` rebuilds the list with the
native cell-list kernels; `forward(xyz)` -> gnn(inputs, xyz)['energy'].

Reference quirk kept for parity: `inputs['offsets']` are the raw integer image offsets and
SchNet subtracts them WITHOUT multiplying by the cell (`pbc_mode='reference'`, default, bug-compatible:
boundary-crossing edges get huge distances and drop out of the filter).  `pbc_mode='correct'` stores
offsets @ cell instead.
"""
import torch

from .interface import GeneralInteraction
from .topology import _exclusion_keys, cell_lengths, generate_nbr_list


def batch_to(batch, device):
    """reference nff/utils/cuda.py:6-10"""
    return {k: (v.to(device) if hasattr(v, "to") else v) for k, v in batch.items()}


class GNNPotentials(GeneralInteraction):
    def __init__(self, system, gnn, cutoff, ex_pairs=None, pbc_mode="reference"):
        super().__init__(system)
        if pbc_mode not in ("reference", "correct"):
            raise ValueError("pbc_mode must be 'reference' or 'correct'")
        self.gnn = gnn
        self.cutoff = cutoff
        self.pbc_mode = pbc_mode
        self.second_order = False
        self.inputs = batch_to(self.system.get_batch(), self.device)
        self.ex_pairs = ex_pairs
        self.to(self.device)
        self._L = cell_lengths(self.cell)
        self._ctx_key = "gnn%d" % id(self)
        self._reset_topology(torch.Tensor(system.get_positions()).to(system.device))

    def _reset_topology(self, xyz):
        nbr, offsets = generate_nbr_list(xyz, self.cutoff, self.cell, ex_pairs=self.ex_pairs, _ctx_key=self._ctx_key)
        self.inputs["nbr_list"] = nbr
        self.inputs["offsets"] = offsets if self.pbc_mode == "reference" else offsets * torch.tensor(self._L, device=offsets.device)
        self.inputs.pop("_native_graph", None)

    def forward(self, xyz):
        if hasattr(self.gnn, "second_order"):
            self.gnn.second_order = self.second_order
        return self.gnn(self.inputs, xyz)["energy"]
