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

from . import _lib
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

    def __del__(self):
        # the native context registered under this instance's key dies with the instance (no leak over hyper-parameter
        # sweeps, and a recycled id() can never inherit a stale context)
        try:
            from . import topology
            for k in [k for k in topology._CTX if k[1] == self._ctx_key]:
                topology._CTX.pop(k, None)
        except Exception:
            pass

    def _reset_topology(self, xyz):
        nbr, offsets = generate_nbr_list(xyz, self.cutoff, self.cell, ex_pairs=self.ex_pairs, _ctx_key=self._ctx_key)
        self.inputs["nbr_list"] = nbr
        self.inputs["offsets"] = offsets if self.pbc_mode == "reference" else offsets * torch.tensor(self._L, device=offsets.device)
        self.inputs.pop("_native_graph", None)

    def _ex_keys(self, device):
        """GNNPotentials(ex_pairs=) as the sorted unique int64 keys of mdg_nbr_build (cached)"""
        if getattr(self, "_exk_cache", None) is None:
            self._exk_cache = (_exclusion_keys(self.inputs["nxyz"].shape[0], self.ex_pairs, device),)
        return self._exk_cache[0]

    # -- native force route (no autograd tape): used by the solvers whenever no graph is being recorded -----------
    def native_ready(self):
        """True when `native_force` covers this model: our SchNet mirror with the default energy readout."""
        from .nffm.schnet import SchNet, shifted_softplus
        g = self.gnn
        if type(g) is not SchNet or self.second_order or g.atomwisereadout.post_readout is not None:
            return False
        ro = g.atomwisereadout.readout
        if list(ro.keys()) != ["energy"] or len(ro["energy"]) != 3:
            return False
        l0, act, l2 = ro["energy"][0], ro["energy"][1], ro["energy"][2]
        if not (isinstance(l0, torch.nn.Linear) and isinstance(act, shifted_softplus) and isinstance(l2, torch.nn.Linear)):
            return False
        if l2.out_features != 1 or l0.bias is None or l2.bias is None:
            return False
        G = g.convolutions[0].moduledict["message_edge_filter"][1].in_features
        return len(g.convolutions) <= _lib.SCHNET_MAX_LAYERS and G <= 64

    def _native_model(self):
        sd = self.gnn.state_dict()
        key = tuple((k, v.data_ptr(), v._version) for k, v in sd.items())
        if getattr(self, "_nm_key", None) != key:
            self._nm = _lib.schnet_model_struct(sd, self.inputs["nxyz"].device)
            self._nm_key = key
        return self._nm

    def native_energy_force(self, xyz, want_force=True):
        """(energy 0-d, forces (N,3)) of the stored list at xyz from ONE native program (mdg_schnet_energy_force):
        SchNet.forward (nff/nn/models/schnet.py:113-171) + the autograd force of md.py:227-228."""
        if getattr(self, "_sn_ctx", None) is None:
            self._sn_ctx = _lib.Context(xyz.device)
            self._z = self.inputs["nxyz"][:, 0].to(torch.int64).contiguous()
        return self._sn_ctx.schnet_energy_force(self._native_model(), self._z, xyz, self.inputs["nbr_list"],
                                                self.inputs["offsets"], (1.0, 1.0, 1.0), want_force=want_force)

    def native_force(self, xyz):
        return self.native_energy_force(xyz)[1]

    def forward(self, xyz):
        if hasattr(self.gnn, "second_order"):
            self.gnn.second_order = self.second_order
        return self.gnn(self.inputs, xyz)["energy"]
