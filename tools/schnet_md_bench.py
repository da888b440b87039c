"""SchNet MD timing (BASELINE configs[2] shape: 64-water box, SchNet A128/F128/G29/L2 rc 5.85 + O-O ExcludedVolume
prior, NoseHooverChain 5 chains, dt 0.5 fs) through the public API on cuda:0.  Prints one JSON line.
The epoch runs on the generic op-level route (native neighbor list, distance and cfconv-aggregation kernels,
cuBLAS dense layers, PyTorch solver loop); the reference's CPU number for the same shape is ~63 steps/s
(SURVEY 6, 8 vCPU)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(steps=100):
    from nff.nn.models.schnet import SchNet
    from torchmd.interface import GNNPotentials, PairPotentials, Stack
    from torchmd.potentials import ExcludedVolume
    from torchmd.md import NoseHooverChain, Simulations
    from torchmd.system import System
    from mdgrad_b200._ase_compat import Atoms, units
    g = np.load(os.path.join(ROOT, "tests", "golden", "schnet_water.npz"))
    system = System(Atoms(numbers=g["numbers"], positions=g["positions"], cell=g["cell"], pbc=True), device=0)
    np.random.seed(0)
    system.set_temperature(298.0 * units.kB)
    torch.manual_seed(0)
    params = {"n_atom_basis": 128, "n_filters": 128, "n_gaussians": 29, "n_convolutions": 2,
              "cutoff": 5.847718540914188, "trainable_gauss": False}
    model = SchNet(params).cuda()
    gnn = GNNPotentials(system, model, cutoff=params["cutoff"])
    oxy = [int(i) for i in np.nonzero(g["numbers"] == 8)[0]]
    prior = PairPotentials(system, ExcludedVolume(2.6, 0.015, 12).cuda(), cutoff=params["cutoff"], index_tuple=(oxy, oxy))
    integ = NoseHooverChain(Stack({"gnn": gnn, "prior": prior}), system, T=298.0 * units.kB, num_chains=5, Q=50.0, adjoint=True)
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    sim.simulate(steps=11, frequency=11, dt=0.5 * units.fs)            # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v, q, pv = sim.simulate(steps=steps + 1, frequency=steps + 1, dt=0.5 * units.fs)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    print(json.dumps({"workload": "64 H2O, SchNet A128 F128 G29 L2 rc 5.85 + O-O ExcludedVolume, NHC M=5, dt 0.5 fs",
                      "steps": steps, "steps_per_s": steps / el, "ns_per_day": steps / el * 0.5e-6 * 86400,
                      "edges": int(gnn.inputs["nbr_list"].shape[0]), "finite": bool(torch.isfinite(q).all())}))


if __name__ == "__main__":
    main()
