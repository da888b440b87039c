"""SchNet MD timing through the public API on cuda:0.  Prints one JSON line.

  --config water   BASELINE configs[2] shape: 64 H2O (192 atoms), SchNet A128/F128/G29/L2 rc 5.85 + O-O ExcludedVolume
                   prior, NoseHooverChain 5 chains, dt 0.5 fs
  --config si      BASELINE configs[4] shape: 4096-atom Si (diamond 8^3, a = 5.45933 A, jitter 0.05 A), SchNet
                   A512/F256/G33/L3 rc 4.9 + ExcludedVolume(0.015, 1.9, 12) prior, NoseHooverChain T = 100 kB, dt 1 fs
  --route engine   (default) the epoch runs on the device engine: mdg_md_run_gnn (per-step exact lists, native SchNet
                   energy+force program, pair prior, fused integrator kernels - no Python inside the loop)
  --route oplevel  the PyTorch solver loop with native forces (integrator.disable_gnn_engine)
  MDG_SCHNET_TC=1  dense layers on tcgen05 (experimental, tools/tc_check.py first)

Weights are random (torch.manual_seed(0), xavier) - there is no network for checkpoints; the reference's CPU number for
the water shape is ~63 steps/s (SURVEY 6, 8 vCPU).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(config):
    from nff.nn.models.schnet import SchNet
    from torchmd.interface import GNNPotentials, PairPotentials, Stack
    from torchmd.potentials import ExcludedVolume
    from torchmd.md import NoseHooverChain, Simulations
    from torchmd.system import System
    from mdgrad_b200._ase_compat import Atoms, units
    dev = int(os.environ.get("LOCAL_RANK", "0"))           # (replica runs under torchrun: one model per GPU)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    np.random.seed(0)
    if config == "water":
        g = np.load(os.path.join(ROOT, "tests", "golden", "schnet_water.npz"))
        system = System(Atoms(numbers=g["numbers"], positions=g["positions"], cell=g["cell"], pbc=True), device=dev)
        T, dt = 298.0 * units.kB, 0.5 * units.fs
        params = {"n_atom_basis": 128, "n_filters": 128, "n_gaussians": 29, "n_convolutions": 2,
                  "cutoff": 5.847718540914188, "trainable_gauss": False}
        oxy = [int(i) for i in np.nonzero(g["numbers"] == 8)[0]]
        prior_args = dict(pot=ExcludedVolume(2.6, 0.015, 12), cutoff=params["cutoff"], index_tuple=(oxy, oxy))
        label = "64 H2O, SchNet A128 F128 G29 L2 rc 5.85 + O-O ExcludedVolume, NHC M=5, dt 0.5 fs"
    else:
        a, nc = 5.45933, 8
        basis = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
        basis = np.concatenate([basis, basis + 0.25])
        cells = np.array([[i, j, k] for i in range(nc) for j in range(nc) for k in range(nc)])
        pos = ((cells[:, None, :] + basis[None, :, :]).reshape(-1, 3)) * a
        pos = pos + np.random.default_rng(3).normal(0, 0.05, pos.shape)
        system = System(Atoms(numbers=[14] * len(pos), positions=pos, cell=[a * nc] * 3, pbc=True), device=dev)
        T, dt = 100.0 * units.kB, 1.0 * units.fs
        params = {"n_atom_basis": 512, "n_filters": 256, "n_gaussians": 33, "n_convolutions": 3, "cutoff": 4.9,
                  "trainable_gauss": False}
        prior_args = dict(pot=ExcludedVolume(1.9, 0.015, 12), cutoff=4.9, index_tuple=None)
        label = "4096 Si (diamond 8^3), SchNet A512 F256 G33 L3 rc 4.9 + ExcludedVolume, NHC M=5, dt 1 fs"
    system.set_temperature(T)
    model = SchNet(params).cuda()
    gnn = GNNPotentials(system, model, cutoff=params["cutoff"])
    prior = PairPotentials(system, prior_args["pot"].cuda(), cutoff=prior_args["cutoff"], index_tuple=prior_args["index_tuple"])
    integ = NoseHooverChain(Stack({"gnn": gnn, "prior": prior}), system, T=T, num_chains=5, Q=50.0, adjoint=True)
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    return sim, integ, gnn, dt, label


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="water", choices=["water", "si"])
    ap.add_argument("--route", default="engine", choices=["engine", "oplevel"])
    ap.add_argument("--steps", type=int, default=0)
    args = ap.parse_args()
    from mdgrad_b200._ase_compat import units
    sim, integ, gnn, dt, label = build(args.config)
    integ.disable_gnn_engine = args.route == "oplevel"
    steps = args.steps or (200 if args.config == "water" else 20)
    sim.simulate(steps=6, frequency=6, dt=dt)                        # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v, q, pv = sim.simulate(steps=steps + 1, frequency=steps + 1, dt=dt)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    st = integ.last_engine_stats or {}
    print(json.dumps({"workload": label, "route": args.route, "tc": os.environ.get("MDG_SCHNET_TC") == "1", "steps": steps,
                      "steps_per_s": steps / el, "ns_per_day": steps / el * (dt / units.fs) * 1e-6 * 86400,
                      "edges": int(gnn.inputs["nbr_list"].shape[0]), "launches_per_step": (st.get("launches", 0) or 0) / max(1, steps),
                      "engine_mode": {0: "synchronous", 1: "asynchronous", 2: "asynchronous + graph replay"}.get(st.get("maxrow_or_K"), None)
                      if args.route == "engine" else None,
                      "finite": bool(torch.isfinite(q).all())}))


if __name__ == "__main__":
    main()
