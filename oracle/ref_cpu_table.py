"""CPU baseline of the UNMODIFIED reference, run in the authoring container (the reference cannot
travel to the GPU box): C1 as specified, LJ scaling points N = 500 / 2048 / 4000 / 8788 with an O(N^2) fit and the labelled
extrapolation to C2 / C4, C3 (192-atom water SchNet MD), C5 at 512 atoms (forward + adjoint) and 4096 atoms (one evaluation).

    python oracle/ref_cpu_table.py            -> profiles/r02_reference_cpu_table.json / .md

TEST / MEASUREMENT INFRASTRUCTURE ONLY (imports /root/reference through oracle/ref_import.py).
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from mdgrad_b200._ase_compat import Atoms, Diamond, FaceCenteredCubic, units  # noqa: E402


def timed_epochs(sim, steps, dt, reps=3, warm=True):
    if warm:
        sim.simulate(steps=steps, frequency=steps, dt=dt)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = sim.simulate(steps=steps, frequency=steps, dt=dt)
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), out


def lj_sim(ref, ncell, a, seed=0):
    atoms = FaceCenteredCubic(symbol="H", size=(ncell,) * 3, latticeconstant=a, pbc=True)
    system = ref.system.System(atoms, device="cpu")
    np.random.seed(seed)
    system.set_temperature(1.0)
    pair = ref.interface.PairPotentials(system, ref.potentials.LennardJones(1.0, 1.0), cutoff=2.5)
    integ = ref.md.NoseHooverChain(pair, system, T=1.0, num_chains=5, Q=50.0 * max(1, len(atoms) / 256), adjoint=True, topology_update_freq=1)
    return ref.md.Simulations(system, integ, wrap=True, method="NH_verlet"), system


def main():
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    rows = {"host": {"cores": threads, "torch": torch.__version__, "where": "authoring container (no GPU)",
                     "protocol": "unmodified reference imported from /root/reference, fp32, adjoint=True, topology_update_freq=1, "
                                 "NH_verlet; 1 warm-up epoch, median of 3 timed epochs, steps = frequency - 1"}}
    with ref_import.active() as ref:
        # C1
        sim, _ = lj_sim(ref, 3, 1.679)
        el, _ = timed_epochs(sim, 50, 0.01)
        rows["c1_108"] = {"atoms": 108, "steps": 49, "seconds": el, "steps_per_s": 49 / el}
        print("C1", rows["c1_108"], flush=True)
        a = (4.0 / 0.845) ** (1.0 / 3.0)
        pts = []
        for ncell, nsteps, reps in ((5, 20, 3), (8, 6, 3), (10, 3, 2), (13, 2, 1)):
            sim, system = lj_sim(ref, ncell, a)
            n = len(system)
            el, _ = timed_epochs(sim, nsteps + 1, 0.005, reps=reps, warm=(ncell < 13))
            rows["lj_%d" % n] = {"atoms": n, "steps": nsteps, "seconds": el, "steps_per_s": nsteps / el, "s_per_step": el / nsteps}
            pts.append((n, el / nsteps))
            print("LJ", rows["lj_%d" % n], flush=True)
        N = np.array([p[0] for p in pts], dtype=float)
        T = np.array([p[1] for p in pts])
        A = np.stack([N ** 2, N, np.ones_like(N)], 1)
        coef, *_ = np.linalg.lstsq(A, T, rcond=None)
        fit = lambda n: float(coef[0] * n * n + coef[1] * n + coef[2])  # noqa: E731
        rows["fit"] = {"model": "s_per_step = a N^2 + b N + c", "a": float(coef[0]), "b": float(coef[1]), "c": float(coef[2]),
                       "c2_256000_extrapolated_s_per_step": fit(256000.0), "c2_256000_extrapolated_steps_per_s": 1.0 / fit(256000.0),
                       "c4_1048576_extrapolated_s_per_step": fit(1048576.0),
                       "note": "EXTRAPOLATION: the reference cannot allocate these boxes (~70 N^2 bytes of (N, N, 3) temporaries)"}
        print("fit", rows["fit"], flush=True)
        # C3: 64 waters, SchNet A128/F128/G29/L2 + O-O ExcludedVolume
        g = np.load(os.path.join(ROOT, "tests", "golden", "schnet_water.npz"))
        atoms = Atoms(numbers=g["numbers"], positions=g["positions"], cell=g["cell"], pbc=True)
        system = ref.system.System(atoms, device="cpu")
        np.random.seed(0)
        system.set_temperature(298.0 * units.kB)
        torch.manual_seed(0)
        params = {"n_atom_basis": 128, "n_filters": 128, "n_gaussians": 29, "n_convolutions": 2, "cutoff": 5.847718540914188,
                  "trainable_gauss": False}
        model = ref.schnet.SchNet(params)
        gnn = ref.interface.GNNPotentials(system, model, cutoff=params["cutoff"])
        oxy = [int(i) for i in np.nonzero(g["numbers"] == 8)[0]]
        prior = ref.interface.PairPotentials(system, ref.potentials.ExcludedVolume(2.6, 0.015, 12), cutoff=params["cutoff"], index_tuple=(oxy, oxy))
        integ = ref.md.NoseHooverChain(ref.interface.Stack({"gnn": gnn, "prior": prior}), system, T=298.0 * units.kB, num_chains=5, Q=50.0, adjoint=True)
        sim = ref.md.Simulations(system, integ, wrap=True, method="NH_verlet")
        el, _ = timed_epochs(sim, 21, 0.5 * units.fs)
        rows["c3_water192"] = {"atoms": 192, "steps": 20, "seconds": el, "steps_per_s": 20 / el}
        print("C3", rows["c3_water192"], flush=True)
        # C5 at 512 atoms: forward MD + adjoint through 5 steps; 4096 atoms: one energy+force evaluation (fixture timing)
        atoms = Diamond("Si", (4, 4, 4), 5.45933)
        atoms.set_positions(atoms.get_positions() + np.random.default_rng(11).normal(0, 0.05, (len(atoms), 3)))
        system = ref.system.System(atoms, device="cpu")
        np.random.seed(0)
        system.set_temperature(100.0 * units.kB)
        torch.manual_seed(1)
        params = {"n_atom_basis": 512, "n_filters": 256, "n_gaussians": 33, "n_convolutions": 3, "cutoff": 4.9, "trainable_gauss": False}
        model = ref.schnet.SchNet(params)
        gnn = ref.interface.GNNPotentials(system, model, cutoff=4.9)
        prior = ref.interface.PairPotentials(system, ref.potentials.ExcludedVolume(1.9, 0.015, 12), cutoff=4.9)
        integ = ref.md.NoseHooverChain(ref.interface.Stack({"gnn": gnn, "prior": prior}), system, T=100.0 * units.kB, num_chains=5, Q=50.0, adjoint=True)
        sim = ref.md.Simulations(system, integ, wrap=True, method="NH_verlet")
        el, _ = timed_epochs(sim, 6, 0.1 * units.fs, reps=2)
        rows["c5_si512_forward"] = {"atoms": 512, "steps": 5, "seconds": el, "steps_per_s": 5 / el}
        t0 = time.perf_counter()
        v, q, pv = sim.simulate(steps=6, frequency=6, dt=0.1 * units.fs)
        loss = (q[-1] ** 2).sum() + (v[-1] ** 2).sum()
        t1 = time.perf_counter()
        loss.backward()
        t2 = time.perf_counter()
        rows["c5_si512_adjoint_5_steps"] = {"forward_s": t1 - t0, "backward_s": t2 - t1}
        print("C5-512", rows["c5_si512_forward"], rows["c5_si512_adjoint_5_steps"], flush=True)
        g4 = np.load(os.path.join(ROOT, "tests", "golden", "schnet_si4096.npz"))
        rows["c5_si4096_one_evaluation"] = {"atoms": 4096, "edges": int(g4["n_edges"]), "energy_force_seconds": float(g4["ref_eval_seconds"]),
                                            "threads": int(g4["ref_threads"]),
                                            "note": "GNNPotentials energy + autograd force on a GIVEN list (oracle/make_golden.py "
                                                    "--schnet-configured); the reference's own list rebuild at this size adds ~3 s"}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "profiles", "r02_reference_cpu_table.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", "r02_reference_cpu_table.md"), "w") as f:
        f.write("# Unmodified reference on CPU (BASELINE.md section 3) - authoring container, %d cores, torch %s\n\n" % (threads, torch.__version__))
        f.write(rows["host"]["protocol"] + "\n\n| case | atoms | steps timed | seconds | steps/s |\n|---|---|---|---|---|\n")
        for k, r in rows.items():
            if isinstance(r, dict) and "steps_per_s" in r:
                f.write("| %s | %d | %d | %.3f | %.4g |\n" % (k, r["atoms"], r["steps"], r["seconds"], r["steps_per_s"]))
        f.write("\nO(N^2) fit: s/step = %.3e N^2 + %.3e N + %.3e; **extrapolated** C2 (256 000 atoms): %.4g s/step = %.3g steps/s; C4 (1 048 576): %.4g s/step "
                "(the reference cannot allocate either box).\n" % (rows["fit"]["a"], rows["fit"]["b"], rows["fit"]["c"],
                                                                 rows["fit"]["c2_256000_extrapolated_s_per_step"],
                                                                 rows["fit"]["c2_256000_extrapolated_steps_per_s"],
                                                                 rows["fit"]["c4_1048576_extrapolated_s_per_step"]))
        f.write("\nC5-512 adjoint through 5 steps: forward %.2f s, backward %.2f s.  C5-4096: one energy + force evaluation on a given list %.2f s (%d threads).\n"
                % (rows["c5_si512_adjoint_5_steps"]["forward_s"], rows["c5_si512_adjoint_5_steps"]["backward_s"],
                   rows["c5_si4096_one_evaluation"]["energy_force_seconds"], rows["c5_si4096_one_evaluation"]["threads"]))
    print("wrote profiles/r02_reference_cpu_table.{json,md}")


if __name__ == "__main__":
    main()
