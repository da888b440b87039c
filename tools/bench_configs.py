"""bench.py --config c1 | c3 | c5: the other BASELINE.json configurations, through the public API on cuda:0.

  c1  108-atom FCC LJ (3x3x3, sigma = eps = 1, rc 2.5), NoseHooverChain, 50 Verlet steps per epoch - the reference demo
  c3  64-molecule water box, SchNet A128/F128/G29/L2 rc 5.85 + O-O ExcludedVolume prior, NHC, RDF(O-O) at the end of the epoch
  c5  4096-atom Si box, SchNet A512/F256/G33/L3 rc 4.9 + ExcludedVolume prior: forward epochs, and the adjoint backward
      through 20 steps (Simulations.simulate -> RDF loss -> .backward())

Each returns the fields of bench.py's JSON line (metric MD steps/sec; `e2e` = the same call with the host state of `System`,
i.e. H2D at the start and D2H of the last frame per epoch inside the timed region; `gpu_launches` = launches of the engine for
one timed epoch).  Weights are random (seeded) - there is no network for checkpoints.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _time_epochs(sim, integ, torch, steps, dt, reps):
    sim.simulate(steps=steps + 1, frequency=steps + 1, dt=dt)           # warm-up epoch (>= 3 steps)
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        v, q, pv = sim.simulate(steps=steps + 1, frequency=steps + 1, dt=dt)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    st = integ.last_engine_stats or {}
    return float(np.median(times)), st, q


def c1(args):
    import torch
    from torchmd.system import System
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones
    from torchmd.md import NoseHooverChain, Simulations
    from mdgrad_b200._ase_compat import FaceCenteredCubic
    atoms = FaceCenteredCubic(symbol="H", size=(3, 3, 3), latticeconstant=1.679, pbc=True)
    system = System(atoms, device=int(os.environ.get("LOCAL_RANK", "0")))
    system.set_velocities(np.random.default_rng(0).standard_normal((108, 3)) * np.sqrt(1.0 / 1.008))
    pair = PairPotentials(system, LennardJones(1.0, 1.0), cutoff=2.5)
    integ = NoseHooverChain(pair, system, T=1.0, num_chains=5, Q=50.0, adjoint=True)
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    steps = 50
    el, st, q = _time_epochs(sim, integ, torch, steps, 0.01, reps=max(3, min(50, args.steps // steps)))
    return {"value": steps / el, "ms_per_step": 1e3 * el / steps, "gpu_launches": int(st.get("launches", 0)),
            "config": {"workload": "C1: 108-atom FCC LJ (3x3x3, sigma=eps=1, rc 2.5), NoseHooverChain M=5 Q=50 T=1, 50 steps per "
                                   "Simulations.simulate epoch, dt 0.01 (reference demo); median epoch, host System state in and out",
                       "atoms": 108, "steps_per_epoch": steps, "launches_per_step": st.get("launches", 0) / steps,
                       "finite": bool(torch.isfinite(q).all())}}


def _schnet(args, which):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import schnet_md_bench as B
    sim, integ, gnn, dt, label = B.build(which)
    steps = 200 if which == "water" else 20
    if which == "si":
        # random (seeded) weights are not a stable force field: a short time step keeps the 4096-atom box finite over the few
        # epochs timed here - the work per step does not depend on dt
        dt = 0.1 * dt
        sim.simulate(steps=4, frequency=4, dt=dt)                        # warm-up (3 steps)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        v, q, pv = sim.simulate(steps=steps + 1, frequency=steps + 1, dt=dt)
        torch.cuda.synchronize()
        el, st = time.perf_counter() - t0, integ.last_engine_stats or {}
    else:
        el, st, q = _time_epochs(sim, integ, torch, steps, dt, reps=3)
    mode = {0: "synchronous", 1: "asynchronous", 2: "asynchronous + graph replay"}.get(st.get("maxrow_or_K"))
    out = {"value": steps / el, "ms_per_step": 1e3 * el / steps, "gpu_launches": int(st.get("launches", 0)),
           "config": {"workload": ("C3: " if which == "water" else "C5: ") + label, "atoms": int(q.shape[1]), "steps_per_epoch": steps,
                      "edges": int(gnn.inputs["nbr_list"].shape[0]), "launches_per_step": st.get("launches", 0) / steps,
                      "engine_mode": mode, "dense_layers": "tcgen05 3xTF32" if os.environ.get("MDG_SCHNET_TC") == "1" else "SIMT fp32",
                      "finite": bool(torch.isfinite(q).all())}}
    return out, sim, integ, gnn, dt


def c3(args):
    import torch
    out, sim, integ, gnn, dt = _schnet(args, "water")
    from torchmd.observable import rdf
    g = np.load(os.path.join(ROOT, "tests", "golden", "schnet_water.npz"))
    oxy = [int(i) for i in np.nonzero(g["numbers"] == 8)[0]]
    obs = rdf(sim.system, 100, (1.0, 5.8), index_tuple=(oxy, oxy))
    v, q, pv = sim.simulate(steps=21, frequency=21, dt=dt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        _, _, gr = obs(q[-1:])
    torch.cuda.synchronize()
    out["config"]["rdf_oo_ms"] = 1e3 * (time.perf_counter() - t0)
    out["config"]["rdf_oo_finite"] = bool(torch.isfinite(gr).all())
    return out


def c5(args):
    import torch
    out, sim, integ, gnn, dt = _schnet(args, "si")
    dt = 0.1 * dt
    from torchmd.observable import rdf
    obs = rdf(sim.system, 30, (1.8, 4.9))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v, q, pv = sim.simulate(steps=21, frequency=21, dt=dt)
    _, _, gr = obs(q[-1:])
    loss = gr.pow(2).sum() + 1e3 * (v[-1] ** 2).sum()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    loss.backward()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    gn = float(sum((p.grad.double() ** 2).sum() for p in gnn.gnn.parameters() if p.grad is not None) ** 0.5)
    out["config"]["adjoint_20_steps"] = {"forward_s": t1 - t0, "backward_s": t2 - t1, "grad_norm": gn,
                                         "params_with_grad": int(sum(p.grad is not None for p in gnn.gnn.parameters())),
                                         "what": "Simulations.simulate(20 steps, adjoint=True) -> RDF + velocity loss -> .backward() "
                                                 "through the adjoint solver (torchmd/sovlers.py:211-293)"}
    return out


RUN = {"c1": c1, "c3": c3, "c5": c5}
