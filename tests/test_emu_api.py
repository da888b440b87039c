"""The UNMODIFIED Python layer (torchmd.* drop-in API: System / PairPotentials / GNNPotentials / Stack / NoseHooverChain /
Simulations / adjoint solver) driven end to end on host tensors through the CPU-emulated kernels (tests/cuemu), against
the committed reference fixtures.  Only four functions of the binding module are swapped (library handle, stream,
device guard, "is this a device tensor"); everything above the C ABI is the product code.

Test infrastructure: functional coverage in the GPU-less container; the `-m gpu` tests remain the parity gate.
"""
import contextlib
import ctypes
import os

import numpy as np
import pytest
import torch

import emu_lib
from mdgrad_b200 import _lib, observable, topology

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(autouse=True)
def emulated_backend(monkeypatch):
    monkeypatch.setattr(_lib, "load", emu_lib.load)
    monkeypatch.setattr(_lib, "Context", lambda device: emu_lib.EmuContext())
    monkeypatch.setattr(_lib, "require_cuda", lambda t, name="tensor": None)
    monkeypatch.setattr(_lib, "on_device", lambda t: True)
    monkeypatch.setattr(_lib, "_stream", lambda device: ctypes.c_void_p(0))
    monkeypatch.setattr(_lib, "_guard", lambda device: contextlib.nullcontext())
    ctxs = {}
    emu_context_for = lambda device, key="default": ctxs.setdefault(key, emu_lib.EmuContext())    # noqa: E731
    monkeypatch.setattr(topology, "context_for", emu_context_for)
    monkeypatch.setattr(observable, "context_for", emu_context_for)     # (imported by name there)
    yield


def _fcc_system(size=3, a=1.679):
    from torchmd.system import System
    from mdgrad_b200._ase_compat import FaceCenteredCubic
    return System(FaceCenteredCubic(symbol="H", size=(size,) * 3, latticeconstant=a, pbc=True), device="cpu")


def test_emu_c1_simulate_vs_reference_fixture():
    """configs[0]: 108-atom LJ NoseHooverChain epoch through Simulations.simulate on the fused engine"""
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones
    from torchmd.md import NoseHooverChain, Simulations
    from torchmd.observable import rdf
    g = np.load(os.path.join(G, "c1_traj.npz"))
    system = _fcc_system()
    system.set_positions(g["q0"])
    system.set_velocities(g["v0"])
    pair = PairPotentials(system, LennardJones(1.0, 1.0), cutoff=2.5)
    integ = NoseHooverChain(pair, system, T=1.0, num_chains=5, Q=50.0, adjoint=True, topology_update_freq=1)
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    v, q, pv = sim.simulate(steps=50, frequency=50, dt=0.01)
    v, q, pv = v.detach(), q.detach(), pv.detach()
    assert integ.last_engine_stats is not None and integ.update_count == 98
    assert np.array_equal(v[0].numpy(), g["v"][0]) and np.array_equal(q[0].numpy(), g["q"][0])
    for k, tol in ((1, 2e-6), (10, 3e-5), (49, 3e-3)):
        assert np.abs(q[k].numpy() - g["q"][k]).max() < tol
        assert np.abs(v[k].numpy() - g["v"][k]).max() < 10 * tol
        assert np.abs(pv[k].numpy() - g["pv"][k]).max() < 30 * tol
    obs = rdf(system, 100, (0.75, 2.0))
    count, bins, gr = obs(torch.tensor(g["q"][-1]))
    np.testing.assert_allclose(gr.numpy(), g["rdf_g"], rtol=1e-5, atol=1e-5 * g["rdf_g"].max())


def _adjoint_grads(native, steps=50, seed_loss=True):
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones
    from torchmd.md import NoseHooverChain, Simulations
    g = np.load(os.path.join(G, "c1_traj.npz"))
    system = _fcc_system()
    system.set_positions(g["q0"])
    system.set_velocities(g["v0"])
    lj = LennardJones(1.0, 1.0)
    integ = NoseHooverChain(PairPotentials(system, lj, cutoff=2.5), system, T=1.0, num_chains=5, Q=50.0, adjoint=True)
    integ.disable_native_adjoint = not native
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    v, q, pv = sim.simulate(steps=steps, frequency=steps, dt=0.01)
    k1, k2 = (20, 30) if steps > 30 else (2, 4)
    loss = (q[-1] ** 2).sum() + (v[k1] * v[k2]).sum() + pv[-1].sum()
    loss.backward()
    return lj.sigma.grad.item(), lj.epsilon.grad.item(), g


def test_emu_adjoint_native_route_vs_reference_fixture():
    """49 reverse steps with the ANALYTIC augmented dynamics (force kernel + mdg_pair_hvp + written-out thermostat
    algebra, no autograd) against the reference's d loss / d sigma, d loss / d epsilon"""
    ds, de, g = _adjoint_grads(native=True)
    assert abs(ds - g["dsigma"][0]) <= 2e-2 * abs(g["dsigma"][0])
    assert abs(de - g["depsilon"][0]) <= 2e-2 * abs(g["depsilon"][0])


def test_emu_adjoint_native_route_equals_autograd_route():
    """short horizon (chaos-free): analytic route == generic double-backward route"""
    a = _adjoint_grads(native=True, steps=8)
    b = _adjoint_grads(native=False, steps=8)
    assert abs(a[0] - b[0]) <= 2e-4 * max(1.0, abs(b[0]))
    assert abs(a[1] - b[1]) <= 2e-4 * max(1.0, abs(b[1]))


def test_emu_nve_adjoint_native_vs_autograd():
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones69
    from torchmd.md import NVE, Simulations
    g = np.load(os.path.join(G, "c1_nve.npz"))
    out = []
    for native in (True, False):
        system = _fcc_system()
        system.set_positions(g["q0"])
        system.set_velocities(g["v0"])
        pot = LennardJones69(1.0, 0.9)
        integ = NVE(PairPotentials(system, pot, cutoff=2.5), system, adjoint=True)
        integ.disable_native_adjoint = not native
        sim = Simulations(system, integ, wrap=True, method="verlet")
        v, q = sim.simulate(steps=8, frequency=8, dt=0.005)
        ((q[-1] ** 2).sum() + (v[3] * v[5]).sum()).backward()
        out.append((pot.sigma.grad.item(), pot.epsilon.grad.item()))
    assert abs(out[0][0] - out[1][0]) <= 2e-4 * max(1.0, abs(out[1][0]))
    assert abs(out[0][1] - out[1][1]) <= 2e-4 * max(1.0, abs(out[1][1]))


def test_emu_schnet_md_native_force_equals_autograd_route():
    """Stack(GNNPotentials(SchNet) + O-O ExcludedVolume prior) under NoseHooverChain (configs[2] shape): the no-grad
    solver route (native SchNet energy+force program, pair-force kernel) vs the op-by-op autograd route"""
    from nff.nn.models.schnet import SchNet
    from test_schnet import _fixture
    from torchmd.interface import GNNPotentials, PairPotentials, Stack
    from torchmd.potentials import ExcludedVolume
    from torchmd.md import NoseHooverChain
    from torchmd.sovlers import odeint, odeint_reuse_force
    from torchmd.system import System
    from mdgrad_b200._ase_compat import Atoms, units
    g, params, sd = _fixture("water")
    system = System(Atoms(numbers=g["numbers"], positions=g["positions"], cell=g["cell"], pbc=True), device="cpu")
    np.random.seed(0)
    system.set_temperature(298.0 * units.kB)
    model = SchNet(params)
    model.load_state_dict(sd)
    gnn = GNNPotentials(system, model, cutoff=params["cutoff"])
    e, f = gnn.native_energy_force(torch.Tensor(system.get_positions()))
    eref = float(g["energy"].reshape(-1)[0])
    assert abs(e.item() - eref) <= 1e-5 * abs(eref)
    assert np.abs(f.numpy() - g["forces"]).max() <= 1e-5 * np.abs(g["forces"]).max()
    oxy = [int(i) for i in np.nonzero(g["numbers"] == 8)[0]]
    prior = PairPotentials(system, ExcludedVolume(2.6, 0.015, 12), cutoff=params["cutoff"], index_tuple=(oxy, oxy))
    integ = NoseHooverChain(Stack({"gnn": gnn, "prior": prior}), system, T=298.0 * units.kB, num_chains=5, Q=50.0, adjoint=True)
    assert integ.model.native_ready()
    y0 = tuple(integ.get_inital_states(True))
    t = torch.Tensor([0.5 * units.fs * i for i in range(4)])
    with torch.no_grad():
        a = odeint_reuse_force(integ, y0, t, "NH_verlet")
    integ.adjoint = False                                          # plain autograd solve (reference md.py:88-91)
    b = odeint(integ, tuple(v.clone().requires_grad_(True) for v in y0), t, method="NH_verlet")
    for xa, xb in zip(a, b):
        assert (xa - xb.detach()).abs().max().item() <= 2e-5 * max(1e-3, xb.detach().abs().max().item())


def test_emu_simulate_device_handoff_equals_host_roundtrip():
    """Simulations.simulate over several epochs: state kept on the device between epochs (bit-identical fp64 wrap)
    vs the reference's host round trip after every epoch - identical log, System and trajectories."""
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones
    from torchmd.md import NoseHooverChain, Simulations
    g = np.load(os.path.join(G, "c1_traj.npz"))
    runs = []
    for handoff in (True, False):
        system = _fcc_system()
        system.set_positions(g["q0"] + 7.0)                 # several atoms outside the box: the wrap matters
        system.set_velocities(g["v0"])
        integ = NoseHooverChain(PairPotentials(system, LennardJones(1.0, 1.0), cutoff=2.5), system, T=1.0, num_chains=5, Q=50.0)
        sim = Simulations(system, integ, wrap=True, method="NH_verlet")
        sim.device_handoff = handoff
        v, q, pv = sim.simulate(steps=4 * 6, frequency=6, dt=0.01)
        v2, q2, pv2 = sim.simulate(steps=6, frequency=6, dt=0.01)          # continues from the logged check point
        runs.append((sim.log, system.get_positions(), system.get_velocities(), v.detach(), q.detach(), pv.detach(), q2.detach()))
    a, b = runs
    assert len(a[0]["positions"]) == 5 and len(b[0]["positions"]) == 5
    for key in ("velocities", "positions", "baths"):
        for x, y in zip(a[0][key], b[0][key]):
            assert x.dtype == y.dtype and np.array_equal(x, y)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for k in range(3, 7):
        assert torch.equal(a[k], b[k])


def test_emu_gnn_engine_epoch_equals_op_level_solver():
    """Simulations.simulate with Stack(GNNPotentials + prior): device engine epoch (mdg_md_run_gnn) vs the op-level
    solver with native forces; NoseHooverChain and NVE"""
    from nff.nn.models.schnet import SchNet
    from test_schnet import _fixture
    from torchmd.interface import GNNPotentials, PairPotentials, Stack
    from torchmd.potentials import ExcludedVolume
    from torchmd.md import NVE, NoseHooverChain, Simulations
    from torchmd.sovlers import odeint_reuse_force
    from torchmd.system import System
    from mdgrad_b200._ase_compat import Atoms, units
    g, params, sd = _fixture("water")
    for which in ("nhc", "nve"):
        system = System(Atoms(numbers=g["numbers"], positions=g["positions"], cell=g["cell"], pbc=True), device="cpu")
        np.random.seed(0)
        system.set_temperature(298.0 * units.kB)
        model = SchNet(params)
        model.load_state_dict(sd)
        gnn = GNNPotentials(system, model, cutoff=params["cutoff"])
        oxy = [int(i) for i in np.nonzero(g["numbers"] == 8)[0]]
        prior = PairPotentials(system, ExcludedVolume(2.6, 0.015, 12), cutoff=params["cutoff"], index_tuple=(oxy, oxy))
        stack = Stack({"gnn": gnn, "prior": prior})
        if which == "nhc":
            integ = NoseHooverChain(stack, system, T=298.0 * units.kB, num_chains=5, Q=50.0, adjoint=True)
            method = "NH_verlet"
        else:
            integ = NVE(stack, system, adjoint=True)
            method = "verlet"
        sim = Simulations(system, integ, wrap=True, method=method)
        out = sim.simulate(steps=5, frequency=5, dt=0.5 * units.fs)
        assert integ.last_engine_stats is not None and integ.update_count == 8
        nbr_after = gnn.inputs["nbr_list"].clone()
        integ.disable_gnn_engine = True
        t = torch.Tensor([0.5 * units.fs * i for i in range(5)])
        with torch.no_grad():
            ref = odeint_reuse_force(integ, tuple(o[0].detach() for o in out), t, method)
        for a, b in zip(out, ref):
            assert (a.detach() - b).abs().max().item() <= 2e-6 * max(1.0, b.abs().max().item())
        assert torch.equal(nbr_after, gnn.inputs["nbr_list"])          # python-visible list = list at the last frame


def test_emu_gnn_engine_async_steps_and_overflow_retry(monkeypatch):
    """The device engine's GNN epochs run their steps without any read-back (pair count consumed on the device, edge buffers sized
    from the first evaluation); a too-small capacity is latched and the epoch repeated on the synchronous path.  All three ways
    (asynchronous, forced synchronous, forced overflow -> retry) and the opt-in CUDA-graph replay give the SAME bits."""
    from nff.nn.models.schnet import SchNet
    from torchmd.interface import GNNPotentials, PairPotentials, Stack
    from torchmd.potentials import ExcludedVolume
    from torchmd.md import NoseHooverChain, Simulations
    from torchmd.system import System
    from mdgrad_b200._ase_compat import Atoms, units
    from test_schnet import _fixture
    g, params, sd = _fixture("water")
    runs = {}
    for mode in ("async", "sync", "overflow", "async_long", "graph"):
        monkeypatch.delenv("MDG_GNN_SYNC", raising=False)
        monkeypatch.delenv("MDG_GNN_MARGIN", raising=False)
        monkeypatch.setenv("MDG_GNN_GRAPH", "0")               # plain asynchronous launches (graph replay is the default)
        if mode == "graph":
            monkeypatch.setenv("MDG_GNN_GRAPH", "1")           # force evaluations captured once, replayed as a CUDA graph
        if mode == "sync":
            monkeypatch.setenv("MDG_GNN_SYNC", "1")
        if mode == "overflow":
            monkeypatch.setenv("MDG_GNN_MARGIN", "-40")        # edge buffers 40 pairs SMALLER than the first count
        system = System(Atoms(numbers=g["numbers"], positions=g["positions"], cell=g["cell"], pbc=True), device="cpu")
        np.random.seed(0)
        system.set_temperature(298.0 * units.kB)
        model = SchNet(params)
        model.load_state_dict(sd)
        gnn = GNNPotentials(system, model, cutoff=params["cutoff"])
        oxy = [int(i) for i in np.nonzero(g["numbers"] == 8)[0]]
        prior = PairPotentials(system, ExcludedVolume(2.6, 0.015, 12), cutoff=params["cutoff"], index_tuple=(oxy, oxy))
        integ = NoseHooverChain(Stack({"gnn": gnn, "prior": prior}), system, T=298.0 * units.kB, num_chains=5, Q=50.0, adjoint=True)
        sim = Simulations(system, integ, wrap=True, method="NH_verlet")
        real_run = emu_lib.EmuContext.md_run_gnn
        syncs = []

        def counted(self, *a, **k):
            c0 = emu_lib.load().cuemu_counter(3)
            r = real_run(self, *a, **k)
            syncs.append(emu_lib.load().cuemu_counter(3) - c0)
            return r
        monkeypatch.setattr(emu_lib.EmuContext, "md_run_gnn", counted)
        nst = 8 if mode == "async_long" else 5
        out = sim.simulate(steps=nst, frequency=nst, dt=0.5 * units.fs)
        monkeypatch.setattr(emu_lib.EmuContext, "md_run_gnn", real_run)
        runs[mode] = ([o.detach().clone() for o in out], integ.last_engine_stats["maxrow_or_K"], syncs[0])
    assert runs["async"][1] == 1 and runs["sync"][1] == 0 and runs["overflow"][1] == 0      # which path completed the epoch
    # host-blocking synchronisations inside the engine call (counted by the emulator's runtime): two per evaluation on the
    # synchronous path; on the asynchronous one a constant (first evaluation, one-time buffer growth, end of the epoch) that
    # does not depend on the number of steps
    assert runs["sync"][2] >= 2 * 5 and runs["async"][2] < runs["sync"][2], (runs["sync"][2], runs["async"][2])
    assert runs["async_long"][1] == 1 and runs["async_long"][2] == runs["async"][2], (runs["async_long"][2], runs["async"][2])
    assert runs["graph"][1] == 2                                 # ... with graph replays (private stream, captured at step 2)
    # a following epoch of the same system reuses the capacity: no pair-count read-back at all any more
    system = System(Atoms(numbers=g["numbers"], positions=g["positions"], cell=g["cell"], pbc=True), device="cpu")
    np.random.seed(0)
    system.set_temperature(298.0 * units.kB)
    model = SchNet(params)
    model.load_state_dict(sd)
    gnn = GNNPotentials(system, model, cutoff=params["cutoff"])
    prior = PairPotentials(system, ExcludedVolume(2.6, 0.015, 12), cutoff=params["cutoff"], index_tuple=(oxy, oxy))
    integ = NoseHooverChain(Stack({"gnn": gnn, "prior": prior}), system, T=298.0 * units.kB, num_chains=5, Q=50.0, adjoint=True)
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    real_run, per_epoch = emu_lib.EmuContext.md_run_gnn, []

    def counted2(self, *a, **k):
        c0 = emu_lib.load().cuemu_counter(3)
        r = real_run(self, *a, **k)
        per_epoch.append(emu_lib.load().cuemu_counter(3) - c0)
        return r
    monkeypatch.setattr(emu_lib.EmuContext, "md_run_gnn", counted2)
    sim.simulate(steps=15, frequency=5, dt=0.5 * units.fs)       # three epochs
    monkeypatch.setattr(emu_lib.EmuContext, "md_run_gnn", real_run)
    assert len(per_epoch) == 3 and per_epoch[1] <= per_epoch[0] - 2 and per_epoch[2] == per_epoch[1], per_epoch
    assert per_epoch[2] <= 3, per_epoch                          # end-of-epoch synchronisation (+ the host copy of the bath trajectory)
    for mode in ("sync", "overflow", "graph"):
        for a, b in zip(runs["async"][0], runs[mode][0]):
            assert torch.equal(a, b), mode


def test_emu_schnet_second_order_through_native_aggregation():
    import schnet_checks
    schnet_checks.check_second_order_through_native_aggregation("cpu")


def test_emu_gnn_adjoint_fit_vs_reference_fixture():
    import schnet_checks
    schnet_checks.check_gnn_adjoint_fit_vs_reference_fixture("cpu")


def test_emu_water_rdf_oo_species_selection():
    import schnet_checks
    schnet_checks.check_water_rdf_oo_species_selection("cpu")


def test_emu_angle_distribution_vs_live_reference():
    """angle_distribution (native neighbor list -> device-side triple enumeration -> smeared histogram) against the
    unmodified reference observable (authoring container only)"""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    from torchmd.observable import angle_distribution
    from torchmd.system import System
    from mdgrad_b200._ase_compat import Atoms
    rng = np.random.default_rng(4)
    n, L = 81, 9.3
    pos = rng.uniform(0, L, (n, 3))
    numbers = np.array(([8, 1, 1] * n)[:n])
    atoms = Atoms(numbers=numbers, positions=pos, cell=[L] * 3, pbc=True)
    oxy = [int(i) for i in np.nonzero(numbers == 8)[0]]
    xyz = torch.tensor(np.stack([pos, pos + 0.1 * rng.standard_normal((n, 3))]), dtype=torch.float32)
    obs = angle_distribution(System(atoms, device="cpu"), nbins=32, angle_range=(0.0, np.pi), cutoff=3.3, index_tuple=(oxy, oxy))
    bins, count, angles = obs(xyz)
    with ref_import.active() as ref:
        robs = ref.observable.angle_distribution(ref.system.System(atoms, device="cpu"), nbins=32, angle_range=(0.0, np.pi),
                                                 cutoff=3.3, index_tuple=(oxy, oxy))
        rbins, rcount, rangles = robs(xyz)
    assert angles.shape == rangles.shape and angles.shape[0] > 50
    assert torch.equal(bins, rbins)
    torch.testing.assert_close(angles, rangles, rtol=0, atol=2e-6)
    torch.testing.assert_close(count, rcount, rtol=1e-5, atol=1e-7)


def test_emu_stack_of_pair_potentials_on_device_engine():
    """scripts/fit_2_comp.py shape: Stack of three species-pair PairPotentials (index_tuples) under NoseHooverChain -
    the epoch runs on the device engine (no SchNet member) and equals the op-level solver"""
    from torchmd.interface import PairPotentials, Stack
    from torchmd.potentials import LennardJones
    from torchmd.md import NoseHooverChain, Simulations
    from torchmd.sovlers import odeint_reuse_force
    system = _fcc_system(size=4)
    n = len(system)
    np.random.seed(1)
    system.set_temperature(1.0)
    A, B = list(range(0, n, 2)), list(range(1, n, 2))
    stack = Stack({"aa": PairPotentials(system, LennardJones(1.0, 1.0), cutoff=2.5, index_tuple=(A, A)),
                   "bb": PairPotentials(system, LennardJones(0.9, 0.8), cutoff=2.2, index_tuple=(B, B)),
                   "ab": PairPotentials(system, LennardJones(0.95, 1.1), cutoff=2.5, index_tuple=(A, B))})
    integ = NoseHooverChain(stack, system, T=1.0, num_chains=3, Q=20.0, adjoint=True)
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    out = sim.simulate(steps=6, frequency=6, dt=0.005)
    assert integ.last_engine_stats is not None and integ.update_count == 10
    integ.disable_gnn_engine = True
    t = torch.Tensor([0.005 * i for i in range(6)])
    with torch.no_grad():
        ref = odeint_reuse_force(integ, tuple(o[0].detach() for o in out), t, "NH_verlet")
    for a, b in zip(out, ref):
        assert (a.detach() - b).abs().max().item() <= 2e-6 * max(1.0, b.abs().max().item())


def test_emu_gnn_engine_cell_list_path_4096_atoms():
    """configs[4] geometry (4096-atom diamond Si, rc 4.9: the cell-list path of the per-step list builds) with a narrow
    SchNet: device engine epoch vs op-level solver"""
    from nff.train import get_model
    from torchmd.interface import GNNPotentials, PairPotentials, Stack
    from torchmd.potentials import ExcludedVolume
    from torchmd.md import NoseHooverChain, Simulations
    from torchmd.sovlers import odeint_reuse_force
    from torchmd.system import System
    from mdgrad_b200._ase_compat import Atoms, units
    a, nc = 5.45933, 8
    basis = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
    basis = np.concatenate([basis, basis + 0.25])
    cells = np.array([[i, j, k] for i in range(nc) for j in range(nc) for k in range(nc)])
    pos = ((cells[:, None, :] + basis[None, :, :]).reshape(-1, 3)) * a
    pos = pos + np.random.default_rng(3).normal(0, 0.05, pos.shape)
    system = System(Atoms(numbers=[14] * len(pos), positions=pos, cell=[a * nc] * 3, pbc=True), device="cpu")
    np.random.seed(0)
    system.set_temperature(100.0 * units.kB)
    torch.manual_seed(0)
    model = get_model({"n_atom_basis": 16, "n_filters": 12, "n_gaussians": 9, "n_convolutions": 2, "cutoff": 4.9, "trainable_gauss": False})
    gnn = GNNPotentials(system, model, cutoff=4.9)
    assert gnn.inputs["nbr_list"].shape[0] == 57344 or gnn.inputs["nbr_list"].shape[0] > 50000        # E = 57 344 on the perfect lattice
    prior = PairPotentials(system, ExcludedVolume(1.9, 0.015, 12), cutoff=4.9)
    integ = NoseHooverChain(Stack({"gnn": gnn, "prior": prior}), system, T=100.0 * units.kB, num_chains=5, Q=50.0, adjoint=True)
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    out = sim.simulate(steps=3, frequency=3, dt=1.0 * units.fs)
    assert integ.last_engine_stats is not None and integ.last_engine_stats["path"] == 0
    integ.disable_gnn_engine = True
    t = torch.Tensor([1.0 * units.fs * i for i in range(3)])
    with torch.no_grad():
        ref = odeint_reuse_force(integ, tuple(o[0].detach() for o in out), t, "NH_verlet")
    for x, y in zip(out, ref):
        assert (x.detach() - y).abs().max().item() <= 2e-6 * max(1.0, y.abs().max().item())


def test_emu_thermo_temperature_and_virial_pressure():
    """Temperature = the reference's algebra; Pressure (working replacement of the reference's broken class) = ideal part
    minus dE/dV of the pair energy, checked against a finite difference under uniform scaling of the box"""
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones
    from torchmd.thermo import Pressure, Temperature
    from torchmd.system import System
    from mdgrad_b200._ase_compat import FaceCenteredCubic
    rng = np.random.default_rng(0)

    def energy(scale):
        atoms = FaceCenteredCubic(symbol="H", size=(4, 4, 4), latticeconstant=1.679 * scale, pbc=True)
        system = System(atoms, device="cpu")
        pos = (base_frac * scale)
        system.set_positions(pos)
        pair = PairPotentials(system, LennardJones(1.0, 1.0), cutoff=2.5 * scale)      # same pair set under scaling
        q = torch.Tensor(pos)
        pair._reset_topology(q)
        return system, pair, q, float(pair(q).detach())

    atoms0 = FaceCenteredCubic(symbol="H", size=(4, 4, 4), latticeconstant=1.679, pbc=True)
    base_frac = atoms0.get_positions() + 0.05 * rng.standard_normal((len(atoms0), 3))
    system, pair, q, e0 = energy(1.0)
    n = len(system)
    v = torch.tensor(rng.standard_normal((n, 3)), dtype=torch.float32)
    T = Temperature(system)(v)
    m = torch.Tensor(system.get_masses())
    assert abs(T.item() - float((m[:, None] * v * v).sum() / (3 * n))) <= 1e-5 * T.item()
    P = Pressure(system, pair)(q, v).item()
    h = 1e-3
    ep, em = energy(1.0 + h)[3], energy(1.0 - h)[3]
    V = system.get_volume()
    dEdV = (ep - em) / (V * ((1 + h) ** 3 - (1 - h) ** 3))
    assert abs(P - (n * T.item() / V - dEdV)) <= 2e-3 * abs(P)


def test_emu_buck_adjoint_native_vs_autograd():
    """three-parameter potential (Buck A, B, C): analytic reverse dynamics == the double-backward route"""
    from torchmd.interface import PairPotentials
    from torchmd.potentials import Buck
    from torchmd.md import NoseHooverChain, Simulations
    g = np.load(os.path.join(G, "c1_traj.npz"))
    out = []
    for native in (True, False):
        system = _fcc_system()
        system.set_positions(g["q0"])
        system.set_velocities(g["v0"])
        pot = Buck(900.0, 3.2, 1.5)
        integ = NoseHooverChain(PairPotentials(system, pot, cutoff=2.5), system, T=1.0, num_chains=3, Q=20.0, adjoint=True)
        integ.disable_native_adjoint = not native
        sim = Simulations(system, integ, wrap=True, method="NH_verlet")
        v, q, pv = sim.simulate(steps=7, frequency=7, dt=0.005)
        ((q[-1] ** 2).sum() + (v[2] * v[5]).sum() + pv[-1].sum()).backward()
        out.append([p.grad.item() for p in (pot.A, pot.B, pot.C)])
    for x, y in zip(*out):
        assert abs(x - y) <= 3e-4 * max(1.0, abs(y)), out


def test_emu_native_schnet_sees_in_place_weight_updates():
    """the library caches derived weight layouts per weights_tag: an optimiser step (in-place update, version bump) must be
    picked up by the next native evaluation"""
    from nff.nn.models.schnet import SchNet
    from test_schnet import _fixture
    from torchmd.interface import GNNPotentials
    from torchmd.system import System
    from mdgrad_b200._ase_compat import Atoms
    from oracle import oracle_torch as O
    g, params, sd = _fixture("water")
    system = System(Atoms(numbers=g["numbers"], positions=g["positions"], cell=g["cell"], pbc=True), device="cpu")
    model = SchNet(params)
    model.load_state_dict(sd)
    gnn = GNNPotentials(system, model, cutoff=params["cutoff"])
    xyz = torch.Tensor(system.get_positions())
    e0, _ = gnn.native_energy_force(xyz)
    e0b, _ = gnn.native_energy_force(xyz)
    assert e0.item() == e0b.item()
    with torch.no_grad():
        model.convolutions[0].moduledict["message_edge_filter"][3].weight.mul_(1.05)
        model.convolutions[1].moduledict["message_edge_filter"][1].weight.add_(0.01)
    e1, f1 = gnn.native_energy_force(xyz)
    z = torch.tensor(g["numbers"], dtype=torch.long)
    x = xyz.clone().requires_grad_(True)
    eo = O.schnet_energy(model.state_dict(), z, x, gnn.inputs["nbr_list"], gnn.inputs["offsets"])
    fo = -torch.autograd.grad(eo, x)[0]
    assert abs(e1.item() - e0.item()) > 1e-4 * abs(e0.item())
    assert abs(e1.item() - eo.item()) <= 1e-5 * abs(eo.item())
    assert (f1 - fo).abs().max().item() <= 2e-5 * fo.abs().max().item()


@pytest.mark.parametrize("tag", ["lj", "buck"])
@pytest.mark.parametrize("route", ["native_hvp", "autograd"])
def test_emu_adjoint_gradients_short_horizon_vs_reference_fixture(tag, route):
    """the body of tests/test_gpu_api.py::test_adjoint_gradients_short_horizon_vs_reference_fixture on the emulated kernels:
    5-step adjoint gradients of every potential parameter within 1e-4 of the unmodified reference's"""
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones, Buck
    from torchmd.md import NoseHooverChain, Simulations
    g = np.load(os.path.join(G, "c1_adjoint_short.npz"))
    system = _fcc_system()
    system.set_positions(g["q0"])
    system.set_velocities(g["v0"])
    pot = LennardJones(1.1, 0.9) if tag == "lj" else Buck(1000.0, 3.5, 2.0)
    integ = NoseHooverChain(PairPotentials(system, pot, cutoff=2.5), system, T=1.0, num_chains=5, Q=50.0, adjoint=True)
    integ.disable_native_adjoint = route == "autograd"
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    v, q, pv = sim.simulate(steps=6, frequency=6, dt=0.01)
    assert np.abs(q.detach().numpy() - g["q_" + tag]).max() <= 2e-6 * 5.04
    assert np.abs(v.detach().numpy() - g["v_" + tag]).max() <= 2e-5 * np.abs(g["v_" + tag]).max()
    loss = (q[-1] ** 2).sum() + (v[2] * v[4]).sum() + pv[-1].sum()
    assert abs(loss.item() - float(g["loss_" + tag])) <= 1e-5 * abs(float(g["loss_" + tag]))
    loss.backward()
    scale = max(abs(float(g["d%s_%s" % (n, tag)].reshape(-1)[0])) for n, _ in pot.named_parameters())
    for name, prm in pot.named_parameters():
        ref = float(g["d%s_%s" % (name, tag)].reshape(-1)[0])
        assert prm.grad is not None, name
        assert abs(prm.grad.item() - ref) <= 1e-4 * max(abs(ref), 1e-2 * scale), (name, prm.grad.item(), ref)


def test_emu_direct_odeint_without_adjoint_keeps_parameter_gradients():
    """the public `odeint` called directly with adjoint=False and states that require grad must put the forces'
    twice-differentiable form on the tape - parameter gradients equal those of the same solve through Simulations
    (which equal the reference's, test_emu_live_matrix)."""
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones
    from torchmd.md import NoseHooverChain, Simulations
    from torchmd.sovlers import odeint
    g = np.load(os.path.join(G, "c1_traj.npz"))

    def build():
        system = _fcc_system()
        system.set_positions(g["q0"])
        system.set_velocities(g["v0"])
        lj = LennardJones(1.0, 1.0)
        integ = NoseHooverChain(PairPotentials(system, lj, cutoff=2.5), system, T=1.0, num_chains=5, Q=50.0, adjoint=False)
        return system, lj, integ

    system, lj, integ = build()
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    v, q, pv = sim.simulate(steps=6, frequency=6, dt=0.01)
    ((q[-1] ** 2).sum() + (v[2] * v[4]).sum()).backward()
    ref = (lj.sigma.grad.item(), lj.epsilon.grad.item())
    system, lj, integ = build()
    y0 = tuple(s.requires_grad_(True) for s in integ.get_inital_states(True))
    t = torch.Tensor([0.01 * i for i in range(6)])
    v, q, pv = odeint(integ, y0, t, method="NH_verlet")
    ((q[-1] ** 2).sum() + (v[2] * v[4]).sum()).backward()
    assert lj.sigma.grad is not None and lj.epsilon.grad is not None
    assert abs(lj.sigma.grad.item() - ref[0]) <= 1e-5 * abs(ref[0]) and abs(lj.epsilon.grad.item() - ref[1]) <= 1e-5 * abs(ref[1])
    assert y0[1].grad is not None and torch.isfinite(y0[1].grad).all()


def test_emu_f2_observables_vs_reference_fixture():
    """bodies of tests/test_gpu_observables.py on the emulated kernels (vacf lag-product kernel, angle distribution over the
    native list, Temperature / virial Pressure)"""
    import observable_checks as C
    cpu = torch.device("cpu")
    C.check_vacf("cpu", cpu)
    C.check_angle_distribution("cpu", cpu)
    C.check_temperature_and_pressure("cpu", cpu)
