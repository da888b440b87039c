"""Differential test of the public API against the UNMODIFIED reference imported in-process (authoring container only; skipped
where /root/reference is absent): a matrix of Simulations settings - integrator, chain length, wrap on/off, several epochs,
species selection / exclusions, potential kind - each run through the reference on the CPU and through this package on the
CPU-emulated kernels, comparing the logged frames, the System state and the returned trajectory.
TEST INFRASTRUCTURE (tests/cuemu)."""
import numpy as np
import pytest
import torch

from oracle import ref_import
from test_emu_api import emulated_backend  # noqa: F401

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")

CASES = [
    # integrator, chains, wrap, potential, kwargs of PairPotentials, epochs x steps, dt
    ("nhc", 5, True, ("LennardJones", (1.0, 1.0)), {}, (2, 5), 0.01),
    ("nhc", 2, False, ("LennardJones", (1.0, 1.0)), {}, (3, 4), 0.005),
    ("nve", 0, True, ("LennardJones69", (1.05, 0.8)), {}, (2, 6), 0.005),
    ("nhc", 3, True, ("ExcludedVolume", (1.0, 0.7, 10)), {"index_tuple": "AB"}, (2, 5), 0.01),
    ("nve", 0, False, ("Buck", (800.0, 3.4, 1.5)), {"ex_pairs": True}, (2, 4), 0.004),
    ("nhc", 4, True, ("LJFamily", (1.0, 0.9, 5, 10)), {"index_tuple": "AB", "ex_pairs": True}, (1, 7), 0.01),
    ("nhc", 5, True, ("LennardJones", (1.0, 1.0)), {"continue": True}, (2, 4), 0.01),
    ("nve", 0, True, ("ModifiedMorse", (6.0, 2.0)), {"continue": True}, (1, 5), 0.004),
]


def _build(ns, case, seed):
    """ns: namespace with .system .interface .potentials .md (reference or ours)"""
    from mdgrad_b200._ase_compat import FaceCenteredCubic
    kind, chains, wrap, (pname, pargs), pkw, (epochs, steps), dt = case
    atoms = FaceCenteredCubic(symbol="H", size=(3, 3, 3), latticeconstant=1.679, pbc=True)
    rng = np.random.default_rng(seed)
    atoms.set_positions(atoms.get_positions() + rng.normal(0, 0.04, (108, 3)) + (0.0 if wrap else 3.0))
    system = ns.system.System(atoms, device="cpu")
    system.set_velocities(rng.standard_normal((108, 3)) * 0.9)
    kw = {}
    if "index_tuple" in pkw:
        kw["index_tuple"] = (list(range(0, 108, 2)), list(range(1, 108, 3)))
    if "ex_pairs" in pkw:
        kw["ex_pairs"] = torch.LongTensor(rng.integers(0, 108, (40, 2)))
    if pname == "LJFamily":
        pot = ns.potentials.LJFamily(pargs[0], pargs[1], attr_pow=pargs[2], rep_pow=pargs[3])
    else:
        pot = getattr(ns.potentials, pname)(*pargs)
    pair = ns.interface.PairPotentials(system, pot, cutoff=2.5, **kw)
    if kind == "nhc":
        integ = ns.md.NoseHooverChain(pair, system, T=1.0, num_chains=chains, Q=50.0, adjoint=True)
        method = "NH_verlet"
    else:
        integ = ns.md.NVE(pair, system, adjoint=True)
        method = "verlet"
    sim = ns.md.Simulations(system, integ, wrap=wrap, method=method)
    out = sim.simulate(steps=epochs * steps, frequency=steps, dt=dt)
    if pkw.get("continue"):                 # a second call continues from the log (md.py:76-79); steps not a multiple of frequency
        out = sim.simulate(steps=2 * steps + 1, frequency=steps, dt=dt)
    loss = (out[1][-1] ** 2).sum() + (out[0][-1] * out[0][1]).sum()
    grads = {}
    if list(pot.parameters()):              # (ModifiedMorse has no parameters: nothing to differentiate)
        loss.backward()                     # adjoint reverse sweep of the LAST simulate() call
        grads = {n: p.grad.detach().numpy().copy() for n, p in pot.named_parameters() if p.grad is not None}
    return system, sim, [o.detach().numpy() for o in out], grads


@pytest.mark.parametrize("k", range(len(CASES)))
def test_emu_simulations_matrix_vs_live_reference(k):
    import types
    import torchmd
    case = CASES[k]
    with ref_import.active() as ref:
        rs, rsim, rout, rgrads = _build(ref, case, seed=100 + k)
    ours = types.SimpleNamespace(system=torchmd.system, interface=torchmd.interface, potentials=torchmd.potentials, md=torchmd.md)
    os_, osim, oout, ograds = _build(ours, case, seed=100 + k)
    assert list(osim.log.keys()) == list(rsim.log.keys())
    for key in rsim.log:
        assert len(osim.log[key]) == len(rsim.log[key]) == case[5][0] + (2 if case[4].get("continue") else 0)
        for a, b in zip(osim.log[key], rsim.log[key]):
            assert a.shape == b.shape and a.dtype == b.dtype
            assert np.abs(a - b).max() <= 3e-5 * max(1.0, np.abs(b).max()), (key, np.abs(a - b).max())
    assert len(oout) == len(rout)
    for a, b in zip(oout, rout):
        assert a.shape == b.shape
        assert np.abs(a - b).max() <= 3e-5 * max(1.0, np.abs(b).max())
    np.testing.assert_allclose(os_.get_positions(), rs.get_positions(), atol=3e-5)
    np.testing.assert_allclose(os_.get_velocities(), rs.get_velocities(), atol=3e-4)
    assert osim.integrator.update_count == rsim.integrator.update_count        # force evaluations counted as the reference does
    # the Python-visible topology of the potential after the run = the list of the reference's last evaluation
    rn, on = rsim.integrator.model.nbr_list, osim.integrator.model.nbr_list
    assert on.shape == rn.shape and torch.equal(on.cpu(), rn.cpu()), (on.shape, rn.shape)
    # parameter gradients through the adjoint solver (analytic second-order route for these potentials)
    assert set(ograds) == set(rgrads)
    for name in rgrads:
        assert np.abs(ograds[name] - rgrads[name]).max() <= 2e-3 * max(np.abs(rgrads[name]).max(), 1e-6), (name, ograds[name], rgrads[name])


def _fit_case(ns, which, seed):
    """short differentiable runs whose reverse sweep goes through autograd (no closed-form route): learned pair potential,
    Stack of pair members; loss = RDF-based + state terms; returns outputs and all parameter gradients"""
    from mdgrad_b200._ase_compat import FaceCenteredCubic
    atoms = FaceCenteredCubic(symbol="H", size=(3, 3, 3), latticeconstant=1.679, pbc=True)
    rng = np.random.default_rng(seed)
    atoms.set_positions(atoms.get_positions() + rng.normal(0, 0.04, (108, 3)))
    system = ns.system.System(atoms, device="cpu")
    system.set_velocities(rng.standard_normal((108, 3)) * 0.8)
    torch.manual_seed(seed)
    if which == "mlp":
        net = ns.potentials.pairMLP(n_gauss=10, r_start=0.0, r_end=2.5, n_layers=1, n_width=12, nonlinear="ELU")
        prior = ns.potentials.ExcludedVolume(1.0, 1.0, 12)
        model = ns.interface.Stack({"mlp": ns.interface.PairPotentials(system, net, cutoff=2.5),
                                    "prior": ns.interface.PairPotentials(system, prior, cutoff=2.5)})
        mods = [net, prior]
    else:
        a = ns.potentials.LennardJones(1.0, 0.6)
        b = ns.potentials.ExcludedVolume(0.9, 0.4, 12)
        A, B = list(range(0, 108, 2)), list(range(1, 108, 2))
        model = ns.interface.Stack({"aa": ns.interface.PairPotentials(system, a, cutoff=2.5, index_tuple=(A, A)),
                                    "ab": ns.interface.PairPotentials(system, b, cutoff=2.0, index_tuple=(A, B))})
        mods = [a, b]
    integ = ns.md.NoseHooverChain(model, system, T=1.0, num_chains=3, Q=50.0, adjoint=True)
    sim = ns.md.Simulations(system, integ, wrap=True, method="NH_verlet")
    v, q, pv = sim.simulate(steps=5, frequency=5, dt=0.005)
    obs = ns.observable.rdf(system, 40, (0.8, 2.0))
    _, _, g = obs(q[-1:])
    loss = (g ** 2).sum() + (v[-1] ** 2).sum() + pv[-1].sum()
    loss.backward()
    grads = []
    for m in mods:
        grads += [p.grad.detach().numpy().copy() for p in m.parameters()]
    return [x.detach().numpy() for x in (v, q, pv, g)], loss.item(), grads


@pytest.mark.parametrize("which", ["mlp", "stack"])
def test_emu_fit_flows_vs_live_reference(which):
    import types
    import torchmd
    with ref_import.active() as ref:
        rout, rloss, rgrads = _fit_case(ref, which, seed=7)
    ours = types.SimpleNamespace(system=torchmd.system, interface=torchmd.interface, potentials=torchmd.potentials, md=torchmd.md,
                                 observable=torchmd.observable)
    oout, oloss, ograds = _fit_case(ours, which, seed=7)
    for a, b in zip(oout, rout):
        assert a.shape == b.shape and np.abs(a - b).max() <= 5e-5 * max(1.0, np.abs(b).max())
    assert abs(oloss - rloss) <= 1e-4 * abs(rloss)
    assert len(ograds) == len(rgrads) and len(rgrads) >= 4
    gmax = max(np.abs(g).max() for g in rgrads)
    for a, b in zip(ograds, rgrads):
        assert a.shape == b.shape
        assert np.abs(a - b).max() <= 2e-3 * max(np.abs(b).max(), 1e-3 * gmax), (np.abs(a - b).max(), np.abs(b).max())


def _optimisation_loop(ns, seed, iters=3):
    """the loop of scripts/fit_rdf_pair.py in miniature: simulate -> RDF loss -> backward -> optimiser step, the simulation
    CONTINUING from the log between iterations while the parameters change in place"""
    from mdgrad_b200._ase_compat import FaceCenteredCubic
    atoms = FaceCenteredCubic(symbol="H", size=(3, 3, 3), latticeconstant=1.679, pbc=True)
    rng = np.random.default_rng(seed)
    system = ns.system.System(atoms, device="cpu")
    system.set_velocities(rng.standard_normal((108, 3)) * 0.8)
    lj = ns.potentials.LennardJones(0.95, 1.1)
    pair = ns.interface.PairPotentials(system, lj, cutoff=2.5)
    integ = ns.md.NoseHooverChain(pair, system, T=1.0, num_chains=5, Q=50.0, adjoint=True)
    sim = ns.md.Simulations(system, integ)
    obs = ns.observable.rdf(system, 50, (0.8, 2.0))
    target = torch.linspace(0.0, 1.5, 50)
    opt = torch.optim.Adam(list(lj.parameters()), lr=0.01)
    hist = []
    for _ in range(iters):
        v, q, pv = sim.simulate(8, dt=0.01, frequency=8)
        _, _, g = obs(q[-2:])
        loss = (g - target).pow(2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        hist.append((loss.item(), lj.sigma.item(), lj.epsilon.item()))
    return np.array(hist), system.get_positions()


def test_emu_optimisation_loop_vs_live_reference():
    import types
    import torchmd
    with ref_import.active() as ref:
        rh, rq = _optimisation_loop(ref, 11)
    ours = types.SimpleNamespace(system=torchmd.system, interface=torchmd.interface, potentials=torchmd.potentials, md=torchmd.md,
                                 observable=torchmd.observable)
    oh, oq = _optimisation_loop(ours, 11)
    np.testing.assert_allclose(oh, rh, rtol=2e-4)                 # losses and the parameter values after every Adam step
    np.testing.assert_allclose(oq, rq, atol=1e-4)


def _gnn_optimisation_loop(ns, schnet_cls, seed, iters=3):
    """demo/fit_rdf_gnn.py in miniature: SchNet + ExcludedVolume Stack, epochs continuing from the log, Adam steps on the
    SchNet weights between them (the native weight cache of the engine must follow the in-place updates)"""
    from mdgrad_b200._ase_compat import Diamond, units
    atoms = Diamond("Si", (2, 2, 2), 5.45933)
    rng = np.random.default_rng(seed)
    atoms.set_positions(atoms.get_positions() + rng.normal(0, 0.08, (len(atoms), 3)))
    system = ns.system.System(atoms, device="cpu")
    system.set_velocities(rng.standard_normal((len(atoms), 3)) * 0.02)
    torch.manual_seed(seed)
    model = schnet_cls({"n_atom_basis": 16, "n_filters": 16, "n_gaussians": 10, "n_convolutions": 2, "cutoff": 4.9,
                        "trainable_gauss": False})
    gnn = ns.interface.GNNPotentials(system, model, cutoff=4.9)
    prior = ns.interface.PairPotentials(system, ns.potentials.ExcludedVolume(1.9, 0.015, 12), cutoff=4.9)
    integ = ns.md.NoseHooverChain(ns.interface.Stack({"gnn": gnn, "prior": prior}), system, Q=50.0, T=600.0 * units.kB,
                                  num_chains=5, adjoint=True)
    sim = ns.md.Simulations(system, integ)
    obs = ns.observable.rdf(system, 30, (1.8, 4.9))
    target = torch.linspace(0.0, 2.0, 30)
    opt = torch.optim.Adam(list(model.parameters()), lr=3e-3)
    hist = []
    for _ in range(iters):
        v, q, pv = sim.simulate(5, dt=1.0 * units.fs, frequency=5)
        _, _, g = obs(q[-1:])
        loss = (g - target).pow(2).mean() + 10.0 * (v[-1] ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        hist.append(loss.item())
    w = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy()
    return np.array(hist), w, system.get_positions()


def test_emu_gnn_optimisation_loop_vs_live_reference():
    import types
    import torchmd
    from nff.nn.models.schnet import SchNet
    with ref_import.active() as ref:
        rh, rw, rq = _gnn_optimisation_loop(ref, ref.schnet.SchNet, 4)
    ours = types.SimpleNamespace(system=torchmd.system, interface=torchmd.interface, potentials=torchmd.potentials, md=torchmd.md,
                                 observable=torchmd.observable)
    oh, ow, oq = _gnn_optimisation_loop(ours, SchNet, 4)
    np.testing.assert_allclose(oh, rh, rtol=5e-4)                 # loss of every iteration
    assert np.abs(ow - rw).max() <= 2e-4 * np.abs(rw).max()       # all SchNet weights after the Adam steps
    np.testing.assert_allclose(oq, rq, atol=2e-4)


def test_emu_topology_and_observable_calls_vs_live_reference():
    """call-level differential of the remaining public functions against the reference: generate_nbr_list for one frame and for
    a stack of frames (get_dis on/off, (3,) and (3,3) cells, masks), compute_dis, rdf over several frames with explicit width,
    vacf, Temperature"""
    import torchmd
    from mdgrad_b200._ase_compat import FaceCenteredCubic
    rng = np.random.default_rng(2)
    cell3 = torch.tensor([7.1, 7.9, 6.6])
    xyz = torch.tensor(rng.uniform(-1.0, 8.0, (150, 3)), dtype=torch.float32)
    frames = torch.tensor(rng.uniform(0.0, 6.5, (3, 90, 3)), dtype=torch.float32)
    A, B = list(range(0, 150, 3)), list(range(1, 150, 2))
    ex = torch.LongTensor(rng.integers(0, 150, (30, 2)))
    with ref_import.active() as ref:
        r1 = ref.topology.generate_nbr_list(xyz, 2.6, cell3, get_dis=True)
        r2 = ref.topology.generate_nbr_list(xyz, 2.6, torch.diag(cell3), index_tuple=(A, B), ex_pairs=ex)
        r3 = ref.topology.generate_nbr_list(frames, 2.2, cell3)
        rd = ref.topology.compute_dis(xyz, r1[0], r1[2], torch.diag(cell3))
        atoms = FaceCenteredCubic(symbol="H", size=(3, 3, 3), latticeconstant=1.679, pbc=True)
        rsys = ref.system.System(atoms, device="cpu")
        traj = torch.tensor(rsys.get_positions()[None] + rng.normal(0, 0.05, (4, 108, 3)), dtype=torch.float32)
        vel = torch.tensor(rng.standard_normal((20, 108, 3)), dtype=torch.float32)
        rg = ref.observable.rdf(rsys, 60, (0.8, 2.0), width=0.05)(traj)
        rv = ref.observable.vacf(rsys, t_range=8)(vel)
    o1 = torchmd.topology.generate_nbr_list(xyz, 2.6, cell3, get_dis=True)
    o2 = torchmd.topology.generate_nbr_list(xyz, 2.6, torch.diag(cell3), index_tuple=(A, B), ex_pairs=ex)
    o3 = torchmd.topology.generate_nbr_list(frames, 2.2, cell3)
    od = torchmd.topology.compute_dis(xyz, o1[0], o1[2], torch.diag(cell3))
    for a, b in ((o1[0], r1[0]), (o1[2], r1[2]), (o2[0], r2[0]), (o2[1], r2[1]), (o3[0], r3[0])):
        assert a.dtype == b.dtype and torch.equal(a.cpu(), b), (a.shape, b.shape)            # indices / offsets: bit-exact
    torch.testing.assert_close(o1[1], r1[1], rtol=2e-6, atol=0)
    torch.testing.assert_close(od, rd, rtol=2e-6, atol=1e-6)
    osys = torchmd.system.System(FaceCenteredCubic(symbol="H", size=(3, 3, 3), latticeconstant=1.679, pbc=True), device="cpu")
    og = torchmd.observable.rdf(osys, 60, (0.8, 2.0), width=0.05)(traj)
    for a, b in zip(og, rg):
        torch.testing.assert_close(torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float(), rtol=2e-5, atol=2e-5 * float(torch.as_tensor(b).abs().max()))
    ov = torchmd.observable.vacf(osys, t_range=8)(vel)
    torch.testing.assert_close(ov, rv, rtol=1e-6, atol=1e-6)


def _fold_optimisation_loop(ns, schnet_cls, seed, iters=2):
    """demo/fold.py in miniature: Stack{gnn: SchNet, prior: BondPotentials, pair: ExcludedVolume with the bonds excluded}, a loss on
    bond lengths and dihedral cosines of the last frame (observable.compute_dihe), adjoint backward, Adam steps"""
    sys_mod = __import__("sys")
    sys_mod.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__file__), "..", "oracle"))
    from make_golden import chain_system
    atoms, bond_top, angle_top = chain_system(n_beads=16, L=8.0, seed=9)
    n = len(atoms)
    system = ns.system.System(atoms, device="cpu")
    rng = np.random.default_rng(seed)
    system.set_velocities(rng.standard_normal((n, 3)) * 0.3)
    torch.manual_seed(seed)
    model = schnet_cls({"n_atom_basis": 16, "n_filters": 16, "n_gaussians": 8, "n_convolutions": 2, "cutoff": 2.5, "trainable_gauss": False})
    bt = torch.LongTensor(bond_top)
    gnn = ns.interface.GNNPotentials(system, model, cutoff=2.5)
    bond = ns.interface.BondPotentials(system, bt, 3.0, 1.3)
    pair = ns.interface.PairPotentials(system, ns.potentials.ExcludedVolume(1.0, 0.8, 10), cutoff=2.5, ex_pairs=bt)
    ff = ns.interface.Stack({"gnn": gnn, "prior": bond, "pair": pair})
    integ = ns.md.NoseHooverChain(ff, system, Q=50.0, T=0.6, num_chains=5, adjoint=True)
    sim = ns.md.Simulations(system, integ)
    dihes = torch.LongTensor([[i, i + 1, i + 2, i + 3] for i in range(n - 3)])
    opt = torch.optim.Adam(list(model.parameters()), lr=2e-3)
    hist = []
    for _ in range(iters):
        v, q, pv = sim.simulate(6, dt=0.002, frequency=6)
        cosphi = ns.observable.compute_dihe(q[-2:], dihes)
        blen = (q[-1][bt[:, 0]] - q[-1][bt[:, 1]]).pow(2).sum(-1)
        loss = (cosphi - 0.3).pow(2).mean() + (blen - 1.2).pow(2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        hist.append(loss.item())
    w = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy()
    return np.array(hist), w, system.get_positions()


def test_emu_fold_optimisation_loop_vs_live_reference():
    import types
    import torchmd
    from nff.nn.models.schnet import SchNet
    with ref_import.active() as ref:
        rh, rw, rq = _fold_optimisation_loop(ref, ref.schnet.SchNet, 3)
    ours = types.SimpleNamespace(system=torchmd.system, interface=torchmd.interface, potentials=torchmd.potentials, md=torchmd.md,
                                 observable=torchmd.observable)
    oh, ow, oq = _fold_optimisation_loop(ours, SchNet, 3)
    np.testing.assert_allclose(oh, rh, rtol=5e-4)
    assert np.abs(ow - rw).max() <= 2e-4 * np.abs(rw).max()
    np.testing.assert_allclose(oq, rq, atol=2e-4)


def test_emu_gnn_potentials_orthorhombic_and_exclusions_vs_live_reference():
    """GNNPotentials on an orthorhombic box with atoms outside the cell and an exclusion list: the reference's raw-offset quirk
    (integer offsets subtracted without the cell, schnet.py:122-127) only cancels on unit cells - energies and forces through
    the native program and through the op-level route must still equal the reference's"""
    import torchmd
    from nff.nn.models.schnet import SchNet
    from mdgrad_b200._ase_compat import Atoms
    rng = np.random.default_rng(12)
    cell = np.array([9.3, 10.4, 8.8])
    n = 70
    pos = rng.uniform(-1.0, 1.0, (n, 3)) + rng.uniform(0, 1, (n, 3)) * cell
    numbers = rng.integers(1, 9, n)
    ex = torch.LongTensor(rng.integers(0, n, (25, 2)))
    params = {"n_atom_basis": 20, "n_filters": 24, "n_gaussians": 11, "n_convolutions": 2, "cutoff": 3.6, "trainable_gauss": False}
    out = {}
    with ref_import.active() as ref:
        torch.manual_seed(2)
        rsys = ref.system.System(Atoms(numbers=numbers, positions=pos, cell=cell, pbc=True), device="cpu")
        rmodel = ref.schnet.SchNet(params)
        rg = ref.interface.GNNPotentials(rsys, rmodel, cutoff=3.6, ex_pairs=ex)
        x = torch.Tensor(rsys.get_positions()).requires_grad_(True)
        e = rg(x)
        out["ref"] = (e.detach().reshape(-1), -torch.autograd.grad(e.sum(), x)[0], rg.inputs["nbr_list"].clone(), rg.inputs["offsets"].clone())
        sd = {k: v.clone() for k, v in rmodel.state_dict().items()}
    osys = torchmd.system.System(Atoms(numbers=numbers, positions=pos, cell=cell, pbc=True), device="cpu")
    omodel = SchNet(params)
    omodel.load_state_dict(sd)
    og = torchmd.interface.GNNPotentials(osys, omodel, cutoff=3.6, ex_pairs=ex)
    assert torch.equal(og.inputs["nbr_list"].cpu(), out["ref"][2]) and torch.equal(og.inputs["offsets"].cpu(), out["ref"][3])
    x = torch.Tensor(osys.get_positions()).requires_grad_(True)
    e = og(x)                                                           # op-level route
    f = -torch.autograd.grad(e.sum(), x)[0]
    en, fn = og.native_energy_force(torch.Tensor(osys.get_positions()))    # native program
    for ee, ff in ((e.detach().reshape(-1), f), (en.reshape(-1), fn)):
        assert (ee - out["ref"][0]).abs().max() <= 1e-5 * out["ref"][0].abs().max()
        assert (ff - out["ref"][1]).abs().max() <= 2e-5 * out["ref"][1].abs().max()


def test_emu_water_fit_loop_vs_live_reference():
    """BASELINE configs[2]'s flow (scripts/run_water.py in miniature): water box, SchNet + O-O ExcludedVolume Stack, NHC epochs, a loss on
    RDF(O-O) through `index_tuple`, adjoint backward, Adam on the SchNet weights - against the reference, same seed"""
    import types
    import torchmd
    from nff.nn.models.schnet import SchNet
    from mdgrad_b200._ase_compat import Atoms, units

    def loop(ns, schnet_cls):
        g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "schnet_water.npz"))
        keep = np.arange(96)                                   # 32 molecules of the 64 (keeps the reference's O(N^2) lists quick)
        atoms = Atoms(numbers=g["numbers"][keep], positions=g["positions"][keep], cell=g["cell"], pbc=True)
        system = ns.system.System(atoms, device="cpu")
        np.random.seed(3)
        system.set_temperature(298.0 * units.kB)
        torch.manual_seed(5)
        model = schnet_cls({"n_atom_basis": 24, "n_filters": 24, "n_gaussians": 12, "n_convolutions": 2, "cutoff": 5.0,
                            "trainable_gauss": False})
        oxy = [int(i) for i in np.nonzero(g["numbers"][keep] == 8)[0]]
        gnn = ns.interface.GNNPotentials(system, model, cutoff=5.0)
        prior = ns.interface.PairPotentials(system, ns.potentials.ExcludedVolume(2.6, 0.015, 12), cutoff=5.0, index_tuple=(oxy, oxy))
        integ = ns.md.NoseHooverChain(ns.interface.Stack({"gnn": gnn, "prior": prior}), system, Q=50.0, T=298.0 * units.kB,
                                      num_chains=5, adjoint=True)
        sim = ns.md.Simulations(system, integ)
        obs = ns.observable.rdf(system, 40, (1.8, 5.5), index_tuple=(oxy, oxy))
        target = torch.linspace(0.0, 1.5, 40)
        opt = torch.optim.Adam(list(model.parameters()), lr=2e-3)
        hist = []
        for _ in range(2):
            v, q, pv = sim.simulate(5, dt=0.5 * units.fs, frequency=5)
            _, _, gr = obs(q[-1:])
            loss = (gr - target).pow(2).mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
            hist.append(loss.item())
        return np.array(hist), torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy(), system.get_positions()

    with ref_import.active() as ref:
        rh, rw, rq = loop(ref, ref.schnet.SchNet)
    ours = types.SimpleNamespace(system=torchmd.system, interface=torchmd.interface, potentials=torchmd.potentials, md=torchmd.md,
                                 observable=torchmd.observable)
    oh, ow, oq = loop(ours, SchNet)
    np.testing.assert_allclose(oh, rh, rtol=5e-4)
    assert np.abs(ow - rw).max() <= 2e-4 * np.abs(rw).max()
    np.testing.assert_allclose(oq, rq, atol=2e-4)
