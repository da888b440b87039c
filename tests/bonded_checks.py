"""Bodies of the bonded / electrostatic member checks, shared by the emulated (CPU, tests/test_emu_bonded.py)
and the GPU (tests/test_gpu_zz_bonded.py) suites.  Everything is compared with tests/golden/bonded_chain.npz, written by
oracle/make_golden.py --bonded from the unmodified reference (torchmd/interface.py:303-361, :406-510).
Tolerance: 1e-5 relative to the largest component (fp32), as for the pair terms."""
import os

import numpy as np
import torch

G = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-5


def _system(dev):
    from torchmd.system import System
    from mdgrad_b200._ase_compat import Atoms
    g = np.load(os.path.join(G, "bonded_chain.npz"))
    atoms = Atoms(numbers=[1] * len(g["positions"]), positions=g["positions"], cell=g["cell"], pbc=True)
    return System(atoms, device=dev), g


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert np.abs(a - b).max() <= rtol * max(np.abs(b).max(), 1e-30), (np.abs(a - b).max(), np.abs(b).max())


def check_bond_angle_energy_force(dev):
    from torchmd.interface import AnglePotentials, BondPotentials
    system, g = _system(dev)
    kb, ro, ka, th0 = [float(x) for x in g["params"]]
    bond = BondPotentials(system, torch.LongTensor(g["bond_top"]), kb, ro)
    angle = AnglePotentials(system, torch.LongTensor(g["angle_top"]), ka, th0)
    raw = torch.tensor(g["positions"], dtype=torch.float32)
    L = torch.tensor(g["cell"], dtype=torch.float32)
    from mdgrad_b200._ase_compat import wrap_positions
    wrap = torch.tensor(wrap_positions(g["positions"], np.diag(g["cell"])), dtype=torch.float32)
    for tag, x in (("raw", raw), ("wrap", wrap)):
        for name, mod in (("bond", bond), ("angle", angle)):
            q = x.to(dev).clone().requires_grad_(True)
            e = mod(q)
            assert e.dim() == 0
            f = -torch.autograd.grad(e, q)[0]
            _close(e.item(), g["e_%s_%s" % (name, tag)])
            _close(f.cpu().numpy(), g["f_%s_%s" % (name, tag)])
            # native force route (what the no-grad solver path and the device engine use)
            _close(mod.native_force(x.to(dev)).cpu().numpy(), g["f_%s_%s" % (name, tag)])
            # pure-torch restatement (double-backward route) agrees
            _close(mod.forward_torch(x.to(dev)).item(), g["e_%s_%s" % (name, tag)])
    assert L.shape == (3,)


def check_bonded_param_grads(dev):
    from torchmd.interface import AnglePotentials, BondPotentials
    from mdgrad_b200._ase_compat import wrap_positions
    system, g = _system(dev)
    kb, ro, ka, th0 = [float(x) for x in g["params"]]
    x = torch.tensor(wrap_positions(g["positions"], np.diag(g["cell"])), dtype=torch.float32).to(dev)
    kt, rt = torch.tensor(kb, requires_grad=True, device=dev), torch.tensor(ro, requires_grad=True, device=dev)
    e = BondPotentials(system, torch.LongTensor(g["bond_top"]), kt, rt)(x)
    _close([t.item() for t in torch.autograd.grad(e, [kt, rt])], g["dp_bond"])
    kt, tt = torch.tensor(ka, requires_grad=True, device=dev), torch.tensor(th0, requires_grad=True, device=dev)
    e = AnglePotentials(system, torch.LongTensor(g["angle_top"]), kt, tt)(x)
    _close([t.item() for t in torch.autograd.grad(e, [kt, tt])], g["dp_angle"])


def check_bonded_second_order_route(dev):
    """double backward goes through the pure-torch restatement (adjoint reverse sweep), same numbers"""
    from torchmd.interface import BondPotentials
    system, g = _system(dev)
    kb, ro = float(g["params"][0]), float(g["params"][1])
    bond = BondPotentials(system, torch.LongTensor(g["bond_top"]), kb, ro)
    bond.second_order = True
    q = torch.tensor(g["positions"], dtype=torch.float32).to(dev).requires_grad_(True)
    f = -torch.autograd.grad(bond(q), q, create_graph=True)[0]
    _close(f.detach().cpu().numpy(), g["f_bond_raw"])
    h = torch.autograd.grad((f ** 2).sum(), q)[0]
    assert torch.isfinite(h).all() and float(h.abs().max()) > 0


def check_electrostatics(dev):
    from torchmd.interface import Electrostatics
    from mdgrad_b200._ase_compat import wrap_positions
    system, g = _system(dev)
    x = torch.tensor(wrap_positions(g["positions"], np.diag(g["cell"])), dtype=torch.float32).to(dev)
    es = Electrostatics(torch.tensor(g["charges"]), g["cell"], device=dev, cutoff=2.5, ex_pairs=torch.LongTensor(g["bond_top"]))
    _close(es.conversion, g["coul_conversion"], 1e-12)
    q = x.clone().requires_grad_(True)
    e = es(q)
    _close(e.item(), g["e_coul"])
    _close((-torch.autograd.grad(e, q)[0]).cpu().numpy(), g["f_coul"])
    # the physical product differs from the reference's q_j^2 (and is what a user should ask for)
    ep = Electrostatics(torch.tensor(g["charges"]), g["cell"], device=dev, cutoff=2.5, ex_pairs=torch.LongTensor(g["bond_top"]),
                        charge_product="physical")(x)
    assert abs(ep.item() - e.item()) > 1e-3 * abs(e.item())


def _stack_sim(dev, engine):
    from torchmd.interface import BondPotentials, PairPotentials, Stack
    from torchmd.potentials import ExcludedVolume
    from torchmd.md import NoseHooverChain, Simulations
    system, g = _system(dev)
    system.set_positions(g["q0"])
    system.set_velocities(g["v0"])
    kb, ro = float(g["params"][0]), float(g["params"][1])
    bond = BondPotentials(system, torch.LongTensor(g["bond_top"]), kb, ro)
    pair = PairPotentials(system, ExcludedVolume(1.0, 0.8, 10), cutoff=2.5, ex_pairs=torch.LongTensor(g["bond_top"])).to(dev)
    ff = Stack({"prior": bond, "pair": pair})
    integ = NoseHooverChain(ff, system, Q=50.0, T=0.6, num_chains=5, adjoint=True).to(dev)
    integ.disable_gnn_engine = not engine
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    v, q, pv = sim.simulate(steps=30, frequency=30, dt=0.002)
    return integ, v.detach().cpu().numpy(), q.detach().cpu().numpy(), pv.detach().cpu().numpy(), g


def check_fold_stack_on_device_engine(dev):
    """Stack{BondPotentials, PairPotentials(ex_pairs=bonds)} NoseHooverChain epoch (the force field of demo/fold.py:130-160
    without its GNN member) on the device engine, against the reference trajectory"""
    integ, v, q, pv, g = _stack_sim(dev, engine=True)
    assert integ.last_engine_stats is not None, "the epoch did not run on the device engine"
    assert np.array_equal(q[0], g["traj_q"][0].astype(np.float32))
    for k, tol in ((1, 3e-6), (10, 5e-5), (29, 5e-4)):
        assert np.abs(q[k] - g["traj_q"][k]).max() < tol, (k, np.abs(q[k] - g["traj_q"][k]).max())
        assert np.abs(v[k] - g["traj_v"][k]).max() < 20 * tol
        assert np.abs(pv[k] - g["traj_pv"][k]).max() < 50 * tol
    # and the op-level solver (bonded autograd Function + pair kernel) gives the same epoch
    integ2, v2, q2, pv2, _ = _stack_sim(dev, engine=False)
    assert np.abs(q2 - q).max() < 2e-5 and np.abs(v2 - v).max() < 2e-4


def _fold_sim(dev, engine):
    from torchmd.interface import BondPotentials, GNNPotentials, PairPotentials, Stack
    from torchmd.potentials import ExcludedVolume
    from torchmd.md import NoseHooverChain, Simulations
    from nff.nn.models.schnet import SchNet
    system, g = _system(dev)
    system.set_positions(g["q0"])
    system.set_velocities(g["v0"])
    params = {k[5:]: (float(g[k]) if k == "gnnp_cutoff" else int(g[k])) for k in g.files if k.startswith("gnnp_")}
    params["trainable_gauss"] = False
    model = SchNet(params)
    model.load_state_dict({k[2:]: torch.tensor(g[k]) for k in g.files if k.startswith("w_")})
    model = model.to(dev)
    kb, ro = float(g["params"][0]), float(g["params"][1])
    gnn = GNNPotentials(system, model, cutoff=params["cutoff"])
    bond = BondPotentials(system, torch.LongTensor(g["bond_top"]), kb, ro)
    pair = PairPotentials(system, ExcludedVolume(1.0, 0.8, 10), cutoff=2.5, ex_pairs=torch.LongTensor(g["bond_top"])).to(dev)
    integ = NoseHooverChain(Stack({"gnn": gnn, "prior": bond, "pair": pair}), system, Q=50.0, T=0.6, num_chains=5, adjoint=True).to(dev)
    integ.disable_gnn_engine = not engine
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    v, q, pv = sim.simulate(steps=12, frequency=12, dt=0.002)
    return integ, v.detach().cpu().numpy(), q.detach().cpu().numpy(), pv.detach().cpu().numpy(), g


def check_fold_force_field_with_gnn(dev):
    """the whole force field of demo/fold.py:130-160 - Stack{gnn: GNNPotentials(SchNet), prior: BondPotentials,
    pair: PairPotentials(ExcludedVolume, ex_pairs=bonds)} - as one device-engine epoch, against the reference trajectory"""
    integ, v, q, pv, g = _fold_sim(dev, engine=True)
    assert integ.last_engine_stats is not None, "the epoch did not run on the device engine"
    for k, tol in ((1, 3e-6), (11, 1e-4)):
        assert np.abs(q[k] - g["fold_q"][k]).max() < tol, (k, np.abs(q[k] - g["fold_q"][k]).max())
        assert np.abs(v[k] - g["fold_v"][k]).max() < 20 * tol
        assert np.abs(pv[k] - g["fold_pv"][k]).max() < 50 * tol


def check_pair_tab_through_pair_potentials(dev):
    """PairPotentials(system, pairTab(...)): tabulated u(r) on the native list + distance op (generic route), energy and
    forces and d/dtab against the oracle's list / distances with the same table"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle_torch as O
    from torchmd.interface import PairPotentials
    from torchmd.potentials import pairTab
    from mdgrad_b200._ase_compat import wrap_positions
    system, g = _system(dev)
    tab = pairTab(nbins=64, rc=2.5, device=dev)
    with torch.no_grad():
        tab.tab.copy_(4.0 * ((1.0 / (tab.x + 0.6)) ** 8 - (1.0 / (tab.x + 0.6)) ** 4))
    pair = PairPotentials(system, tab, cutoff=2.5)
    xw = torch.tensor(wrap_positions(g["positions"], np.diag(g["cell"])), dtype=torch.float32)
    pair._reset_topology(xw.to(dev))
    q = xw.to(dev).clone().requires_grad_(True)
    e = pair(q)
    gq, gt = torch.autograd.grad(e, [q, tab.tab])
    # oracle: reference-order list and distances on the CPU, same spline module on the CPU
    cell = torch.tensor(g["cell"], dtype=torch.float32)
    nbr, off = O.neighbor_list(xw, 2.5, cell)
    assert torch.equal(pair.nbr_list.cpu(), nbr)
    tab_c = pairTab(nbins=64, rc=2.5)
    with torch.no_grad():
        tab_c.tab.copy_(tab.tab.detach().cpu())
    qc = xw.clone().requires_grad_(True)
    ec = tab_c(O.pair_distance(qc, nbr, off, torch.diag(cell))).sum()
    gqc, gtc = torch.autograd.grad(ec, [qc, tab_c.tab])
    _close(e.item(), ec.item())
    _close(gq.cpu().numpy(), gqc.numpy())
    _close(gt.cpu().numpy(), gtc.numpy())


def check_fold_force_field_graph_replay(dev):
    """the same epoch with the force evaluations (SchNet + bonds + excluded volume: ~50 launches) replayed as a CUDA graph
    (the default; MDG_GNN_GRAPH=0 = plain asynchronous launches): bit-identical to the plain asynchronous epoch"""
    old = os.environ.pop("MDG_GNN_GRAPH", None)
    try:
        os.environ["MDG_GNN_GRAPH"] = "0"
        integ, v, q, pv, g = _fold_sim(dev, engine=True)
        assert integ.last_engine_stats["maxrow_or_K"] == 1
        os.environ["MDG_GNN_GRAPH"] = "1"
        integ2, v2, q2, pv2, _ = _fold_sim(dev, engine=True)
        assert integ2.last_engine_stats["maxrow_or_K"] == 2, "the force evaluations were not replayed from a graph"
        assert np.array_equal(q, q2) and np.array_equal(v, v2) and np.array_equal(pv, pv2)
    finally:
        os.environ.pop("MDG_GNN_GRAPH", None)
        if old is not None:
            os.environ["MDG_GNN_GRAPH"] = old


def check_fold_engine_sync_equals_async(dev):
    """the epoch with a pair-count read-back per list build (MDG_GNN_SYNC=1) and the default asynchronous epoch: same bits"""
    old = os.environ.pop("MDG_GNN_SYNC", None)
    try:
        integ, v, q, pv, g = _fold_sim(dev, engine=True)
        assert integ.last_engine_stats["maxrow_or_K"] in (1, 2), "the default epoch did not complete on the asynchronous path"
        os.environ["MDG_GNN_SYNC"] = "1"
        integ2, v2, q2, pv2, _ = _fold_sim(dev, engine=True)
        assert integ2.last_engine_stats["maxrow_or_K"] == 0
        assert np.array_equal(q, q2) and np.array_equal(v, v2) and np.array_equal(pv, pv2)
    finally:
        os.environ.pop("MDG_GNN_SYNC", None)
        if old is not None:
            os.environ["MDG_GNN_SYNC"] = old


def check_generic_route_configs(dev):
    """The configurations that stay on the op-level solver: stale lists (topology_update_freq = 3: the list is rebuilt at every
    third EVALUATION, reference md.py:200-204), method='rk4', and adjoint=False (the whole trajectory on the autograd tape,
    .backward() differentiates the forces again); trajectories and parameter gradients against the reference
    (tests/golden/c1_generic.npz, oracle/make_golden.py --generic)"""
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones
    from torchmd.md import NoseHooverChain, Simulations
    from torchmd.system import System
    from mdgrad_b200._ase_compat import FaceCenteredCubic
    g = np.load(os.path.join(G, "c1_generic.npz"))
    for tag, kw, method, steps, dt in (("freq3", dict(topology_update_freq=3, adjoint=True), "NH_verlet", 13, 0.01),
                                       ("rk4", dict(topology_update_freq=1, adjoint=True), "rk4", 9, 0.005),
                                       ("tape", dict(topology_update_freq=1, adjoint=False), "NH_verlet", 7, 0.01)):
        system = System(FaceCenteredCubic(symbol="H", size=(3, 3, 3), latticeconstant=1.679, pbc=True), device=dev)
        system.set_positions(g["q0"])
        system.set_velocities(g["v0"])
        lj = LennardJones(1.0, 1.0).to(dev)
        integ = NoseHooverChain(PairPotentials(system, lj, cutoff=2.5), system, T=1.0, num_chains=3, Q=50.0, **kw).to(dev)
        sim = Simulations(system, integ, wrap=True, method=method)
        v, q, pv = sim.simulate(steps=steps, frequency=steps, dt=dt)
        assert np.abs(q.detach().cpu().numpy() - g["q_" + tag]).max() < 2e-5, tag
        assert np.abs(v.detach().cpu().numpy() - g["v_" + tag]).max() < 2e-4, tag
        assert np.abs(pv.detach().cpu().numpy() - g["pv_" + tag]).max() < 2e-3, tag
        loss = (q[-1] ** 2).sum() + pv[-1].sum()
        loss.backward()
        for name, got in (("dsigma_", lj.sigma.grad), ("depsilon_", lj.epsilon.grad)):
            ref = float(g[name + tag][0])
            assert abs(got.item() - ref) <= 5e-3 * abs(ref), (tag, name, got.item(), ref)
        # number of force evaluations (and with it the list-rebuild cadence) of forward + reverse sweep, as the reference
        assert integ.update_count == int(g["update_count_" + tag]), (tag, integ.update_count, int(g["update_count_" + tag]))


def check_tpair_potentials_vs_reference_fixture(dev):
    """TPairPotentials + TpairMLP (temperature-conditioned learned pair potential, reference interface.py:139-215,
    potentials.py:208-217): the reference's state_dict loads unchanged; energy and forces through the native list and
    distance op"""
    from torchmd.interface import TPairPotentials
    from torchmd.potentials import TpairMLP
    from torchmd.system import System
    from mdgrad_b200._ase_compat import FaceCenteredCubic
    g = np.load(os.path.join(G, "c1_generic.npz"))
    system = System(FaceCenteredCubic(symbol="H", size=(3, 3, 3), latticeconstant=1.679, pbc=True), device=dev)
    net = TpairMLP(n_gauss=12, r_start=0.0, r_end=2.5, n_layers=2, n_width=16, nonlinear="ELU")
    missing = net.load_state_dict({k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("tw_")}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    tp = TPairPotentials(system, net.to(dev), T=1.3, cutoff=2.5)
    xyz = torch.tensor(g["tpair_xyz"]).to(dev)
    tp._reset_topology(xyz)
    q = xyz.clone().requires_grad_(True)
    e = tp(q)
    f = -torch.autograd.grad(e, q)[0]
    _close(e.item(), g["tpair_e"], 2e-5)
    _close(f.cpu().numpy(), g["tpair_f"], 2e-5)


def check_stack_adjoint_native_equals_autograd(dev):
    """Stack of analytic species-pair members (scripts/fit_2_comp.py's force field): the closed-form reverse dynamics
    (mdg_pair_hvp per member) and the generic double-backward route give the same parameter gradients"""
    from torchmd.interface import PairPotentials, Stack
    from torchmd.potentials import ExcludedVolume, LennardJones
    from torchmd.md import NoseHooverChain, Simulations
    from torchmd.system import System
    from mdgrad_b200._ase_compat import FaceCenteredCubic
    res = []
    for native in (True, False):
        rng = np.random.default_rng(7)
        atoms = FaceCenteredCubic(symbol="H", size=(3, 3, 3), latticeconstant=1.679, pbc=True)
        atoms.set_positions(atoms.get_positions() + rng.normal(0, 0.04, (108, 3)))
        system = System(atoms, device=dev)
        system.set_velocities(rng.standard_normal((108, 3)) * 0.8)
        a, b = LennardJones(1.0, 0.6).to(dev), ExcludedVolume(0.9, 0.4, 12).to(dev)
        A, B = list(range(0, 108, 2)), list(range(1, 108, 2))
        model = Stack({"aa": PairPotentials(system, a, cutoff=2.5, index_tuple=(A, A)),
                       "ab": PairPotentials(system, b, cutoff=2.0, index_tuple=(A, B))})
        integ = NoseHooverChain(model, system, T=1.0, num_chains=3, Q=50.0, adjoint=True).to(dev)
        integ.disable_native_adjoint = not native
        v, q, pv = Simulations(system, integ).simulate(steps=5, frequency=5, dt=0.005)
        ((q[-1] ** 2).sum() + (v[-1] * v[1]).sum() + pv[-1].sum()).backward()
        res.append([p.grad.item() for m in (a, b) for p in m.parameters() if p.grad is not None])
    assert len(res[0]) == len(res[1]) >= 4
    for x, y in zip(*res):
        assert abs(x - y) <= 5e-4 * max(1.0, abs(y)), (res[0], res[1])


def check_bonded_edge_cases(dev):
    """empty topologies (no terms: zero energy, zero forces, no launch with an empty grid), a topology naming a non-existent
    atom (rejected at construction), duplicate bonds (each listed term counts, as in the reference's sum)"""
    import pytest
    from torchmd.interface import AnglePotentials, BondPotentials
    system, g = _system(dev)
    n = len(system)
    x = torch.tensor(g["positions"], dtype=torch.float32).to(dev)
    for mod in (BondPotentials(system, torch.zeros((0, 2), dtype=torch.long), 3.0, 1.3),
                AnglePotentials(system, torch.zeros((0, 3), dtype=torch.long), 2.0, 1.9)):
        q = x.clone().requires_grad_(True)
        e = mod(q)
        assert e.item() == 0.0
        assert float(mod.native_force(x).abs().max()) == 0.0 and mod.native_force(x).shape == (n, 3)
    with pytest.raises(IndexError):
        BondPotentials(system, torch.LongTensor([[0, n]]), 3.0, 1.3)
    one = BondPotentials(system, torch.LongTensor(g["bond_top"][:5]), 3.0, 1.3)
    two = BondPotentials(system, torch.LongTensor(np.concatenate([g["bond_top"][:5], g["bond_top"][:5]])), 3.0, 1.3)
    _close(two(x).item(), 2.0 * one(x).item(), 1e-6)
    _close(two.native_force(x).cpu().numpy(), 2.0 * one.native_force(x).cpu().numpy(), 1e-6)
