"""GPU: the drop-in Python API (torchmd.*) against the committed reference fixtures (tests/golden,
produced from the UNMODIFIED reference) and the oracle.  These read like the reference's own usage."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle_torch as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _fcc_system(device=0, size=3, a=1.679):
    from torchmd.system import System
    from mdgrad_b200._ase_compat import FaceCenteredCubic
    return System(FaceCenteredCubic(symbol="H", size=(size,) * 3, latticeconstant=a, pbc=True), device=device)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_generate_nbr_list_vs_reference_fixture(tag):
    from torchmd.topology import generate_nbr_list, compute_dis
    g = np.load(os.path.join(G, "nbr_%s.npz" % tag))
    xyz, cell = torch.tensor(g["xyz"]).cuda(), torch.tensor(g["cell"]).cuda()
    nbr, dis, off = generate_nbr_list(xyz, float(g["rc"]), cell, get_dis=True)        # reference return order
    assert np.array_equal(nbr.cpu().numpy(), g["nbr"].astype(np.int64))
    assert np.array_equal(off.cpu().numpy(), g["off"].astype(np.float32))
    np.testing.assert_allclose(dis.cpu().numpy(), g["dis"], rtol=3e-7)
    nbr_m, off_m = generate_nbr_list(xyz, float(g["rc"]), torch.diag(cell), index_tuple=(g["sel_a"], g["sel_b"]),
                                     ex_pairs=torch.tensor(g["ex"].astype(np.int64)))
    assert np.array_equal(nbr_m.cpu().numpy(), g["nbr_m"].astype(np.int64))
    assert np.array_equal(off_m.cpu().numpy(), g["off_m"].astype(np.float32))
    # compute_dis + its hand-written backward vs torch autograd on the same formula
    q = xyz.clone().requires_grad_(True)
    d = compute_dis(q, nbr, off, cell)
    assert d.shape == (nbr.shape[0], 1)
    np.testing.assert_allclose(d.detach().cpu().numpy()[:, 0], g["dis"], rtol=1e-6)
    w = torch.linspace(0.5, 1.5, d.shape[0], device="cuda")[:, None]
    (d * w).sum().backward()
    q2 = xyz.clone().requires_grad_(True)
    d2 = (q2[nbr[:, 0]] - q2[nbr[:, 1]] - off * cell).pow(2).sum(1).sqrt()[:, None]
    (d2 * w).sum().backward()
    assert (q.grad - q2.grad).abs().max().item() < 1e-4 * q2.grad.abs().max().item()
    # batched input: leading frame column
    nb2, _ = generate_nbr_list(torch.stack([xyz, xyz]), float(g["rc"]), cell)
    assert nb2.shape == (2 * nbr.shape[0], 3) and int(nb2[:, 0].max()) == 1


@pytest.mark.parametrize("name", ["lj", "ljfam", "lj69", "exv", "buck", "morse"])
def test_pair_potentials_vs_reference_fixture(name):
    from torchmd import potentials as P
    from torchmd.interface import PairPotentials
    g = np.load(os.path.join(G, "pair_fcc108.npz"))
    pots = {"lj": P.LennardJones(1.0, 1.0), "ljfam": P.LJFamily(1.0, 0.8, attr_pow=5, rep_pow=10),
            "lj69": P.LennardJones69(1.1, 0.7), "exv": P.ExcludedVolume(1.0, 0.5, 12),
            "buck": P.Buck(1000.0, 3.5, 2.0), "morse": P.ModifiedMorse(6.0, 2.0)}
    system = _fcc_system()
    pair = PairPotentials(system, pots[name].cuda() if name != "morse" else pots[name], cutoff=2.5)
    assert pair.nbr_list.device.type == "cpu" and pair.nbr_list.shape[0] == int(g["fcc_pairs"])
    xyz = torch.tensor(g["xyz"]).cuda()
    pair._reset_topology(xyz)
    q = xyz.clone().requires_grad_(True)
    e = pair(q)
    params = list(pair.model.parameters())
    grads = torch.autograd.grad(e, [q] + params, allow_unused=True)
    eref, fref = float(g["e_" + name]), g["f_" + name]
    assert abs(e.item() - eref) <= 1e-5 * abs(eref)                              # 1e-5 relative fp32 (north_star)
    assert np.abs(-grads[0].cpu().numpy() - fref).max() <= 1e-5 * np.abs(fref).max()
    for gr, ref in zip(grads[1:], g["dp_" + name]):
        assert abs(gr.item() - ref) <= 2e-5 * max(1.0, abs(ref))


def test_learned_pair_potential_through_native_distance_op():
    """pairMLP u(r) keeps its torch graph; distances + their backward are the native kernels."""
    from torchmd.potentials import pairMLP
    from torchmd.interface import PairPotentials
    torch.manual_seed(0)
    system = _fcc_system()
    mlp = pairMLP(16, 0.5, 2.5, 1, 32, "ELU").cuda()
    pair = PairPotentials(system, mlp, cutoff=2.5)
    xyz = torch.Tensor(system.get_positions()).cuda() + 0.05 * torch.randn(108, 3, device="cuda")
    pair._reset_topology(xyz)
    q = xyz.clone().requires_grad_(True)
    e = pair(q)
    f, = torch.autograd.grad(e, q)
    q2 = xyz.clone().requires_grad_(True)
    cell = pair.cell.detach()
    nbr = pair.nbr_list.cuda()
    d = (q2[nbr[:, 0]] - q2[nbr[:, 1]] - pair.offsets.matmul(cell)).pow(2).sum(1).sqrt()[:, None]
    e2 = mlp(d).sum()
    f2, = torch.autograd.grad(e2, q2)
    assert abs(e.item() - e2.item()) <= 1e-5 * abs(e2.item()) + 1e-6
    assert (f - f2).abs().max().item() <= 2e-5 * f2.abs().max().item() + 1e-7


def test_c1_simulate_vs_reference_fixture():
    """BASELINE configs[0]: 108-atom FCC LJ, NoseHooverChain, simulate(steps=50, frequency=50, dt=0.01):
    full (v, q, p_v) trajectory against the reference's own output."""
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
    assert v.requires_grad                      # adjoint=True: outputs are attached to the adjoint solve
    v, q, pv = v.detach(), q.detach(), pv.detach()
    assert integ.last_engine_stats is not None and integ.update_count == 98
    assert v.shape == (50, 108, 3) and q.shape == (50, 108, 3) and pv.shape == (50, 5)
    assert np.array_equal(v[0].cpu().numpy(), g["v"][0]) and np.array_equal(q[0].cpu().numpy(), g["q"][0])
    for k, tol in ((1, 2e-6), (10, 3e-5), (49, 3e-3)):           # chaos-limited growth of rounding differences
        assert np.abs(q[k].cpu().numpy() - g["q"][k]).max() < tol
        assert np.abs(v[k].cpu().numpy() - g["v"][k]).max() < 10 * tol
        assert np.abs(pv[k].cpu().numpy() - g["pv"][k]).max() < 30 * tol
    assert len(sim.log["positions"]) == 1 and sim.log["positions"][0].shape == (108, 3)
    # RDF observable on the reference's final frame: 1e-5 (north_star)
    obs = rdf(system, 100, (0.75, 2.0))
    count, bins, gr = obs(torch.tensor(g["q"][-1]).cuda())
    assert np.array_equal(bins.numpy(), g["rdf_bins"])
    np.testing.assert_allclose(gr.cpu().numpy(), g["rdf_g"], rtol=1e-5, atol=1e-5 * g["rdf_g"].max())
    # a second call continues from the wrapped check point
    v2, q2, pv2 = sim.simulate(steps=10, frequency=10, dt=0.01)
    assert len(sim.log["positions"]) == 2 and q2.shape == (10, 108, 3)


def test_nve_simulate_vs_reference_fixture():
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones
    from torchmd.md import NVE, Simulations
    g = np.load(os.path.join(G, "c1_nve.npz"))
    system = _fcc_system()
    system.set_positions(g["q0"])
    system.set_velocities(g["v0"])
    integ = NVE(PairPotentials(system, LennardJones(1.0, 1.0), cutoff=2.5), system, adjoint=True)
    sim = Simulations(system, integ, wrap=True, method="verlet")
    v, q = sim.simulate(steps=20, frequency=20, dt=0.005)
    v, q = v.detach(), q.detach()
    assert np.abs(q[-1].cpu().numpy() - g["q"][-1]).max() < 2e-5
    assert np.abs(v[-1].cpu().numpy() - g["v"][-1]).max() < 2e-4


def test_adjoint_gradients_vs_reference_fixture():
    """Adjoint backward through the 49 steps (generic second-order route) against the reference's
    d loss / d sigma, d loss / d epsilon."""
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones
    from torchmd.md import NoseHooverChain, Simulations
    g = np.load(os.path.join(G, "c1_traj.npz"))
    system = _fcc_system()
    system.set_positions(g["q0"])
    system.set_velocities(g["v0"])
    lj = LennardJones(1.0, 1.0).cuda()
    integ = NoseHooverChain(PairPotentials(system, lj, cutoff=2.5), system, T=1.0, num_chains=5, Q=50.0, adjoint=True)
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    v, q, pv = sim.simulate(steps=50, frequency=50, dt=0.01)
    loss = (q[-1] ** 2).sum() + (v[20] * v[30]).sum() + pv[-1].sum()
    loss.backward()
    assert abs(lj.sigma.grad.item() - g["dsigma"][0]) <= 2e-2 * abs(g["dsigma"][0])
    assert abs(lj.epsilon.grad.item() - g["depsilon"][0]) <= 2e-2 * abs(g["depsilon"][0])


@pytest.mark.parametrize("tag", ["lj", "buck"])
@pytest.mark.parametrize("route", ["native_hvp", "autograd"])
def test_adjoint_gradients_short_horizon_vs_reference_fixture(tag, route):
    """SHORT-horizon adjoint (5 NH-Verlet steps, tests/golden/c1_adjoint_short.npz from the unmodified reference): every
    parameter gradient within 1e-4 relative on both reverse routes (analytic second-order kernel mdg_pair_hvp / autograd double
    backward); chaos has not amplified rounding differences over 5 steps, so this is the tight bar the 49-step check cannot give."""
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones, Buck
    from torchmd.md import NoseHooverChain, Simulations
    g = np.load(os.path.join(G, "c1_adjoint_short.npz"))
    system = _fcc_system()
    system.set_positions(g["q0"])
    system.set_velocities(g["v0"])
    pot = (LennardJones(1.1, 0.9) if tag == "lj" else Buck(1000.0, 3.5, 2.0)).cuda()
    integ = NoseHooverChain(PairPotentials(system, pot, cutoff=2.5), system, T=1.0, num_chains=5, Q=50.0, adjoint=True)
    integ.disable_native_adjoint = route == "autograd"
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    v, q, pv = sim.simulate(steps=6, frequency=6, dt=0.01)
    assert np.abs(q.detach().cpu().numpy() - g["q_" + tag]).max() <= 2e-6 * 5.04
    assert np.abs(v.detach().cpu().numpy() - g["v_" + tag]).max() <= 2e-5 * np.abs(g["v_" + tag]).max()
    loss = (q[-1] ** 2).sum() + (v[2] * v[4]).sum() + pv[-1].sum()
    assert abs(loss.item() - float(g["loss_" + tag])) <= 1e-5 * abs(float(g["loss_" + tag]))
    loss.backward()
    scale = max(abs(float(g["d%s_%s" % (n, tag)].reshape(-1)[0])) for n, _ in pot.named_parameters())
    for name, prm in pot.named_parameters():
        ref = float(g["d%s_%s" % (name, tag)].reshape(-1)[0])
        assert prm.grad is not None, name
        assert abs(prm.grad.item() - ref) <= 1e-4 * max(abs(ref), 1e-2 * scale), (name, prm.grad.item(), ref)


def test_stack_and_masks():
    from torchmd.interface import PairPotentials, Stack
    from torchmd.potentials import LennardJones, ExcludedVolume
    system = _fcc_system(size=4)
    n = len(system)
    A, B = list(range(0, n, 2)), list(range(1, n, 2))
    ex = np.array([[0, 1], [2, 3], [10, 200]])
    p1 = PairPotentials(system, LennardJones(1.0, 1.0), cutoff=2.5, index_tuple=(A, B))
    p2 = PairPotentials(system, ExcludedVolume(0.9, 0.5, 12), cutoff=2.0, ex_pairs=torch.tensor(ex))
    st = Stack({"a": p1, "b": p2})
    xyz = torch.Tensor(system.get_positions()).cuda() + 0.03 * torch.randn(n, 3, device="cuda")
    st._reset_topology(xyz)
    e = st(xyz)
    assert e.shape == (1,)
    cell = torch.Tensor(np.diag(system.get_cell()))
    n1, o1 = O.neighbor_list(xyz.cpu(), 2.5, cell, index_tuple=(A, B))
    n2, o2 = O.neighbor_list(xyz.cpu(), 2.0, cell, ex_pairs=ex)
    assert torch.equal(p1.nbr_list, n1) and torch.equal(p2.nbr_list, n2)
    e1 = O.pair_energy_forces(xyz.cpu(), n1, o1, cell, "lj", (1.0, 1.0))[0]
    e2 = O.pair_energy_forces(xyz.cpu(), n2, o2, cell, "exv", (0.9, 0.5, 12))[0]
    assert abs(e.item() - (e1 + e2).item()) <= 1e-5 * abs((e1 + e2).item())


def test_simulate_frequency_one_integrates_zero_steps():
    """frequency=1 (the default) integrates zero steps but still logs one frame per epoch."""
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones
    from torchmd.md import NoseHooverChain, Simulations
    system = _fcc_system()
    q_before = system.get_positions().copy()
    integ = NoseHooverChain(PairPotentials(system, LennardJones(1.0, 1.0), cutoff=2.5), system, T=1.0, num_chains=5, Q=50.0)
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    v, q, pv = sim.simulate(steps=3, frequency=1, dt=0.01)
    assert v.shape == (1, 108, 3) and q.shape == (1, 108, 3) and pv.shape == (1, 5)
    assert len(sim.log["positions"]) == 3 and integ.update_count == 0
    assert np.allclose(system.get_positions(), q_before, atol=1e-5)
    with pytest.raises(UnboundLocalError):
        sim.simulate(steps=2, frequency=5, dt=0.01)
