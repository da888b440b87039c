"""Bodies of SchNet op-level checks shared by the emulated (CPU) and GPU suites."""
import torch


def check_second_order_through_native_aggregation(dev):
    """The adjoint solver's reverse sweep differentiates the forces once more.  The cfconv aggregation Functions are closed
    under differentiation (backward of agg = agg / edge-grad, backward of edge-grad = agg), so the second-order graph runs on
    the native kernels: gradients of a force-dependent scalar w.r.t. positions and ALL parameters against the oracle's
    pure-torch chain."""
    from nff.nn.models.schnet import SchNet
    from oracle import oracle_torch as O
    from test_schnet import _fixture
    g, params, sd = _fixture("water")
    torch.manual_seed(3)
    n = 60
    idx = torch.arange(n)
    xyz = torch.Tensor(g["positions"])[idx]
    z = torch.tensor(g["numbers"], dtype=torch.long)[idx]
    cell = torch.Tensor(g["cell"])
    nbr, off = O.neighbor_list(xyz, 4.0, cell)
    avec = torch.randn(n, 3)
    model = SchNet(params)
    model.load_state_dict(sd)
    model = model.to(dev)
    model.second_order = True
    batch = {"nxyz": torch.cat([z[:, None].float(), xyz], dim=1).to(dev), "num_atoms": torch.tensor([n]), "nbr_list": nbr.to(dev),
             "offsets": off.to(dev), "energy": torch.zeros(1, device=dev)}
    x = xyz.clone().to(dev).requires_grad_(True)
    e = model(batch, x)["energy"].sum()
    f = -torch.autograd.grad(e, x, create_graph=True)[0]
    assert batch.get("_native_graph") is not None                       # the native aggregation was on the tape
    loss = (f * avec.to(dev)).sum() + 0.1 * e
    plist = [p for p in model.parameters() if p.requires_grad]
    grads = torch.autograd.grad(loss, [x] + plist, allow_unused=True)
    # oracle: the reference op chain (gather / multiply / scatter_add), pure torch
    sd_o = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    xo = xyz.clone().requires_grad_(True)
    eo = O.schnet_energy(sd_o, z, xo, nbr, off)
    fo = -torch.autograd.grad(eo, xo, create_graph=True)[0]
    loss_o = (fo * avec).sum() + 0.1 * eo
    names = [k for k, p in model.named_parameters() if p.requires_grad]
    grads_o = torch.autograd.grad(loss_o, [xo] + [sd_o[k] for k in names], allow_unused=True)
    assert (f.detach().cpu() - fo.detach()).abs().max() <= 2e-5 * fo.detach().abs().max()
    for name, a, b in zip(["xyz"] + names, grads, grads_o):
        if b is None:
            assert a is None or float(a.abs().max()) == 0.0, name
            continue
        a = a.detach().cpu()
        assert (a - b).abs().max().item() <= 5e-4 * max(1e-6, b.abs().max().item()), (name, (a - b).abs().max().item(), b.abs().max().item())




def check_gnn_adjoint_fit_vs_reference_fixture(dev):
    """BASELINE configs[4]'s flow at small size: Stack{SchNet, ExcludedVolume}, NoseHooverChain(adjoint=True) epoch on the device
    engine, RDF-based loss on the last frame, `.backward()` through the adjoint solver - the gradient of EVERY SchNet parameter
    against the reference's (tests/golden/gnn_adjoint.npz, oracle/make_golden.py --gnn-adjoint)"""
    import os
    import numpy as np
    from nff.nn.models.schnet import SchNet
    from torchmd.interface import GNNPotentials, PairPotentials, Stack
    from torchmd.md import NoseHooverChain, Simulations
    from torchmd.observable import rdf
    from torchmd.potentials import ExcludedVolume
    from torchmd.system import System
    from mdgrad_b200._ase_compat import Atoms, units
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "gnn_adjoint.npz"))
    system = System(Atoms(numbers=g["numbers"], positions=g["q0"], cell=g["cell"], pbc=True), device=dev)
    system.set_velocities(g["v0"])
    params = {k[5:]: (float(g[k]) if k == "gnnp_cutoff" else int(g[k])) for k in g.files if k.startswith("gnnp_")}
    params["trainable_gauss"] = False
    model = SchNet(params)
    model.load_state_dict({k[2:]: torch.tensor(g[k]) for k in g.files if k.startswith("w_")})
    model = model.to(dev)
    gnn = GNNPotentials(system, model, cutoff=params["cutoff"])
    prior = PairPotentials(system, ExcludedVolume(1.9, 0.015, 12).to(dev), cutoff=4.9)
    integ = NoseHooverChain(Stack({"gnn": gnn, "prior": prior}), system, Q=50.0, T=600.0 * units.kB, num_chains=5, adjoint=True).to(dev)
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    v, q, pv = sim.simulate(steps=6, frequency=6, dt=1.0 * units.fs)
    assert integ.last_engine_stats is not None, "the forward epoch did not run on the device engine"
    assert np.abs(q.detach().cpu().numpy() - g["traj_q"]).max() < 2e-5
    assert np.abs(v.detach().cpu().numpy() - g["traj_v"]).max() < 2e-5 * max(1.0, np.abs(g["traj_v"]).max() * 1e2)
    obs = rdf(system, 30, (1.8, 4.9))
    _, bins, gr = obs(q[-1:])
    np.testing.assert_allclose(gr.detach().cpu().numpy(), g["rdf_g"], rtol=2e-4, atol=2e-4 * g["rdf_g"].max())
    loss = gr.pow(2).sum() + 1e3 * (v[-1] ** 2).sum()
    assert abs(loss.item() - float(g["loss"])) <= 2e-4 * abs(float(g["loss"]))
    loss.backward()
    checked = 0
    gnorm = float(np.sqrt(sum((g[k].astype(np.float64) ** 2).sum() for k in g.files if k.startswith("g_"))))
    for name, p in model.named_parameters():
        key = "g_" + name
        if key not in g.files:
            continue
        assert p.grad is not None, name
        err = np.abs(p.grad.detach().cpu().numpy() - g[key]).max()
        assert err <= 1e-4 * max(np.abs(g[key]).max(), 1e-3 * gnorm), (name, err, np.abs(g[key]).max())      # measured: ~3e-7
        checked += 1
    assert checked >= 20


def check_water_rdf_oo_species_selection(dev):
    """BASELINE configs[2]'s observable: RDF(O-O) of the 64-water box through `rdf(..., index_tuple=(O, O))` - the RDF kernel's
    species selection (no-grad path) and the differentiable path - against the oracle's restatement of observable.py:33-76"""
    import os
    import numpy as np
    from oracle import oracle_torch as O
    from torchmd.observable import rdf
    from torchmd.system import System
    from mdgrad_b200._ase_compat import Atoms
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "schnet_water.npz"))
    system = System(Atoms(numbers=g["numbers"], positions=g["positions"], cell=g["cell"], pbc=True), device=dev)
    oxy = [int(i) for i in np.nonzero(g["numbers"] == 8)[0]]
    hyd = [int(i) for i in np.nonzero(g["numbers"] == 1)[0]]
    xyz = torch.Tensor(system.get_positions(wrap=True))
    for sel in ((oxy, oxy), (oxy, hyd), None):
        obs = rdf(system, 100, (1.8, 5.7), index_tuple=sel)
        co, bo, go = O.rdf(xyz, [float(x) for x in g["cell"]], 100, (1.8, 5.7), index_tuple=sel)
        c, b, gr = obs(xyz.to(dev))                                     # kernel path
        torch.testing.assert_close(gr.cpu(), go, rtol=1e-5, atol=1e-5 * go.max().item())
        x = xyz.to(dev).clone().requires_grad_(True)                     # differentiable path (fitting loss)
        c2, _, gr2 = obs(x)
        torch.testing.assert_close(gr2.detach().cpu(), go, rtol=2e-5, atol=2e-5 * go.max().item())
        (gr2 ** 2).sum().backward()
        assert torch.isfinite(x.grad).all() and float(x.grad.abs().max()) > 0
