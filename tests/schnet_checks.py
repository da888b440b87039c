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


