"""CPU-emulated run of the analytic pair Hessian-vector kernel (mdg_pair_hvp) against torch double backward through
the oracle's pair energy - the quantity the reference adjoint obtains with autograd (sovlers.py:211-293)."""
import numpy as np
import pytest
import torch

from emu_lib import EmuContext
from mdgrad_b200 import _lib
from oracle import oracle_torch as O


@pytest.fixture(scope="module")
def ectx():
    return EmuContext()


CASES = [("lj", 0, (1.0, 1.0)), ("ljfam", 1, (1.05, 0.8, 10, 5)), ("lj69", 2, (1.1, 0.7)), ("exv", 3, (1.0, 0.5, 12)),
         ("ljfam", 1, (0.95, 1.2, 11.5, 5.5)), ("buck", 4, (1000.0, 3.5, 2.0)), ("morse", 5, (6.0, 2.0)), ("morse", 5, (4.0, -1.5))]


@pytest.mark.parametrize("name,kind,params", CASES)
@pytest.mark.parametrize("ncell", [3, 10])
def test_emu_pair_hvp_vs_double_backward(ectx, name, kind, params, ncell):
    pos, _, L = O.lj_system(ncell, rho=0.845, jitter=0.05, seed=3, a=1.679 if ncell == 3 else None)
    xyz = torch.tensor(pos, dtype=torch.float64)
    cell = torch.tensor([L, L, L], dtype=torch.float64)
    rc = 2.5
    nbr, off = O.neighbor_list(xyz.float(), rc, cell.float())
    g = torch.Generator().manual_seed(1)
    a = torch.randn(xyz.shape, generator=g, dtype=torch.float64)
    # oracle in fp64: F = -dE/dx with a graph, then (dF/dx)^T a and (dF/dtheta)^T a
    x = xyz.clone().requires_grad_(True)
    sig = torch.tensor(float(params[0]), dtype=torch.float64, requires_grad=True)
    eps = torch.tensor(float(params[1]), dtype=torch.float64, requires_grad=True)
    third = torch.tensor(float(params[2]) if len(params) > 2 else 0.0, dtype=torch.float64, requires_grad=True)
    r = (x[nbr[:, 0]] - x[nbr[:, 1]] - off.double() * cell).pow(2).sum(1).sqrt()
    s = sig / r
    if name == "buck":          # (A, B, C) take the places of (sigma, eps, third)
        u = sig * torch.exp(-eps * r) - third / r ** 6
    elif name == "morse":
        am, phi = float(params[0]), float(params[1])
        A0 = 0.0 if phi >= 0 else float(np.exp(2 * am / phi) - 2 * np.exp(am / phi))
        ex = am * (1 - r ** phi) / phi
        u = (torch.exp(2 * ex) - 2 * torch.exp(ex) - A0) / (1 + A0) + 0.0 * (sig + eps)
    elif name == "lj":
        u = 4 * eps * (s ** 12 - s ** 6)
    elif name == "lj69":
        u = 4 * eps * (s ** 9 - s ** 6)
    elif name == "ljfam":
        u = 4 * eps * (s ** params[2] - s ** params[3])
    else:
        u = 4 * eps * s ** params[2]
    F = -torch.autograd.grad(u.sum(), x, create_graph=True)[0]
    hv_o, ds_o, de_o, d3_o = torch.autograd.grad((F * a).sum(), (x, sig, eps, third), allow_unused=True)
    if name == "morse":
        ds_o, de_o = torch.zeros(()), torch.zeros(())
    d3_o = torch.zeros((), dtype=torch.float64) if d3_o is None else d3_o
    ectx.nbr_list(xyz.float(), [float(np.float32(L))] * 3, rc)
    hv, dth = ectx.pair_hvp(kind, [float(p) for p in params], xyz.float(), a.float())
    scale = hv_o.abs().max().item()
    assert (hv.double() - hv_o).abs().max().item() <= 2e-5 * scale
    # the parameter products are signed sums over all pairs with heavy cancellation (random a): the fp32 error scales
    # with the sum of the |per-pair terms|, not with the result
    with torch.no_grad():
        rda = ((x[nbr[:, 1]] - x[nbr[:, 0]] + off.double() * cell) * (a[nbr[:, 0]] - a[nbr[:, 1]])).sum(1).abs()
    cond = float((rda / r.detach() ** 2).sum()) * 4 * abs(params[1]) * 12 * 12     # ~ sum |dg/dtheta| |r.da| (s <~ 1)
    if name == "buck":
        cond = float((rda * (1.0 / r.detach() ** 8 + torch.exp(-params[1] * r.detach()) * (1 + params[0]) / r.detach())).sum()) * 8
        assert abs(dth[2].item() - d3_o.item()) <= 2e-5 * max(1.0, abs(d3_o.item())) + 3e-7 * cond
    assert abs(dth[0].item() - ds_o.item()) <= 2e-5 * max(1.0, abs(ds_o.item())) + 3e-7 * cond
    assert abs(dth[1].item() - de_o.item()) <= 2e-5 * max(1.0, abs(de_o.item())) + 3e-7 * cond


def test_emu_pair_hvp_rejects_unknown_kind(ectx):
    xyz = torch.tensor(O.fcc_positions(3, 1.679), dtype=torch.float32)
    ectx.nbr_list(xyz, [3 * 1.679] * 3, 2.5)
    with pytest.raises(_lib.MdgError):
        ectx.pair_hvp(17, [1.0, 1.0], xyz, torch.zeros_like(xyz))
