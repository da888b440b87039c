"""CPU-emulated run (tests/cuemu) of the native SchNet energy+force program against the oracle (autograd forces) and
against the reference fixtures.  Test infrastructure - the GPU parity test of the same op is in test_schnet.py."""
import numpy as np
import pytest
import torch

from emu_lib import EmuContext
from mdgrad_b200 import _lib
from oracle import oracle_torch as O
from test_schnet import _fixture


@pytest.fixture(scope="module")
def ectx():
    return EmuContext()


def _rand_sd(A, F, G, L, R, cutoff, seed, bias=True):
    g = torch.Generator().manual_seed(seed)
    sd = {"atom_embed.weight": torch.randn(100, A, generator=g)}
    mu = torch.linspace(0.0, cutoff, G)
    lin = lambda o, i: torch.randn(o, i, generator=g) * (1.5 / np.sqrt(i))      # noqa: E731
    b = lambda o: (torch.randn(o, generator=g) * 0.3 if bias else torch.zeros(o))  # noqa: E731
    for l in range(L):
        pre = "convolutions.%d.moduledict." % l
        sd[pre + "message_edge_filter.0.width"] = (mu[1] - mu[0]) * torch.ones(G)
        sd[pre + "message_edge_filter.0.offsets"] = mu.clone()
        sd[pre + "message_edge_filter.1.weight"], sd[pre + "message_edge_filter.1.bias"] = lin(G, G), b(G)
        sd[pre + "message_edge_filter.3.weight"], sd[pre + "message_edge_filter.3.bias"] = lin(F, G), b(F)
        sd[pre + "message_node_filter.weight"], sd[pre + "message_node_filter.bias"] = lin(F, A), b(F)
        sd[pre + "update_function.0.weight"], sd[pre + "update_function.0.bias"] = lin(A, F), b(A)
        sd[pre + "update_function.2.weight"], sd[pre + "update_function.2.bias"] = lin(A, A), b(A)
    ro = "atomwisereadout.readout.energy."
    sd[ro + "linear0.weight"], sd[ro + "linear0.bias"] = lin(R, A), b(R)
    sd[ro + "linear2.weight"], sd[ro + "linear2.bias"] = lin(1, R), b(1)
    return sd


@pytest.mark.parametrize("n,A,F,G,L,R,mode", [
    (40, 16, 24, 7, 1, 8, "reference"),          # ragged sizes: every tile guard of the GEMM / edge kernels
    (97, 36, 52, 13, 2, 18, "correct"),
    (150, 64, 64, 29, 3, 32, "reference"),
    (64, 512, 256, 33, 3, 256, "correct"),       # the layer widths of configs[4] (a-Si: A512 F256 G33 L3)
    (50, 64, 512, 64, 1, 32, "reference"),       # F > one block of threads, G at the supported maximum
])
def test_emu_schnet_random_model(ectx, n, A, F, G, L, R, mode):
    rng = np.random.default_rng(n)
    box = 9.0
    xyz = torch.tensor(rng.uniform(0, box, (n, 3)), dtype=torch.float32)
    z = torch.tensor(rng.integers(1, 9, n), dtype=torch.long)
    cell = torch.tensor([box, box * 1.1, box * 0.95])
    rc = 3.2
    nbr, off = O.neighbor_list(xyz, rc, cell)
    sd = _rand_sd(A, F, G, L, R, rc, seed=n)
    x = xyz.clone().requires_grad_(True)
    e_o = O.schnet_energy(sd, z, x, nbr, off, cell=torch.diag(cell), pbc_mode=mode)
    f_o = -torch.autograd.grad(e_o, x)[0]
    model = _lib.schnet_model_struct(sd, "cpu")
    scale = (1.0, 1.0, 1.0) if mode == "reference" else tuple(cell.tolist())
    e, f = ectx.schnet_energy_force(model, z, xyz, nbr, off, scale)
    assert abs(e.item() - e_o.item()) <= 1e-5 * max(1.0, abs(e_o.item()))
    assert (f - f_o).abs().max().item() <= 2e-5 * f_o.abs().max().item()
    e2, _ = ectx.schnet_energy_force(model, z, xyz, nbr, off, scale, want_force=False)
    assert e2.item() == e.item()


@pytest.mark.parametrize("tag", ["water", "si"])
def test_emu_schnet_vs_reference_fixture(ectx, tag):
    g, params, sd = _fixture(tag)
    xyz = torch.Tensor(g["positions"])
    cell = torch.Tensor(g["cell"])
    nbr, off = O.neighbor_list(xyz, params["cutoff"], cell)
    z = torch.tensor(g["numbers"], dtype=torch.long)
    model = _lib.schnet_model_struct(sd, "cpu")
    e, f = ectx.schnet_energy_force(model, z, xyz, nbr, off)
    eref = float(g["energy"].reshape(-1)[0])
    assert abs(e.item() - eref) <= 1e-5 * abs(eref)
    assert np.abs(f.numpy() - g["forces"]).max() <= 1e-5 * np.abs(g["forces"]).max()


def test_emu_schnet_no_edges_and_single_atom(ectx):
    """isolated atoms: energy = sum of the per-atom readout of the embedding path, zero forces"""
    sd = _rand_sd(16, 24, 7, 2, 8, 3.0, seed=1)
    model = _lib.schnet_model_struct(sd, "cpu")
    for n in (1, 5):
        xyz = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) * 50.0
        z = torch.arange(1, n + 1, dtype=torch.long)
        nbr = torch.zeros((0, 2), dtype=torch.int64)
        off = torch.zeros((0, 3), dtype=torch.float32)
        e_o = O.schnet_energy(sd, z, xyz, nbr, off)
        e, f = ectx.schnet_energy_force(model, z, xyz, nbr, off)
        assert abs(e.item() - e_o.item()) <= 1e-5 * max(1.0, abs(e_o.item()))
        assert float(f.abs().max()) == 0.0


@pytest.mark.parametrize("n,A,F,G,L,R", [
    (40, 16, 24, 7, 1, 8),            # every tile guard: M, N far below one 128 x 64 tile, K = 16 / 24 (not a multiple of the 32-wide k-block)
    (150, 64, 64, 29, 2, 32),         # N = 64 -> the 64-column tile exactly
    (200, 132, 68, 13, 2, 36),        # N = 132 / 68: the 128-column tile with a ragged second tile; M spans two row tiles
    (64, 512, 256, 33, 3, 256),       # the layer widths of configs[4]
])
def test_emu_tc_dense_layers_equal_simt(ectx, monkeypatch, n, A, F, G, L, R):
    """MDG_SCHNET_TC=1: the dense layers go through k_sn_gemm_tc - the tcgen05 kernel's logic (canonical K-major staging, shared
    memory / instruction descriptors, 3xTF32 split, K loop, tile guards, TMEM epilogue mapping) executed on a functional model of
    the tensor-core primitives (csrc/schnet_tc.cuh) - and must reproduce the SIMT path within the parity bar"""
    rng = np.random.default_rng(n + A)
    box = 9.0
    xyz = torch.tensor(rng.uniform(0, box, (n, 3)), dtype=torch.float32)
    z = torch.tensor(rng.integers(1, 9, n), dtype=torch.long)
    cell = torch.tensor([box, box * 1.1, box * 0.95])
    nbr, off = O.neighbor_list(xyz, 3.0, cell)
    sd = _rand_sd(A, F, G, L, R, 3.0, seed=n)
    model = _lib.schnet_model_struct(sd, "cpu")
    monkeypatch.delenv("MDG_SCHNET_TC", raising=False)
    e0, f0 = ectx.schnet_energy_force(model, z, xyz, nbr, off)
    l0 = ectx.stats()["launches"]
    monkeypatch.setenv("MDG_SCHNET_TC", "1")
    e1, f1 = ectx.schnet_energy_force(model, z, xyz, nbr, off)
    monkeypatch.delenv("MDG_SCHNET_TC", raising=False)
    assert abs(e1.item() - e0.item()) <= 1e-5 * max(1.0, abs(e0.item())), (e0.item(), e1.item())
    assert (f1 - f0).abs().max().item() <= 2e-5 * f0.abs().max().item()
    assert not torch.equal(f1, f0)                    # (a different arithmetic did run: 3xTF32 is not bit-identical to fp32 FMA)
