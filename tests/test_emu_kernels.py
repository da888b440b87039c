"""CPU-emulated runs of the CUDA kernels (tests/cuemu) against the oracle - functional coverage of the kernels and of the
host-side launch logic in the GPU-less container.  The test BODIES are the GPU parity tests of test_gpu_kernels.py,
re-run with an emulated context and CPU tensors (smaller size lists where the emulation would take minutes).

This is test infrastructure: it proves nothing about the GPU build except that the same source computes the right
thing when executed thread by thread; the `-m gpu` tests stay the parity gate.
"""
import numpy as np
import pytest
import torch

import test_gpu_kernels as G
from emu_lib import EmuContext
from oracle import oracle_torch as O


@pytest.fixture(scope="module")
def ectx():
    return EmuContext()


@pytest.fixture(autouse=True)
def _cpu_device(monkeypatch):
    monkeypatch.setattr(G, "_dev", lambda: torch.device("cpu"))


@pytest.mark.parametrize("n,L,rc,seed", [
    (2, 5.0, 2.5, 0), (108, 5.037, 2.5, 1), (400, 7.3, 2.5, 2),        # all-pairs path
    (4000, 16.8, 2.5, 5),                                              # cell path
])
def test_emu_nbr_list_bit_exact(ectx, n, L, rc, seed):
    G.test_nbr_list_bit_exact.__wrapped__(ectx, n, L, rc, seed) if hasattr(G.test_nbr_list_bit_exact, "__wrapped__") \
        else G.test_nbr_list_bit_exact(ectx, n, L, rc, seed)


def test_emu_nbr_list_masks(ectx):
    G.test_nbr_list_masks(ectx)


def test_emu_nbr_known_answers(ectx):
    G.test_nbr_list_fcc_known_answer(ectx)
    G.test_known_answers_fcc500(ectx)


@pytest.mark.parametrize("name,kind,params", G.POTS)
def test_emu_pair_force_parity(ectx, name, kind, params):
    G.test_pair_force_parity(ectx, name, kind, params, 3)


def test_emu_pair_force_cells(ectx):
    G.test_pair_force_parity(ectx, "lj", 0, (1.0, 1.0), 10)


@pytest.mark.parametrize("ncell,skin,K,nsteps", [(3, 0.0, 1, 20), (10, 0.0, 1, 4), (10, 0.3, 4, 9)])
def test_emu_nhc_epoch_vs_oracle(ectx, ncell, skin, K, nsteps):
    G.test_nhc_epoch_vs_oracle(ectx, ncell, skin, K, nsteps)


def test_emu_nve_epoch_vs_oracle(ectx):
    G.test_nve_epoch_vs_oracle(ectx)


def test_emu_skin_list_equals_fresh_list(ectx):
    """fast builder + re-test kernel (pure and image-coded rows) vs rebuilding every step"""
    pos, vel, L = O.lj_system(12, jitter=0.03, seed=2)
    n = pos.shape[0]
    L32 = float(np.float32(L))
    q0 = torch.tensor(pos, dtype=torch.float32)
    v0 = torch.tensor(vel, dtype=torch.float32)
    mass = torch.full((n,), 1.008)
    t = O.time_grid(0.005, 10).tolist()
    pa, _ = G._md_params(1, L32, n, skin=0.0, K=1)
    pb, _ = G._md_params(1, L32, n, skin=0.4, K=8)
    a = ectx.md_run(pa, mass, v0, q0, [0.0] * 5, t)
    b = ectx.md_run(pb, mass, v0, q0, [0.0] * 5, t)
    vs = a[0].abs().max().item()
    assert (a[0] - b[0]).abs().max().item() <= 5e-5 * vs
    assert (a[1] - b[1]).abs().max().item() <= 5e-6 * L
    assert ectx.stats()["rebuilds"] < 10


def test_emu_rdf_parity(ectx):
    G.test_rdf_parity(ectx, 3)
    G.test_rdf_parity(ectx, 10)


def test_emu_edges(ectx):
    G.test_edge_empty_and_single_atom(ectx)
    G.test_edge_zero_step_epoch(ectx)
    G.test_edge_dense_cluster_grows_row_capacity(ectx)
