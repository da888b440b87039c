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


# ------------------------------------------------------------------------------------------
# engine corner cases that the GPU suite only touches at the op level: fast builder + pair filters, non-cubic boxes,
# unwrapped coordinates, other potential kinds under the re-test kernel, capacity growth and the skin retry
# ------------------------------------------------------------------------------------------
def _oracle_filtered_force(cell, rc, kind_name, params, index_tuple=None, ex_pairs=None):
    def force(q):
        nbr, off = O.neighbor_list(q.detach(), rc, cell, index_tuple=index_tuple, ex_pairs=ex_pairs)
        return O.pair_energy_forces(q.detach(), nbr, off, cell, kind_name, params)[1]
    return force


@pytest.mark.parametrize("kind_name,kind,pp", [("lj", 0, (1.0, 1.0)), ("lj69", 2, (1.05, 0.8)), ("exv", 3, (1.0, 0.5, 12))])
def test_emu_engine_skin_list_noncubic_unwrapped_filtered(ectx, kind_name, kind, pp):
    rng = np.random.default_rng(7)
    pos, vel, L = O.lj_system(12, jitter=0.03, seed=2)
    n = pos.shape[0]
    scale = np.array([1.0, 1.07, 0.94])
    Ls = (np.float32(L) * scale).astype(np.float32)
    pos = pos * scale
    pos = pos + rng.integers(-2, 3, (n, 1)) * Ls          # whole-box displacements: raw coordinates far outside
    q0 = torch.tensor(pos, dtype=torch.float32)
    v0 = torch.tensor(vel, dtype=torch.float32)
    mass = torch.full((n,), 1.008)
    A = list(range(0, n, 2))
    B = list(range(0, n, 3))
    ex = rng.integers(0, n, (300, 2))
    ex = ex[ex[:, 0] != ex[:, 1]]
    sa = torch.zeros(n, dtype=torch.uint8); sa[A] = 1
    sb = torch.zeros(n, dtype=torch.uint8); sb[B] = 1
    lo, hi = np.minimum(ex[:, 0], ex[:, 1]), np.maximum(ex[:, 0], ex[:, 1])
    keys = torch.tensor(np.unique(lo.astype(np.int64) * n + hi), dtype=torch.int64)
    p, Qb = G._md_params(1, 1.0, n, skin=0.35, K=3, kind=kind, pp=pp)
    for k in range(3):
        p.cell[k] = float(Ls[k])
    t = O.time_grid(0.004, 7)
    cell = torch.tensor(Ls)
    force = _oracle_filtered_force(cell, 2.5, kind_name, pp, index_tuple=(A, B), ex_pairs=torch.tensor(ex))
    vo, qo, po = O.nh_verlet_trajectory(force, v0, q0, torch.zeros(5), t, mass, Qb, 1.0, 3 * n)
    ectx.set_pair_filter(sa, sb, keys)
    try:
        tv, tq, tpv, _ = ectx.md_run(p, mass, v0, q0, [0.0] * 5, t.tolist())
    finally:
        ectx.set_pair_filter(None, None, None)
    assert ectx.stats()["rebuilds"] >= 2 and ectx.stats()["path"] == 0
    vs = vo.abs().max().item()
    assert (tv[1] - vo[1]).abs().max().item() <= 3e-6 * vs
    assert (tv[-1] - vo[-1]).abs().max().item() <= 1e-4 * vs
    assert (tq[-1] - qo[-1]).abs().max().item() <= 2e-5 * float(Ls.max()) * 3


def test_emu_engine_capacity_growth_and_skin_retry(ectx):
    """a dense blob makes the row-capacity estimate overflow (grow + redo); a far too optimistic rebuild interval with fast
    atoms violates the skin criterion (halve K + redo) - both inside mdg_md_run, the result still equals the oracle"""
    rng = np.random.default_rng(3)
    blob = rng.normal(20.0, 1.1, (500, 3))
    gas = rng.uniform(0, 40.0, (3300, 3))
    pos = np.concatenate([blob, gas])
    # push apart overlapping pairs a little so that forces stay finite-ish
    q0 = torch.tensor(pos, dtype=torch.float32)
    n = q0.shape[0]
    v0 = torch.tensor(rng.normal(0, 1.0, (n, 3)), dtype=torch.float32)
    v0[:20] *= 40.0                                        # a few very fast atoms: skin violation with K = 16
    mass = torch.full((n,), 1.0)
    L = 40.0
    p, Qb = G._md_params(0, L, n, skin=0.3, K=16, kind=3, pp=(1.0, 0.01, 4))     # soft ExcludedVolume: overlaps are harmless
    t = O.time_grid(0.002, 9)
    cell = torch.tensor([L] * 3)
    force = _oracle_filtered_force(cell, 2.5, "exv", (1.0, 0.01, 4))
    vo, qo = O.verlet_trajectory(force, v0, q0, t)
    tv, tq, tpv, _ = ectx.md_run(p, mass, v0, q0, [], t.tolist())
    st = ectx.stats()
    assert st["maxrow_or_K"] < 16                          # the engine had to shorten the rebuild interval
    assert (tq[-1] - qo[-1]).abs().max().item() <= 2e-5 * L
    assert (tv[-1] - vo[-1]).abs().max().item() <= 2e-4 * vo.abs().max().item()


def test_emu_nbr_list_fuzz(ectx):
    """seeded random sweep over atom counts (around the all-pairs / cell-list switch), densities, anisotropic boxes, cutoffs,
    unwrapped coordinates, coincident atoms, species selections and exclusions: indices, order and offsets bit-exact"""
    rng = np.random.default_rng(11)
    for it in range(14):
        n = int(rng.choice([1, 3, 33, 500, 3072, 3073, 3500, 4200]))
        dens = rng.uniform(0.05, 1.1)
        L = max(1.5, (n / dens) ** (1 / 3))
        Ls = np.array([L, L * rng.uniform(0.7, 1.4), L * rng.uniform(0.7, 1.4)], dtype=np.float32)
        rc = float(rng.uniform(0.3, max(0.35, min(4.5, 0.49 * Ls.min()))))
        spread = float(rng.choice([0.0, 0.3, 2.5]))
        xyz = torch.tensor(rng.uniform(-spread * Ls, (1 + spread) * Ls, (n, 3)), dtype=torch.float32)
        if n > 10:
            xyz[3] = xyz[8]
        kw, sel, keys = {}, (None, None), None
        if n > 10 and rng.random() < 0.5:
            A = sorted(set(rng.integers(0, n, n // 2).tolist()))
            B = sorted(set(rng.integers(0, n, n // 3).tolist()))
            sa = torch.zeros(n, dtype=torch.uint8); sa[A] = 1
            sb = torch.zeros(n, dtype=torch.uint8); sb[B] = 1
            kw["index_tuple"], sel = (A, B), (sa, sb)
        if n > 10 and rng.random() < 0.5:
            ex = rng.integers(0, n, (50, 2))
            ex = ex[ex[:, 0] != ex[:, 1]]
            kw["ex_pairs"] = torch.tensor(ex)
            lo, hi = np.minimum(ex[:, 0], ex[:, 1]), np.maximum(ex[:, 0], ex[:, 1])
            keys = torch.tensor(np.unique(lo.astype(np.int64) * n + hi), dtype=torch.int64)
        nbr_o, off_o = O.neighbor_list(xyz, rc, torch.tensor(Ls), block=256, **kw)
        nbr, off = ectx.nbr_list(xyz, Ls.tolist(), rc, sel_a=sel[0], sel_b=sel[1], ex_keys=keys)
        assert torch.equal(nbr, nbr_o) and torch.equal(off, off_o), (it, n, Ls, rc, spread)


def test_emu_tile_list_engine_equals_row_list_engine(monkeypatch):
    """body of test_gpu_kernels.py::test_tile_list_engine_equals_row_list_engine on the emulated kernels (the emulation replaces
    the TMA bulk copies / mbarrier pipeline of k_force_tiles by plain copies + block barriers: layout, builder and arithmetic
    are what is checked here)"""
    monkeypatch.setattr(G, "_new_ctx", lambda: EmuContext())
    G.test_tile_list_engine_equals_row_list_engine(monkeypatch)


def test_emu_fullsize_property_checks_at_reduced_size(ectx):
    """the property checks of tests/test_gpu_zfullsize.py (run there at 256 000 atoms) executed here on a 4 000-atom box"""
    import fullsize_checks as F
    cpu = torch.device("cpu")
    P = F.check_list_structure(ectx, cpu, 10)
    assert 100_000 < P < 120_000
    pairs, shifted = F.check_list_rows_bit_exact(ectx, cpu, 10, nrandom=1024, nboundary=1024)
    assert pairs > 40_000 and shifted > 5_000
    F.check_forces_against_c_oracle(ectx, cpu, 10, nrows=512)
    F.check_engine_invariants(ectx, cpu, 10, nsteps=12)
    F.check_rdf_two_paths(ectx, cpu, 10)
