"""GPU parity tests proper: CUDA path (through the C ABI) vs the CPU oracle on seeded inputs."""
import math

import numpy as np
import pytest
import torch

from oracle import oracle_torch as O

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


def _rand_system(n, L, seed, spread=0.3):
    rng = np.random.default_rng(seed)
    Ls = np.array([L, L * 1.1, L * 0.9])
    xyz = rng.uniform(-spread * Ls, (1 + spread) * Ls, (n, 3))
    return torch.tensor(xyz, dtype=torch.float32), torch.tensor(Ls, dtype=torch.float32)


# ------------------------------------------------------------------------------------------
# K1: neighbor list bit-exact (indices, order, offsets); distances to 2 ulp
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,L,rc,seed", [
    (2, 5.0, 2.5, 0), (108, 5.037, 2.5, 1), (400, 7.3, 2.5, 2), (900, 11.1, 3.1, 3),
    (1500, 21.84, 4.9, 4),                   # all-pairs path
    (4000, 16.8, 2.5, 5), (6000, 19.0, 2.5, 6), (12000, 30.0, 3.3, 7),   # cell path
])
def test_nbr_list_bit_exact(ctx, n, L, rc, seed):
    xyz, cell = _rand_system(n, L, seed)
    if n > 10:
        xyz[5] = xyz[7]                      # coincident atoms are dropped by d2 != 0
    nbr_o, dis_o, off_o = O.neighbor_list(xyz, rc, cell, get_dis=True)
    nbr, off, dis = ctx.nbr_list(xyz.to(_dev()), cell.tolist(), rc, get_dis=True)
    assert nbr.dtype == torch.int64 and off.dtype == torch.float32
    assert torch.equal(nbr.cpu(), nbr_o), "neighbor indices/order differ"
    assert torch.equal(off.cpu(), off_o), "image offsets differ"
    # torch-CPU sqrt is not always correctly rounded: tolerance on distances
    torch.testing.assert_close(dis.cpu(), dis_o, rtol=3e-7, atol=0)
    st = ctx.stats()
    assert st["path"] == (0 if n > 3072 else 1)


def test_nbr_list_masks(ctx):
    n = 5000
    xyz, cell = _rand_system(n, 18.0, 11)
    rng = np.random.default_rng(3)
    A = list(range(0, n, 3))
    B = list(range(1, n, 2))
    ex = rng.integers(0, n, (400, 2))
    ex = ex[ex[:, 0] != ex[:, 1]]
    nbr_o, off_o = O.neighbor_list(xyz, 2.5, cell, index_tuple=(A, B), ex_pairs=torch.tensor(ex))
    dev = _dev()
    sa = torch.zeros(n, dtype=torch.uint8); sa[A] = 1
    sb = torch.zeros(n, dtype=torch.uint8); sb[B] = 1
    lo, hi = np.minimum(ex[:, 0], ex[:, 1]), np.maximum(ex[:, 0], ex[:, 1])
    keys = torch.tensor(np.unique(lo.astype(np.int64) * n + hi), dtype=torch.int64)
    nbr, off = ctx.nbr_list(xyz.to(dev), cell.tolist(), 2.5, sel_a=sa.to(dev), sel_b=sb.to(dev), ex_keys=keys.to(dev))
    assert torch.equal(nbr.cpu(), nbr_o)
    assert torch.equal(off.cpu(), off_o)


def test_nbr_list_fcc_known_answer(ctx):
    # reference torchmd/topology.py:126-147: FCC 3x3x3, a=1.679, rc=2.5 -> 5832 directed pairs
    pos = O.fcc_positions(3, 1.679)
    xyz = torch.tensor(pos, dtype=torch.float32)
    nbr, off = ctx.nbr_list(xyz.to(_dev()), [3 * 1.679] * 3, 2.5)
    assert 2 * nbr.shape[0] == 5832
    assert int((off.abs().sum(1) != 0).sum()) == 1362


# ------------------------------------------------------------------------------------------
# K2+K3: energies / forces / parameter gradients, 1e-5 relative
# ------------------------------------------------------------------------------------------
POTS = [
    ("lj", 0, (1.0, 1.0)), ("ljfam", 1, (1.0, 0.8, 10, 5)), ("lj69", 2, (1.1, 0.7)),
    ("exv", 3, (1.0, 0.5, 12)), ("buck", 4, (1000.0, 3.5, 2.0)), ("morse", 5, (6.0, 2.0)),
]


@pytest.mark.parametrize("name,kind,params", POTS)
@pytest.mark.parametrize("ncell", [3, 10])
def test_pair_force_parity(ctx, name, kind, params, ncell):
    pos, _, L = O.lj_system(ncell, rho=0.845, jitter=0.04, seed=5)
    xyz = torch.tensor(pos, dtype=torch.float32)
    cell = torch.tensor([L, L, L], dtype=torch.float32)
    rc = 2.5
    nbr_o, off_o = O.neighbor_list(xyz, rc, cell)
    need_dp = name in ("lj", "lj69", "buck")
    out = O.pair_energy_forces(xyz, nbr_o, off_o, cell, name, params, need_param_grads=need_dp)
    e_o, f_o = out[0], out[1]
    dev = _dev()
    ctx.nbr_list(xyz.to(dev), cell.tolist(), rc)
    e, f, dp = ctx.pair_force(kind, [float(p) for p in params], xyz.to(dev), want_dparams=need_dp)
    assert abs(e.item() - e_o.item()) <= 1e-5 * abs(e_o.item())
    fmax = f_o.abs().max().item()
    assert (f.cpu() - f_o).abs().max().item() <= 1e-5 * fmax
    if need_dp:
        for k, g in enumerate(out[2]):
            assert abs(dp[k].item() - g.item()) <= 2e-5 * max(1.0, abs(g.item()))


def test_pair_force_fcc_golden_energy(ctx):
    # FCC 3x3x3 a=1.679 LJ(1,1) rc 2.5 -> E = -732.372681 (ref fp32) / -732.372746 (fp64)
    xyz = torch.tensor(O.fcc_positions(3, 1.679), dtype=torch.float32).to(_dev())
    ctx.nbr_list(xyz, [3 * 1.679] * 3, 2.5)
    e, f, _ = ctx.pair_force(0, [1.0, 1.0], xyz)
    assert abs(e.item() - (-732.372746)) < 1e-5 * 732.4
    assert f.abs().max().item() < 1e-3


# ------------------------------------------------------------------------------------------
# K4 + driver: fused epoch vs oracle trajectory
# ------------------------------------------------------------------------------------------
def _md_params(integrator, L, n, M=5, Qv=50.0, T=1.0, skin=0.0, K=1, kind=0, pp=(1.0, 1.0), rc=2.5):
    from mdgrad_b200 import _lib
    p = _lib.MdParams()
    p.integrator = integrator
    p.pot_kind = kind
    for i, v in enumerate(pp):
        p.pot_params[i] = float(v)
    p.cutoff = rc
    for k in range(3):
        p.cell[k] = L
    p.n_chains = M
    Qb = O.nhc_bath_masses(Qv, n, M)
    for k in range(M):
        p.Q[k] = float(Qb[k])
    p.T = T
    p.ndof = 3 * n
    p.skin = skin
    p.rebuild_every = K
    p.traj_stride = 1
    return p, Qb


@pytest.mark.parametrize("ncell,skin,K,nsteps", [(3, 0.0, 1, 50), (5, 0.0, 1, 20), (10, 0.0, 1, 12), (10, 0.3, 4, 12)])
def test_nhc_epoch_vs_oracle(ctx, ncell, skin, K, nsteps):
    a = 1.679 if ncell == 3 else None
    pos, vel, L = O.lj_system(ncell, jitter=0.02, seed=9, a=a)
    n = pos.shape[0]
    L32 = float(np.float32(L))
    q0 = torch.tensor(pos, dtype=torch.float32)
    v0 = torch.tensor(vel, dtype=torch.float32)
    mass = torch.full((n,), 1.008)
    p, Qb = _md_params(1, L32, n, skin=skin, K=K)
    t = O.time_grid(0.005, nsteps)
    sysO = O.PairSystemOracle(torch.tensor([L32] * 3), 2.5, "lj", (1.0, 1.0))
    vo, qo, po = O.nh_verlet_trajectory(sysO.force, v0, q0, torch.zeros(5), t, mass, Qb, 1.0, 3 * n)
    dev = _dev()
    tv, tq, tpv, e = ctx.md_run(p, mass.to(dev), v0.to(dev), q0.to(dev), [0.0] * 5, t.tolist(), want_energy=True)
    assert tv.shape == vo.shape and tq.shape == qo.shape and tpv.shape == po.shape
    assert torch.equal(tv[0].cpu(), v0) and torch.equal(tq[0].cpu(), q0)
    # one step from identical states: tight; whole short trajectory: chaos-limited
    vscale = vo.abs().max().item()
    assert (tv[1].cpu() - vo[1]).abs().max().item() <= 2e-6 * vscale
    assert (tq[1].cpu() - qo[1]).abs().max().item() <= 2e-6 * L
    assert (tv[-1].cpu() - vo[-1]).abs().max().item() <= 2e-4 * vscale
    assert (tq[-1].cpu() - qo[-1]).abs().max().item() <= 2e-5 * L
    torch.testing.assert_close(tpv.cpu(), po, rtol=2e-4, atol=2e-4 * po.abs().max().item())
    assert abs(e - sysO.last_energy.item()) <= 2e-5 * abs(sysO.last_energy.item())


def test_nve_epoch_vs_oracle(ctx):
    pos, vel, L = O.lj_system(5, jitter=0.02, seed=4)
    n = pos.shape[0]
    L32 = float(np.float32(L))
    q0 = torch.tensor(pos, dtype=torch.float32)
    v0 = torch.tensor(vel, dtype=torch.float32)
    mass = torch.full((n,), 1.008)
    p, _ = _md_params(0, L32, n)
    t = O.time_grid(0.005, 20)
    sysO = O.PairSystemOracle(torch.tensor([L32] * 3), 2.5, "lj", (1.0, 1.0))
    vo, qo = O.verlet_trajectory(sysO.force, v0, q0, t)
    dev = _dev()
    tv, tq, tpv, _ = ctx.md_run(p, mass.to(dev), v0.to(dev), q0.to(dev), [], t.tolist())
    assert tpv is None
    assert (tv[-1].cpu() - vo[-1]).abs().max().item() <= 2e-4 * vo.abs().max().item()
    assert (tq[-1].cpu() - qo[-1]).abs().max().item() <= 2e-5 * L


def test_skin_list_equals_fresh_list(ctx):
    """The Verlet-skin engine must give the same trajectory as rebuilding every step."""
    pos, vel, L = O.lj_system(12, jitter=0.03, seed=2)
    n = pos.shape[0]
    L32 = float(np.float32(L))
    dev = _dev()
    q0 = torch.tensor(pos, dtype=torch.float32).to(dev)
    v0 = torch.tensor(vel, dtype=torch.float32).to(dev)
    mass = torch.full((n,), 1.008).to(dev)
    t = O.time_grid(0.005, 30).tolist()
    pa, _ = _md_params(1, L32, n, skin=0.0, K=1)
    pb, _ = _md_params(1, L32, n, skin=0.4, K=8)
    a = ctx.md_run(pa, mass, v0, q0, [0.0] * 5, t)
    b = ctx.md_run(pb, mass, v0, q0, [0.0] * 5, t)
    # identical pair sets at every evaluation; only the summation order (sort history) differs
    vs = a[0].abs().max().item()
    assert (a[0] - b[0]).abs().max().item() <= 5e-5 * vs
    assert (a[1] - b[1]).abs().max().item() <= 5e-6 * L
    assert (a[2] - b[2]).abs().max().item() <= 5e-5 * a[2].abs().max().item()
    assert ctx.stats()["rebuilds"] < 30


def _new_ctx():
    from mdgrad_b200 import _lib
    return _lib.Context(_dev())


def test_tile_list_engine_equals_row_list_engine(monkeypatch):
    """MDG_TILES=1 (engine skin list in the block-local tile form: TMA-staged stencil in shared memory, 16-bit rows,
    persistent warp-specialised k_force_tiles) against the default row form and against rebuilding every step: same pair set
    at every evaluation (exact re-test in both), only the summation order differs.  Boundary blocks (image-shifted
    entries), a non-cubic box and a pair filter are covered."""
    pos, vel, L = O.lj_system(12, jitter=0.03, seed=2)
    n = pos.shape[0]
    L32 = float(np.float32(L))
    dev = _dev()
    q0 = torch.tensor(pos, dtype=torch.float32).to(dev)
    v0 = torch.tensor(vel, dtype=torch.float32).to(dev)
    mass = torch.full((n,), 1.008).to(dev)
    t = O.time_grid(0.005, 30).tolist()
    monkeypatch.setenv("MDG_TILES", "1")
    tctx = _new_ctx()
    monkeypatch.delenv("MDG_TILES")
    rctx = _new_ctx()
    pb, _ = _md_params(1, L32, n, skin=0.4, K=8)
    pa, _ = _md_params(1, L32, n, skin=0.0, K=1)
    a = rctx.md_run(pa, mass, v0, q0, [0.0] * 5, t, want_energy=True)
    r = rctx.md_run(pb, mass, v0, q0, [0.0] * 5, t, want_energy=True)
    assert rctx.stats()["path"] == 0
    b = tctx.md_run(pb, mass, v0, q0, [0.0] * 5, t, want_energy=True)
    assert tctx.stats()["path"] == 2 and tctx.stats()["rebuilds"] < 30
    vs = a[0].abs().max().item()
    for x in (a, r):
        assert (x[0] - b[0]).abs().max().item() <= 5e-5 * vs
        assert (x[1] - b[1]).abs().max().item() <= 5e-6 * L
        assert (x[2] - b[2]).abs().max().item() <= 5e-5 * a[2].abs().max().item()
        assert abs(x[3] - b[3]) <= 2e-5 * abs(a[3])
    # other analytic kinds share the kernel template
    for kind, pp in ((1, (1.0, 0.8, 10.0, 5.0)), (3, (1.0, 0.5, 12.0))):
        pk, _ = _md_params(1, L32, n, skin=0.4, K=8, kind=kind, pp=pp)
        x = rctx.md_run(pk, mass, v0, q0, [0.0] * 5, t[:12])
        y = tctx.md_run(pk, mass, v0, q0, [0.0] * 5, t[:12])
        assert (x[1] - y[1]).abs().max().item() <= 5e-6 * L


# ------------------------------------------------------------------------------------------
# K6: RDF
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ncell", [3, 10])
def test_rdf_parity(ctx, ncell):
    pos, _, L = O.lj_system(ncell, jitter=0.06, seed=8, a=1.679 if ncell == 3 else None)
    xyz = torch.tensor(pos, dtype=torch.float32)
    end = 2.0 if ncell == 3 else 3.3
    co, bo, go = O.rdf(xyz, [L] * 3, 100, (0.75, end))
    dev = _dev()
    count = torch.zeros(100, device=dev)
    ctx.rdf_accumulate(xyz.to(dev), [L] * 3, 0.75, end, 100, None, count)
    c = count / count.sum()
    torch.testing.assert_close(c.cpu(), co, rtol=1e-5, atol=1e-5 * co.max().item())


# ------------------------------------------------------------------------------------------
# edge cases: empty / single-atom / no-pairs inputs, zero-step epochs, capacity growth
# ------------------------------------------------------------------------------------------
def test_edge_empty_and_single_atom(ctx):
    dev = _dev()
    for n in (0, 1):
        xyz = torch.rand(n, 3, device=dev) * 4.0
        nbr, off, dis = ctx.nbr_list(xyz, [5.0, 5.0, 5.0], 2.5, get_dis=True)
        assert nbr.shape == (0, 2) and off.shape == (0, 3) and dis.shape == (0,)
        if n:
            e, f, _ = ctx.pair_force(0, [1.0, 1.0], xyz)
            assert e.item() == 0.0 and float(f.abs().max()) == 0.0
    # two atoms exactly at the cutoff distance are NOT neighbors (strict <), one ulp inside they are
    rc = 2.5
    a = torch.tensor([[1.0, 1.0, 1.0], [1.0 + rc, 1.0, 1.0]], device=dev)
    assert ctx.nbr_list(a, [20.0] * 3, rc)[0].shape[0] == 0
    b = a.clone()
    b[1, 0] = torch.nextafter(b[1, 0], torch.tensor(0.0, device=dev))
    nbr_o, _ = O.neighbor_list(b.cpu(), rc, torch.tensor([20.0] * 3))
    assert ctx.nbr_list(b, [20.0] * 3, rc)[0].shape[0] == nbr_o.shape[0]


def test_edge_no_pairs_in_range_and_dilute_cells(ctx):
    # very dilute box on the cell path: most cells empty, few pairs
    rng = np.random.default_rng(0)
    n = 4000
    xyz = torch.tensor(rng.uniform(0, 200.0, (n, 3)), dtype=torch.float32)
    nbr_o, off_o = O.neighbor_list(xyz, 2.5, torch.tensor([200.0] * 3), block=512)
    nbr, off = ctx.nbr_list(xyz.to(_dev()), [200.0] * 3, 2.5)
    assert torch.equal(nbr.cpu(), nbr_o) and torch.equal(off.cpu(), off_o)
    e, f, _ = ctx.pair_force(0, [1.0, 1.0], xyz.to(_dev()))
    eo = O.pair_energy_forces(xyz, nbr_o, off_o, torch.tensor([200.0] * 3), "lj", (1.0, 1.0))[0] if nbr_o.shape[0] else torch.tensor(0.0)
    assert abs(e.item() - eo.item()) <= 1e-5 * max(1.0, abs(eo.item()))


def test_edge_dense_cluster_grows_row_capacity(ctx):
    # a dense blob in a large box: the mean-density capacity estimate is far too small -> internal growth + retry
    rng = np.random.default_rng(1)
    blob = rng.normal(10.0, 0.9, (600, 3))
    gas = rng.uniform(0, 60.0, (3400, 3))
    xyz = torch.tensor(np.concatenate([blob, gas]), dtype=torch.float32)
    nbr_o, off_o = O.neighbor_list(xyz, 2.5, torch.tensor([60.0] * 3), block=512)
    nbr, off = ctx.nbr_list(xyz.to(_dev()), [60.0] * 3, 2.5)
    assert torch.equal(nbr.cpu(), nbr_o) and torch.equal(off.cpu(), off_o)
    assert int(torch.bincount(nbr_o.reshape(-1)).max()) > 200          # rows far beyond the mean-density estimate


def test_edge_zero_step_epoch(ctx):
    pos, vel, L = O.lj_system(3, jitter=0.02, seed=9, a=1.679)
    n = pos.shape[0]
    dev = _dev()
    q0 = torch.tensor(pos, dtype=torch.float32, device=dev)
    v0 = torch.tensor(vel, dtype=torch.float32, device=dev)
    mass = torch.full((n,), 1.008, device=dev)
    p, _ = _md_params(1, float(np.float32(L)), n)
    tv, tq, tpv, e = ctx.md_run(p, mass, v0, q0, [0.1, 0.2, 0.3, 0.4, 0.5], [0.0], want_energy=True)
    assert tv.shape == (1, n, 3) and torch.equal(tv[0], v0) and torch.equal(tq[0], q0)
    assert tpv.shape == (1, 5) and abs(tpv[0, 2].item() - 0.3) < 1e-7
    nbr_o, off_o = O.neighbor_list(q0.cpu(), 2.5, torch.tensor([float(np.float32(L))] * 3))
    eo = O.pair_energy_forces(q0.cpu(), nbr_o, off_o, torch.tensor([float(np.float32(L))] * 3), "lj", (1.0, 1.0))[0]
    assert abs(e - eo.item()) <= 2e-5 * abs(eo.item())


def test_edge_atoms_outside_the_box_and_unwrapped(ctx):
    # atoms several boxes away from the cell (unwrapped trajectories): same list as the oracle
    xyz, cell = _rand_system(5000, 18.0, 21, spread=2.2)
    nbr_o, off_o = O.neighbor_list(xyz, 2.5, cell, block=512)
    nbr, off = ctx.nbr_list(xyz.to(_dev()), cell.tolist(), 2.5)
    assert torch.equal(nbr.cpu(), nbr_o) and torch.equal(off.cpu(), off_o)


def test_known_answers_fcc500(ctx):
    a = (4 / 0.8442) ** (1 / 3)
    xyz = torch.tensor(O.fcc_positions(5, a), dtype=torch.float32).to(_dev())
    nbr, off = ctx.nbr_list(xyz, [5 * a] * 3, 2.5)
    assert nbr.shape[0] == 13500
    e, f, _ = ctx.pair_force(0, [1.0, 1.0], xyz)
    assert abs(e.item() - (-3386.684)) < 1e-5 * 3386.7 + 1e-3
