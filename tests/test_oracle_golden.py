"""CPU: the oracle (torch restatement + C restatement) against the committed golden fixtures that
oracle/make_golden.py produced from the UNMODIFIED reference, and against the reference's own known
answer.  No GPU, no /root/reference needed."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle_c as C
from oracle import oracle_torch as O

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_neighbor_list_vs_reference_fixture(tag):
    g = _load("nbr_%s.npz" % tag)
    xyz, cell, rc = torch.tensor(g["xyz"]), torch.tensor(g["cell"]), float(g["rc"])
    nbr, dis, off = O.neighbor_list(xyz, rc, cell, get_dis=True, block=256)
    assert np.array_equal(nbr.numpy(), g["nbr"].astype(np.int64))            # bit-exact indices and order
    assert np.array_equal(off.numpy(), g["off"].astype(np.float32))
    assert np.array_equal(dis.numpy(), g["dis"])                             # same torch sqrt -> same bits
    nbr_m, off_m = O.neighbor_list(xyz, rc, torch.diag(cell), index_tuple=(g["sel_a"], g["sel_b"]),
                                   ex_pairs=g["ex"].astype(np.int64))
    assert np.array_equal(nbr_m.numpy(), g["nbr_m"].astype(np.int64))
    assert np.array_equal(off_m.numpy(), g["off_m"].astype(np.float32))
    # C restatement: same list
    nbr_c, off_c, dis_c = C.nbr_list(g["xyz"], g["cell"], rc)
    assert np.array_equal(nbr_c, g["nbr"].astype(np.int64))
    assert np.array_equal(off_c, g["off"].astype(np.float32))
    np.testing.assert_allclose(dis_c, g["dis"], rtol=2e-7)                   # torch-CPU sqrt is not always correctly rounded


def test_reference_known_answer_fcc():
    # reference torchmd/topology.py:126-147 prints 5832 directed (ASE) = 2 x 2916 pairs
    xyz = torch.tensor(O.fcc_positions(3, 1.679), dtype=torch.float32)
    nbr, off = O.neighbor_list(xyz, 2.5, torch.tensor([3 * 1.679] * 3))
    assert 2 * nbr.shape[0] == 5832
    g = _load("pair_fcc108.npz")
    assert int(g["fcc_pairs"]) == 2916
    assert abs(float(g["fcc_energy"]) - (-732.372681)) < 2e-4


@pytest.mark.parametrize("name,params", [
    ("lj", (1.0, 1.0)), ("ljfam", (1.0, 0.8, 10, 5)), ("lj69", (1.1, 0.7)), ("exv", (1.0, 0.5, 12)),
    ("buck", (1000.0, 3.5, 2.0)), ("morse", (6.0, 2.0))])
def test_pair_energy_forces_vs_reference_fixture(name, params):
    g = _load("pair_fcc108.npz")
    xyz, cell = torch.tensor(g["xyz"]), torch.tensor(g["cell"])
    nbr, off = O.neighbor_list(xyz, 2.5, cell)
    need = name in ("lj", "lj69", "buck", "ljfam", "exv")
    out = O.pair_energy_forces(xyz, nbr, off, cell, name, params, need_param_grads=need)
    assert np.array_equal(out[0].numpy(), g["e_" + name])                    # same op chain -> same bits
    assert np.array_equal(out[1].numpy(), g["f_" + name])
    if need:
        np.testing.assert_allclose([x.item() for x in out[2]], g["dp_" + name], rtol=1e-6)


def test_c1_trajectory_vs_reference_fixture():
    g = _load("c1_traj.npz")
    n = 108
    cell = torch.tensor([3 * 1.679] * 3)
    v0, q0 = torch.Tensor(g["v0"]), torch.Tensor(g["q0"])
    mass, Qb = torch.full((n,), 1.008), O.nhc_bath_masses(50.0, n, 5)
    t = O.time_grid(0.01, 50)
    for reuse in (True, False):       # one force evaluation per step is bitwise the reference's two
        sysO = O.PairSystemOracle(cell, 2.5, "lj", (1.0, 1.0))
        v, q, pv = O.nh_verlet_trajectory(sysO.force, v0, q0, torch.zeros(5), t, mass, Qb, 1.0, 3 * n, reuse_force=reuse)
        assert np.array_equal(v.numpy(), g["v"]) and np.array_equal(q.numpy(), g["q"]) and np.array_equal(pv.numpy(), g["pv"])
        assert sysO.n_eval == (50 if reuse else 98)
    # reference bookkeeping: 98 forward evaluations (two per step, md.py:200-204) + 147 in the adjoint
    # backward pass the fixture script ran afterwards (three per step, sovlers.py:211-293)
    assert int(g["update_count"]) == 98 + 147
    # C restatement of the epoch (double accumulation of forces -> tolerance)
    dts = (t[1:] - t[:-1]).numpy()
    for k in (1, 10, 49):              # chaos amplifies the rounding differences with the step count
        vc, qc, pvc, _ = C.nhc_md(g["v0"], g["q0"], np.zeros(5), mass.numpy(), cell.numpy(), 2.5, 1.0, 1.0, Qb.numpy(),
                                  1.0, 3 * n, dts[:k])
        tol = {1: 2e-6, 10: 2e-5, 49: 2e-3}[k]
        assert np.abs(qc - g["q"][k]).max() < tol and np.abs(vc - g["v"][k]).max() < 10 * tol


def test_nve_trajectory_vs_reference_fixture():
    g = _load("c1_nve.npz")
    sysO = O.PairSystemOracle(torch.tensor([3 * 1.679] * 3), 2.5, "lj", (1.0, 1.0))
    v, q = O.verlet_trajectory(sysO.force, torch.Tensor(g["v0"]), torch.Tensor(g["q0"]), O.time_grid(0.005, 20))
    assert np.array_equal(v.numpy(), g["v"]) and np.array_equal(q.numpy(), g["q"])


def test_rdf_vs_reference_fixture():
    g = _load("c1_traj.npz")
    count, bins, gr = O.rdf(torch.tensor(g["q"][-1]), [3 * 1.679] * 3, 100, (0.75, 2.0))
    assert np.array_equal(bins.numpy(), g["rdf_bins"])
    np.testing.assert_allclose(count.numpy(), g["rdf_count"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(gr.numpy(), g["rdf_g"], rtol=1e-6, atol=1e-7)


def test_c_oracle_forces_vs_torch_oracle():
    pos, _, L = O.lj_system(8, jitter=0.05, seed=1)
    xyz, cell = torch.tensor(pos, dtype=torch.float32), torch.tensor([L] * 3, dtype=torch.float32)
    nbr, off = O.neighbor_list(xyz, 2.5, cell)
    e, f = O.pair_energy_forces(xyz, nbr, off, cell, "lj", (1.0, 1.0))
    ec, fc = C.lj_forces(xyz.numpy(), cell.numpy(), 2.5)
    assert abs(ec - e.item()) < 1e-5 * abs(e.item())
    assert np.abs(fc - f.numpy()).max() < 1e-5 * np.abs(f.numpy()).max()


def test_known_answers_fcc500():
    # FCC 5x5x5 at rho 0.8442 -> 13 500 pairs (54 neighbours per atom), E = -3386.684 (fp32 / fp64)
    a = (4 / 0.8442) ** (1 / 3)
    xyz = torch.tensor(O.fcc_positions(5, a), dtype=torch.float32)
    cell = torch.tensor([5 * a] * 3, dtype=torch.float32)
    nbr, off = O.neighbor_list(xyz, 2.5, cell)
    assert nbr.shape[0] == 13500 and int(torch.bincount(nbr.reshape(-1)).min()) == 54
    e = O.pair_energy_forces(xyz, nbr, off, cell, "lj", (1.0, 1.0))[0]
    assert abs(e.item() - (-3386.684)) < 2e-3
    ec, _ = C.lj_forces(xyz.numpy(), cell.numpy(), 2.5)
    assert abs(ec - (-3386.684)) < 2e-3


def test_bonded_members_vs_reference_fixture():
    """oracle restatement of BondPotentials / AnglePotentials / Electrostatics against the reference's outputs
    (tests/golden/bonded_chain.npz, oracle/make_golden.py --bonded)"""
    from mdgrad_b200._ase_compat import wrap_positions
    g = _load("bonded_chain.npz")
    cell = torch.tensor(g["cell"], dtype=torch.float32)
    kb, ro, ka, th0 = [float(x) for x in g["params"]]
    bt, at = torch.tensor(g["bond_top"]), torch.tensor(g["angle_top"])
    raw = torch.tensor(g["positions"], dtype=torch.float32)
    wrap = torch.tensor(wrap_positions(g["positions"], np.diag(g["cell"])), dtype=torch.float32)
    for tag, x in (("raw", raw), ("wrap", wrap)):
        for name, fn, args in (("bond", O.bond_energy, (bt, cell, kb, ro)), ("angle", O.angle_energy, (at, cell, ka, th0))):
            q = x.clone().requires_grad_(True)
            e = fn(q, *args)
            f = -torch.autograd.grad(e, q)[0]
            assert e.item() == float(g["e_%s_%s" % (name, tag)])                       # same ops, same bits
            assert np.array_equal(f.numpy(), g["f_%s_%s" % (name, tag)])
    q = wrap.clone().requires_grad_(True)
    e = O.coulomb_energy(q, torch.tensor(g["charges"]), cell, 2.5, float(g["coul_conversion"]), ex_pairs=g["bond_top"])
    np.testing.assert_allclose(e.item(), float(g["e_coul"]), rtol=2e-6)
    np.testing.assert_allclose((-torch.autograd.grad(e, q)[0]).numpy(), g["f_coul"], rtol=1e-5, atol=1e-5 * np.abs(g["f_coul"]).max())
