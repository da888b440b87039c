"""CPU, authoring container only: live diff of the oracle restatement and of the generic solvers
against the UNMODIFIED reference imported from /root/reference (skipped where it is absent)."""
import numpy as np
import pytest
import torch

from oracle import oracle_torch as O
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


def test_neighbor_list_live():
    ref = ref_import.load()
    rng = np.random.default_rng(5)
    for n, L, rc in [(300, 6.5, 2.5), (700, 14.0, 4.9)]:
        xyz = torch.tensor(rng.uniform(-0.4 * L, 1.4 * L, (n, 3)), dtype=torch.float32)
        cell = torch.tensor([L, 1.2 * L, 0.8 * L], dtype=torch.float32)
        a = ref.topology.generate_nbr_list(xyz, rc, cell, get_dis=True)
        b = O.neighbor_list(xyz, rc, cell, get_dis=True, block=128)
        assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_adjoint_solver_live():
    """My odeint_adjoint driving the REFERENCE's own NoseHooverChain module reproduces the reference's
    odeint_adjoint bit for bit (trajectory and d loss / d sigma, epsilon)."""
    from mdgrad_b200 import sovlers as S
    from mdgrad_b200._ase_compat import FaceCenteredCubic
    with ref_import.active() as ref:
        res = []
        for impl in (ref.sovlers.odeint_adjoint, S.odeint_adjoint):
            atoms = FaceCenteredCubic(symbol="H", size=(2, 2, 2), latticeconstant=1.679 * 1.5, pbc=True)
            system = ref.system.System(atoms, device="cpu")
            np.random.seed(0)
            system.set_temperature(1.0)
            lj = ref.potentials.LennardJones(1.0, 1.0)
            integ = ref.md.NoseHooverChain(ref.interface.PairPotentials(system, lj, cutoff=2.5), system, T=1.0,
                                           num_chains=5, Q=50.0, adjoint=True)
            t = torch.Tensor([0.01 * i for i in range(8)])
            v, q, pv = impl(integ, tuple(integ.get_inital_states(True)), t, method="NH_verlet")
            ((q[-1] ** 2).sum() + (v[3] * v[5]).sum() + pv[-1].sum()).backward()
            res.append((v.detach(), q.detach(), pv.detach(), lj.sigma.grad.clone(), lj.epsilon.grad.clone()))
        for a, b in zip(*res):
            assert torch.equal(a, b)


def test_vacf_live():
    """the vacf mirror (device-agnostic tensor algebra) returns the reference's un-normalised values bit for bit"""
    from mdgrad_b200 import observable as Ob
    from mdgrad_b200._ase_compat import FaceCenteredCubic
    from mdgrad_b200.system import System
    with ref_import.active() as ref:
        atoms = FaceCenteredCubic(symbol="H", size=(2, 2, 2), latticeconstant=1.679, pbc=True)
        vel = torch.randn(40, 32, 3, generator=torch.Generator().manual_seed(3))
        a = ref.observable.vacf(ref.system.System(atoms, device="cpu"), t_range=15)(vel)
    b = Ob.vacf(System(atoms, device="cpu"), t_range=15)(vel)
    assert a.shape == (15,) and torch.equal(a, b)


def test_angle_list_and_angles_live():
    """generate_angle_list enumerated from a CSR on the device == the reference's (2P x 2P) mask construction, row for
    row; compute_angle / the smeared angle histogram agree with the reference on the same list"""
    from mdgrad_b200 import observable as Ob
    from mdgrad_b200 import topology as T
    ref = ref_import.load()
    rng = np.random.default_rng(2)
    n, L = 60, 7.0
    xyz = torch.tensor(rng.uniform(0, L, (3, n, 3)), dtype=torch.float32)          # 3 frames
    cell = torch.tensor([L, L, L])
    frames = []
    for f in range(3):
        nb, _ = O.neighbor_list(xyz[f], 2.2, cell)
        frames.append(torch.cat([torch.full((nb.shape[0], 1), f, dtype=torch.int64), nb], 1))
    nbr = torch.cat(frames)
    a = ref.topology.generate_angle_list(nbr)
    b = T.generate_angle_list(nbr)
    assert a.shape[0] > 500 and torch.equal(a, b)
    ca = ref.observable.compute_angle(xyz, a, cell, N=n)
    cb = Ob.compute_angle(xyz, b, cell, N=n)
    assert torch.equal(ca, cb)
    assert T.generate_angle_list(nbr[:0]).shape == (0, 4)


@pytest.mark.parametrize("maker,fname", [("bonded_golden", "bonded_chain.npz"), ("gnn_adjoint_golden", "gnn_adjoint.npz"),
                                         ("generic_route_golden", "c1_generic.npz")])
def test_committed_fixtures_are_reproducible_from_the_reference(maker, fname, tmp_path, monkeypatch):
    """provenance of tests/golden/: re-running oracle/make_golden.py against the reference tree reproduces the committed arrays"""
    import os
    import sys
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import make_golden
    monkeypatch.setattr(make_golden, "OUT", str(tmp_path))
    getattr(make_golden, maker)()
    new = np.load(os.path.join(str(tmp_path), fname))
    old = np.load(os.path.join(os.path.dirname(__file__), "golden", fname))
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        assert new[k].shape == old[k].shape, k
        if new[k].dtype.kind == "f":
            np.testing.assert_allclose(new[k], old[k], rtol=1e-6, atol=1e-6 * max(1e-30, float(np.abs(old[k]).max())), err_msg=k)
        else:
            assert np.array_equal(new[k], old[k]), k
