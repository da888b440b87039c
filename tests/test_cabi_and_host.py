"""CPU: the C-ABI library loads and exports every symbol include/mdgrad_b200.h declares (no compute
calls without a GPU); host-side logic of the Python mirror (System, time grid, wrap, log semantics,
generic solvers + adjoint against reference fixtures); loud failure on CPU tensors."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mdgrad_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mdgrad_b200.h")).read()
    declared = set(re.findall(r"\b(mdg_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no symbols parsed from the header"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libmdgrad_b200.so does not export %s" % name
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    assert _lib.load().mdg_version() >= 100


def test_md_params_struct_matches_header_layout():
    from mdgrad_b200 import _lib
    # int, int, float[4], double, float[3], int, float[16], double, int, float, int, int
    assert ctypes.sizeof(_lib.MdParams) == 136
    assert _lib.MdParams.cutoff.offset == 24 and _lib.MdParams.T.offset == 112


def test_no_cpu_fallback():
    from mdgrad_b200 import _lib
    from torchmd.topology import generate_nbr_list, compute_dis
    xyz = torch.rand(10, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        generate_nbr_list(xyz, 2.5, torch.tensor([5.0, 5.0, 5.0]))
    with pytest.raises(RuntimeError, match="CUDA"):
        _lib.Context("cpu")
    with pytest.raises(RuntimeError):
        compute_dis(xyz, torch.zeros(1, 2, dtype=torch.long), torch.zeros(1, 3), torch.tensor([5.0, 5.0, 5.0]))


def test_triclinic_cell_is_rejected_loudly():
    from torchmd.topology import cell_lengths
    with pytest.raises(NotImplementedError):
        cell_lengths(torch.tensor([[5.0, 0.5, 0.0], [0.0, 5.0, 0.0], [0.0, 0.0, 5.0]]))
    assert cell_lengths(torch.tensor([4.0, 5.0, 6.0])) == [4.0, 5.0, 6.0]


def test_system_and_ase_compat():
    from torchmd.system import System, check_system
    from mdgrad_b200._ase_compat import FaceCenteredCubic, Diamond, wrap_positions, units
    atoms = FaceCenteredCubic(symbol="H", size=(3, 3, 3), latticeconstant=1.679, pbc=True)
    s = System(atoms, device="cpu")
    assert len(s) == 108 and s.get_nxyz().shape == (108, 4) and s.dim == 3
    assert np.allclose(s.get_cell_len(), 3 * 1.679)
    assert s.get_batch()["num_atoms"].item() == 108
    check_system(s)
    with pytest.raises(TypeError):
        check_system(atoms)
    np.random.seed(0)
    s.set_temperature(1.0)
    ke = 0.5 * (s.get_masses()[:, None] * s.get_velocities() ** 2).sum()
    assert 0.6 < ke / (1.5 * 108) < 1.4
    assert len(Diamond("Si", (2, 2, 2), 5.43)) == 64
    w = wrap_positions(np.array([[-0.1, 5.2, 2.0]]), np.diag([5.0, 5.0, 5.0]))
    assert np.allclose(w, [[4.9, 0.2, 2.0]])
    assert abs(units.fs - 0.09822694788) < 1e-9


def test_atoms_defers_the_fp64_conversion_of_logged_frames():
    """set_positions / set_velocities with an fp32 frame (what Simulations.update_states hands over, reference md.py:54-58) keep
    a reference and convert at the first read; every reader sees exactly the eager values."""
    from mdgrad_b200._ase_compat import Atoms
    rng = np.random.default_rng(3)
    a = Atoms(numbers=[1, 8, 1, 8, 1], positions=rng.random((5, 3)), cell=[3.0, 3.0, 3.0], pbc=True)
    eager = Atoms(a)
    p32, v32 = (rng.random((5, 3)) * 7 - 2).astype(np.float32), rng.standard_normal((5, 3)).astype(np.float32)
    a.set_positions(p32)
    a.set_velocities(v32)
    assert a.__dict__["_pos_src"] is p32 and a.__dict__["_vel_src"] is v32          # nothing converted yet
    eager.positions = np.array(p32, dtype=float)
    eager._momenta = np.asarray(v32, dtype=float) * eager.get_masses()[:, None]
    assert a.get_positions().dtype == np.float64 and np.array_equal(a.get_positions(), eager.get_positions())
    assert a.__dict__["_pos_src"] is None
    assert np.array_equal(a.get_momenta(), eager.get_momenta()) and np.array_equal(a.get_velocities(), eager.get_velocities())
    assert a.get_kinetic_energy() == eager.get_kinetic_energy()
    a.set_positions(p32)
    assert np.array_equal(a.get_positions(wrap=True), eager.get_positions(wrap=True))
    a.set_positions(p32)
    assert np.array_equal(Atoms(a).positions, eager.positions)                        # the copy constructor reads through it
    a.set_velocities(v32)
    a.set_masses([2.0] * 5)                                                            # momenta were defined with the OLD masses
    assert np.array_equal(a.get_momenta(), eager.get_momenta())
    a.set_positions([[0.0, 0.0, 0.0]] * 5)                                             # anything else converts eagerly, as before
    assert a.__dict__["_pos_src"] is None and a.get_positions().sum() == 0.0


def test_time_grid_and_simulate_bookkeeping():
    """frequency grid points -> frequency-1 steps; fp32 grid; log keeps the last frame per epoch."""
    from torchmd.md import Simulations
    from torchmd.system import System
    from mdgrad_b200._ase_compat import Atoms

    class Free(torch.nn.Module):          # free flight: dv/dt = 0, dq/dt = v, through the generic solver
        adjoint, state_keys = False, ["velocities", "positions"]

        def __init__(self, system):
            super().__init__()
            self.system = system

        def forward(self, t, state):
            return (torch.zeros_like(state[0]), state[0])

        def get_inital_states(self, wrap=True):
            return [torch.Tensor(self.system.get_velocities()), torch.Tensor(self.system.get_positions(wrap=wrap))]

    s = System(Atoms(numbers=[1, 1], positions=[[0.5, 0.5, 0.5], [1.0, 1.0, 1.0]], cell=[4.0] * 3, pbc=True), device="cpu")
    s.set_velocities(np.array([[1.0, 0, 0], [0, -2.0, 0]]))
    sim = Simulations(s, Free(s), wrap=True, method="verlet")
    v, q = sim.simulate(steps=20, frequency=5, dt=0.1)
    assert v.shape == (5, 2, 3) and len(sim.log["positions"]) == 4                # 4 epochs x 4 steps
    assert np.allclose(s.get_positions()[0], [0.5 + 1.6, 0.5, 0.5], atol=1e-5)
    assert np.allclose(sim.get_check_point()[1][1].numpy(), [1.0, (1.0 - 3.2) % 4.0, 1.0], atol=1e-5)   # wrapped on the host
    v, q = sim.simulate(steps=3, frequency=1, dt=0.1)                              # frequency=1 integrates zero steps
    assert v.shape == (1, 2, 3) and len(sim.log["positions"]) == 7
    with pytest.raises(UnboundLocalError):
        sim.simulate(steps=2, frequency=5, dt=0.1)                                  # zero epochs, as the reference


def test_generic_solvers_with_toy_module():
    from torchmd.sovlers import odeint, odeint_adjoint

    class Spring(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.k = torch.nn.Parameter(torch.tensor([2.0]))

        def forward(self, t, y):
            v, q = y
            return (-self.k * q, v)

    f = Spring()
    t = torch.Tensor([0.01 * i for i in range(101)])
    v, q = odeint_adjoint(f, (torch.tensor([0.0]), torch.tensor([1.0])), t, method="verlet")
    w = np.sqrt(2.0)
    assert abs(q[-1].item() - np.cos(w * 1.0)) < 1e-4
    q[-1].sum().backward()
    # d cos(sqrt(k) T)/dk = -sin(sqrt(k) T) T / (2 sqrt(k))
    assert abs(f.k.grad.item() - (-np.sin(w) / (2 * w))) < 6e-3     # the reference's reverse-midpoint adjoint scheme is only ~1% accurate at dt=0.01
    v2, q2 = odeint(f, (torch.tensor([0.0]), torch.tensor([1.0])), t, method="rk4")
    assert abs(q2[-1].item() - np.cos(w)) < 1e-6
    with pytest.raises(KeyError):
        odeint(f, (torch.tensor([0.0]), torch.tensor([1.0])), t, method="dopri5")


def test_potential_modules_match_reference_formulas():
    from torchmd import potentials as P
    r = torch.linspace(0.8, 2.5, 50)[:, None]
    assert torch.allclose(P.LennardJones(1.1, 0.7)(r), 4 * 0.7 * ((1.1 / r) ** 12 - (1.1 / r) ** 6))
    assert torch.allclose(P.ExcludedVolume(1.0, 0.5, 12)(r), 4 * 0.5 * (1.0 / r) ** 12)
    assert torch.allclose(P.Buck(10.0, 2.0, 3.0)(r), 10.0 * torch.exp(-2.0 * r) - 3.0 / r ** 6)
    assert P.LennardJones().native_spec()[0] == 0 and list(P.LennardJones().state_dict()) == ["sigma", "epsilon"]
    m = P.pairMLP(8, 0.5, 2.5, 1, 16, "ELU")
    assert m(r).shape == (50, 1)
    from torchmd.interface import PairPotentials
    assert P.PairPotentials is PairPotentials          # exported from both modules (SURVEY naming trap)


def test_force_reuse_loop_is_bitwise_the_two_evaluation_solver():
    """odeint_reuse_force (one force evaluation per step) == odeint (two, like the reference), bit for bit,
    for NoseHooverChain and NVE over a CPU-capable toy interaction; bookkeeping counts two per step."""
    from torchmd.md import NVE, NoseHooverChain
    from torchmd.sovlers import odeint, odeint_reuse_force
    from torchmd.system import System
    from mdgrad_b200._ase_compat import Atoms

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.k = torch.nn.Parameter(torch.tensor([1.3]))
            self.resets = 0

        def _reset_topology(self, q):
            self.resets += 1

        def forward(self, q):
            d = q[:, None, :] - q[None, :, :]
            r2 = (d ** 2).sum(-1) + torch.eye(q.shape[0])
            return (self.k * torch.exp(-r2)).sum() + 0.1 * (q ** 2).sum()

    rng = np.random.default_rng(0)
    s = System(Atoms(numbers=[1] * 12, positions=rng.uniform(0, 3, (12, 3)), cell=[3.0] * 3, pbc=True), device="cpu")
    s.set_velocities(rng.normal(0, 1, (12, 3)))
    t = torch.Tensor([0.01 * i for i in range(9)])
    for cls, method, kw in ((NoseHooverChain, "NH_verlet", dict(T=1.0, num_chains=3, Q=5.0)), (NVE, "verlet", {})):
        a_model, b_model = Toy(), Toy()
        a, b = cls(a_model, s, **kw), cls(b_model, s, **kw)
        ya = tuple(x.clone() for x in a.get_inital_states(True))
        yb = tuple(x.clone() for x in b.get_inital_states(True))
        with torch.no_grad():
            ra = odeint(a, ya, t, method=method)
            rb = odeint_reuse_force(b, yb, t, method)
        assert rb is not None and all(torch.equal(x, y) for x, y in zip(ra, rb))
        assert a.update_count == b.update_count == 16
        assert a_model.resets == 16 and b_model.resets == 9          # half the neighbor-list rebuilds
    c = NoseHooverChain(Toy(), s, T=1.0, num_chains=3, Q=5.0, topology_update_freq=3)
    assert odeint_reuse_force(c, tuple(c.get_inital_states(True)), t, "NH_verlet") is None


def _wrap_published(positions, cell, pbc=(True, True, True), center=(0.5, 0.5, 0.5), eps=1e-7):
    """ASE 3.20 `wrap_positions` as published (solve + per-column `%=`): the checker for the vectorised product path."""
    shift = np.asarray(center, dtype=float) - 0.5 - eps
    pbc = np.asarray(pbc, dtype=bool)
    shift[~pbc] = 0.0
    fractional = np.linalg.solve(np.asarray(cell, dtype=float).T, np.asarray(positions, dtype=float).T).T - shift
    for i, periodic in enumerate(pbc):
        if periodic:
            fractional[:, i] %= 1.0
            fractional[:, i] += shift[i]
    return np.dot(fractional, cell)


def test_wrap_positions_vectorised_path_matches_the_published_formula():
    from mdgrad_b200._ase_compat import wrap_positions
    rng = np.random.default_rng(7)
    for trial in range(4):
        L = rng.uniform(2.0, 90.0, 3)
        cell = np.diag(L)
        pos = (rng.random((20000, 3)) * 7 - 3) * L          # several boxes either side
        pos[:8, 0] = [-0.0, 0.0, -1e-20, 1e-20, -L[0], L[0], 2.5 * L[0], -1e-300]
        for pbc in [(True, True, True), (True, False, True)]:
            got = wrap_positions(pos, cell, pbc=pbc)
            ref = _wrap_published(pos, cell, pbc=pbc)
            # the state that enters the device is the fp32 rounding of this array: it must not change; in fp64 allow the
            # last bit (LAPACK's triangular solve may divide or multiply by the reciprocal, depending on the BLAS build)
            assert np.array_equal(got.astype(np.float32), ref.astype(np.float32))
            assert np.max(np.abs(got - ref)) <= 4 * np.finfo(float).eps * L.max()
    # a sheared cell takes the published route itself
    cell = np.array([[10.0, 0, 0], [2.0, 9.0, 0], [0, 1.0, 8.0]])
    pos = rng.random((100, 3)) * 30 - 10
    assert np.array_equal(wrap_positions(pos, cell), _wrap_published(pos, cell))


def test_host_to_device_is_the_legacy_constructor_value_for_value():
    from mdgrad_b200.md import _host_to_device
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5000, 3)) * 50
    assert torch.equal(_host_to_device(x, "cpu"), torch.Tensor(x))
    x32 = x.astype(np.float32)[::2]                            # non-contiguous fp32 view
    assert torch.equal(_host_to_device(x32, "cpu"), torch.Tensor(x32))
    assert torch.equal(_host_to_device([0.0, 1.5], "cpu"), torch.Tensor([0.0, 1.5]))


def test_nhc_vjp_algebra_vs_autograd():
    """The written-out vector-Jacobian products of the Nose-Hoover-chain derivative (used by the analytic adjoint
    route, md.py nhc_vjp_algebra) against torch autograd of the oracle's restatement of md.py:221-240."""
    from mdgrad_b200.md import nhc_vjp_algebra
    from oracle import oracle_torch as O
    torch.manual_seed(0)
    for M in (2, 3, 5):
        n = 7
        v = torch.randn(n, 3, dtype=torch.float64, requires_grad=True)
        pv = torch.randn(M, dtype=torch.float64, requires_grad=True)
        F = torch.randn(n, 3, dtype=torch.float64)
        mass = torch.rand(n, dtype=torch.float64) + 0.5
        Qb = torch.rand(M, dtype=torch.float64) + 0.5
        dv, dq, dpv = O.nhc_derivative(v, F, pv, mass, Qb, 0.7, 3 * n)
        cv, cq, cp = torch.randn(n, 3, dtype=torch.float64), torch.randn(n, 3, dtype=torch.float64), torch.randn(M, dtype=torch.float64)
        gv_o, gpv_o = torch.autograd.grad((dv * cv).sum() + (dq * cq).sum() + (dpv * cp).sum(), (v, pv))
        a_q, gv, gpv = nhc_vjp_algebra(v.detach(), pv.detach(), mass[:, None], Qb, cv, cq, cp)
        torch.testing.assert_close(gv, gv_o, rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(gpv, gpv_o, rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(a_q, cv / mass[:, None])       # dv = F/m + ...: the force enters through c_v / m


def test_device_wrap_is_bit_identical():
    """The epoch hand-off wraps positions with separate fp64 tensor ops instead of the host numpy path: same bits."""
    from mdgrad_b200.md import _device_wrap, _host_to_device
    from mdgrad_b200._ase_compat import wrap_positions
    rng = np.random.default_rng(5)
    for L in ([67.16363, 67.16363, 67.16363], [5.037, 6.1, 4.4]):
        q = (rng.uniform(-2.5, 3.5, (20000, 3)) * np.array(L)).astype(np.float32)
        q[:50] = np.array([0.0, L[1], -L[2]], dtype=np.float32)          # exact boundaries
        q[50:60] = np.float32(-1e-7)
        host = _host_to_device(wrap_positions(q, np.diag(L)), "cpu")
        dev = _device_wrap(torch.from_numpy(q), np.diag(L))
        assert dev.dtype == torch.float32 and torch.equal(dev, host)
    assert _device_wrap(torch.zeros(4, 3), np.array([[5.0, 0.1, 0], [0, 5.0, 0], [0, 0, 5.0]])) is None


def test_solver_classes_match_odeint():
    """NHVerlet / Verlet / RK4 class forms (reference sovlers.py:11-19, tinydiffeq.py:88-95) = the functional grid loop"""
    from torchmd.sovlers import NHVerlet, RK4, Verlet, odeint
    from torchmd import tinydiffeq
    assert tinydiffeq.RK4 is RK4 and hasattr(tinydiffeq, "FixedGridODESolver")
    torch.manual_seed(0)
    v0, q0 = torch.randn(5, 3), torch.randn(5, 3)
    t = torch.linspace(0, 0.1, 6)
    f2 = lambda tt, y: (-y[1], y[0])                                           # noqa: E731
    f3 = lambda tt, y: (-y[1] - y[2][0] * y[0], y[0], torch.stack([y[0].pow(2).sum() - 1.0, -y[2][0]]))   # noqa: E731
    for cls, name, func, y0 in ((Verlet, "verlet", f2, (v0, q0)), (RK4, "rk4", f2, (v0, q0)),
                                (NHVerlet, "NH_verlet", f3, (v0, q0, torch.zeros(2)))):
        a = cls(func, y0).integrate(t)
        b = odeint(func, y0, t, method=name)
        assert all(torch.equal(x, y) for x, y in zip(a, b))
    with pytest.raises(NotImplementedError):
        Verlet(f2, (v0, q0), step_size=0.01)


def test_compute_dihe_matches_reference_formula():
    """observable.compute_dihe (reference observable.py:181-197) without the (F,N,N,3) tensor: same cosines"""
    from torchmd.observable import compute_dihe
    torch.manual_seed(1)
    x = torch.randn(6, 10, 3)
    dihes = torch.tensor([[0, 1, 2, 3], [4, 5, 6, 7], [2, 5, 8, 9], [3, 2, 1, 0]])
    D = x[:, :, None, :] - x[:, None, :, :]                                   # D[f, i, j] = x_i - x_j
    c1 = torch.cross(D[:, dihes[:, 1], dihes[:, 0]], D[:, dihes[:, 1], dihes[:, 2]], dim=-1)
    c2 = torch.cross(D[:, dihes[:, 2], dihes[:, 1]], D[:, dihes[:, 2], dihes[:, 3]], dim=-1)
    ref = (c1 * c2).sum(-1) / (c1.pow(2).sum(-1) * c2.pow(2).sum(-1)).sqrt()
    torch.testing.assert_close(compute_dihe(x, dihes), ref, rtol=1e-6, atol=1e-6)


def test_pair_tab_is_the_not_a_knot_cubic_spline():
    """pairTab (reference potentials.py:152-160 -> xitorch Interp1D, an absent dependency: parity unpinned) restates the
    documented default, a not-a-knot cubic spline; anchored on scipy's implementation of the same published algorithm"""
    from scipy.interpolate import CubicSpline
    from torchmd.potentials import pairTab
    torch.manual_seed(0)
    for nb in (4, 7, 200):
        p = pairTab(nbins=nb, rc=2.5)
        assert p.tab.shape == (nb,) and isinstance(p.tab, torch.nn.Parameter) and torch.equal(p.x, torch.linspace(0.0, 2.5, nb))
        with torch.no_grad():
            p.tab.copy_(torch.sin(3 * p.x) * torch.exp(-p.x) + 0.05 * torch.randn(nb))
        r = (torch.rand(300, 1) * 2.5).requires_grad_(True)
        u = p(r)
        assert u.shape == (300, 1)
        cs = CubicSpline(p.x.double().numpy(), p.tab.detach().double().numpy())
        rr = r.detach().double().numpy().ravel()
        assert np.abs(u.detach().numpy().ravel() - cs(rr)).max() < 2e-6
        du = torch.autograd.grad(u.sum(), r, create_graph=True)[0]
        assert np.abs(du.detach().numpy().ravel() - cs(rr, 1)).max() < 2e-4 * max(1.0, np.abs(cs(rr, 1)).max())
        g = torch.autograd.grad(du.pow(2).sum(), p.tab)[0]                     # second order w.r.t. the table: adjoint route
        assert torch.isfinite(g).all() and float(g.abs().max()) > 0
    # interpolates the knots exactly
    p = pairTab(nbins=12, rc=3.0)
    with torch.no_grad():
        p.tab.copy_(torch.randn(12))
    torch.testing.assert_close(p(p.x[:, None]).squeeze(-1), p.tab.detach(), rtol=1e-5, atol=1e-6)


def test_3xtf32_split_reaches_fp32_parity_model():
    """Numerical model of the tcgen05 dense-layer design (csrc/schnet_tc.cuh): x = hi + lo with hi = tf32(x), lo = tf32(x - hi),
    A B ~ Ah Bh + Ah Bl + Al Bh accumulated in fp32.  On SchNet-sized layers (K = 512) the result must sit within the 1e-5
    parity bar of the fp32 product, while a single TF32 product does not - which is why the kernel issues three MMAs."""
    def tf32(x):                                    # cvt.rna.tf32.f32: round to nearest, ties away, 10 explicit mantissa bits
        b = x.view(np.uint32).astype(np.uint64)
        b = (b + 0x1000) & 0xFFFFE000
        return b.astype(np.uint32).view(np.float32)
    rng = np.random.default_rng(0)
    A = rng.standard_normal((256, 512)).astype(np.float32)
    B = (rng.standard_normal((512, 256)) / np.sqrt(512)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    Ah, Bh = tf32(A), tf32(B)
    Al, Bl = tf32(A - Ah), tf32(B - Bh)
    one = (Ah @ Bh).astype(np.float32)
    three = (Ah @ Bh + Ah @ Bl + Al @ Bh).astype(np.float32)
    scale = np.abs(ref).max()
    assert np.abs(one - ref).max() / scale > 1e-4                    # a single TF32 pass misses the bar by > 10x
    assert np.abs(three - ref).max() / scale < 2e-6                  # the 3-pass split is at fp32 accumulation noise
    fp32 = (A @ B)
    assert np.abs(three - ref).max() < 4 * np.abs(fp32 - ref).max() + 1e-7 * scale


def test_header_is_plain_c(tmp_path):
    """include/mdgrad_b200.h is the C ABI: it must compile as C99 without any C++ / CUDA / torch type"""
    import subprocess
    src = tmp_path / "h.c"
    src.write_text('#include "mdgrad_b200.h"\nint main(void) { mdg_bonded_terms t; mdg_gnn_md_params p; (void)t; (void)p; return MDG_OK; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
