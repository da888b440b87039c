"""GPU: API-level tests of the paths added after the round-1 GPU budget was spent (device-side epoch hand-off, analytic
adjoint route, Stack on the device engine).  Each mirrors a test that passes under the CPU emulation (tests/test_emu_api.py);
kept in a file that sorts after the suites that had already run on hardware, so that `-x` reaches those first."""
import os

import numpy as np
import pytest
import torch

from test_gpu_api import G, _fcc_system

pytestmark = pytest.mark.gpu


def test_device_wrap_bit_identical_on_gpu():
    """epoch hand-off: the fp64 wrap as separate CUDA tensor ops gives the bits of the host (ASE) wrap"""
    from mdgrad_b200.md import _device_wrap, _host_to_device
    from mdgrad_b200._ase_compat import wrap_positions
    rng = np.random.default_rng(5)
    for L in ([67.16363, 67.16363, 67.16363], [5.037, 6.1, 4.4]):
        q = (rng.uniform(-2.5, 3.5, (200000, 3)) * np.array(L)).astype(np.float32)
        q[:50] = np.array([0.0, L[1], -L[2]], dtype=np.float32)
        host = _host_to_device(wrap_positions(q, np.diag(L)), "cpu")
        dev = _device_wrap(torch.from_numpy(q).cuda(), np.diag(L))
        assert torch.equal(dev.cpu(), host)


def test_simulate_device_handoff_equals_host_roundtrip():
    from torchmd.interface import PairPotentials
    from torchmd.potentials import LennardJones
    from torchmd.md import NoseHooverChain, Simulations
    g = np.load(os.path.join(G, "c1_traj.npz"))
    runs = []
    for handoff in (True, False):
        system = _fcc_system()
        system.set_positions(g["q0"] + 7.0)
        system.set_velocities(g["v0"])
        integ = NoseHooverChain(PairPotentials(system, LennardJones(1.0, 1.0), cutoff=2.5), system, T=1.0, num_chains=5, Q=50.0)
        sim = Simulations(system, integ, wrap=True, method="NH_verlet")
        sim.device_handoff = handoff
        v, q, pv = sim.simulate(steps=4 * 6, frequency=6, dt=0.01)
        runs.append((sim.log, system.get_positions(), system.get_velocities(), q.detach()))
    a, b = runs
    for key in ("velocities", "positions", "baths"):
        assert len(a[0][key]) == 4
        for x, y in zip(a[0][key], b[0][key]):
            assert np.array_equal(x, y)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and torch.equal(a[3], b[3])


def _adjoint_grads_gpu(native, pot_factory, steps=8):
    from torchmd.interface import PairPotentials
    from torchmd.md import NoseHooverChain, Simulations
    g = np.load(os.path.join(G, "c1_traj.npz"))
    system = _fcc_system()
    system.set_positions(g["q0"])
    system.set_velocities(g["v0"])
    pot = pot_factory().cuda()
    integ = NoseHooverChain(PairPotentials(system, pot, cutoff=2.5), system, T=1.0, num_chains=5, Q=50.0, adjoint=True)
    integ.disable_native_adjoint = not native
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    v, q, pv = sim.simulate(steps=steps, frequency=steps, dt=0.01)
    ((q[-1] ** 2).sum() + (v[2] * v[4]).sum() + pv[-1].sum()).backward()
    return [p.grad.item() for p in pot.parameters()]


@pytest.mark.parametrize("name", ["lj", "buck"])
def test_adjoint_native_route_equals_autograd_route(name):
    """short horizon (chaos-free): the analytic reverse dynamics (mdg_pair_hvp + written-out thermostat algebra) equal the
    generic double-backward route of the reference"""
    from torchmd.potentials import Buck, LennardJones
    factory = (lambda: LennardJones(1.0, 1.0)) if name == "lj" else (lambda: Buck(900.0, 3.2, 1.5))
    a = _adjoint_grads_gpu(True, factory)
    b = _adjoint_grads_gpu(False, factory)
    for x, y in zip(a, b):
        assert abs(x - y) <= 3e-4 * max(1.0, abs(y)), (a, b)


def test_stack_of_pair_potentials_on_device_engine():
    """scripts/fit_2_comp.py shape: Stack of three species-pair PairPotentials under NoseHooverChain on the device engine
    (every member on its own exact per-step list) vs the op-level solver"""
    from torchmd.interface import PairPotentials, Stack
    from torchmd.potentials import LennardJones
    from torchmd.md import NoseHooverChain, Simulations
    from torchmd.sovlers import odeint_reuse_force
    system = _fcc_system(size=4)
    n = len(system)
    np.random.seed(1)
    system.set_temperature(1.0)
    A, B = list(range(0, n, 2)), list(range(1, n, 2))
    stack = Stack({"aa": PairPotentials(system, LennardJones(1.0, 1.0), cutoff=2.5, index_tuple=(A, A)),
                   "bb": PairPotentials(system, LennardJones(0.9, 0.8), cutoff=2.2, index_tuple=(B, B)),
                   "ab": PairPotentials(system, LennardJones(0.95, 1.1), cutoff=2.5, index_tuple=(A, B))})
    integ = NoseHooverChain(stack, system, T=1.0, num_chains=3, Q=20.0, adjoint=True)
    sim = Simulations(system, integ, wrap=True, method="NH_verlet")
    out = sim.simulate(steps=6, frequency=6, dt=0.005)
    assert integ.last_engine_stats is not None and integ.update_count == 10
    integ.disable_gnn_engine = True
    t = torch.Tensor([0.005 * i for i in range(6)]).cuda()
    with torch.no_grad():
        ref = odeint_reuse_force(integ, tuple(o[0].detach() for o in out), t, "NH_verlet")
    for a, b in zip(out, ref):
        assert (a.detach() - b).abs().max().item() <= 2e-6 * max(1.0, b.abs().max().item())
