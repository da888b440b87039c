"""Bonded / electrostatic Stack members (BondPotentials, AnglePotentials, Electrostatics - SURVEY 8f-4) through the
CPU-emulated kernels against the reference fixture.  TEST INFRASTRUCTURE; the GPU run of the same bodies is
tests/test_gpu_zz_bonded.py."""
import pytest

import bonded_checks as B
from test_emu_api import emulated_backend  # noqa: F401  (autouse fixture: swaps the four binding hooks)


@pytest.mark.parametrize("check", [B.check_bond_angle_energy_force, B.check_bonded_param_grads, B.check_bonded_second_order_route,
                                   B.check_electrostatics, B.check_fold_stack_on_device_engine,
                                   B.check_fold_force_field_with_gnn, B.check_pair_tab_through_pair_potentials,
                                   B.check_fold_force_field_graph_replay, B.check_fold_engine_sync_equals_async, B.check_generic_route_configs,
                                   B.check_tpair_potentials_vs_reference_fixture, B.check_stack_adjoint_native_equals_autograd,
                                   B.check_bonded_edge_cases], ids=lambda f: f.__name__)
def test_emu_bonded(check):
    check("cpu")
