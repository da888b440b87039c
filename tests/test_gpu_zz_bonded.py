"""GPU: bonded / electrostatic Stack members against the reference fixture (same bodies as tests/test_emu_bonded.py).
Sorted after the suites that already ran on hardware."""
import pytest

import bonded_checks as B

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("check", [B.check_bond_angle_energy_force, B.check_bonded_param_grads, B.check_bonded_second_order_route,
                                   B.check_electrostatics, B.check_fold_stack_on_device_engine,
                                   B.check_fold_force_field_with_gnn, B.check_pair_tab_through_pair_potentials,
                                   B.check_fold_engine_sync_equals_async, B.check_generic_route_configs,
                                   B.check_tpair_potentials_vs_reference_fixture, B.check_stack_adjoint_native_equals_autograd,
                                   B.check_bonded_edge_cases], ids=lambda f: f.__name__)
def test_gpu_bonded(check):
    check("cuda")
