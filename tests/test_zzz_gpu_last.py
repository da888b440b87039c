"""GPU: the opt-in CUDA-graph replay of the device engine's force evaluations (MDG_GNN_GRAPH=1) - first hardware run pending - and the end-to-end checks added in the same GPU-less session.
Kept in the file that sorts LAST, so that `-x` reaches every other GPU test first."""
import pytest

import bonded_checks as B

pytestmark = pytest.mark.gpu


def test_gpu_fold_force_field_graph_replay():
    B.check_fold_force_field_graph_replay("cuda")


def test_gpu_schnet_second_order_through_native_aggregation():
    import schnet_checks
    schnet_checks.check_second_order_through_native_aggregation("cuda")


def test_gpu_gnn_adjoint_fit_vs_reference_fixture():
    import schnet_checks
    schnet_checks.check_gnn_adjoint_fit_vs_reference_fixture("cuda")


def test_gpu_water_rdf_oo_species_selection():
    import schnet_checks
    schnet_checks.check_water_rdf_oo_species_selection("cuda")
