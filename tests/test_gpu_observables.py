"""GPU: f2 observables on the device against the reference fixture (bodies in tests/observable_checks.py)."""
import pytest
import torch

import observable_checks as C

pytestmark = pytest.mark.gpu


def test_vacf_kernel_and_algebra_vs_reference_fixture():
    C.check_vacf(0, torch.device("cuda", 0))


def test_angle_distribution_vs_reference_fixture():
    C.check_angle_distribution(0, torch.device("cuda", 0))


def test_temperature_and_virial_pressure():
    C.check_temperature_and_pressure(0, torch.device("cuda", 0))
