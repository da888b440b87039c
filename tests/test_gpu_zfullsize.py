"""GPU: size-independent properties at BASELINE configs[1]'s FULL size (256 000-atom LJ box, FCC 40^3, rho 0.845, rc 2.5) -
list layout invariants, forces of a block of rows against the C restatement of the reference, Newton's third law, skin-list
vs per-step-list invariance of the engine, momentum conservation, RDF kernel vs list-based histogram."""
import pytest
import torch

import fullsize_checks as F

pytestmark = pytest.mark.gpu
NCELL = 40


def _dev():
    return torch.device("cuda", 0)


def test_fullsize_list_structure(ctx):
    P = F.check_list_structure(ctx, _dev(), NCELL)
    assert 6_900_000 < P < 7_250_000          # P ~ 7.07 M at 256 000 atoms


def test_fullsize_list_rows_bit_exact(ctx):
    pairs, shifted = F.check_list_rows_bit_exact(ctx, _dev(), NCELL)
    assert pairs > 250_000 and shifted > 20_000


def test_fullsize_forces_vs_c_oracle_rows(ctx):
    F.check_forces_against_c_oracle(ctx, _dev(), NCELL)


def test_fullsize_engine_invariants(ctx):
    F.check_engine_invariants(ctx, _dev(), NCELL)


def test_fullsize_rdf_two_paths(ctx):
    F.check_rdf_two_paths(ctx, _dev(), NCELL)
