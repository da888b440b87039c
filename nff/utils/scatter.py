from mdgrad_b200.md import compute_grad  # noqa: F401
from mdgrad_b200.nffm.schnet import scatter_add  # noqa: F401
