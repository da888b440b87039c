from mdgrad_b200.nffm.schnet import Dense, GaussianSmearing  # noqa: F401
