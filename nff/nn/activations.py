from mdgrad_b200.nffm.schnet import shifted_softplus  # noqa: F401
