from mdgrad_b200.nffm.schnet import SchNet  # noqa: F401
