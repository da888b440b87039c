from mdgrad_b200.nffm.schnet import SchNet, SchNetConv  # noqa: F401
