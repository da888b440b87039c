from mdgrad_b200.nffm.schnet import NodeMultiTaskReadOut, SchNetConv  # noqa: F401
