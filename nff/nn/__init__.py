from mdgrad_b200.nffm.schnet import (Dense, GaussianSmearing, MessagePassingModule, NodeMultiTaskReadOut,  # noqa: F401
                                     SchNetConv, shifted_softplus)
