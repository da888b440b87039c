"""B200-backed mirror of the parts of the reference's `nff` package that the MD hot path uses
(SURVEY 2a #9/#10): nff.nn.{layers, activations, modules.SchNetConv, graphconv, graphop},
nff.nn.models.schnet.SchNet, nff.utils.{scatter, cuda.batch_to}."""
