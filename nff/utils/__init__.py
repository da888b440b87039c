from mdgrad_b200.gnn import batch_to  # noqa: F401
