"""`nff.train.get_model` of the reference (nff/train/builders/model.py:92-106), SchNet only - the one entry point the
MD scripts use (scripts/fit_rdf_gnn.py:138, demo/fit_rdf_gnn.py).  Training infrastructure (Trainer, hooks, loss builders)
is out of scope of this repository."""
from mdgrad_b200.nffm.schnet import SchNet

_REQUIRED = ("n_atom_basis", "n_filters", "n_gaussians", "n_convolutions", "cutoff")


def get_model(params, model_type="SchNet", **kwargs):
    if model_type != "SchNet":
        raise NotImplementedError("only the SchNet force field of the MD hot path is provided (got %r)" % (model_type,))
    missing = [k for k in _REQUIRED if k not in params]
    if missing:
        raise ValueError("Parameter(s) %s missing for SchNet" % missing)
    return SchNet(params, **kwargs)
