"""TEST INFRASTRUCTURE ONLY - never imported by the product path.

Imports the UNMODIFIED reference (torchmd/mdgrad, read-only at /root/reference)
in-process so that (a) the oracle restatement in `oracle/oracle_torch.py` can be
validated against it and (b) `oracle/make_golden.py` can generate the committed
fixtures under tests/golden/.  /root/reference does not exist on the GPU box, so
nothing that runs there (gpu tests, smoke, bench) may call this module.

The reference does `from ase import ...` / `from xitorch.interpolate import Interp1D`
at module top (torchmd/system.py:5-6, interface.py:7-8, md.py:5-7,
potentials.py:8-10); neither is installed here, so the ASE restatement in
`mdgrad_b200/_ase_compat.py` and a one-class xitorch stub are registered in
`sys.modules` first.  The repo's own `torchmd`/`nff` mirror packages share the
reference's package names, so the reference is loaded under a scrubbed
`sys.modules` and handed back as a namespace of module objects.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("MDGRAD_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "torchmd"))


_cache = {}


def load():
    """Return a namespace with the reference modules:
    .topology .interface .potentials .md .sovlers .tinydiffeq .system .observable
    .schnet (nff.nn.models.schnet) .nff_layers .nff_scatter"""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, repo_root)
    from mdgrad_b200 import _ase_compat
    _ase_compat.install_as_ase()
    if "xitorch" not in sys.modules:
        xi = types.ModuleType("xitorch")
        xin = types.ModuleType("xitorch.interpolate")

        class Interp1D:  # only reached by pairTab, which is out of scope
            def __init__(self, *a, **k):
                raise NotImplementedError("xitorch stub")
        xin.Interp1D = Interp1D
        xi.interpolate = xin
        sys.modules["xitorch"] = xi
        sys.modules["xitorch.interpolate"] = xin

    # stash this repo's same-named mirror packages, import the reference's, then restore
    stash = {k: v for k, v in sys.modules.items()
             if k == "torchmd" or k.startswith("torchmd.") or k == "nff" or k.startswith("nff.")}
    for k in stash:
        del sys.modules[k]
    saved_path = list(sys.path)
    sys.path[:] = [REF_ROOT] + [p for p in saved_path
                                if os.path.abspath(p or ".") != repo_root]
    try:
        ns = types.SimpleNamespace()
        ns.topology = importlib.import_module("torchmd.topology")
        ns.tinydiffeq = importlib.import_module("torchmd.tinydiffeq")
        ns.sovlers = importlib.import_module("torchmd.sovlers")
        ns.system = importlib.import_module("torchmd.system")
        ns.potentials = importlib.import_module("torchmd.potentials")
        ns.interface = importlib.import_module("torchmd.interface")
        ns.md = importlib.import_module("torchmd.md")
        ns.observable = importlib.import_module("torchmd.observable")
        ns.schnet = importlib.import_module("nff.nn.models.schnet")
        ns.nff_layers = importlib.import_module("nff.nn.layers")
        ns.nff_scatter = importlib.import_module("nff.utils.scatter")
        ref_mods = {k: v for k, v in sys.modules.items()
                    if k == "torchmd" or k.startswith("torchmd.") or k == "nff" or k.startswith("nff.")}
        ns._modules = ref_mods
    finally:
        for k in [k for k in sys.modules
                  if k == "torchmd" or k.startswith("torchmd.") or k == "nff" or k.startswith("nff.")]:
            del sys.modules[k]
        sys.modules.update(stash)
        sys.path[:] = saved_path
    _cache["ns"] = ns
    return ns


class active:
    """Context manager: temporarily expose the reference's `torchmd`/`nff` modules under
    their own names (the reference does a late `import torchmd` in system.py:11-14)."""

    def __enter__(self):
        ns = load()
        self._stash = {k: v for k, v in sys.modules.items()
                       if k == "torchmd" or k.startswith("torchmd.") or k == "nff" or k.startswith("nff.")}
        for k in self._stash:
            del sys.modules[k]
        sys.modules.update(ns._modules)
        return ns

    def __exit__(self, *exc):
        for k in list(sys.modules):
            if k == "torchmd" or k.startswith("torchmd.") or k == "nff" or k.startswith("nff."):
                del sys.modules[k]
        sys.modules.update(self._stash)
        return False
