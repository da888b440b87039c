"""Drop-in package: the reference's `torchmd.*` module names backed by mdgrad_b200 (sm_100a).

`from torchmd.system import System`, `from torchmd.interface import PairPotentials`,
`from torchmd.md import Simulations, NoseHooverChain`, ... resolve to the B200-native
implementation (SURVEY.md 8b Python surface).
"""
import importlib
import sys

_MAP = ["system", "topology", "potentials", "interface", "md", "sovlers", "observable", "thermo"]
for _name in _MAP:
    _mod = importlib.import_module("mdgrad_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod
sys.modules[__name__ + ".tinydiffeq"] = sys.modules[__name__ + ".sovlers"]
