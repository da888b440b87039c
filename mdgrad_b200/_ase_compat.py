"""Minimal stand-in for the subset of ASE that the MD hot path touches.

The reference state container `System` derives from `ase.Atoms`
(reference torchmd/system.py:16) and the driver calls
`ase.geometry.wrap_positions` and `ase.units` (reference torchmd/md.py:5-7).
ASE (pinned 3.20.1, reference requirements.txt:1) is an un-vendored third-party
dependency that is not installed in this image, so this module restates the
handful of published ASE 3.20 behaviours the path relies on:

* `Atoms`: numpy fp64 positions/momenta/cell/masses container,
* `units`: CODATA-2014 derived eV/Angstrom/amu unit system (ASE's default),
* `wrap_positions(pos, cell, center=(.5,.5,.5), eps=1e-7)`,
* `MaxwellBoltzmannDistribution(atoms, temp)`,
* `FaceCenteredCubic` / `Diamond` cubic lattice builders.

When a real ASE is importable, `mdgrad_b200.system.System` derives from it
instead and nothing here is used.  Parity note: no reference test pins these
host-side fp64 formulas ("parity unpinned" for the ASE edge).
"""
import math
import types

import numpy as np

# ----------------------------------------------------------------------------
# units (ASE `ase.units`, CODATA 2014 = ASE 3.20 default)
# ----------------------------------------------------------------------------
_c = 299792458.0
_mu0 = 4.0e-7 * math.pi
_Grav = 6.67408e-11
_hplanck = 6.626070040e-34
_e = 1.6021766208e-19
_me = 9.10938356e-31
_mp = 1.672621898e-27
_Nav = 6.022140857e23
_k = 1.38064852e-23
_amu = 1.660539040e-27

units = types.ModuleType("ase.units")
units.eV = 1.0
units.Ang = units.Angstrom = 1.0
units.nm = 10.0
units.kB = _k / _e
units.J = 1.0 / _e
units.kJ = 1000.0 * units.J
units.kcal = 4.184 * units.kJ
units.mol = _Nav
units.m = 1e10
units.kg = 1.0 / _amu
units.C = 1.0 / _e
units.second = units.s = 1e10 * math.sqrt(_e / _amu)
units.fs = 1e-15 * units.second
units._e, units._amu, units._k = _e, _amu, _k

# standard atomic weights for the species the configs use (ASE `ase.data.atomic_masses`)
_MASSES = {1: 1.008, 2: 4.002602, 6: 12.011, 7: 14.007, 8: 15.999, 14: 28.085,
           18: 39.948}
_SYMBOLS = {"H": 1, "He": 2, "C": 6, "N": 7, "O": 8, "Si": 14, "Ar": 18}
_NUM2SYM = {v: k for k, v in _SYMBOLS.items()}


def _symbols_to_numbers(symbols, n=None):
    if isinstance(symbols, str):
        # either a single symbol ('H') or a formula like 'H2O' / 'Si8'; only the
        # plain repeated-symbol form is needed by the path
        import re
        out = []
        for sym, cnt in re.findall(r"([A-Z][a-z]?)(\d*)", symbols):
            out += [_SYMBOLS[sym]] * (int(cnt) if cnt else 1)
        return np.array(out, dtype=int)
    return np.array([_SYMBOLS[s] if isinstance(s, str) else int(s) for s in symbols],
                    dtype=int)


class Atoms:
    """fp64 numpy state container with the `ase.Atoms` calls the path makes
    (reference torchmd/system.py:39-66, torchmd/md.py:54-66,150-156,242-249)."""

    def __init__(self, symbols=None, positions=None, numbers=None, cell=None,
                 pbc=None, momenta=None, masses=None, velocities=None):
        if isinstance(symbols, Atoms):
            src = symbols
            numbers = src.numbers.copy() if numbers is None else numbers
            positions = src.positions.copy() if positions is None else positions
            cell = src.cell.copy() if cell is None else cell
            pbc = src.pbc.copy() if pbc is None else pbc
            momenta = src._momenta.copy() if momenta is None else momenta
            masses = src._masses.copy() if masses is None else masses
            symbols = None
        if numbers is None and symbols is not None:
            numbers = _symbols_to_numbers(symbols)
        if numbers is None:
            numbers = np.zeros(0 if positions is None else len(positions), dtype=int)
        self.numbers = np.asarray(numbers, dtype=int).copy()
        n = len(self.numbers)
        self.positions = (np.zeros((n, 3)) if positions is None
                          else np.array(positions, dtype=float).reshape(n, 3))
        self.set_cell(np.zeros((3, 3)) if cell is None else cell)
        self.set_pbc(False if pbc is None else pbc)
        if masses is None:
            masses = np.array([_MASSES.get(int(z), 1.0) for z in self.numbers], dtype=float)
        self._masses = np.asarray(masses, dtype=float).copy()
        self._momenta = np.zeros((n, 3))
        if momenta is not None:
            self.set_momenta(momenta)
        if velocities is not None:
            self.set_velocities(velocities)

    # -- lazily converted state -------------------------------------------
    # `positions` / `_momenta` are fp64 (n, 3) arrays, as in ase.Atoms.  set_positions / set_velocities with an array that is
    # not fp64 already (the fp32 log frame Simulations.update_states hands over after every epoch, reference md.py:54-58) keep
    # a reference to it and convert at the first READ: the same values as the eager np.array(..., dtype=float), without two
    # 6 MB single-threaded conversions per epoch at 256 000 atoms (5 ms of a 19 ms simulate() call in the r02 e2e profile).
    # The source arrays are log entries, which nobody mutates.
    @property
    def positions(self):
        src = self.__dict__.get("_pos_src")
        if src is not None:
            self.__dict__["_pos"] = np.array(src, dtype=float).reshape(len(self.numbers), 3)
            self.__dict__["_pos_src"] = None
        return self.__dict__["_pos"]

    @positions.setter
    def positions(self, value):
        self.__dict__["_pos"] = value
        self.__dict__["_pos_src"] = None

    @property
    def _momenta(self):
        src = self.__dict__.get("_vel_src")
        if src is not None:
            self.__dict__["_mom"] = np.asarray(src, dtype=float).reshape(len(self.numbers), 3) * self._masses[:, None]
            self.__dict__["_vel_src"] = None
        return self.__dict__["_mom"]

    @_momenta.setter
    def _momenta(self, value):
        self.__dict__["_mom"] = value
        self.__dict__["_vel_src"] = None

    @staticmethod
    def _deferrable(a, n):
        return isinstance(a, np.ndarray) and a.dtype == np.float32 and a.size == 3 * n

    # -- geometry ----------------------------------------------------------
    def set_cell(self, cell):
        cell = np.array(cell, dtype=float)
        if cell.shape == (3,):
            cell = np.diag(cell)
        elif cell.shape == ():
            cell = np.eye(3) * float(cell)
        self.cell = cell.reshape(3, 3)

    def get_cell(self):
        return self.cell.copy()

    def set_pbc(self, pbc):
        if isinstance(pbc, (bool, np.bool_, int)):
            pbc = (bool(pbc),) * 3
        self.pbc = np.array(pbc, dtype=bool)

    def get_pbc(self):
        return self.pbc.copy()

    def get_volume(self):
        return abs(float(np.linalg.det(self.cell)))

    # -- per-atom arrays ---------------------------------------------------
    def __len__(self):
        return len(self.numbers)

    def get_number_of_atoms(self):
        return len(self.numbers)

    get_global_number_of_atoms = get_number_of_atoms

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_chemical_symbols(self):
        return [_NUM2SYM.get(int(z), "X") for z in self.numbers]

    def get_masses(self):
        return self._masses.copy()

    def set_masses(self, masses):
        _ = self._momenta                                # (pending velocities are momenta with the masses of THEIR call)
        self._masses = np.array(masses, dtype=float).reshape(len(self))

    def get_positions(self, wrap=False, **wrap_kw):
        if wrap:
            if "pbc" not in wrap_kw:
                wrap_kw["pbc"] = self.pbc
            return wrap_positions(self.positions, self.cell, **wrap_kw)
        return self.positions.copy()

    def set_positions(self, pos):
        if self._deferrable(pos, len(self)):
            self.__dict__["_pos_src"] = pos              # converted at the first read (see `positions`)
            return
        self.positions = np.array(pos, dtype=float).reshape(len(self), 3)

    def get_momenta(self):
        return self._momenta.copy()

    def set_momenta(self, p):
        self._momenta = np.array(p, dtype=float).reshape(len(self), 3)

    def get_velocities(self):
        return self._momenta / self._masses[:, None]

    def set_velocities(self, v):
        # one conversion and one product (same values as ASE's set_momenta(v * masses[:, None]); the product is a fresh array)
        if self._deferrable(v, len(self)):
            self.__dict__["_vel_src"] = v                # converted at the first read (see `_momenta`); masses as of that read
            return
        self._momenta = np.asarray(v, dtype=float).reshape(len(self), 3) * self._masses[:, None]

    def get_kinetic_energy(self):
        return 0.5 * float(np.sum(self._momenta ** 2 / self._masses[:, None]))

    def get_temperature(self):
        return self.get_kinetic_energy() / (1.5 * units.kB * max(len(self), 1))

    def copy(self):
        return Atoms(self)

    def wrap(self, **kw):
        self.positions = self.get_positions(wrap=True, **kw)


def wrap_positions(positions, cell, pbc=True, center=(0.5, 0.5, 0.5), eps=1e-7):
    """ASE 3.20 `ase.geometry.wrap_positions`: fractional = solve(cell^T, pos^T)^T - shift;
    periodic axes `%= 1`, `+= shift`, with shift = center - 0.5 - eps."""
    if not hasattr(center, "__len__"):
        center = (center,) * 3
    if isinstance(pbc, (bool, np.bool_, int)):
        pbc = (bool(pbc),) * 3
    pbc = np.asarray(pbc, dtype=bool)
    shift = np.asarray(center, dtype=float) - 0.5 - eps
    shift[~pbc] = 0.0
    cell = np.asarray(cell, dtype=float)
    if cell.shape == (3,):
        cell = np.diag(cell)
    pos = np.asarray(positions, dtype=float)
    off = cell - np.diag(np.diag(cell))
    ortho = pos.ndim == 2 and not off.any() and np.all(np.diag(cell) != 0)
    if ortho:
        # Orthorhombic cell (the only kind the pair path supports): ASE's `solve(cell^T, pos^T)^T` is an LU solve
        # whose back-substitution (OpenBLAS dtrsm) MULTIPLIES by the reciprocal of the diagonal - a plain division
        # differs in the last bit, the reciprocal product is bit-identical (tests/test_cabi_and_host.py) - and it
        # skips numpy's per-column gufunc marshalling, which costs ~60 ms for 256k atoms.
        fractional = pos * (1.0 / np.diag(cell))
        fractional -= shift
    else:
        fractional = np.ascontiguousarray(np.linalg.solve(cell.T, pos.T).T) - shift
    # `x % 1.0` == `x - floor(x)` bit for bit (fmod is exact, both round a - floor(a) once); floor vectorises
    if pbc.all():
        fractional -= np.floor(fractional)
        fractional += shift
    else:
        for i, periodic in enumerate(pbc):
            if periodic:
                fractional[:, i] -= np.floor(fractional[:, i])
                fractional[:, i] += shift[i]
    if ortho:       # dot with a diagonal matrix = f_i * c_ii + 0 + 0: the same bits without a threaded BLAS call
        fractional *= np.diag(cell)
        return fractional
    return np.dot(fractional, cell)


def MaxwellBoltzmannDistribution(atoms, temp=None, *, temperature_K=None, rng=None,
                                 communicator=None, force_temp=False):
    """ASE 3.20: momenta = N(0,1) * sqrt(m * temp), temp in energy units."""
    if temperature_K is not None:
        temp = temperature_K * units.kB
    if rng is None:
        rng = np.random
    masses = atoms.get_masses()
    xi = rng.standard_normal((len(masses), 3))
    atoms.set_momenta(xi * np.sqrt(masses * temp)[:, None])


def _cubic_lattice(basis, symbol, size, latticeconstant, pbc=True):
    """Cubic Bravais lattice with a basis; ASE ordering: unit cells in
    (i outer, j, k inner) order, basis atoms innermost."""
    if isinstance(size, int):
        size = (size,) * 3
    a = float(latticeconstant)
    basis = np.asarray(basis, dtype=float)
    pos = []
    for i in range(size[0]):
        for j in range(size[1]):
            for k in range(size[2]):
                for b in basis:
                    pos.append((np.array([i, j, k], dtype=float) + b) * a)
    pos = np.array(pos)
    z = _SYMBOLS[symbol] if isinstance(symbol, str) else int(symbol)
    return Atoms(numbers=[z] * len(pos), positions=pos,
                 cell=np.diag([size[0] * a, size[1] * a, size[2] * a]), pbc=pbc)


_FCC = [(0, 0, 0), (0.5, 0.5, 0), (0.5, 0, 0.5), (0, 0.5, 0.5)]
_DIA = _FCC + [(0.25, 0.25, 0.25), (0.75, 0.75, 0.25), (0.75, 0.25, 0.75), (0.25, 0.75, 0.75)]


def FaceCenteredCubic(symbol="H", size=(1, 1, 1), latticeconstant=1.0, pbc=True,
                      directions=None, **_):
    return _cubic_lattice(_FCC, symbol, size, latticeconstant, pbc)


def Diamond(symbol="Si", size=(1, 1, 1), latticeconstant=1.0, pbc=True,
            directions=None, **_):
    return _cubic_lattice(_DIA, symbol, size, latticeconstant, pbc)


def install_as_ase():
    """Register this module's objects under the `ase.*` names in `sys.modules`.

    TEST/ORACLE INFRASTRUCTURE ONLY: lets `oracle/ref_import.py` import the
    unmodified reference (which does `from ase import Atoms, units` at module
    top) in a container without ASE.  Never called by the product path.
    """
    import sys
    if "ase" in sys.modules and not getattr(sys.modules["ase"], "_mdgrad_standin", False):
        return sys.modules["ase"]
    ase = types.ModuleType("ase")
    ase._mdgrad_standin = True
    ase.Atoms = Atoms
    ase.units = units
    geometry = types.ModuleType("ase.geometry")
    geometry.wrap_positions = wrap_positions
    md = types.ModuleType("ase.md")
    veldist = types.ModuleType("ase.md.velocitydistribution")
    veldist.MaxwellBoltzmannDistribution = MaxwellBoltzmannDistribution
    md.velocitydistribution = veldist
    lattice = types.ModuleType("ase.lattice")
    cubic = types.ModuleType("ase.lattice.cubic")
    cubic.FaceCenteredCubic = FaceCenteredCubic
    cubic.Diamond = Diamond
    lattice.cubic = cubic
    ase.geometry, ase.md, ase.lattice = geometry, md, lattice
    for name, mod in [("ase", ase), ("ase.units", units), ("ase.geometry", geometry),
                      ("ase.md", md), ("ase.md.velocitydistribution", veldist),
                      ("ase.lattice", lattice), ("ase.lattice.cubic", cubic)]:
        sys.modules[name] = mod
    return ase
