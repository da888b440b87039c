"""Pair energy functions u(r) - mirror of reference torchmd/potentials.py (analytic family
:61-93, :317-365; learned pairMLP/TpairMLP/MLP :163-217, :368-391).

Every class keeps the reference constructor signature, parameter names (`sigma`, `epsilon`, `A`,
`B`, `C` as `nn.Parameter` of shape (1,)) and a differentiable torch `forward(r)`.  The analytic
classes additionally expose `native_spec()` -> (kind, values, parameter tensors) which lets
`PairPotentials` and the fused MD engine evaluate energy + force + dE/dparam in ONE sm_100a
kernel (mdg_pair_force / mdg_md_run) instead of u(r) + autograd.
"""
import math

import numpy as np
import torch
from torch import nn

from . import _lib

nlr_dict = {
    "ReLU": nn.ReLU(), "ELU": nn.ELU(), "Tanh": nn.Tanh(), "LeakyReLU": nn.LeakyReLU(),
    "ReLU6": nn.ReLU6(), "SELU": nn.SELU(), "CELU": nn.CELU(), "Tanhshrink": nn.Tanhshrink(),
}

LJPARAMS = {"epsilon": 1.0, "sigma": 1.0}
MLPPARAMS = {"D_in": 1, "H": 128, "num_layers": 3, "act": "relu", "D_out": 1}


def _p(v):
    return nn.Parameter(torch.Tensor([v]))


class LJFamily(nn.Module):
    """u = 4 eps ((sigma/r)^rep - (sigma/r)^attr)   (reference potentials.py:61-73)"""

    def __init__(self, sigma=1.0, epsilon=1.0, attr_pow=6, rep_pow=12):
        super().__init__()
        self.sigma, self.epsilon = _p(sigma), _p(epsilon)
        self.attr_pow, self.rep_pow = attr_pow, rep_pow

    def LJ(self, r, sigma, epsilon):
        return 4 * epsilon * ((sigma / r) ** self.rep_pow - (sigma / r) ** self.attr_pow)

    def forward(self, x):
        return self.LJ(x, self.sigma, self.epsilon)

    def native_spec(self):
        return _lib.POT_LJFAM, [self.sigma.item(), self.epsilon.item(), float(self.rep_pow), float(self.attr_pow)], \
            [self.sigma, self.epsilon]


class LennardJones(nn.Module):
    """u = 4 eps ((sigma/r)^12 - (sigma/r)^6)   (reference potentials.py:317-327)"""

    def __init__(self, sigma=1.0, epsilon=1.0):
        super().__init__()
        self.sigma, self.epsilon = _p(sigma), _p(epsilon)

    def LJ(self, r, sigma, epsilon):
        return 4 * epsilon * ((sigma / r) ** 12 - (sigma / r) ** 6)

    def forward(self, x):
        return self.LJ(x, self.sigma, self.epsilon)

    def native_spec(self):
        return _lib.POT_LJ, [self.sigma.item(), self.epsilon.item()], [self.sigma, self.epsilon]


class LennardJones69(nn.Module):
    """u = 4 eps ((sigma/r)^9 - (sigma/r)^6)   (reference potentials.py:329-339)"""

    def __init__(self, sigma=1.0, epsilon=1.0):
        super().__init__()
        self.sigma, self.epsilon = _p(sigma), _p(epsilon)

    def LJ(self, r, sigma, epsilon):
        return 4 * epsilon * ((sigma / r) ** 9 - (sigma / r) ** 6)

    def forward(self, x):
        return self.LJ(x, self.sigma, self.epsilon)

    def native_spec(self):
        return _lib.POT_LJ69, [self.sigma.item(), self.epsilon.item()], [self.sigma, self.epsilon]


class ExcludedVolume(nn.Module):
    """u = 4 eps (sigma/r)^power   (reference potentials.py:341-352)"""

    def __init__(self, sigma=1.0, epsilon=1.0, power=12):
        super().__init__()
        self.sigma, self.epsilon = _p(sigma), _p(epsilon)
        self.power = power

    def LJ(self, r, sigma, epsilon):
        return 4 * epsilon * ((sigma / r) ** self.power)

    def forward(self, x):
        return self.LJ(x, self.sigma, self.epsilon)

    def native_spec(self):
        return _lib.POT_EXV, [self.sigma.item(), self.epsilon.item(), float(self.power)], [self.sigma, self.epsilon]


class Buck(nn.Module):
    """u = A exp(-B r) - C / r^6   (reference potentials.py:354-365)"""

    def __init__(self, A=1.0, B=1.0, C=1.0):
        super().__init__()
        self.A, self.B, self.C = _p(A), _p(B), _p(C)

    def Buckingham(self, r, A, B, C):
        return A * torch.exp(-B * r) - C / r ** 6

    def forward(self, x):
        return self.Buckingham(x, self.A, self.B, self.C)

    def native_spec(self):
        return _lib.POT_BUCK, [self.A.item(), self.B.item(), self.C.item()], [self.A, self.B, self.C]


class ModifiedMorse(nn.Module):
    """u = (exp(2x) - 2 exp(x) - A) / (1 + A), x = a (1 - r^phi) / phi   (reference potentials.py:75-93)"""

    def __init__(self, a, phi):
        super().__init__()
        self.a, self.phi = a, phi
        self.A = 0 if phi >= 0 else np.exp(2 * a / phi) - 2 * np.exp(a / phi)

    def forward(self, r):
        exponent = self.a * (1 - r ** self.phi) / self.phi
        return (torch.exp(2 * exponent) - 2 * torch.exp(exponent) - self.A) / (1 + self.A)

    def native_spec(self):
        return _lib.POT_MORSE, [float(self.a), float(self.phi)], []


# ---------------------------------------------------------------------------------------------
# learned u(r): plain torch modules; PairPotentials feeds them with the native compute_dis op
# ---------------------------------------------------------------------------------------------
class GaussianSmearing(nn.Module):
    """exp(-0.5/w^2 (d - mu_k)^2), mu = linspace(start, stop, n)  (reference nff/nn/layers.py:34-83)"""

    def __init__(self, start, stop, n_gaussians, width=None, centered=False, trainable=False):
        super().__init__()
        offset = torch.linspace(start, stop, n_gaussians)
        widths = torch.FloatTensor(((offset[1] - offset[0]) if width is None else width) * torch.ones_like(offset))
        if trainable:
            self.width = nn.Parameter(widths)
            self.offsets = nn.Parameter(offset)
        else:
            self.register_buffer("width", widths)
            self.register_buffer("offsets", offset)
        self.centered = centered

    def forward(self, distances):
        if not self.centered:
            coeff = -0.5 / torch.pow(self.width, 2)
            diff = distances - self.offsets
        else:
            coeff = -0.5 / torch.pow(self.offsets, 2)
            diff = distances
        return torch.exp(coeff * torch.pow(diff, 2))


class pairMLP(nn.Module):
    """Gaussian-expanded distance -> MLP -> u(r)   (reference potentials.py:163-206)"""

    def __init__(self, n_gauss, r_start, r_end, n_layers, n_width, nonlinear, res=False):
        super().__init__()
        nlr = nlr_dict[nonlinear]
        self.smear = GaussianSmearing(start=r_start, stop=r_end, n_gaussians=n_gauss, trainable=True)
        self.layers = nn.ModuleList([nn.Linear(n_gauss, n_gauss), nlr, nn.Linear(n_gauss, n_width), nlr])
        for _ in range(n_layers):
            self.layers.append(nn.Linear(n_width, n_width))
            self.layers.append(nlr)
        self.layers.append(nn.Linear(n_width, n_gauss))
        self.layers.append(nlr)
        self.layers.append(nn.Linear(n_gauss, 1))
        self.res = res

    def forward(self, r):
        r = self.smear(r)
        for layer in self.layers:
            if self.res is False:
                r = layer(r)
            else:
                dr = layer(r)
                r = r + dr if dr.shape[-1] == r.shape[-1] else dr
        return r


class TpairMLP(nn.Module):
    """u(r, T) = energy(r) - T entropy(r)   (reference potentials.py:208-217)"""

    def __init__(self, n_gauss, r_start, r_end, n_layers, n_width, nonlinear, res=False):
        super().__init__()
        self.energy = pairMLP(n_gauss, r_start, r_end, n_layers, n_width, nonlinear, res=res)
        self.entropy = pairMLP(n_gauss, r_start, r_end, n_layers, n_width, nonlinear, res=res)

    def forward(self, r, T):
        return self.energy(r) - T * self.entropy(r)


class MLP(nn.Module):
    """ReLU MLP on r with optional (0.6/r)^12 core   (reference potentials.py:368-391)"""

    def __init__(self, D_in=1, H=128, D_out=1, num_layers=3, act="relu", excluded_vol=True):
        super().__init__()
        self.NN = nn.ModuleList([nn.Linear(D_in, H), nn.ReLU()])
        for _ in range(num_layers):
            self.NN.append(nn.Linear(H, H))
            self.NN.append(nn.ReLU())
        self.NN.append(nn.Linear(H, 1))
        self.excluded_vol = excluded_vol

    def forward(self, x):
        u_ex = (0.6 / x) ** 12 if self.excluded_vol else 0.0
        for layer in self.NN:
            x = layer(x)
        return u_ex + x


class pairTab(nn.Module):
    """pairTab(nbins=1000, rc=2.5, device='cpu'): tabulated u(r), `tab` (nbins,) a Parameter on the knots
    x = linspace(0, rc, nbins), evaluated by cubic-spline interpolation (reference potentials.py:152-160, which calls
    `xitorch.interpolate.Interp1D(x, tab)(r)`).

    xitorch is an un-vendored dependency that is not installed here: PARITY UNPINNED.  This class restates
    Interp1D's documented default - a C2 cubic spline with not-a-knot end conditions (the same algorithm as
    scipy.interpolate.CubicSpline, against which tests/test_cabi_and_host.py checks it) - as torch ops: the spline's
    second derivatives are LINEAR in the table, so one (nbins x nbins) solve at construction gives a fixed matrix and
    every forward is a matmul + a gather, differentiable in both `tab` and r (to any order - the adjoint route works).
    r outside [0, rc] is extrapolated with the end polynomials (the list cutoff keeps r <= rc)."""

    def __init__(self, nbins=1000, rc=2.5, device="cpu"):
        super().__init__()
        self.tab = nn.Parameter(torch.zeros(nbins).to(device))
        self.x = torch.linspace(0.0, rc, nbins).to(device)
        n = int(nbins)
        if n < 4:
            raise ValueError("pairTab needs at least 4 knots")
        x = torch.linspace(0.0, rc, nbins, dtype=torch.float64)
        h = x[1:] - x[:-1]
        A = torch.zeros(n, n, dtype=torch.float64)       # A m = B y,  m = second derivatives at the knots
        B = torch.zeros(n, n, dtype=torch.float64)
        for i in range(1, n - 1):
            A[i, i - 1], A[i, i], A[i, i + 1] = h[i - 1], 2 * (h[i - 1] + h[i]), h[i]
            B[i, i - 1], B[i, i], B[i, i + 1] = 6 / h[i - 1], -6 / h[i - 1] - 6 / h[i], 6 / h[i]
        # not-a-knot: the third derivative is continuous across the second and the second-to-last knot
        A[0, 0], A[0, 1], A[0, 2] = h[1], -(h[0] + h[1]), h[0]
        A[n - 1, n - 3], A[n - 1, n - 2], A[n - 1, n - 1] = h[n - 2], -(h[n - 3] + h[n - 2]), h[n - 3]
        self.register_buffer("_m_of_y", torch.linalg.solve(A, B).to(torch.float32).to(device), persistent=False)

    def forward(self, r):
        shape = r.shape
        rq = r.reshape(-1)
        x = self.x.to(rq.device)
        y = self.tab
        m = self._m_of_y.to(rq.device) @ y
        n = x.shape[0]
        i = torch.clamp(torch.searchsorted(x, rq.detach(), right=True) - 1, 0, n - 2)
        h = x[i + 1] - x[i]
        a = (x[i + 1] - rq) / h
        b = (rq - x[i]) / h
        u = a * y[i] + b * y[i + 1] + ((a ** 3 - a) * m[i] + (b ** 3 - b) * m[i + 1]) * (h * h) / 6.0
        return u.reshape(shape) if len(shape) and shape[-1] == 1 else u.unsqueeze(-1)


def _out_of_scope(name, why):
    """Names of the reference's potentials.py that are outside the MD hot path (SURVEY 8 'out of scope'): importable, so that
    the reference's scripts (`from torchmd.potentials import ..., SplineOverlap, ...` at module top of scripts/data.py) load
    against this package, and loud when constructed."""
    def __init__(self, *a, **k):
        raise NotImplementedError("mdgrad_b200: %s is outside the MD hot path this package implements (%s)" % (name, why))
    return type(name, (nn.Module,), {"__init__": __init__, "__doc__": "out-of-scope placeholder, see potentials._out_of_scope"})


SplineOverlap = _out_of_scope("SplineOverlap", "needs torchcubicspline; reference potentials.py")
BoltzmannInversionSpline = _out_of_scope("BoltzmannInversionSpline", "needs torchcubicspline; reference potentials.py")
Harmonic1D = _out_of_scope("Harmonic1D", "toy potential of the isomerisation demos")
toy2d = _out_of_scope("toy2d", "toy potential of the isomerisation demos")
leps = _out_of_scope("leps", "toy potential of the isomerisation demos")
MLP2d = _out_of_scope("MLP2d", "toy potential of the isomerisation demos")


def __getattr__(name):
    # `PairPotentials` lives in interface.py in the reference, but BASELINE.json / README name it
    # `torchmd.potentials.PairPotentials` (SURVEY naming trap): export it from both.
    if name == "PairPotentials":
        from .interface import PairPotentials
        return PairPotentials
    raise AttributeError(name)
