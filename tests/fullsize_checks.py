"""Size-independent property checks of the hot path, written once and run (a) on the GPU at BASELINE configs[1]'s full size
(256 000 atoms, tests/test_gpu_zfullsize.py) and (b) under the CPU emulation at a reduced size (tests/test_emu_kernels.py),
which is what validates the checks themselves in the GPU-less container.

ctx: a Context (GPU or emulated); dev: the device the tensors go to."""
import numpy as np
import torch

from oracle import oracle_c as OC
from oracle import oracle_torch as O

RC = 2.5


def system(ncell, dev):
    pos, vel, L = O.lj_system(ncell, rho=0.845, jitter=0.05, seed=1)
    L32 = float(np.float32(L))
    pos32 = np.ascontiguousarray(pos, dtype=np.float32)
    return pos32, np.ascontiguousarray(vel, dtype=np.float32), L32, torch.from_numpy(pos32).to(dev), torch.from_numpy(
        np.ascontiguousarray(vel, dtype=np.float32)).to(dev)


def check_list_structure(ctx, dev, ncell):
    """reference layout invariants of the exported list: i < j, rows strictly increasing in (i, j) (torch.nonzero order,
    no duplicates), offsets in {-1,0,1}, count = half the directed entries, liquid-density coordination"""
    pos32, _, L32, xyz, _ = system(ncell, dev)
    n = xyz.shape[0]
    nbr, off = ctx.nbr_list(xyz, [L32] * 3, RC)
    P = nbr.shape[0]
    assert nbr.dtype == torch.int64 and off.dtype == torch.float32 and off.shape == (P, 3)
    assert bool((nbr[:, 0] < nbr[:, 1]).all()) and int(nbr.min()) >= 0 and int(nbr.max()) < n
    key = nbr[:, 0] * n + nbr[:, 1]
    assert bool((key[1:] > key[:-1]).all()), "list is not in the reference's row-major order"
    assert bool(((off == -1) | (off == 0) | (off == 1)).all())
    assert ctx.stats()["entries"] == 2 * P
    coord = 2.0 * P / n
    assert abs(coord - 4.18879 * RC ** 3 * 0.845) < 0.06 * coord, coord      # ~55 neighbors inside 2.5 sigma
    # distances of the listed pairs recomputed from the offsets are inside the cutoff
    d = (xyz[nbr[:, 0]] - xyz[nbr[:, 1]] - off * L32).pow(2).sum(1)
    assert float(d.max()) < RC * RC * (1 + 1e-6) and float(d.min()) > 0.0
    return P


def check_list_rows_bit_exact(ctx, dev, ncell, nrandom=6144, nboundary=6144):
    """neighbor INDICES, ORDER and image OFFSETS of sampled rows of the exported list, bit-exact against the C restatement of the
    reference arithmetic (topology.py:59-68) run over the whole box for those rows: random rows plus rows of atoms within
    one cutoff of a box face (the rows whose pairs carry image shifts), the first and the last atoms.  At 256 000 atoms the
    sample holds ~330 000 listed pairs; the reference itself cannot allocate this box (70 N^2 bytes)."""
    pos32, _, L32, xyz, _ = system(ncell, dev)
    n = xyz.shape[0]
    nbr, off = ctx.nbr_list(xyz, [L32] * 3, RC)
    nbr_h, off_h = nbr.cpu().numpy(), off.cpu().numpy()
    rng = np.random.default_rng(11)
    near_face = np.nonzero(((pos32 < RC) | (pos32 > L32 - RC)).any(1))[0]
    sel = np.unique(np.concatenate([rng.choice(n, size=min(nrandom, n), replace=False),
                                    rng.choice(near_face, size=min(nboundary, len(near_face)), replace=False),
                                    np.arange(min(64, n)), np.arange(max(0, n - 64), n)]))
    cnt, jj, oo = OC.nbr_rows_upper(pos32, [L32] * 3, RC, sel, cap=96)
    lo = np.searchsorted(nbr_h[:, 0], sel, side="left")
    hi = np.searchsorted(nbr_h[:, 0], sel, side="right")
    assert np.array_equal(hi - lo, cnt), "row lengths differ from the reference arithmetic at %d rows" % int((hi - lo != cnt).sum())
    rows = np.repeat(np.arange(len(sel)), cnt)
    cols = np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    flat = np.repeat(lo, cnt) + cols
    assert np.array_equal(nbr_h[flat, 1], jj[rows, cols]), "neighbor indices differ"
    assert np.array_equal(off_h[flat], oo[rows, cols]), "image offsets differ"
    shifted = int((oo[rows, cols] != 0).any(1).sum())
    assert shifted > 0 or ncell < 4
    return int(cnt.sum()), shifted


def check_forces_against_c_oracle(ctx, dev, ncell, nrows=2048):
    """forces / per-row energy of a block of rows vs the C restatement of the reference (all-pairs over the whole box,
    fp64 accumulation), 1e-5; Newton's third law over the whole box"""
    pos32, _, L32, xyz, _ = system(ncell, dev)
    n = xyz.shape[0]
    ctx.nbr_list(xyz, [L32] * 3, RC)
    e, f, _ = ctx.pair_force(0, [1.0, 1.0], xyz)
    i0 = n // 3
    nrows = min(nrows, n - i0)
    e_rows, f_rows = OC.lj_forces(pos32, [L32] * 3, RC, rows=(i0, i0 + nrows))
    fmax = float(np.abs(f_rows).max())
    assert np.abs(f[i0:i0 + nrows].cpu().numpy() - f_rows).max() <= 1e-5 * fmax
    fsum = f.double().sum(0).abs().max().item()
    assert fsum <= 1e-6 * f.double().abs().sum().item(), fsum
    return float(e)


def check_engine_invariants(ctx, dev, ncell, nsteps=24):
    """(a) skin list + exact re-test == rebuilding the exact list every step (same pair set at every evaluation),
    (b) NVE: total momentum (sum v: dv/dt = f, md.py:146) conserved"""
    from mdgrad_b200 import _lib
    pos32, vel32, L32, q0, v0 = system(ncell, dev)
    n = q0.shape[0]
    mass = torch.full((n,), 1.008, device=dev)

    def params(integrator, skin, K):
        p = _lib.MdParams()
        p.integrator, p.pot_kind = integrator, 0
        p.pot_params[0], p.pot_params[1] = 1.0, 1.0
        p.cutoff = RC
        for k in range(3):
            p.cell[k] = L32
        p.n_chains = 5
        Qb = O.nhc_bath_masses(50.0 * n / 256.0, n, 5)
        for k in range(5):
            p.Q[k] = float(Qb[k])
        p.T, p.ndof, p.skin, p.rebuild_every, p.traj_stride = 1.0, 3 * n, skin, K, nsteps
        return p

    t = O.time_grid(0.002, nsteps + 1).tolist()           # jittered lattice start: small step, no equilibration needed
    a = ctx.md_run(params(1, 0.0, 1), mass, v0, q0, [0.0] * 5, t)
    b = ctx.md_run(params(1, 0.45, 6), mass, v0, q0, [0.0] * 5, t)
    assert ctx.stats()["rebuilds"] < nsteps
    vs = a[0].abs().max().item()
    assert torch.isfinite(a[1]).all() and torch.isfinite(b[1]).all()
    assert (a[0][-1] - b[0][-1]).abs().max().item() <= 1e-4 * vs
    assert (a[1][-1] - b[1][-1]).abs().max().item() <= 1e-5 * L32
    c = ctx.md_run(params(0, 0.45, 6), mass, v0, q0, [], t)
    p0, p1 = c[0][0].double().sum(0), c[0][-1].double().sum(0)
    assert (p1 - p0).abs().max().item() <= 1e-6 * c[0][-1].double().abs().sum().item()


def check_rdf_two_paths(ctx, dev, ncell, nbins=100, r_range=(0.75, 3.3)):
    """the cell-traversal RDF kernel vs the same smeared histogram evaluated from the exported list's distances"""
    pos32, _, L32, xyz, _ = system(ncell, dev)
    start, end = r_range
    count = torch.zeros(nbins, device=dev)
    ctx.rdf_accumulate(xyz, [L32] * 3, start, end, nbins, None, count)
    nbr, off, dis = ctx.nbr_list(xyz, [L32] * 3, end + 0.5, get_dis=True)
    mu = torch.linspace(start, end, nbins, device=dev)
    w = float(mu[1] - mu[0])
    ref = torch.zeros(nbins, dtype=torch.float64, device=dev)
    for chunk in torch.split(dis, 2_000_000):
        ref += torch.exp(-0.5 / w ** 2 * (chunk[:, None].double() - mu.double()) ** 2).sum(0)
    assert (count.double() - ref).abs().max().item() <= 2e-5 * ref.max().item()
