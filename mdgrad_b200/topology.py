"""Neighbor-list / PBC geometry ops - mirror of reference torchmd/topology.py, executed by the
sm_100a kernels behind the C ABI (mdg_nbr_build / mdg_nbr_export / mdg_pair_dis_*).

Same call signatures and return layouts as the reference:
  generate_nbr_list(xyz, cutoff, cell, index_tuple=None, ex_pairs=None, get_dis=False)  topology.py:30-73
  compute_dis(xyz, nbr_list, offsets, cell)                                              topology.py:5-12
  generate_pair_index(N, index_tuple)                                                    topology.py:15-27
  get_offsets(vecs, cell, device)                                                        topology.py:75-80
"""
import numpy as np
import torch

from . import _lib

_CTX = {}


def context_for(device, key="default"):
    """One native context per (device, key)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("mdgrad_b200: the MD hot path runs on CUDA (sm_100a) only; got device %s" % device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    k = (idx, key)
    if k not in _CTX:
        _CTX[k] = _lib.Context(torch.device("cuda", idx))
    return _CTX[k]


def cell_lengths(cell):
    """(Lx, Ly, Lz) python floats (fp32 values) of an orthorhombic cell given as (3,) or (3,3)."""
    c = cell.detach().to("cpu", torch.float32) if isinstance(cell, torch.Tensor) else torch.tensor(np.asarray(cell), dtype=torch.float32)
    if c.dim() == 1:
        return [float(v) for v in c]
    if c.shape != (3, 3):
        raise ValueError("cell must have shape (3,) or (3,3)")
    if float((c - torch.diag(torch.diag(c))).abs().max()) != 0.0:
        raise NotImplementedError(
            "mdgrad_b200 supports orthorhombic (diagonal) cells only; a triclinic cell was given "
            "(no silent fallback)")
    return [float(v) for v in torch.diag(c)]


def _selection_flags(n, index_tuple, device):
    if index_tuple is None:
        return None, None
    a = torch.zeros(n, dtype=torch.uint8)
    b = torch.zeros(n, dtype=torch.uint8)
    a[torch.as_tensor(np.asarray(index_tuple[0]), dtype=torch.long)] = 1
    b[torch.as_tensor(np.asarray(index_tuple[1]), dtype=torch.long)] = 1
    return a.to(device), b.to(device)


def _exclusion_keys(n, ex_pairs, device):
    if ex_pairs is None:
        return None
    ex = torch.as_tensor(ex_pairs).to("cpu", torch.int64).reshape(-1, 2)
    ex = ex[ex[:, 0] != ex[:, 1]]          # a self pair has d2 == 0 anyway
    if ex.numel() == 0:
        return None
    lo = torch.minimum(ex[:, 0], ex[:, 1])
    hi = torch.maximum(ex[:, 0], ex[:, 1])
    return torch.unique(lo * n + hi).to(device)   # sorted unique


def generate_pair_index(N, index_tuple):
    """Dense (N,N) species-selection mask, kept for API compatibility (reference topology.py:15-27).
    The native list builder consumes per-atom flags instead of this O(N^2) mask."""
    m = torch.zeros(N, N)
    if index_tuple is not None:
        a = torch.as_tensor(np.asarray(index_tuple[0]), dtype=torch.long)
        b = torch.as_tensor(np.asarray(index_tuple[1]), dtype=torch.long)
        ia, ib = a.repeat_interleave(len(b)), b.repeat(len(a))
        m[ia, ib] = 1
        m[ib, ia] = 1
    return m


def generate_nbr_list(xyz, cutoff, cell, index_tuple=None, ex_pairs=None, get_dis=False, _ctx_key="default"):
    """Minimum-image neighbor list with the reference's exact membership, order and layout.

    xyz (N,3) or (F,N,3) fp32 CUDA tensor.  Returns (nbr_list, offsets) or, with get_dis,
    (nbr_list, pair_dis, offsets) - note the reference's return order (topology.py:70-73).
    For batched input nbr_list has a leading frame column (F index, i, j) like torch.nonzero on
    the (F,N,N) mask; offsets are per listed pair (the reference's own batched offsets indexing,
    topology.py:71,73, is only meaningful un-batched - SURVEY 8a a1)."""
    _lib.require_cuda(xyz, "xyz")
    L = cell_lengths(cell)
    ctx = context_for(xyz.device, _ctx_key)
    if xyz.dim() == 2:
        n = xyz.shape[0]
        sa, sb = _selection_flags(n, index_tuple, xyz.device)
        keys = _exclusion_keys(n, ex_pairs, xyz.device)
        out = ctx.nbr_list(xyz, L, cutoff, sa, sb, keys, get_dis=get_dis)
        return (out[0], out[2], out[1]) if get_dis else out
    if xyz.dim() != 3:
        raise ValueError("xyz must be (N,3) or (F,N,3)")
    n = xyz.shape[1]
    sa, sb = _selection_flags(n, index_tuple, xyz.device)
    keys = _exclusion_keys(n, ex_pairs, xyz.device)
    nbrs, offs, diss = [], [], []
    for f in range(xyz.shape[0]):
        out = ctx.nbr_list(xyz[f], L, cutoff, sa, sb, keys, get_dis=get_dis)
        fcol = torch.full((out[0].shape[0], 1), f, dtype=torch.int64, device=xyz.device)
        nbrs.append(torch.cat([fcol, out[0]], 1))
        offs.append(out[1])
        if get_dis:
            diss.append(out[2])
    nbr, off = torch.cat(nbrs), torch.cat(offs)
    return (nbr, torch.cat(diss), off) if get_dis else (nbr, off)


class _PairDis(torch.autograd.Function):
    """compute_dis with a hand-written backward (scatter of dE/dr * unit vector)."""

    @staticmethod
    def forward(ctx, xyz, nbr, offsets, L):
        xyz_c = xyz.detach().to(torch.float32).contiguous()
        nbr_c = nbr.to(xyz.device, torch.int64).contiguous()
        off_c = offsets.detach().to(xyz.device, torch.float32).contiguous()
        dis = _lib.pair_dis_fwd(xyz_c, nbr_c, off_c, L)
        ctx.save_for_backward(xyz_c, nbr_c, off_c, dis)
        ctx.L = L
        return dis[:, None]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        xyz, nbr, off, dis = ctx.saved_tensors
        gx = _lib.pair_dis_bwd(xyz, nbr, off, ctx.L, dis, g.reshape(-1).contiguous())
        return gx, None, None, None


def compute_dis(xyz, nbr_list, offsets, cell):
    """|x_i - x_j - offsets @ cell| for the listed pairs, (P,1)  (reference topology.py:5-12).
    First-order differentiable w.r.t. xyz through the native backward kernel; when a second-order
    graph is needed (adjoint / create_graph) use `compute_dis_torch`."""
    return _PairDis.apply(xyz, nbr_list, offsets, cell_lengths(cell))


def compute_dis_torch(xyz, nbr_list, offsets, cell):
    """Same quantity expressed in differentiable torch ops (for double backward)."""
    if cell.dim() == 1:
        cell = torch.diag(cell)
    nbr_list = nbr_list.to(xyz.device)
    return (xyz[nbr_list[:, 0]] - xyz[nbr_list[:, 1]] - offsets.matmul(cell)).pow(2).sum(1).sqrt()[:, None]


def get_offsets(vecs, cell, device):
    """reference topology.py:75-80"""
    return -vecs.ge(0.5 * cell).to(torch.float).to(device) + vecs.lt(-0.5 * cell).to(torch.float).to(device)


def make_directed(nbr_list):
    """both directions of every listed pair: rows (f, a, b) followed by rows (f, b, a)  (reference topology.py:108-122)"""
    rev = torch.stack([nbr_list[:, 0], nbr_list[:, 2], nbr_list[:, 1]], dim=-1)
    return torch.cat([nbr_list, rev], dim=0)


def generate_angle_list(nbr_list):
    """All angle triples (frame, a, b, c): directed bond a->b followed by a directed bond b->c with c != a, in the
    reference's row order (reference topology.py:83-106).  The reference materialises a (2P x 2P) boolean mask on the
    CPU (numpy repeat); here the same list is enumerated on the tensor's own device from a per-(frame, atom) CSR of
    the directed list: O(#angles) memory, no host round trip."""
    assert nbr_list.shape[1] == 3
    d = make_directed(nbr_list)
    R = d.shape[0]
    dev = d.device
    if R == 0:
        return torch.zeros((0, 4), dtype=d.dtype, device=dev)
    n = int(d[:, 1:].max()) + 1
    key_first = d[:, 0] * n + d[:, 1]                       # (frame, first atom) of every directed row
    order = torch.argsort(key_first, stable=True)           # CSR order = ascending row index inside a segment
    nseg = int(key_first.max()) + 1
    counts = torch.bincount(key_first, minlength=nseg)
    start = torch.cumsum(counts, 0) - counts
    key_second = d[:, 0] * n + d[:, 2]                      # rows q with (frame, first atom) == (frame_p, b_p)
    cnt_p = counts[key_second]
    rep = torch.repeat_interleave(torch.arange(R, device=dev), cnt_p)
    within = torch.arange(rep.shape[0], device=dev) - torch.repeat_interleave(torch.cumsum(cnt_p, 0) - cnt_p, cnt_p)
    q = order[start[key_second][rep] + within]
    third = d[q, 2]
    keep = third != d[rep, 1]                               # c != a
    return torch.cat([d[rep[keep]], third[keep].reshape(-1, 1)], dim=1)
