"""ctypes binding of libmdgrad_b200.so (the C ABI declared in include/mdgrad_b200.h).

This is the stub a reference maintainer would add: raw `tensor.data_ptr()`
device pointers and PyTorch's current CUDA stream go straight into the `extern "C"` entry
points.  There is NO fallback: if the shared library is missing or a tensor is not on a CUDA
device the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# MDG_LIB_VARIANT=<name> loads an experimental build (mdgrad_b200/build.py VARIANTS) for A/B measurements
_VARIANT = os.environ.get("MDG_LIB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, "libmdgrad_b200%s.so" % (("_" + _VARIANT) if _VARIANT else ""))

MDG_OK, MDG_E_BADARG, MDG_E_CUDA, MDG_E_CAPACITY, MDG_E_STATE, MDG_E_SKIN, MDG_E_NCCL = 0, -1, -2, -3, -4, -5, -6
MDG_E_NUMERIC = -7
POT_LJ, POT_LJFAM, POT_LJ69, POT_EXV, POT_BUCK, POT_MORSE = range(6)
INT_NVE, INT_NHC = 0, 1
MAX_POT_PARAMS = 4
MAX_CHAINS = 16

# every symbol include/mdgrad_b200.h declares (checked by tests/test_cabi_symbols.py)
SYMBOLS = [
    "mdg_version", "mdg_last_error", "mdg_create", "mdg_destroy", "mdg_nbr_build", "mdg_nbr_export",
    "mdg_pair_force", "mdg_pair_dis_fwd", "mdg_pair_dis_bwd", "mdg_rdf_accumulate", "mdg_md_run",
    "mdg_get_stats", "mdg_set_pair_filter", "mdg_set_profile", "mdg_get_profile",
    "mdg_slab_plan", "mdg_dist_unique_id", "mdg_dist_init", "mdg_dist_finalize",
    "mdg_graph_build", "mdg_cfconv_agg", "mdg_cfconv_edge_grad", "mdg_schnet_energy_force", "mdg_pair_hvp", "mdg_md_run_gnn",
    "mdg_bonded_force", "mdg_vacf",
]


class MdParams(ctypes.Structure):
    """mirror of struct mdg_md_params"""
    _fields_ = [
        ("integrator", ctypes.c_int),
        ("pot_kind", ctypes.c_int),
        ("pot_params", ctypes.c_float * MAX_POT_PARAMS),
        ("cutoff", ctypes.c_double),
        ("cell", ctypes.c_float * 3),
        ("n_chains", ctypes.c_int),
        ("Q", ctypes.c_float * MAX_CHAINS),
        ("T", ctypes.c_double),
        ("ndof", ctypes.c_int),
        ("skin", ctypes.c_float),
        ("rebuild_every", ctypes.c_int),
        ("traj_stride", ctypes.c_int),
    ]


SCHNET_MAX_LAYERS = 8


class SchnetLayer(ctypes.Structure):
    """mirror of struct mdg_schnet_layer (device pointers)"""
    _fields_ = [(k, ctypes.c_void_p) for k in
                ("mu", "width", "We1", "be1", "We2", "be2", "Wn", "bn", "Wu1", "bu1", "Wu2", "bu2")]


class SchnetModel(ctypes.Structure):
    """mirror of struct mdg_schnet_model"""
    _fields_ = [("n_atom_basis", ctypes.c_int), ("n_filters", ctypes.c_int), ("n_gaussians", ctypes.c_int),
                ("n_convolutions", ctypes.c_int), ("n_readout", ctypes.c_int),
                ("embed", ctypes.c_void_p), ("layers", SchnetLayer * SCHNET_MAX_LAYERS),
                ("Wr1", ctypes.c_void_p), ("br1", ctypes.c_void_p), ("Wr2", ctypes.c_void_p), ("br2", ctypes.c_void_p),
                ("weights_tag", ctypes.c_uint64)]


MAX_PRIORS = 4


class PriorSpec(ctypes.Structure):
    """mirror of struct mdg_prior_spec"""
    _fields_ = [("ctx", ctypes.c_void_p), ("kind", ctypes.c_int), ("params", ctypes.c_float * MAX_POT_PARAMS),
                ("n_params", ctypes.c_int), ("cutoff", ctypes.c_double), ("d_sel_a", ctypes.c_void_p),
                ("d_sel_b", ctypes.c_void_p), ("d_ex_keys", ctypes.c_void_p), ("n_ex", ctypes.c_int)]


class BondedTerms(ctypes.Structure):
    """mirror of struct mdg_bonded_terms"""
    _fields_ = [("d_bond_top", ctypes.c_void_p), ("n_bonds", ctypes.c_int), ("k_bond", ctypes.c_float), ("r0", ctypes.c_float),
                ("d_angle_top", ctypes.c_void_p), ("n_angles", ctypes.c_int), ("k_angle", ctypes.c_float),
                ("theta0", ctypes.c_float), ("d_ref_start", ctypes.c_void_p), ("d_refs", ctypes.c_void_p)]


class GnnMdParams(ctypes.Structure):
    """mirror of struct mdg_gnn_md_params"""
    _fields_ = [("integrator", ctypes.c_int), ("n_chains", ctypes.c_int), ("Q", ctypes.c_float * MAX_CHAINS),
                ("T", ctypes.c_double), ("ndof", ctypes.c_int), ("cell", ctypes.c_float * 3), ("cutoff", ctypes.c_double),
                ("off_scale", ctypes.c_float * 3), ("d_ex_keys", ctypes.c_void_p), ("n_ex", ctypes.c_int),
                ("n_priors", ctypes.c_int), ("priors", PriorSpec * MAX_PRIORS), ("traj_stride", ctypes.c_int),
                ("bonded", BondedTerms)]


def schnet_model_struct(sd, device):
    """mdg_schnet_model from a reference-layout SchNet `state_dict` (nff/nn/models/schnet.py); returns the struct and
    the list of fp32 contiguous tensors it points into (keep them alive while the struct is used)."""
    keep = []

    def dp(key):
        t = sd[key].detach().to(device, torch.float32).contiguous()
        keep.append(t)
        return t.data_ptr()

    L = len({k.split(".")[1] for k in sd if k.startswith("convolutions.")})
    if L > SCHNET_MAX_LAYERS:
        raise ValueError("SchNet with %d convolutions: the native path supports up to %d" % (L, SCHNET_MAX_LAYERS))
    m = SchnetModel()
    m.n_atom_basis = sd["atom_embed.weight"].shape[1]
    m.n_convolutions = L
    m.embed = dp("atom_embed.weight")
    for l in range(L):
        pre = "convolutions.%d.moduledict." % l
        y = m.layers[l]
        y.mu, y.width = dp(pre + "message_edge_filter.0.offsets"), dp(pre + "message_edge_filter.0.width")
        y.We1, y.be1 = dp(pre + "message_edge_filter.1.weight"), dp(pre + "message_edge_filter.1.bias")
        y.We2, y.be2 = dp(pre + "message_edge_filter.3.weight"), dp(pre + "message_edge_filter.3.bias")
        y.Wn, y.bn = dp(pre + "message_node_filter.weight"), dp(pre + "message_node_filter.bias")
        y.Wu1, y.bu1 = dp(pre + "update_function.0.weight"), dp(pre + "update_function.0.bias")
        y.Wu2, y.bu2 = dp(pre + "update_function.2.weight"), dp(pre + "update_function.2.bias")
    m.n_gaussians = sd["convolutions.0.moduledict.message_edge_filter.1.weight"].shape[1]
    m.n_filters = sd["convolutions.0.moduledict.message_edge_filter.3.weight"].shape[0]
    ro = "atomwisereadout.readout.energy."
    m.n_readout = sd[ro + "linear0.weight"].shape[0]
    m.Wr1, m.br1 = dp(ro + "linear0.weight"), dp(ro + "linear0.bias")
    m.Wr2, m.br2 = dp(ro + "linear2.weight"), dp(ro + "linear2.bias")
    # value tag: changes whenever a parameter tensor is replaced or modified in place (torch bumps `_version`)
    m.weights_tag = (hash(tuple((sd[k].data_ptr(), sd[k]._version) for k in sd)) & 0x7FFFFFFFFFFFFFFF) | 1
    return m, keep


class MdgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmdgrad_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """dlopen the in-tree library; raises (never falls back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libmdgrad_b200.so not found at %s - build it with `python -m mdgrad_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback for the MD hot path." % LIB_PATH)
    _lib = bind(ctypes.CDLL(LIB_PATH))
    return _lib


def bind(lib):
    """Declares the argument / result types of every entry point of include/mdgrad_b200.h on a loaded library."""
    vp, ip, dbl, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int64
    fp = ctypes.POINTER(ctypes.c_float)
    lib.mdg_version.restype = ip
    lib.mdg_last_error.restype = ctypes.c_char_p
    lib.mdg_create.argtypes = [ip, ctypes.POINTER(vp)]
    lib.mdg_destroy.argtypes = [vp]
    lib.mdg_nbr_build.argtypes = [vp, vp, ip, fp, dbl, vp, vp, vp, ip, vp, ctypes.POINTER(i64)]
    lib.mdg_nbr_export.argtypes = [vp, vp, vp, vp, vp]
    lib.mdg_pair_force.argtypes = [vp, ip, fp, ip, vp, ip, vp, vp, vp, vp]
    lib.mdg_pair_dis_fwd.argtypes = [vp, ip, vp, vp, i64, fp, vp, vp]
    lib.mdg_pair_dis_bwd.argtypes = [vp, ip, vp, vp, i64, fp, vp, vp, vp, vp]
    lib.mdg_rdf_accumulate.argtypes = [vp, vp, ip, fp, dbl, dbl, ip, dbl, vp, vp, vp, vp]
    lib.mdg_vacf.argtypes = [vp, vp, ip, ip, ip, ip, vp, vp]
    lib.mdg_md_run.argtypes = [vp, ctypes.POINTER(MdParams), ip, vp, vp, vp, fp, fp, ip, vp, vp, fp, fp, vp]
    lib.mdg_get_stats.argtypes = [vp, ctypes.POINTER(i64)]
    lib.mdg_set_pair_filter.argtypes = [vp, vp, vp, vp, ip]
    lib.mdg_set_profile.argtypes = [vp, ip]
    lib.mdg_slab_plan.argtypes = [ip, ip, ip, ctypes.POINTER(ip)]
    lib.mdg_dist_unique_id.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    lib.mdg_dist_init.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, ip, ip]
    lib.mdg_dist_finalize.argtypes = [vp]
    lib.mdg_graph_build.argtypes = [vp, vp, i64, ip, vp]
    lib.mdg_cfconv_agg.argtypes = [vp, vp, vp, ip, ip, vp, vp]
    lib.mdg_cfconv_edge_grad.argtypes = [vp, vp, vp, ip, ip, vp, vp]
    lib.mdg_get_profile.argtypes = [vp, ctypes.POINTER(dbl)]
    lib.mdg_md_run_gnn.argtypes = [vp, ctypes.POINTER(GnnMdParams), ctypes.POINTER(SchnetModel), vp, ip, vp, vp, vp, fp, fp, ip,
                                   vp, vp, fp, fp, vp]
    lib.mdg_pair_hvp.argtypes = [vp, ip, fp, ip, vp, ip, vp, vp, vp, vp]
    lib.mdg_bonded_force.argtypes = [vp, ctypes.POINTER(BondedTerms), vp, ip, fp, vp, vp, vp, vp]
    lib.mdg_schnet_energy_force.argtypes = [vp, ctypes.POINTER(SchnetModel), vp, vp, ip, vp, vp, i64, fp, vp, vp, vp]
    for name in SYMBOLS:
        if name not in ("mdg_last_error",):
            getattr(lib, name).restype = ip
    lib.mdg_last_error.restype = ctypes.c_char_p
    return lib


def check(status):
    if status != MDG_OK:
        raise MdgError(status, load().mdg_last_error().decode("utf-8", "replace"))


def _farr(vals, n=None):
    vals = [float(v) for v in vals]
    n = len(vals) if n is None else n
    return (ctypes.c_float * n)(*(vals + [0.0] * (n - len(vals))))


def require_cuda(t, name="tensor"):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(
            "mdgrad_b200: %s must live on a CUDA device (got %s); the MD hot path is sm_100a CUDA "
            "only and has no CPU fallback" % (name, getattr(t, "device", type(t))))


def on_device(t):
    """True for tensors the native path can take (CUDA).  One function so that the CPU emulation harness of the test
    suite (tests/cuemu) can run the unmodified Python layer on host tensors."""
    return bool(t.is_cuda)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _guard(device):
    return torch.cuda.device(device)


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


class Context:
    """One mdg_ctx per (device, use).  Not thread-safe."""

    # hooks (the CPU emulation harness under tests/cuemu subclasses these; the product path is CUDA only)
    def _api(self):
        return load()

    def _check(self, status):
        check(status)

    def _require(self, t, name="tensor"):
        require_cuda(t, name)

    def _stream(self, device):
        return _stream(device)

    def _guard(self, device):
        return _guard(device)

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("mdgrad_b200.Context needs a CUDA device, got %s" % device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self._h = ctypes.c_void_p()
        self._check(self._api().mdg_create(idx, ctypes.byref(self._h)))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self._api().mdg_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    # -- K1 ---------------------------------------------------------------------------------
    def nbr_list(self, xyz, cell3, cutoff, sel_a=None, sel_b=None, ex_keys=None, get_dis=False):
        """generate_nbr_list for one frame: returns (nbr int64 (P,2), offsets fp32 (P,3)[, dis (P,)])."""
        self._require(xyz, "xyz")
        xyz = xyz.detach().to(torch.float32).contiguous()
        n = xyz.shape[0]
        dev = xyz.device
        npairs = ctypes.c_int64(0)
        with self._guard(dev):
            self._check(self._api().mdg_nbr_build(self._h, _ptr(xyz), n, _farr(cell3, 3), float(cutoff), _ptr(sel_a), _ptr(sel_b),
                                       _ptr(ex_keys), 0 if ex_keys is None else int(ex_keys.numel()), self._stream(dev),
                                       ctypes.byref(npairs)))
            P = npairs.value
            nbr = torch.empty((P, 2), dtype=torch.int64, device=dev)
            off = torch.empty((P, 3), dtype=torch.float32, device=dev)
            dis = torch.empty((P,), dtype=torch.float32, device=dev) if get_dis else None
            self._check(self._api().mdg_nbr_export(self._h, _ptr(nbr), _ptr(off), _ptr(dis), self._stream(dev)))
        self._keepalive = (xyz, sel_a, sel_b, ex_keys)
        return (nbr, off, dis) if get_dis else (nbr, off)

    # -- K2+K3 ------------------------------------------------------------------------------
    def pair_force(self, kind, params, xyz, want_force=True, want_dparams=False):
        """E (0-d), F (N,3) or None, dE/dparams (4,) or None over the list of the last nbr_list()."""
        self._require(xyz, "xyz")
        xyz = xyz.detach().to(torch.float32).contiguous()
        n = xyz.shape[0]
        dev = xyz.device
        e = torch.empty((), dtype=torch.float32, device=dev)
        f = torch.empty((n, 3), dtype=torch.float32, device=dev) if want_force else None
        dp = torch.empty((MAX_POT_PARAMS,), dtype=torch.float32, device=dev) if want_dparams else None
        with self._guard(dev):
            self._check(self._api().mdg_pair_force(self._h, int(kind), _farr(params, MAX_POT_PARAMS), len(params), _ptr(xyz), n,
                                        _ptr(e), _ptr(f), _ptr(dp), self._stream(dev)))
        return e, f, dp

    def pair_hvp(self, kind, params, xyz, avec, want_dtheta=True):
        """((dF/dxyz)^T a (N,3), (dF/dparams)^T a (4,) or None) over the list of the last nbr_list() - analytic
        second-order products for the adjoint solver (power-law kinds; raises MdgError(BADARG) otherwise)."""
        self._require(xyz, "xyz")
        xyz = xyz.detach().to(torch.float32).contiguous()
        avec = avec.detach().to(xyz.device, torch.float32).contiguous()
        n = xyz.shape[0]
        hv = torch.empty((n, 3), dtype=torch.float32, device=xyz.device)
        dth = torch.empty((MAX_POT_PARAMS,), dtype=torch.float32, device=xyz.device) if want_dtheta else None
        with self._guard(xyz.device):
            self._check(self._api().mdg_pair_hvp(self._h, int(kind), _farr(params, MAX_POT_PARAMS), len(params), _ptr(xyz), n,
                                              _ptr(avec), _ptr(hv), _ptr(dth), self._stream(xyz.device)))
        return hv, dth

    # -- K6 ---------------------------------------------------------------------------------
    def rdf_accumulate(self, xyz, cell3, start, end, nbins, width, count, sel_a=None, sel_b=None):
        self._require(xyz, "xyz")
        xyz = xyz.detach().to(torch.float32).contiguous()
        dev = xyz.device
        with self._guard(dev):
            self._check(self._api().mdg_rdf_accumulate(self._h, _ptr(xyz), xyz.shape[0], _farr(cell3, 3), float(start), float(end),
                                            int(nbins), float(width) if width else 0.0, _ptr(sel_a), _ptr(sel_b),
                                            _ptr(count), self._stream(dev)))
        self._keepalive = (xyz, sel_a, sel_b)

    def vacf(self, vel, t_range):
        """velocity autocorrelation of a (frames, N, dim) trajectory for lags 0 .. t_range - 1 (observable.py:153-163)"""
        self._require(vel, "vel")
        vel = vel.detach().to(torch.float32).contiguous()
        assert vel.dim() == 3
        dev = vel.device
        out = torch.empty((int(t_range),), dtype=torch.float32, device=dev)
        with self._guard(dev):
            self._check(self._api().mdg_vacf(self._h, _ptr(vel), vel.shape[0], vel.shape[1], vel.shape[2], int(t_range), _ptr(out),
                                             self._stream(dev)))
        self._keepalive = (vel,)
        return out

    # -- K4 + driver ------------------------------------------------------------------------
    def md_run(self, params, mass, v0, q0, pv0, tgrid, want_energy=False, out=None):
        """Runs one epoch; returns (traj_v, traj_q, traj_pv or None, last_energy or None).
        `out=(traj_v, traj_q)` reuses caller-allocated (n_frames, N, 3) fp32 CUDA buffers."""
        for t, nm in ((mass, "mass"), (v0, "v0"), (q0, "q0")):
            self._require(t, nm)
        dev = q0.device
        n = q0.shape[0]
        n_grid = len(tgrid)
        stride = max(1, params.traj_stride)
        n_frames = (n_grid - 1) // stride + 1
        if out is not None:
            tv, tq = out
            assert tv.shape == (n_frames, n, 3) and tq.shape == (n_frames, n, 3) and tv.is_contiguous() and tq.is_contiguous()
        else:
            tv = torch.empty((n_frames, n, 3), dtype=torch.float32, device=dev)
            tq = torch.empty((n_frames, n, 3), dtype=torch.float32, device=dev)
        M = params.n_chains if params.integrator == INT_NHC else 0
        hpv = (ctypes.c_float * max(1, n_frames * M))()
        hpv0 = _farr(pv0 if M else [0.0], max(1, M))
        tg = _farr(tgrid)
        e = ctypes.c_float(0.0)
        with self._guard(dev):
            self._check(self._api().mdg_md_run(self._h, ctypes.byref(params), n, _ptr(mass), _ptr(v0), _ptr(q0), hpv0, tg, n_grid,
                                    _ptr(tv), _ptr(tq), hpv if M else None,
                                    ctypes.byref(e) if want_energy else None, self._stream(dev)))
        tpv = None
        if M:
            tpv = torch.tensor(list(hpv), dtype=torch.float32).reshape(n_frames, M).to(dev)
        return tv, tq, tpv, (e.value if want_energy else None)

    def md_run_gnn(self, params, model, z, mass, v0, q0, pv0, tgrid):
        """One epoch with a SchNet (+ pair priors) force field on the device engine (mdg_md_run_gnn); `params` is a
        filled GnnMdParams, `model` = (SchnetModel, keepalive).  Returns (traj_v, traj_q, traj_pv or None)."""
        for t, nm in ((mass, "mass"), (v0, "v0"), (q0, "q0")):
            self._require(t, nm)
        dev = q0.device
        n = q0.shape[0]
        n_grid = len(tgrid)
        stride = max(1, params.traj_stride)
        n_frames = (n_grid - 1) // stride + 1
        tv = torch.empty((n_frames, n, 3), dtype=torch.float32, device=dev)
        tq = torch.empty((n_frames, n, 3), dtype=torch.float32, device=dev)
        M = params.n_chains if params.integrator == INT_NHC else 0
        hpv = (ctypes.c_float * max(1, n_frames * M))()
        hpv0 = _farr(pv0 if M else [0.0], max(1, M))
        z = z.to(dev, torch.int64).contiguous() if z is not None else None
        with self._guard(dev):
            self._check(self._api().mdg_md_run_gnn(self._h, ctypes.byref(params),
                                                ctypes.byref(model[0]) if model is not None else None, _ptr(z), n, _ptr(mass),
                                                _ptr(v0), _ptr(q0), hpv0, _farr(tgrid), n_grid, _ptr(tv), _ptr(tq),
                                                hpv if M else None, None, self._stream(dev)))
        tpv = torch.tensor(list(hpv), dtype=torch.float32).reshape(n_frames, M).to(dev) if M else None
        return tv, tq, tpv

    def set_pair_filter(self, sel_a=None, sel_b=None, ex_keys=None):
        self._check(self._api().mdg_set_pair_filter(self._h, _ptr(sel_a), _ptr(sel_b), _ptr(ex_keys),
                                         0 if ex_keys is None else int(ex_keys.numel())))
        self._filter_keepalive = (sel_a, sel_b, ex_keys)

    # -- K5: SchNet cfconv aggregation ---------------------------------------------------------
    def graph_build(self, nbr, n):
        self._require(nbr, "nbr_list")
        nbr = nbr.to(torch.int64).contiguous()
        with self._guard(nbr.device):
            self._check(self._api().mdg_graph_build(self._h, _ptr(nbr), nbr.shape[0], int(n), self._stream(nbr.device)))
        self._graph_keepalive = nbr
        self._graph_n = int(n)

    def cfconv_agg(self, h, W):
        self._require(h, "h")
        h, W = h.contiguous(), W.contiguous()
        out = torch.empty_like(h)
        with self._guard(h.device):
            self._check(self._api().mdg_cfconv_agg(self._h, _ptr(h), _ptr(W), h.shape[0], h.shape[1], _ptr(out), self._stream(h.device)))
        return out

    def cfconv_edge_grad(self, h, g, n_edges):
        h, g = h.contiguous(), g.contiguous()
        gW = torch.empty((n_edges, h.shape[1]), dtype=torch.float32, device=h.device)
        with self._guard(h.device):
            self._check(self._api().mdg_cfconv_edge_grad(self._h, _ptr(h), _ptr(g), h.shape[0], h.shape[1], _ptr(gW), self._stream(h.device)))
        return gW

    def schnet_energy_force(self, model, z, xyz, nbr, offsets, off_scale=(1.0, 1.0, 1.0), want_force=True):
        """SchNet energy (0-d) and forces (N,3) over the given reference-layout list; `model` = (SchnetModel, keepalive)
        from schnet_model_struct().  off_scale (1,1,1) = the reference's raw-offset quirk, cell lengths = true PBC."""
        self._require(xyz, "xyz")
        xyz = xyz.detach().to(torch.float32).contiguous()
        z = z.to(xyz.device, torch.int64).contiguous()
        nbr = nbr.to(xyz.device, torch.int64).contiguous()
        offsets = offsets.detach().to(xyz.device, torch.float32).contiguous()
        n = xyz.shape[0]
        e = torch.empty((), dtype=torch.float32, device=xyz.device)
        f = torch.empty((n, 3), dtype=torch.float32, device=xyz.device) if want_force else None
        with self._guard(xyz.device):
            self._check(self._api().mdg_schnet_energy_force(self._h, ctypes.byref(model[0]), _ptr(z), _ptr(xyz), n, _ptr(nbr),
                                                         _ptr(offsets), nbr.shape[0], _farr(off_scale, 3), _ptr(e), _ptr(f),
                                                         self._stream(xyz.device)))
        self._schnet_keepalive = (z, xyz, nbr, offsets)
        return e, f

    # -- bonded terms -------------------------------------------------------------------------
    def bonded_force(self, terms, xyz, cell3, want_force=True, want_dparams=False):
        """(E (2,) = (E_bond, E_angle), F (N,3) or None, dE/d(k_bond, ro, k_angle, theta0) (4,) or None) for a filled
        BondedTerms struct (its tensors kept alive by the caller) - mdg_bonded_force."""
        self._require(xyz, "xyz")
        xyz = xyz.detach().to(torch.float32).contiguous()
        n = xyz.shape[0]
        e = torch.empty((2,), dtype=torch.float32, device=xyz.device)
        f = torch.empty((n, 3), dtype=torch.float32, device=xyz.device) if want_force else None
        dp = torch.empty((4,), dtype=torch.float32, device=xyz.device) if want_dparams else None
        with self._guard(xyz.device):
            self._check(self._api().mdg_bonded_force(self._h, ctypes.byref(terms), _ptr(xyz), n, _farr(cell3, 3), _ptr(e), _ptr(f),
                                                  _ptr(dp), self._stream(xyz.device)))
        return e, f, dp

    # -- multi-GPU ----------------------------------------------------------------------------
    def dist_init(self, group=None):
        """Join this context to an NCCL communicator spanning torch.distributed's (default) group: rank 0
        creates the NCCL unique id, it is broadcast through torch.distributed, every rank calls
        mdg_dist_init.  After this, md_run integrates only this rank's slab (see include/mdgrad_b200.h)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        path = nccl_library_path().encode()
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            self._check(self._api().mdg_dist_unique_id(path, buf))
        t = torch.tensor(list(buf.raw), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            t = t.to(self.device)
        dist.broadcast(t, src=0, group=group)
        ident = bytes(t.cpu().tolist())
        with self._guard(self.device):
            self._check(self._api().mdg_dist_init(self._h, path, ident, rank, world))
        self.rank, self.world = rank, world

    def dist_finalize(self):
        self._check(self._api().mdg_dist_finalize(self._h))

    def set_profile(self, enable):
        self._check(self._api().mdg_set_profile(self._h, int(bool(enable))))

    def get_profile(self):
        out = (ctypes.c_double * 2)()
        self._check(self._api().mdg_get_profile(self._h, out))
        return {"force_ms": out[0], "force_launches": int(out[1])}

    def stats(self):
        out = (ctypes.c_int64 * 8)()
        self._check(self._api().mdg_get_stats(self._h, out))
        keys = ["launches", "rebuilds", "entries", "maxrow_or_K", "ncx", "ncy", "ncz", "path"]
        return dict(zip(keys, list(out)))


def nccl_library_path():
    """The libnccl.so.2 that torch itself uses (one NCCL runtime per process - SURVEY A9)."""
    sp = os.path.dirname(os.path.dirname(torch.__file__))
    cand = os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so.2")
    return cand if os.path.exists(cand) else "libnccl.so.2"


def slab_plan(ncz, world, rank):
    """(zlo, zhi, rank_below, rank_above): host-only logic of the slab decomposition."""
    out = (ctypes.c_int * 4)()
    check(load().mdg_slab_plan(int(ncz), int(world), int(rank), out))
    return tuple(out)


def pair_dis_fwd(xyz, nbr, offsets, cell3):
    require_cuda(xyz, "xyz")
    P = nbr.shape[0]
    dis = torch.empty((P,), dtype=torch.float32, device=xyz.device)
    with _guard(xyz.device):
        check(load().mdg_pair_dis_fwd(_ptr(xyz), xyz.shape[0], _ptr(nbr), _ptr(offsets), P, _farr(cell3, 3), _ptr(dis),
                                      _stream(xyz.device)))
    return dis


def pair_dis_bwd(xyz, nbr, offsets, cell3, dis, grad_dis):
    require_cuda(xyz, "xyz")
    g = torch.empty_like(xyz)
    with _guard(xyz.device):
        check(load().mdg_pair_dis_bwd(_ptr(xyz), xyz.shape[0], _ptr(nbr), _ptr(offsets), nbr.shape[0], _farr(cell3, 3),
                                      _ptr(dis), _ptr(grad_dis), _ptr(g), _stream(xyz.device)))
    return g
