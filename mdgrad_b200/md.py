"""Simulation driver and equations of motion - mirror of reference torchmd/md.py:
Simulations :14-96, NVE :98-156, NoseHooverChain :158-249.

API, state layout, log/check-point semantics (fp64 host wrap between epochs, last frame per epoch
logged, `frequency` grid points = frequency-1 steps) are the reference's.  What changes is where the
epoch runs: for NVE / NoseHooverChain over an analytic `PairPotentials` with
topology_update_freq == 1 the whole epoch (list rebuilds, forces, integrator, trajectory capture)
is ONE call into the fused device engine `mdg_md_run`; every other combination goes through the
generic op-level solvers in `sovlers.py`, exactly like the reference.
"""
import math

import numpy as np
import torch

from . import _lib
from .sovlers import odeint, odeint_adjoint, second_order
from .system import HAVE_ASE

if HAVE_ASE:  # pragma: no cover
    from ase import units
    from ase.geometry import wrap_positions
else:
    from ._ase_compat import units, wrap_positions


def compute_grad(inputs, output, create_graph=True, retain_graph=True):
    """d output / d inputs (reference nff/utils/scatter.py:5-21)."""
    assert inputs.requires_grad
    g, = torch.autograd.grad(output, inputs, grad_outputs=output.data.new(output.shape).fill_(1),
                             create_graph=create_graph, retain_graph=retain_graph)
    return g


def _host_to_device(x, device):
    """`torch.Tensor(x).to(device)` of the reference (md.py:66,69,156,249), value for value: fp32 rounding of the
    host fp64 state.  numpy does the cast (torch's legacy constructor walks a float64 array element by element,
    ~60 ms for 256k atoms) and the copy starts from that fp32 buffer."""
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(device)
    return torch.Tensor(x).to(device)


def _device_wrap(q, cell, eps=1e-7):
    """`wrap_positions(q, cell)` of the epoch hand-off (reference md.py:60-71 -> ase.geometry.wrap_positions, center 0.5,
    eps 1e-7, all axes periodic) evaluated ON THE DEVICE, bit for bit: the host version is six separately rounded fp64
    elementwise operations (x * (1/L), - shift, - floor, + shift, * L) on the fp32 frame and a final rounding to fp32;
    the same six IEEE operations as separate fp64 tensor ops give the same bits (no fusion between eager ops), so the
    state never has to leave the device between the epochs of one `simulate()` call
    (tests/test_cabi_and_host.py::test_device_wrap_is_bit_identical).  Orthorhombic cells only -> None otherwise."""
    c = np.asarray(cell, dtype=float)
    if c.shape == (3,):
        c = np.diag(c)
    if c.shape != (3, 3) or (c - np.diag(np.diag(c))).any() or not np.all(np.diag(c) != 0):
        return None
    L = np.diag(c)
    shift = np.asarray((0.5, 0.5, 0.5), dtype=float) - 0.5 - eps
    t64 = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device=q.device)      # noqa: E731
    f = q.detach().to(torch.float64) * t64(1.0 / L)
    sh = t64(shift)
    f -= sh
    f -= torch.floor(f)
    f += sh
    f *= t64(L)
    return f.to(torch.float32)


def _torch_device(device):
    return torch.device("cuda:%d" % device) if isinstance(device, int) else torch.device(device)


class Simulations():
    """Simulations(system, integrator, wrap=True, method="NH_verlet")   (reference md.py:28-40)."""

    def __init__(self, system, integrator, wrap=True, method="NH_verlet"):
        self.system = system
        self.device = system.device
        self.integrator = integrator
        self.solvemethod = method
        self.wrap = wrap
        self.keys = self.integrator.state_keys
        self.initialize_log()

    def initialize_log(self):
        self.log = {key: [] for key in self.keys}

    def update_log(self, trajs):
        """keep only the LAST frame of the epoch, as numpy (reference md.py:47-52)"""
        for i, key in enumerate(self.keys):
            self.log[key].append(trajs[i][-1].detach().cpu().numpy())

    def update_states(self):
        if "positions" in self.log:
            self.system.set_positions(self.log["positions"][-1])
        if "velocities" in self.log:
            self.system.set_velocities(self.log["velocities"][-1])

    def get_check_point(self):
        """Last logged frame as device tensors; positions re-wrapped on the host in fp64
        (reference md.py:60-71)."""
        if hasattr(self, "log"):
            states = [_host_to_device(self.log[key][-1], self.device) for key in self.log]
            if self.wrap:
                # the fp64 wrap of the logged (fp32) frame: on the device when the cell is orthorhombic - the same six IEEE
                # operations as the host version, bit for bit (tests/test_cabi_and_host.py::test_device_wrap_is_bit_identical),
                # without 8 ms of host arithmetic per call at 256k atoms - else on the host as the reference
                k = list(self.log.keys()).index("positions")
                wrapped = _device_wrap(states[k], self.system.get_cell()) if states[k].is_cuda else None
                if wrapped is None:
                    wrapped = _host_to_device(wrap_positions(self.log["positions"][-1], self.system.get_cell()), self.device)
                states[k] = wrapped
            return states
        raise ValueError("No log available")

    def _simulate(self, steps, dt, frequency):
        if self.log["positions"] == []:
            states = self.integrator.get_inital_states(self.wrap)
        else:
            states = self.get_check_point()
        sim_epochs = int(steps // frequency)
        t = torch.Tensor([dt * i for i in range(frequency)]).to(self.device)
        # Epoch hand-off.  The reference logs the last frame to the host, updates the System and re-reads the log
        # (fp64 wrap on the host) after EVERY epoch (md.py:92-95).  None of that is observable before simulate()
        # returns, and the next epoch's start state is a pure function of the last frame, so between epochs the
        # state stays on the device (bit-identical fp64 wrap there) and the host-side log / System are brought up to
        # date in one flush - same log, same System, same trajectory, without a host round trip per epoch.
        pending = self._pending
        for epoch in range(sim_epochs):
            # Only the LAST epoch's stacked trajectory is returned (and can be back-propagated through: the epochs before it
            # are cut off by the host log in the reference, md.py:92-95).  An earlier epoch is needed for its last frame only,
            # so the fused engine is asked for just that (first + last frame): no (frequency, N, 3) x 2 allocation and no
            # per-step frame writes for it.
            last_only = (self.device_handoff and epoch + 1 < sim_epochs and self.integrator.adjoint and
                         hasattr(self.integrator, "_native_forward"))
            trajs = None
            if last_only:
                with torch.no_grad():
                    self.integrator._traj_last_only = True
                    try:
                        trajs = self.integrator._native_forward(tuple(states), t, self.solvemethod)
                    finally:
                        self.integrator._traj_last_only = False
            if trajs is not None:
                pass
            elif self.integrator.adjoint:
                trajs = odeint_adjoint(self.integrator, states, t, method=self.solvemethod)
            else:
                # adjoint=False: the whole trajectory goes on the autograd tape (reference md.py:84-88) and the forces on it
                # are differentiated again by .backward() - so every interaction module records its twice-differentiable
                # form while the epoch is integrated (the fused first-order kernels would cut the parameters off the tape)
                for var in states:
                    var.requires_grad = True
                with second_order(self.integrator):
                    trajs = odeint(self.integrator, tuple(states), t, method=self.solvemethod)
            if self.device_handoff and epoch + 1 == sim_epochs and self._device_wrap_possible():
                nxt = ()                              # last epoch: nobody consumes a next start state - only its frames are logged
            else:
                nxt = self._device_check_point(trajs) if self.device_handoff else None
            if nxt is None:                           # reference order of operations, on the host
                self._flush_log(pending)
                self.update_log(trajs)
                self.update_states()
                states = self.get_check_point()
                continue
            pending.append(self._stage_frames([tr[-1] for tr in trajs]))
            if len(pending) >= self._flush_every:     # bound the device / pinned memory held by un-logged frames
                self._flush_log(pending)
            states = nxt
        self._flush_log(pending)
        return trajs

    def simulate(self, steps=1, dt=1.0 * units.fs, frequency=1):
        """steps//frequency epochs of frequency-1 integration steps each; returns the stacked
        trajectory of the LAST epoch (reference md.py:73-96)."""
        self._pending = []
        try:
            return self._simulate(steps, dt, frequency)
        finally:
            # epochs that finished before an exception (capacity / skin / non-finite / OOM / KeyboardInterrupt) are logged,
            # as in the reference, which logs after every epoch
            self._flush_log(self._pending)

    device_handoff = True     # False: host round trip after every epoch, literally as the reference
    _flush_every = 8          # epochs whose last frames wait on the device before one pinned-memory flush

    def _device_wrap_possible(self):
        """True when `_device_check_point` would succeed (no wrap requested, or an orthorhombic cell)."""
        if not (self.wrap and "positions" in self.keys):
            return True
        c = np.asarray(self.system.get_cell(), dtype=float)
        if c.shape == (3,):
            c = np.diag(c)
        return bool(c.shape == (3, 3) and not (c - np.diag(np.diag(c))).any() and np.all(np.diag(c) != 0))

    def _device_check_point(self, trajs):
        """The states `get_check_point()` would return after logging `trajs`, computed on the device; None if the
        configuration needs the host path (non-orthorhombic cell)."""
        last = [tr[-1].detach().clone() for tr in trajs]
        if self.wrap and "positions" in self.keys:
            k = self.keys.index("positions")
            wrapped = _device_wrap(last[k], self.system.get_cell())
            if wrapped is None:
                return None
            last[k] = wrapped
        return last

    def _stage_frames(self, frames):
        """Last frames of a finished epoch on their way to the host log.  On CUDA the copy into pinned memory is issued right
        away on a side stream, so it overlaps the next epoch's kernels; `_flush_log` only waits for that stream and moves the
        pinned data into the numpy arrays of the log."""
        if not frames[0].is_cuda:
            return [fr.detach().clone() for fr in frames]
        dev = frames[0].device
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(dev)
            self._pin, self._pin_used = None, 0
        nb_epoch = sum(fr.numel() * fr.element_size() for fr in frames)
        if self._pin is None or self._pin_used + nb_epoch > self._pin.numel():
            if self._pin_used:                       # (frames of another size mid-interval: drain what is staged first)
                self._flush_log(self._pending)
            if self._pin is None or nb_epoch > self._pin.numel():
                # sized once for a full flush interval (page-locking costs ~0.4 ms per MB: it must not recur per call)
                self._pin = torch.empty(nb_epoch * self._flush_every, dtype=torch.uint8, pin_memory=True)
            self._pin_used = 0
        self._side.wait_stream(torch.cuda.current_stream(dev))
        views = []
        with torch.cuda.stream(self._side):
            for fr in frames:
                fr = fr.detach()
                nb = fr.numel() * fr.element_size()
                v = self._pin[self._pin_used:self._pin_used + nb].view(fr.dtype).view(fr.shape)
                v.copy_(fr, non_blocking=True)
                fr.record_stream(self._side)         # the trajectory block may be recycled by the next epoch's allocation
                views.append(v)
                self._pin_used += nb
        return views

    def _flush_log(self, pending):
        """append the deferred last frames to the log (numpy, as update_log) and update the System once"""
        if not pending:
            return
        if getattr(self, "_side", None) is not None and pending[0][0].is_pinned():
            self._side.synchronize()
            for frames in pending:
                for key, v in zip(self.keys, frames):
                    dst = np.empty(tuple(v.shape), dtype=v.numpy().dtype)
                    torch.from_numpy(dst).copy_(v)          # torch's multi-threaded host copy (numpy's is one thread: ~3 GB/s)
                    self.log[key].append(dst)
            self._pin_used = 0
        else:
            for frames in pending:
                for key, fr in zip(self.keys, frames):
                    self.log[key].append(fr.cpu().numpy())
        pending.clear()
        self.update_states()


class _EOM(torch.nn.Module):
    """Shared plumbing of the two equations of motion."""

    engine_skin = None          # Verlet skin of the fused engine; None = 0.12 * cutoff
    engine_rebuild_every = None  # None = chosen from the initial velocities, halved on violation

    def _init_common(self, potentials, system, adjoint, topology_update_freq):
        self.model = potentials
        self.system = system
        self.device = system.device
        self.mass = torch.Tensor(system.get_masses()).to(self.device)
        self.N_dof = self.mass.shape[0] * system.dim
        self.dim = system.dim
        self.adjoint = adjoint
        self.topology_update_freq = topology_update_freq
        self.update_count = 0
        self._engine_ctx = None
        self._engine_K = None
        self.last_engine_stats = None

    def update_topology(self, q):
        """rebuild every topology_update_freq-th EVALUATION (reference md.py:200-204)"""
        if self.update_count % self.topology_update_freq == 0:
            self.model._reset_topology(q)
        self.update_count += 1

    def force(self, q):
        """-dU/dq at q with the reference's side effects (requires_grad on q, topology update, md.py:216-228)."""
        if not torch.is_grad_enabled() and _lib.on_device(q) and getattr(self.model, "native_ready", lambda: False)():
            # nobody records a graph (forward pass of the adjoint solver, plain MD): forces come straight from the
            # native programs (pair-force kernel / SchNet energy+force), no autograd tape
            self.update_topology(q)
            return self.model.native_force(q.detach())
        with torch.set_grad_enabled(True):
            if self.adjoint:
                q.requires_grad = True
            self.update_topology(q)
            u = self.model(q)
            return -compute_grad(inputs=q, output=u.sum(-1), create_graph=_needs_graph(self.model))

    # -- analytic adjoint dynamics -----------------------------------------------------------------------------------
    _POWER_LAW = (_lib.POT_LJ, _lib.POT_LJFAM, _lib.POT_LJ69, _lib.POT_EXV, _lib.POT_BUCK, _lib.POT_MORSE)   # kinds with closed-form second order

    def _native_second_order(self, q):
        """The pair members when every member's force has closed-form second-order products (mdg_pair_hvp): a single
        PairPotentials or a Stack of PairPotentials over the analytic kinds (e.g. the species-pair members of
        scripts/fit_2_comp.py:182); else None."""
        from .interface import PairPotentials, Stack
        if not _lib.on_device(q) or getattr(self, "disable_native_adjoint", False):
            return None
        members = list(self.model.models.values()) if type(self.model) is Stack else [self.model]
        for m in members:
            if type(m) is not PairPotentials:
                return None
            spec = m.native_kind()
            if spec is None or spec[0] not in self._POWER_LAW:
                return None
        return members or None

    def native_augmented(self, tt, y_aug, n):
        """Augmented adjoint dynamics (f, vjp_y, vjp_t, vjp_params) at (y, adj) WITHOUT autograd: the reference
        differentiates the autograd force a second time (sovlers.py:216-236); here the force comes from the force
        kernel, (dF/dq)^T a and (dF/dtheta)^T a from the analytic Hessian-vector kernel (summed over the members of a
        Stack, each on its own list), the thermostat algebra is written out.  Returns None when this configuration is
        not covered (the caller then uses autograd)."""
        y, adj = y_aug[:n], y_aug[n:2 * n]
        members = self._native_second_order(y[1])
        if members is None:
            return None
        with torch.no_grad():
            v, q = y[0], y[1]
            self.update_topology(q)                      # same evaluation count / list refresh as func(t, y)
            F = None
            for m in members:
                Fm = m.native_force(q)
                F = Fm if F is None else F + Fm
            f = self.derivative(tt, y, F)
            cot = tuple(-a for a in adj)
            a_q, gv, gpv = self._vjp_algebra(y, cot)
            hv, dtheta = None, {}
            for m in members:
                kind, values, ptensors = m.native_kind()
                hvm, dth = m._ctx.pair_hvp(kind, values, q, a_q)
                hv = hvm if hv is None else hv + hvm
                for k, pt in enumerate(ptensors):
                    dtheta[id(pt)] = dth[k].reshape(-1) + dtheta.get(id(pt), 0.0)
            vjp_y = (gv, hv) + ((gpv,) if gpv is not None else ())
            parts = [dtheta[id(p_)] if id(p_) in dtheta else torch.zeros(p_.numel(), device=q.device) for p_ in self.parameters()]
            vjp_p = torch.cat(parts) if parts else torch.tensor(0.).to(q)
            return (*f, *vjp_y, torch.zeros_like(tt), vjp_p)

    # -- fused engine -------------------------------------------------------------------------
    def _native_spec(self, method):
        from .interface import PairPotentials
        m = self.model
        if type(m) is not PairPotentials or m.native_kind() is None:
            return None
        if self.topology_update_freq != 1 or method != self._native_method:
            return None
        if any(p.requires_grad and p.grad_fn is not None for p in m.parameters()):
            return None
        return m

    def _gnn_members(self, method):
        """(GNNPotentials, [PairPotentials priors]) when the model is a native-ready GNNPotentials or a Stack of exactly
        one GNNPotentials and analytic PairPotentials priors, else None (device engine mdg_md_run_gnn)."""
        from .gnn import GNNPotentials
        from .interface import PairPotentials, Stack
        if self.topology_update_freq != 1 or method != self._native_method or getattr(self, "disable_gnn_engine", False):
            return None
        from .bonded import AnglePotentials, BondPotentials
        members = list(self.model.models.values()) if type(self.model) is Stack else [self.model]
        gnns = [m for m in members if type(m) is GNNPotentials]
        priors = [m for m in members if type(m) is PairPotentials]
        bonds = [m for m in members if type(m) is BondPotentials]
        angles = [m for m in members if type(m) is AnglePotentials]
        if len(gnns) > 1 or len(bonds) > 1 or len(angles) > 1 or len(priors) > _lib.MAX_PRIORS:
            return None
        if len(gnns) + len(priors) + len(bonds) + len(angles) != len(members):
            return None
        if not gnns and (type(self.model) is not Stack or not (priors or bonds or angles)):
            return None                               # a single PairPotentials runs on the skin-list engine (mdg_md_run)
        if not all(m.native_ready() for m in members):
            return None
        if any(p.requires_grad and p.grad_fn is not None for p in self.model.parameters()):
            return None
        self._bonded_members = (bonds[0] if bonds else None, angles[0] if angles else None)
        return (gnns[0] if gnns else None), priors

    def _native_forward_gnn(self, y0, t, method):
        mem = self._gnn_members(method)
        if mem is None or len(t) < 1 or (len(t) > 1 and not bool((t[1:] > t[:-1]).all())):
            return None
        gnn, priors = mem
        v0, q0 = y0[0], y0[1]
        if not _lib.on_device(q0):
            return None
        p = _lib.GnnMdParams()
        p.integrator = self._native_integrator
        if self._native_integrator == _lib.INT_NHC:
            p.n_chains = int(self.num_chains)
            Qh = self.Q.detach().cpu()
            for k in range(self.num_chains):
                p.Q[k] = float(Qh[k])
            p.T = float(self.T)
        p.ndof = int(self.N_dof)
        bond, angle = self._bonded_members
        L = next(m._L for m in (gnn, *priors, bond, angle) if m is not None)
        for k in range(3):
            p.cell[k] = L[k]
            p.off_scale[k] = 1.0            # reference quirk: raw integer offsets, not multiplied by the cell
        exk = None
        if gnn is not None:
            p.cutoff = float(gnn.cutoff)
            if gnn.pbc_mode != "reference":
                for k in range(3):
                    p.off_scale[k] = L[k]
            exk = gnn._ex_keys(q0.device)
        p.d_ex_keys = 0 if exk is None else exk.data_ptr()
        p.n_ex = 0 if exk is None else int(exk.numel())
        p.n_priors = len(priors)
        for k, pr in enumerate(priors):
            kind, values, _ = pr.native_kind()
            s_ = p.priors[k]
            s_.ctx = pr._ctx._h.value
            s_.kind = kind
            for i, v in enumerate(values):
                s_.params[i] = float(v)
            s_.n_params = len(values)
            s_.cutoff = float(pr.cutoff)
            s_.d_sel_a = 0 if pr._sel[0] is None else pr._sel[0].data_ptr()
            s_.d_sel_b = 0 if pr._sel[1] is None else pr._sel[1].data_ptr()
            s_.d_ex_keys = 0 if pr._exk is None else pr._exk.data_ptr()
            s_.n_ex = 0 if pr._exk is None else int(pr._exk.numel())
        p.traj_stride = 1
        bonded_keep = None
        if bond is not None or angle is not None:
            # one reference list over both members: bond slots first, then two slots per angle (include/mdgrad_b200.h)
            from .bonded import term_refs
            key = (id(bond), id(angle))
            if getattr(self, "_bonded_cache", (None,))[0] != key:
                bt = bond.top.cpu().numpy() if bond is not None else None
                at = angle.top.cpu().numpy() if angle is not None else None
                rs, rf = term_refs(q0.shape[0], bt, at)
                self._bonded_cache = (key, torch.from_numpy(rs).to(q0.device), torch.from_numpy(rf).to(q0.device),
                                      bond.top.contiguous() if bond is not None else None,
                                      angle.top.contiguous() if angle is not None else None)
            _, rs_d, rf_d, bt_d, at_d = self._bonded_cache
            b_ = p.bonded
            if bond is not None:
                b_.d_bond_top, b_.n_bonds = bt_d.data_ptr(), int(bt_d.shape[0])
                b_.k_bond, b_.r0 = bond._scalars()
            if angle is not None:
                b_.d_angle_top, b_.n_angles = at_d.data_ptr(), int(at_d.shape[0])
                b_.k_angle, b_.theta0 = angle._scalars()
            b_.d_ref_start, b_.d_refs = rs_d.data_ptr(), (rf_d.data_ptr() if rf_d.numel() else 0)
            bonded_keep = self._bonded_cache
        if self._engine_ctx is None:
            self._engine_ctx = _lib.Context(q0.device)
        ctx = self._engine_ctx
        tl = [float(x) for x in t.detach().cpu()]
        pv0 = [float(x) for x in y0[2].detach().cpu()] if len(y0) > 2 else []
        mass = self.mass.to(q0.device, torch.float32).contiguous()
        z = gnn.inputs["nxyz"][:, 0].to(torch.int64).contiguous() if gnn is not None else None
        tv, tq, tpv = ctx.md_run_gnn(p, gnn._native_model() if gnn is not None else None, z, mass,
                                     v0.detach().to(torch.float32).contiguous(),
                                     q0.detach().to(torch.float32).contiguous(), pv0, tl)
        self._gnn_keepalive = (exk, z, mass, bonded_keep)
        self.last_engine_stats = ctx.stats()
        if len(tl) > 1:
            self.update_count += 2 * (len(tl) - 1)       # two evaluations per step in the reference
            self.model._reset_topology(tq[-1])            # python-visible lists = those of the reference's last evaluation
        return (tv, tq, tpv) if tpv is not None else (tv, tq)

    def _native_forward(self, y0, t, method):
        """Run the epoch on the fused engine; returns the stacked trajectory or None if this
        configuration is not covered (caller then takes the generic route)."""
        m = self._native_spec(method)
        if m is None:
            return self._native_forward_gnn(y0, t, method)
        if len(t) < 1:
            return None
        if len(t) > 1 and not bool((t[1:] > t[:-1]).all()):
            return None
        v0, q0 = y0[0], y0[1]
        if not _lib.on_device(v0):
            _lib.require_cuda(v0, "state tensors")
        kind, values, _ = m.native_kind()
        n = q0.shape[0]
        p = _lib.MdParams()
        p.integrator = self._native_integrator
        p.pot_kind = kind
        for i, v in enumerate(values):
            p.pot_params[i] = float(v)
        p.cutoff = float(m.cutoff)
        L = m._L
        for k in range(3):
            p.cell[k] = L[k]
        if self._native_integrator == _lib.INT_NHC:
            p.n_chains = int(self.num_chains)
            Qh = self.Q.detach().cpu()
            for k in range(self.num_chains):
                p.Q[k] = float(Qh[k])
            p.T = float(self.T)
        p.ndof = int(self.N_dof)
        tl = [float(x) for x in t.detach().cpu()]
        skin = self.engine_skin if self.engine_skin is not None else 0.12 * float(m.cutoff)
        p.skin = float(skin)
        K = self.engine_rebuild_every or self._engine_K
        if K is None:
            dtmax = max([b - a for a, b in zip(tl[:-1], tl[1:])] + [0.0])
            vmax = float(v0.detach().norm(dim=1).max()) if n else 0.0
            K = 1 if vmax * dtmax <= 0 else int(max(1, min(64, math.floor(0.5 * skin / (1.1 * vmax * dtmax)))))
        p.rebuild_every = int(K)
        p.traj_stride = max(1, len(tl) - 1) if getattr(self, "_traj_last_only", False) else 1
        if self._engine_ctx is None:
            self._engine_ctx = _lib.Context(q0.device)
        ctx = self._engine_ctx
        ctx.set_pair_filter(m._sel[0], m._sel[1], m._exk)
        pv0 = [float(x) for x in y0[2].detach().cpu()] if len(y0) > 2 else []
        mass = self.mass.to(q0.device, torch.float32).contiguous()
        tv, tq, tpv, _ = ctx.md_run(p, mass, v0.detach().to(torch.float32).contiguous(),
                                    q0.detach().to(torch.float32).contiguous(), pv0, tl)
        st = ctx.stats()
        self.last_engine_stats = st
        self._engine_K = int(st["maxrow_or_K"]) if skin > 0 else None
        self.update_count += 2 * (len(tl) - 1)       # two evaluations per step in the reference
        if len(tl) > 1:
            m._mark_stale(tq[-1])                    # python-visible list = that of the reference's last evaluation (built lazily)
        return (tv, tq, tpv) if tpv is not None else (tv, tq)


class NVE(_EOM):
    """NVE(potentials, system, adjoint=True, topology_update_freq=1)   (reference md.py:112)."""
    _native_method = "verlet"
    _native_integrator = _lib.INT_NVE

    def __init__(self, potentials, system, adjoint=True, topology_update_freq=1):
        super().__init__()
        self._init_common(potentials, system, adjoint, topology_update_freq)
        self.state_keys = ["velocities", "positions"]

    def forward(self, t, state):
        """dv/dt = f (not divided by the mass), dq/dt = v   (reference md.py:131-148)"""
        return self.derivative(t, state, self.force(state[1]))

    def derivative(self, t, state, f):
        return (f, state[0])

    def _vjp_algebra(self, y, cot):
        """dv/dt = F(q), dq/dt = v:  vector for the force products = c_v;  d/dv = c_q"""
        return cot[0], cot[1], None

    def get_inital_states(self, wrap=True):
        states = [self.system.get_velocities(), self.system.get_positions(wrap=wrap)]
        return [_host_to_device(var, self.system.device) for var in states]


class NoseHooverChain(_EOM):
    """NoseHooverChain(potentials, system, T, num_chains=2, Q=1.0, adjoint=True,
    topology_update_freq=1)   (reference md.py:179-198)."""
    _native_method = "NH_verlet"
    _native_integrator = _lib.INT_NHC

    def __init__(self, potentials, system, T, num_chains=2, Q=1.0, adjoint=True, topology_update_freq=1):
        super().__init__()
        self._init_common(potentials, system, adjoint, topology_update_freq)
        self.T = T
        self.target_ke = 0.5 * self.N_dof * T
        self.num_chains = num_chains
        Qs = np.array([Q, *[Q / len(system)] * (num_chains - 1)])     # md.py:191-193
        self.Q = torch.Tensor(Qs).to(self.device)
        self.state_keys = ["velocities", "positions", "baths"]

    def update_T(self, T):
        self.T = T

    def forward(self, t, state):
        """reference md.py:210-240"""
        return self.derivative(t, state, self.force(state[1]))

    def derivative(self, t, state, f):
        """thermostat algebra of md.py:221-240 for a given force"""
        with torch.set_grad_enabled(True):
            v, q, p_v = state[0], state[1], state[2]
            m = self.mass[:, None]
            p = v * m
            sys_ke = 0.5 * (p.pow(2) / m).sum()
            coupled = (p_v[0] * p.reshape(-1) / self.Q[0]).reshape(-1, 3)
            dpdt = f - coupled
            d0 = 2 * (sys_ke - self.T * self.N_dof * 0.5) - p_v[0] * p_v[1] / self.Q[1]
            dmid = (p_v[:-2].pow(2) / self.Q[:-2] - self.T) - p_v[2:] * p_v[1:-1] / self.Q[2:]
            dlast = p_v[-2].pow(2) / self.Q[-2] - self.T
            dvdt = dpdt / m
        return (dvdt, v, torch.cat((d0[None], dmid, dlast[None])))

    def _vjp_algebra(self, y, cot):
        """Vector-Jacobian products of `derivative` (md.py:221-240) w.r.t. (v, p_v) for cotangents (c_v, c_q, c_p), and
        the vector a_q = c_v / m that multiplies dF/dq and dF/dtheta:
            dv    = F/m - (p_0/Q_0) v                 dp_0 = 2 (ke - T ndof/2) - p_0 p_1 / Q_1,  ke = 1/2 sum m v^2
            dp_k  = (p_{k-1}^2/Q_{k-1} - T) - p_{k+1} p_k / Q_{k+1}        dp_last = p_{M-2}^2/Q_{M-2} - T"""
        return nhc_vjp_algebra(y[0], y[2], self.mass[:, None], self.Q, cot[0], cot[1], cot[2])

    def get_inital_states(self, wrap=True):
        states = [self.system.get_velocities(), self.system.get_positions(wrap=wrap), [0.0] * self.num_chains]
        return [_host_to_device(var, self.system.device) for var in states]


def nhc_vjp_algebra(v, pv, m, Q, cv, cq, cp):
    """(a_q, g_v, g_pv) for NoseHooverChain.derivative - see NoseHooverChain._vjp_algebra.  Pure tensor algebra
    (device agnostic) so that it can be checked against autograd on the CPU (tests/test_cabi_and_host.py)."""
    M = pv.shape[0]
    a_q = cv / m
    g_v = cq - (pv[0] / Q[0]) * cv + (2.0 * cp[0]) * (m * v)
    g_pv = torch.zeros_like(pv)
    g_pv[0] = -(cv * v).sum() / Q[0] - cp[0] * pv[1] / Q[1] + cp[1] * 2.0 * pv[0] / Q[0]
    for k in range(1, M):
        t = -cp[k - 1] * pv[k - 1] / Q[k]
        if k <= M - 2:
            t = t - cp[k] * pv[k + 1] / Q[k + 1]
        if k + 1 <= M - 1:
            t = t + cp[k + 1] * 2.0 * pv[k] / Q[k]
        g_pv[k] = t
    return a_q, g_v, g_pv


def _needs_graph(model):
    """The reference always builds the force with create_graph=True (scatter.py:18-19).  The fused
    first-order kernels cannot be differentiated again, so a second-order graph is only requested
    when the adjoint solver has switched the interaction modules to their pure-torch distance op."""
    mods = [m for m in model.modules() if hasattr(m, "second_order")]
    if not mods:
        return True
    return any(m.second_order for m in mods)


class Isomerization(torch.nn.Module):
    """reference md.py `Isomerization` (1-D toy dynamics of the isomerisation demos): outside the MD hot path - importable so that
    the reference's scripts load, loud when constructed."""

    def __init__(self, *a, **k):
        raise NotImplementedError("mdgrad_b200: Isomerization is outside the MD hot path this package implements")
