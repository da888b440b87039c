#!/usr/bin/env python
"""bench.py - MD steps/s of the B200-native hot path on BASELINE.json's configuration.

  python bench.py --gpus N --steps K --warmup W            (N>1 under torchrun: one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one NH-Verlet MD step of the whole box (list upkeep + pair forces + NHC integrator).
N=1 workload = BASELINE.json configs[1]: 256 000-atom LJ fluid (40^3 FCC, rho 0.845, jitter 0.05a),
LennardJones(1,1) cutoff 2.5, NoseHooverChain(Q=50, T=1, 5 chains), dt 0.005.
N>1: weak scaling - the box grows along z with the GPU count (256 000 atoms per GPU); one spatial slab of
whole cell layers per rank, NCCL ghost-layer halo exchange + one 2-double all-reduce per step (dist.cu).

Prints ONE JSON line (rank 0).  `value` = device-resident whole-job steps/s (inputs in HBM, CUDA
events, max over ranks); `e2e` = the same metric through the public API
(Simulations.simulate with host numpy state -> H2D, epochs, D2H of every epoch's last frame, fp64
host wrap) ; `roofline` = the pair-force kernel's algorithmic bytes / its measured launch time vs
MEASURED_PEAKS.json ; `cpu_baseline` = the C oracle (port of the reference's all-pairs algorithm)
on this box's host cores on a bounded sample.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import faulthandler

# keep stdout = ONE JSON line: libraries (e.g. NCCL's version banner) write to fd 1 behind python's back, so
# fd 1 is pointed at stderr for the whole run and the JSON line goes to a private duplicate of the real stdout
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


faulthandler.enable()
if os.environ.get("BENCH_WATCHDOG"):      # dump all python stacks and exit if the run takes longer than this
    faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
_T0 = time.time()


def log(msg):
    if os.environ.get("BENCH_VERBOSE", "1") != "0" and int(os.environ.get("RANK", "0")) == 0:
        print("[bench %7.1fs] %s" % (time.time() - _T0, msg), file=sys.stderr, flush=True)


RHO, RC, DT, TEMP, QBATH_256, CHAINS, MASS = 0.845, 2.5, 0.005, 1.0, 50.0, 5, 1.008


def qbath(n):
    """Heat-bath mass.  The reference scripts use Q=50 for their 256-atom boxes with bath masses
    [Q, Q/N, ...] (torchmd/md.py:191-193).  Taken literally at N=256000 that chain diverges within ~20
    steps - in the reference formulation itself (reproduced with the CPU oracle at N=6912) - because the
    chain starts at pv=0 and the anti-friction rate of the second link grows like N/Q.  Keeping the
    reference's equations and Q_k = Q/N rule, Q is scaled with N so the thermostat frequency equals the
    reference's own 256-atom runs: Q = 50 * N / 256.  (Same kernels, same bytes per step.)"""
    return QBATH_256 * max(n, 256) / 256.0
NCELL_DEFAULT = 40          # 40^3 FCC cells = 256 000 atoms


def make_system(ncell, seed=1, zmult=1):
    """FCC ncell x ncell x (ncell*zmult) box at rho 0.845, jitter 0.05a, Maxwell velocities at T=1.
    Returns positions, velocities (fp64) and the box edge L of the cubic unit (Lz = L * zmult)."""
    a = (4.0 / RHO) ** (1.0 / 3.0)
    basis = np.array([(0, 0, 0), (0.5, 0.5, 0), (0.5, 0, 0.5), (0, 0.5, 0.5)], dtype=np.float64)
    g = np.stack(np.meshgrid(np.arange(ncell), np.arange(ncell), np.arange(ncell * zmult), indexing="ij"), -1).reshape(-1, 1, 3)
    pos = ((g + basis[None]) * a).reshape(-1, 3)
    pos = pos + np.random.default_rng(seed).normal(0.0, 0.05 * a, pos.shape)
    vel = np.random.default_rng(seed + 1).standard_normal(pos.shape) * math.sqrt(TEMP / MASS)
    return pos, vel, ncell * a


def ncu_traffic(kernel_substr):
    """DRAM bytes per launch of the dominant kernel from the committed ncu summary of THIS round (profiles/r02_force_ncu.json,
    written by tools/ncu_summary.py from an `ncu --set full` capture of the same bench command); None if it does not list
    the kernel the engine actually ran."""
    p = os.path.join(ROOT, "profiles", "r02_force_ncu.json")
    try:
        for k in json.load(open(p))["kernels"]:
            if kernel_substr in k["name"]:
                return float(k["dram_bytes_read"]) + float(k["dram_bytes_write"]), \
                    "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per launch (%s, kernel %s)" % (
                        os.path.relpath(p, ROOT), k["name"])
    except Exception:
        pass
    return None, "no ncu capture of this kernel committed"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        """Start of the timed region: only rows that arrive from here on are used (nvidia-smi needs > 150 ms to come up on an
        8-GPU box, so it is started before the warm-up)."""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        first = getattr(self, "first", 0)
        t_end = time.time() + 3.0
        time.sleep(0.15)
        while len(self.rows) <= first and time.time() < t_end:      # at least one sample taken after the start of the timed region
            time.sleep(0.02)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[first:]:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for nm, val in zip(names, c[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def md_params(_lib, n, L, skin, K, zmult=1):
    p = _lib.MdParams()
    p.integrator = _lib.INT_NHC
    p.pot_kind = _lib.POT_LJ
    p.pot_params[0], p.pot_params[1] = 1.0, 1.0
    p.cutoff = RC
    for k in range(3):
        p.cell[k] = L
    p.cell[2] = float(np.float32(L * zmult))
    p.n_chains = CHAINS
    Q = np.array([qbath(n), *[qbath(n) / n] * (CHAINS - 1)]).astype(np.float32)
    for k in range(CHAINS):
        p.Q[k] = float(Q[k])
    p.T = TEMP
    p.ndof = 3 * n
    p.skin = skin
    p.rebuild_every = K
    p.traj_stride = 1
    return p


def tgrid(nsteps, dt=DT):
    return [float(np.float32(dt * i)) for i in range(nsteps + 1)]


def assemble(torch, frame, world):
    """multi-GPU: frames hold the owned atoms only (zero elsewhere) -> sum over ranks."""
    if world > 1:
        import torch.distributed as dist
        frame = frame.clone()
        dist.all_reduce(frame)
    return frame


def equilibrate(ctx, _lib, torch, n, L, mass, v, q, skin, log, zmult=1, world=1):
    """UNTIMED set-up: melt / thermalise the jittered FCC start with NVE epochs + velocity rescaling
    (through the same fused engine), because the reference's Nose-Hoover chain - bath masses
    Q/N (torchmd/md.py:191-193) - is stiff for 256k atoms and diverges (also in the CPU oracle) when
    started far from equilibrium.  Returns (v, q) at T ~= TEMP."""
    p = md_params(_lib, n, L, skin, 4, zmult)
    p.integrator = _lib.INT_NVE
    schedule = [(0.001, 50)] * 4 + [(0.0025, 50)] * 6 + [(DT, 50)] * 30
    for i, (dt, nst) in enumerate(schedule):
        p.traj_stride = nst
        tv, tq, _, _ = ctx.md_run(p, mass, v, q, [], tgrid(nst, dt))
        v, q = assemble(torch, tv[-1], world), assemble(torch, tq[-1], world)
        t_inst = float((mass[:, None] * v * v).sum() / (3 * n))
        v = v * math.sqrt(TEMP / t_inst)
        del tv, tq
        if i % 8 == 0 or i == len(schedule) - 1:
            log("  equilibration epoch %d: dt=%.4f T_inst=%.4f K=%d" % (i, dt, t_inst, ctx.stats()["maxrow_or_K"]))
    return v, q


def dist_parity(ctx, _lib, torch, dist, dev, rank, world):
    """UNTIMED multi-GPU parity statement inside the driver-run line: a small box (6912 atoms per rank) runs 40 NHC
    steps on the slab-decomposed engine of all ranks and, on rank 0, on the single-GPU engine; both sum the forces of a row
    in the same order, so they agree to the rounding of the kinetic-energy all-reduce.  Returns relative differences."""
    ncell = 12
    pos, vel, L = make_system(ncell, seed=7, zmult=world)
    n = pos.shape[0]
    L32 = float(np.float32(L))
    q0 = torch.tensor(pos, dtype=torch.float32, device=dev)
    v0 = torch.tensor(vel * 0.5, dtype=torch.float32, device=dev)
    mass = torch.full((n,), MASS, dtype=torch.float32, device=dev)
    p = md_params(_lib, n, L32, 0.4, 5, world)
    nsteps = 40
    p.traj_stride = 10
    t = tgrid(nsteps, 0.002)
    tv, tq, tpv, e = ctx.md_run(p, mass, v0, q0, [0.0] * CHAINS, t, want_energy=True)
    tv, tq = tv.clone(), tq.clone()
    dist.all_reduce(tv)
    dist.all_reduce(tq)
    out = None
    if rank == 0:
        sctx = _lib.Context(dev)
        sv, sq, spv, se = sctx.md_run(p, mass, v0, q0, [0.0] * CHAINS, t, want_energy=True)
        out = {"atoms": n, "steps": nsteps, "rebuilds": int(ctx.stats()["rebuilds"]),
               "dv": (tv - sv).abs().max().item() / sv.abs().max().item(),
               "dq": (tq - sq).abs().max().item() / L,
               "dpv": (tpv - spv).abs().max().item() / max(1e-9, spv.abs().max().item()),
               "dE": abs(e - se) / abs(se),
               "frame0_equal": bool(torch.equal(tq[0], q0) and torch.equal(tv[0], v0)),
               "what": "world-%d slab engine vs the single-GPU engine on rank 0, relative max differences of v, q (per box edge), "
                       "bath momenta, potential energy after %d NHC steps" % (world, nsteps)}
        del sctx
    return out


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the C oracle (port of the reference's per-evaluation all-pairs
# algorithm) on the host cores, on a bounded row sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_reference_steps_per_s(ncell, steps, warmup, budget_s, zmult=1):
    from oracle import oracle_c as C
    pos, vel, L = make_system(ncell, zmult=zmult)
    n = pos.shape[0]
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    C.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    cores = C.num_threads()
    cell3 = np.array([L, L, L * zmult], dtype=np.float32)
    x = pos.astype(np.float32)
    # calibrate the row sample so that (steps + warmup) evaluations fit the budget
    t0 = time.perf_counter()
    probe = min(n, 64 * cores)
    C.lj_forces(x, cell3, RC, rows=(0, probe))
    per_row = (time.perf_counter() - t0) / probe
    rows = int(max(cores, min(n, budget_s / max(1, steps + warmup) / per_row)))
    mass = np.full(n, MASS, np.float32)
    Q = np.array([qbath(n), *[qbath(n) / n] * (CHAINS - 1)]).astype(np.float32)
    dts = np.diff(np.array(tgrid(max(steps, warmup, 1)), dtype=np.float32))
    if warmup:
        C.nhc_md(vel, pos, np.zeros(CHAINS), mass, cell3, RC, 1.0, 1.0, Q, TEMP, 3 * n, dts[:warmup], rows=rows)
    t0 = time.perf_counter()
    C.nhc_md(vel, pos, np.zeros(CHAINS), mass, cell3, RC, 1.0, 1.0, Q, TEMP, 3 * n, dts[:steps], rows=rows)
    el = time.perf_counter() - t0
    # one evaluation of the full box costs n/rows times the sampled rows (the integrator part is O(N) and
    # already complete); steps/s of the FULL workload:
    integ = 0.0
    full_step = (el / steps) * (n / rows) if rows < n else el / steps
    sample = ("C oracle (all-pairs list rebuilt per evaluation, as the reference), %d of %d atom rows per step, "
              "time scaled x%.1f; reference torch path cannot allocate this box (70*N^2 B)" % (rows, n, n / rows))
    detail = {"extrapolated": rows < n, "rows_measured_per_step": int(rows), "rows_total": int(n), "steps_measured": int(steps),
              "seconds_measured": el, "scale_factor": float(n) / rows}
    return zmult / full_step, cores, sample, n, detail     # 256k-atom-box equivalents per second (see value_definition)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    v, cores, sample, n, detail = cpu_reference_steps_per_s(args.ncell, args.steps, args.warmup, budget_s=90.0, zmult=max(1, args.gpus))
    line = {
        "impl": "reference", "metric": "MD steps/sec", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LJ fluid %d atoms (FCC %d^3, rho 0.845), rc 2.5, NoseHooverChain Q=50*N/256 T=1 M=5, dt 0.005"
                               % (n, args.ncell), "atoms": n},
        "cpu_baseline": dict({"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}, **detail),
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "extrapolated": detail["extrapolated"],
    }
    emit(line)


def time_config(ctx, _lib, torch, dist, dev, world, ncell, steps, warmup, skin, log):
    """Device-resident steps/s of an ncell x ncell x (ncell * world) FCC LJ box on the same engine (used for the north_star's
    configs[3] size, 131 072 atoms per GPU, next to the headline configuration): set-up, warm-up, CUDA-event timing with
    barrier + synchronize on both sides, max over ranks."""
    pos, vel, L = make_system(ncell, zmult=world)
    n = pos.shape[0]
    L32 = float(np.float32(L))
    q0 = torch.tensor(pos, dtype=torch.float32, device=dev)
    v0 = torch.tensor(vel, dtype=torch.float32, device=dev)
    mass = torch.full((n,), MASS, dtype=torch.float32, device=dev)
    v0, q0 = equilibrate(ctx, _lib, torch, n, L32, mass, v0, q0, skin, lambda m: None, world, world)
    vmax = float(v0.norm(dim=1).max())
    p = md_params(_lib, n, L32, skin, int(max(1, min(64, math.floor(0.5 * skin / (1.1 * vmax * DT))))), world)

    def run(nsteps, vv, qq, pv):
        p.traj_stride = nsteps
        tv, tq, tpv, _ = ctx.md_run(p, mass, vv, qq, pv, tgrid(nsteps))
        return assemble(torch, tv[-1], world), assemble(torch, tq[-1], world), [float(x) for x in tpv[-1]]

    v1, q1, pv1 = run(warmup, v0, q0, [0.0] * CHAINS)
    p.rebuild_every = int(ctx.stats()["maxrow_or_K"])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    v2, q2, _ = run(steps, v1, q1, pv1)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"atoms": n, "atoms_per_gpu": n // world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
            "box_steps_per_s": steps / (ms / 1000.0), "atom_steps_per_s": n * steps / (ms / 1000.0),
            "rebuild_every": int(p.rebuild_every), "finite": bool(torch.isfinite(q2).all() and torch.isfinite(v2).all()),
            "trajectory": "first and last frame only"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ncell", type=int, default=NCELL_DEFAULT, help="FCC cells per axis (40 -> 256000 atoms per GPU)")
    ap.add_argument("--skin", type=float, default=0.45)
    ap.add_argument("--rebuild-every", type=int, default=0, help="0 = auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-dist-parity", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the extra 131 072-atoms-per-GPU measurement (configs[3] size)")
    ap.add_argument("--config", default="c2", choices=["c2", "c1", "c3", "c5"],
                    help="c2 (default) = BASELINE configs[1], the headline; c1 / c3 / c5 = configs[0] / [2] / [4] through the public API")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.config != "c2":
        # These configurations are a few hundred / thousand atoms: they do not shard ("replicas only") -
        # under torchrun every rank runs its own independent replica on its GPU and the values are summed (no collective on
        # the data path; the sum itself goes over gloo).
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        r = bench_configs.RUN[args.config](args)
        if world > 1:
            import torch
            import torch.distributed as dist
            dist.init_process_group("gloo")
            t = torch.tensor([r["value"], r["ms_per_step"]], dtype=torch.float64)
            tsum, tmax = t.clone(), t.clone()
            dist.all_reduce(tsum)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            r["config"]["replicas"] = world
            r["config"]["per_replica_steps_per_s"] = r["value"]
            r["value"], r["ms_per_step"] = float(tsum[0]), float(tmax[1])
            dist.destroy_process_group()
        if rank != 0:
            return
        line = {"metric": "MD steps/sec", "value": r["value"], "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": r["config"], "gpu_launches": r["gpu_launches"],
                "e2e": {"value": r["value"], "unit": "steps/s", "note": "this configuration is timed through Simulations.simulate "
                        "with the host state of System: H2D at the start of the epoch, D2H of the last frame at its end"}}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from mdgrad_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the MD hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert args.warmup >= 3, "W >= 3 warm-up steps required"

    # N = 1: BASELINE configs[1] (256 000 atoms).  N > 1: weak scaling - the box grows along z with the GPU count
    # (256 000 atoms per GPU, one spatial slab per rank, NCCL ghost-layer halo + 2-double KE all-reduce per step).
    pos, vel, L = make_system(args.ncell, zmult=world)
    n = pos.shape[0]
    L32 = float(np.float32(L))
    q0 = torch.tensor(pos, dtype=torch.float32, device=dev)
    v0 = torch.tensor(vel, dtype=torch.float32, device=dev)
    mass = torch.full((n,), MASS, dtype=torch.float32, device=dev)
    ctx = _lib.Context(dev)
    parity = None
    if world > 1:
        ctx.dist_init()
        if not args.no_dist_parity:
            parity = dist_parity(ctx, _lib, torch, dist, dev, rank, world)
            log("dist parity: %s" % parity)

    p = md_params(_lib, n, L32, args.skin, 4, world)
    log("set-up: equilibrating %d atoms (NVE + rescale, untimed)" % n)
    v0, q0 = equilibrate(ctx, _lib, torch, n, L32, mass, v0, q0, args.skin, log, world, world)
    K = args.rebuild_every
    if K <= 0:   # rebuild cadence from v_max; the engine halves it (and redoes the epoch) on a skin violation
        vmax = float(v0.norm(dim=1).max())
        K = int(max(1, min(64, math.floor(0.5 * args.skin / (1.1 * vmax * DT))))) if args.skin > 0 else 1
    p.rebuild_every = K
    # single GPU: every step captured (reference semantics).  multi GPU: first + last frame only (a 1000-frame
    # trajectory of the whole multi-million-atom box would not fit beside the state).
    full_traj = world == 1

    def run(nsteps, vv, qq, pv, out=None):
        p.traj_stride = 1 if full_traj else nsteps
        tv, tq, tpv, _ = ctx.md_run(p, mass, vv, qq, pv, tgrid(nsteps), out=out)
        return tv, tq, tpv

    log("system ready: n=%d L=%.3f K=%d skin=%.2f world=%d; warm-up %d steps" % (n, L, K, args.skin, world, args.warmup))
    sampler = ClockSampler(local_rank)
    sampler.start()
    tv, tq, tpv = run(args.warmup, v0, q0, [0.0] * CHAINS)
    log("warm-up done: %s" % ctx.stats())
    p.rebuild_every = int(ctx.stats()["maxrow_or_K"])
    v1, q1, pv1 = assemble(torch, tv[-1], world), assemble(torch, tq[-1], world), [float(x) for x in tpv[-1]]
    del tv, tq

    # ---- timed: device-resident K steps, CUDA events on the launch stream, barrier + sync both sides
    nfr = args.steps + 1 if full_traj else 2
    out = (torch.zeros((nfr, n, 3), dtype=torch.float32, device=dev), torch.zeros((nfr, n, 3), dtype=torch.float32, device=dev))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    tv, tq, tpv = run(args.steps, v1, q1, pv1, out=out)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    log("timed region done: %.3f ms for %d steps" % (ms, args.steps))
    clocks = sampler.stop()
    stats = ctx.stats()
    log("clocks %s stats %s" % (clocks, stats))
    launches = int(stats["launches"])
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    box_steps_per_s = args.steps / (ms / 1000.0)
    value = box_steps_per_s * world          # 256k-atom-box equivalents per second (= atom-steps/s / 256000)
    q_end, v_end = assemble(torch, tq[-1], world), assemble(torch, tv[-1], world)
    finite = bool(torch.isfinite(q_end).all() and torch.isfinite(v_end).all())
    del tv, tq
    out = None
    torch.cuda.empty_cache()

    res = {
        "metric": "MD steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LJ fluid, %d atoms per GPU (FCC %dx%dx%d, rho 0.845, jitter 0.05a), LennardJones(1,1) rc 2.5, "
                               "NoseHooverChain Q=50*N/256 T=1.0 M=5, dt 0.005; start equilibrated by untimed NVE+rescale set-up"
                               % (n // world, args.ncell, args.ncell, args.ncell * world),
                   "atoms": n, "atoms_per_gpu": n // world,
                   "parallelism": ("1 GPU" if world == 1 else
                                   "spatial slabs x%d along z in the global cell-sorted index space; per step: one peer-to-peer push "
                                   "kernel (ghost layers + kinetic energies over NVLink stores, CUDA-IPC mapped; MDG_DIST_P2P=0: NCCL "
                                   "send/recv + all-reduce), boundary layers on a side stream beside the interior rows; per rebuild: "
                                   "two-layer state exchange + layer totals over NCCL" % world),
                   "value_definition": "steps/s of the whole box x n_gpus (the box grows with the GPU count: 256000 atoms per GPU), "
                                       "i.e. atom-steps/s / 256000; box_steps_per_s is the raw rate of the %d-atom box" % n,
                   "box_steps_per_s": box_steps_per_s,
                   "skin": args.skin, "rebuild_every": int(p.rebuild_every), "rebuilds": int(stats["rebuilds"]),
                   "l2": "no explicit flush: per-step working set (neighbor rows ~%.0f MB + state %.0f MB per GPU) is streamed "
                         "from HBM and exceeds the 126 MB L2" % (n / world * 96 * 4 / 1e6, n * 16 * 6 / 1e6),
                   "trajectory": "every step captured (reference semantics, stride 1)" if full_traj else "first and last frame only",
                   "finite": finite},
        "gpu_launches": launches, "clocks": clocks,
        "tau_per_day": box_steps_per_s * DT * 86400.0,
    }
    if not args.no_c4:
        # BASELINE configs[3]: 1 048 576-atom LJ box over 8 GPUs = 131 072 atoms per GPU (FCC 32 x 32 x 32 per GPU); reported at
        # every N so that the weak scaling of THAT size can be read off the same lines (efficiency = box_steps_per_s(N) / (N=1))
        res["c4"] = time_config(ctx, _lib, torch, dist, dev, world, 32, min(args.steps, 400), max(3, min(args.warmup, 100)), args.skin, log)
        res["c4"]["what"] = "configs[3] size: 131 072 atoms per GPU (FCC 32^3 per GPU, box grows along z with the GPU count), same engine / skin / thermostat"
        log("c4: %s" % res["c4"])
    if world > 1:
        res["dist_parity"] = parity
        if rank == 0:
            emit(res)
        ctx.dist_finalize()
        dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (pair force): algorithmic bytes 32 N + 8 P over measured duration
    ctx.set_profile(True)
    tvp, tqp, _ = run(min(args.steps, 200), v1, q1, pv1)
    prof = ctx.get_profile()
    ctx.set_profile(False)
    del tvp, tqp
    log("force-kernel profile pass done: %s" % prof)
    cctx = _lib.Context(dev)
    P_rc = int(cctx.nbr_list(q_end, [L32] * 3, RC)[0].shape[0])      # pairs inside the cutoff (reference list size)
    P_list = int(cctx.nbr_list(q_end, [L32] * 3, RC + args.skin)[0].shape[0]) if args.skin > 0 else P_rc
    del cctx
    force_ms = prof["force_ms"] / max(1, prof["force_launches"])
    alg_bytes = 32.0 * n + 8.0 * P_rc
    peak, peak_src = peaks()
    achieved = alg_bytes / (force_ms * 1e-3) / 1e9
    log("pair counts done: P_rc=%d P_list=%d force %.4f ms" % (P_rc, P_list, force_ms))
    tiles = int(stats["path"]) == 2
    kname = ("k_force_tiles<LJ> (persistent CTA per block of cells, TMA-staged stencil positions in shared memory, 16-bit local skin rows)"
             if tiles else "k_force_rows<LJ,RETEST,4 lanes/row> (pair force over the fixed-capacity skin rows)")
    traffic, traffic_src = ncu_traffic("k_force_tiles" if tiles else "k_force_rows")
    res["roofline"] = {"bound": "hbm", "kernel": kname, "achieved": achieved,
                       "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                       "traffic_source": traffic_src,
                       "peak_source": peak_src, "algorithmic_bytes": alg_bytes, "pairs_in_cutoff": P_rc,
                       "pairs_in_skin_list": P_list, "streamed_bytes_with_skin": 32.0 * n + 8.0 * P_list,
                       "kernel_ms": force_ms, "kernel_launches_timed": prof["force_launches"],
                       "limiter": "not HBM: L1TEX / LSU wavefronts 83 % and instruction issue 63 % in the ncu capture (one 16-byte gather per "
                                  "list entry, >= 23 instructions per entry with the reference's exact membership arithmetic) - "
                                  "profiles/r02_force_kernel.md",
                       "share_of_step": force_ms / (ms / args.steps)}

    # ---- e2e through the public API with HOST state (numpy in System, H2D per epoch, D2H last frame per epoch)
    if not args.no_e2e:
        from torchmd.system import System
        from torchmd.interface import PairPotentials
        from torchmd.potentials import LennardJones
        from torchmd.md import NoseHooverChain, Simulations
        from mdgrad_b200._ase_compat import Atoms
        atoms = Atoms(numbers=[1] * n, positions=q_end.cpu().numpy().astype(np.float64), cell=[L] * 3, pbc=True)
        system = System(atoms, device=local_rank)
        system.set_velocities(v_end.cpu().numpy().astype(np.float64))
        pair = PairPotentials(system, LennardJones(1.0, 1.0), cutoff=RC)
        integ = NoseHooverChain(pair, system, T=TEMP, num_chains=CHAINS, Q=qbath(n), adjoint=True)
        integ.engine_skin = args.skin
        sim = Simulations(system, integ, wrap=True, method="NH_verlet")
        per_epoch = 100
        freq = per_epoch + 1
        n_epochs = max(1, args.steps // per_epoch)
        log("e2e: API objects built, warm-up epoch")
        sim.simulate(steps=freq, frequency=freq, dt=DT)                  # warm-up epoch (100 steps)
        torch.cuda.synchronize()
        log("e2e: timed epochs")
        prof = None
        if os.environ.get("BENCH_PROFILE_E2E"):
            import cProfile
            prof = cProfile.Profile()
            prof.enable()
        t0 = time.perf_counter()
        sim.simulate(steps=freq * n_epochs, frequency=freq, dt=DT)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        if prof is not None:
            import pstats
            prof.disable()
            pstats.Stats(prof, stream=sys.stderr).sort_stats("cumulative").print_stats(22)
            pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(14)
        e2e_steps = n_epochs * per_epoch
        state_bytes = n * 3 * 4 * 2 + CHAINS * 4
        res["e2e"] = {"value": e2e_steps / el, "unit": "steps/s",
                      "h2d_bytes_per_step": state_bytes / e2e_steps, "d2h_bytes_per_step": state_bytes / per_epoch,
                      "api": "Simulations.simulate(steps=%d, frequency=%d, dt=0.005): %d epochs x %d steps; host numpy state "
                             "(System) -> H2D at the start of the call, every epoch's last frame D2H into the host log and "
                             "System update before the call returns; between epochs the state is handed over on the device "
                             "(fp64 wrap there, bit-identical to the reference's host wrap)"
                             % (freq * n_epochs, freq, n_epochs, per_epoch)}

    # ---- CPU baseline beside it (bounded sample, rank 0, N=1 only)
    if not args.no_cpu_baseline:
        log("cpu baseline (C oracle, bounded sample)")
        v, cores, sample, _, detail = cpu_reference_steps_per_s(args.ncell, steps=3, warmup=1, budget_s=20.0)
        log("cpu baseline done")
        res["cpu_baseline"] = dict({"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}, **detail)
    emit(res)


if __name__ == "__main__":
    main()
