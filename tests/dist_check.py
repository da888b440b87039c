"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py
Every rank runs the slab-decomposed engine on the same box; rank 0 also runs the single-GPU engine.
Forces are summed in the same order in both -> trajectories agree to the KE all-reduce rounding."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mdgrad_b200 import _lib  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ncell, zmult = 14, 2
    pos, vel, L = bench.make_system(ncell, zmult=zmult)
    pos = np.asarray(pos)
    n = pos.shape[0]
    L32 = float(np.float32(L))
    q0 = torch.tensor(pos, dtype=torch.float32, device=dev)
    v0 = torch.tensor(vel * 0.5, dtype=torch.float32, device=dev)
    mass = torch.full((n,), bench.MASS, dtype=torch.float32, device=dev)
    p = bench.md_params(_lib, n, L32, 0.4, 5, zmult)
    p.integrator = _lib.INT_NHC
    nsteps = 40
    p.traj_stride = 10
    t = bench.tgrid(nsteps, 0.002)
    dctx = _lib.Context(dev)
    dctx.dist_init()
    tv, tq, tpv, e = dctx.md_run(p, mass, v0, q0, [0.0] * bench.CHAINS, t, want_energy=True)
    tv, tq = tv.clone(), tq.clone()
    dist.all_reduce(tv)
    dist.all_reduce(tq)
    st = dctx.stats()
    ok = True
    if rank == 0:
        sctx = _lib.Context(dev)
        sv, sq, spv, se = sctx.md_run(p, mass, v0, q0, [0.0] * bench.CHAINS, t, want_energy=True)
        dv = (tv - sv).abs().max().item() / sv.abs().max().item()
        dq = (tq - sq).abs().max().item() / L
        dp = (tpv - spv).abs().max().item() / max(1e-9, spv.abs().max().item())
        de = abs(e - se) / abs(se)
        print("dist_check world=%d n=%d rebuilds=%d K=%d : dv=%.2e dq=%.2e dpv=%.2e dE=%.2e frames0_equal=%s"
              % (world, n, st["rebuilds"], st["maxrow_or_K"], dv, dq, dp, de,
                 bool(torch.equal(tq[0], q0) and torch.equal(tv[0], v0))))
        ok = dv < 2e-4 and dq < 2e-6 and dp < 1e-3 and de < 1e-5 and torch.equal(tq[0], q0)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dctx.dist_finalize()
    dist.destroy_process_group()
    if not flag.item():
        raise SystemExit("dist_check FAILED")
    if rank == 0:
        print("dist_check OK")


if __name__ == "__main__":
    main()
