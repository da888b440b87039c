"""One emulated RANK of the slab-decomposed engine (TEST INFRASTRUCTURE, spawned by tests/test_emu_dist.py):
    python tests/emu_dist_worker.py <rank> <world> <id_file> <out.npz> [nsteps [rebuild_every [velocity_scale]]]
Loads the CPU-emulated library (tests/cuemu), joins the fake NCCL communicator (tests/cuemu/fake_nccl.cpp, shared memory
between the rank processes) and runs mdg_md_run on the common box; world == 1 runs the single-device engine."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "cuemu"))
import build_emu  # noqa: E402
import emu_lib  # noqa: E402
from mdgrad_b200 import _lib  # noqa: E402

RHO, RC, SKIN, CHAINS, MASS = 0.845, 2.5, 0.4, 5, 1.008


def system(nx=9, nz=14, seed=1):
    a = (4.0 / RHO) ** (1.0 / 3.0)
    basis = np.array([(0, 0, 0), (0.5, 0.5, 0), (0.5, 0, 0.5), (0, 0.5, 0.5)], dtype=np.float64)
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(nx), np.arange(nz), indexing="ij"), -1).reshape(-1, 1, 3)
    pos = ((g + basis[None]) * a).reshape(-1, 3)
    pos = pos + np.random.default_rng(seed).normal(0.0, 0.05 * a, pos.shape)
    vel = np.random.default_rng(seed + 1).standard_normal(pos.shape) * np.sqrt(1.0 / MASS) * 0.5
    return pos.astype(np.float32), vel.astype(np.float32), (nx * a, nx * a, nz * a)


def main():
    rank, world, id_file, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    nsteps = int(sys.argv[5]) if len(sys.argv) > 5 else 12
    K0 = int(sys.argv[6]) if len(sys.argv) > 6 else 4            # requested rebuild interval
    vscale = float(sys.argv[7]) if len(sys.argv) > 7 else 1.0
    lib = emu_lib.load()
    pos, vel, L = system()
    vel = (vel * vscale).astype(np.float32)
    n = pos.shape[0]
    p = _lib.MdParams()
    p.integrator = _lib.INT_NHC
    p.pot_kind = _lib.POT_LJ
    p.pot_params[0], p.pot_params[1] = 1.0, 1.0
    p.cutoff = RC
    for k in range(3):
        p.cell[k] = float(np.float32(L[k]))
    p.n_chains = CHAINS
    q_b = 50.0 * n / 256.0
    for k in range(CHAINS):
        p.Q[k] = float(np.float32(q_b if k == 0 else q_b / n))
    p.T = 1.0
    p.ndof = 3 * n
    p.skin = SKIN
    p.rebuild_every = K0
    p.traj_stride = 3
    t = [float(np.float32(0.002 * i)) for i in range(nsteps + 1)]
    ctx = emu_lib.EmuContext()
    if world > 1:
        fake = build_emu.build_fake_nccl().encode()
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            assert lib.mdg_dist_unique_id(fake, buf) == 0, lib.mdg_last_error()
            with open(id_file + ".tmp", "wb") as f:
                f.write(buf.raw)
            os.rename(id_file + ".tmp", id_file)
        t0 = time.time()
        while not os.path.exists(id_file):
            time.sleep(0.05)
            assert time.time() - t0 < 120, "no unique id from rank 0"
        ident = open(id_file, "rb").read()
        assert lib.mdg_dist_init(ctx._h, fake, ident, rank, world) == 0, lib.mdg_last_error()
    n_frames = nsteps // p.traj_stride + 1
    tv = torch.full((n_frames, n, 3), float("nan"))          # poisoned: the engine owns every element it reports
    tq = torch.full((n_frames, n, 3), float("nan"))
    tv, tq, tpv, e = ctx.md_run(p, torch.full((n,), MASS), torch.from_numpy(vel), torch.from_numpy(pos), [0.0] * CHAINS, t,
                                want_energy=True, out=(tv, tq))
    st = ctx.stats()
    np.savez(out, tv=tv.numpy(), tq=tq.numpy(), tpv=tpv.numpy(), e=np.array(e), rebuilds=st["rebuilds"], K=st["maxrow_or_K"],
             q0=pos, v0=vel)
    if world > 1:
        assert lib.mdg_dist_finalize(ctx._h) == 0
    print("rank %d/%d done: rebuilds=%d K=%d E=%.6f" % (rank, world, st["rebuilds"], st["maxrow_or_K"], e))


if __name__ == "__main__":
    main()
