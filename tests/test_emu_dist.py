"""The slab-decomposed MULTI-RANK engine (halo send/recv in place, KE / flag / layer-count all-reduces on a side stream
overlapping the interior forces, local rebuild - engine.cu, nbr.cu, dist.cu) at world size 2, 3 and 4 in the GPU-less container:
every rank is a process running the CPU-emulated library, NCCL is tests/cuemu/fake_nccl.cpp (shared memory, stream-ordered
through the emulator's queues).  Same assertions as tests/dist_check.py, which does this on real GPUs with real NCCL.
TEST INFRASTRUCTURE - functional coverage of the distributed host logic and kernels, nothing about NVLink."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(world, tmp, nsteps, K=4, vscale=1.0):
    idf = os.path.join(tmp, "id_%d_%d" % (world, K))
    outs = [os.path.join(tmp, "w%d_r%d_k%d.npz" % (world, r, K)) for r in range(world)]
    env = dict(os.environ, FAKE_NCCL_TIMEOUT="150")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "emu_dist_worker.py"), str(r), str(world), idf, outs[r], str(nsteps),
                               str(K), str(vscale)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env) for r in range(world)]
    logs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=900)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(out)
    for p, log in zip(procs, logs):
        assert p.returncode == 0, log[-3000:]
    return [np.load(o) for o in outs]


@pytest.fixture(scope="module")
def single_device_run(tmp_path_factory):
    return _run(1, str(tmp_path_factory.mktemp("single")), 12)[0]


@pytest.mark.parametrize("world", [2, 3, 4])
def test_emu_slab_engine_matches_single_device(world, tmp_path, single_device_run):
    nsteps = 12
    single = single_device_run
    ranks = _run(world, str(tmp_path), nsteps)
    # frames hold the owned atoms only and are ZERO elsewhere (the caller sums over ranks): no NaN poison may survive
    for r in ranks:
        assert np.isfinite(r["tv"]).all() and np.isfinite(r["tq"]).all()
    tv = sum(r["tv"] for r in ranks)
    tq = sum(r["tq"] for r in ranks)
    # each atom is reported by exactly one rank in every frame
    owners = sum((r["tq"] != 0).any(-1).astype(np.int32) for r in ranks)
    assert (owners == 1).all()
    assert np.array_equal(tq[0], single["q0"]) and np.array_equal(tv[0], single["v0"])
    L = 15.0
    dv = np.abs(tv - single["tv"]).max() / np.abs(single["tv"]).max()
    dq = np.abs(tq - single["tq"]).max() / L
    dp = np.abs(ranks[0]["tpv"] - single["tpv"]).max() / max(1e-9, np.abs(single["tpv"]).max())
    de = abs(float(ranks[0]["e"]) - float(single["e"])) / abs(float(single["e"]))
    assert dv < 2e-4 and dq < 2e-6 and dp < 1e-3 and de < 1e-5, (dv, dq, dp, de)      # tests/dist_check.py's bars
    # bath trajectory and energy are replicated: identical bits on every rank
    for r in ranks[1:]:
        assert np.array_equal(r["tpv"], ranks[0]["tpv"]) and float(r["e"]) == float(ranks[0]["e"])
    assert int(ranks[0]["rebuilds"]) >= 2      # the run crossed at least one distributed (local) rebuild


def test_emu_slab_engine_skin_retry_is_collective(tmp_path):
    """A rebuild interval that violates the skin criterion: the violation flag is max-all-reduced, so ALL ranks halve K and
    repeat the epoch together (a rank deciding alone would deadlock the halo exchange - the fake NCCL would time out)."""
    nsteps = 16
    single = _run(1, str(tmp_path), nsteps, K=16, vscale=3.0)[0]
    ranks = _run(2, str(tmp_path), nsteps, K=16, vscale=3.0)
    assert int(single["K"]) < 16 and all(int(r["K"]) == int(ranks[0]["K"]) for r in ranks) and int(ranks[0]["K"]) < 16
    tq = sum(r["tq"] for r in ranks)
    assert np.abs(tq - single["tq"]).max() / 15.0 < 5e-6
    for r in ranks[1:]:
        assert np.array_equal(r["tpv"], ranks[0]["tpv"])
