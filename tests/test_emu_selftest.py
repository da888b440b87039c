"""The CPU emulation harness (tests/cuemu) is only worth something if it FAILS on the mistakes a GPU would punish.
Each case of tests/cuemu/selftest.cu commits one such mistake (host dereference of a device pointer, reading a pinned
buffer before the stream synchronised, a missing cudaStreamWaitEvent, zero-sized grids, > 48 KB dynamic shared memory
without the opt-in, use after free, out-of-bounds writes, thread-order dependence); the strict emulator must expose it.
TEST INFRASTRUCTURE - nothing here touches the product path."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "cuemu"))


@pytest.fixture(scope="module")
def exe():
    import build_emu
    return build_emu.build_selftest()


def _run(exe, case, **env):
    e = dict(os.environ)
    e.update(env)
    return subprocess.run([exe, case], capture_output=True, text=True, env=e, timeout=120)


@pytest.mark.parametrize("case", ["ok", "pinned_before_sync", "pinned_reuse", "missing_wait", "event_wait_ok", "zero_grid",
                                  "grid_y", "smem_optin", "legacy_stream", "graph_ok", "graph_legacy",
                                  "graph_sync_inside"])
def test_strict_emulator_exposes(exe, case):
    r = _run(exe, case)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip().splitlines()[-1] == "caught", (case, r.stdout, r.stderr)


@pytest.mark.parametrize("case,needle", [("host_deref", "HOST code"), ("use_after_free", "FREED device memory"),
                                         ("oob", "out-of-bounds device memory")])
def test_strict_emulator_faults(exe, case, needle):
    r = _run(exe, case)
    assert r.returncode != 0 and "missed" not in r.stdout, (case, r.stdout)
    assert "cuemu:" in r.stderr and needle in r.stderr, r.stderr


def test_relaxed_mode_hides_them(exe):
    """the old synchronous mode (CUEMU_STRICT=0) does hide these - which is why strict is the default"""
    for case in ("pinned_before_sync", "missing_wait", "host_deref"):
        r = _run(exe, case, CUEMU_STRICT="0")
        assert r.returncode == 0 and r.stdout.strip().splitlines()[-1] == "missed", (case, r.stdout, r.stderr)


def test_thread_order_modes_expose_order_dependence(exe):
    vals = {o: _run(exe, "racy", CUEMU_ORDER=o).stdout.strip() for o in ("fwd", "rev")}
    assert vals["fwd"] != vals["rev"], vals
