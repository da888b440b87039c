"""Round-2 bring-up of the tcgen05 dense layers (mdgrad_b200/csrc/schnet_tc.cuh, MDG_SCHNET_TC=1).

Runs the native SchNet energy+force program on the reference fixtures and on a model with the configs[4] layer widths,
once per process mode, and writes energies / forces to an .npz:

    python tools/tc_check.py simt out_simt.npz                      # default SIMT dense layers
    MDG_SCHNET_TC=1 timeout 120 python tools/tc_check.py tc out_tc.npz   # tensor-core dense layers (ALWAYS under a timeout:
                                                                          never executed on hardware when it was written)
    python tools/tc_check.py compare out_simt.npz out_tc.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(out):
    from mdgrad_b200 import _lib
    from oracle import oracle_torch as O
    from test_schnet import _fixture
    from test_emu_schnet import _rand_sd
    ctx = _lib.Context(torch.device("cuda", 0))
    res = {}
    cases = []
    for tag in ("water", "si"):
        g, params, sd = _fixture(tag)
        xyz = torch.Tensor(g["positions"])
        nbr, off = O.neighbor_list(xyz, params["cutoff"], torch.Tensor(g["cell"]))
        cases.append((tag, sd, torch.tensor(g["numbers"], dtype=torch.long), xyz, nbr, off))
    rng = np.random.default_rng(0)
    n, box, rc = 700, 16.0, 3.4
    xyz = torch.tensor(rng.uniform(0, box, (n, 3)), dtype=torch.float32)
    nbr, off = O.neighbor_list(xyz, rc, torch.tensor([box] * 3))
    cases.append(("wide", _rand_sd(512, 256, 33, 3, 256, rc, seed=5), torch.tensor(rng.integers(1, 9, n), dtype=torch.long), xyz, nbr, off))
    for tag, sd, z, xyz, nbr, off in cases:
        model = _lib.schnet_model_struct(sd, "cuda")
        e, f = ctx.schnet_energy_force(model, z.cuda(), xyz.cuda(), nbr.cuda(), off.cuda())
        torch.cuda.synchronize()
        res["e_" + tag] = np.array(e.item())
        res["f_" + tag] = f.cpu().numpy()
    np.savez(out, **res)
    print("wrote", out, {k: float(v) for k, v in res.items() if k.startswith("e_")})


def compare(a, b):
    A, B = np.load(a), np.load(b)
    ok = True
    for k in A.files:
        if k.startswith("e_"):
            err = abs(float(A[k]) - float(B[k])) / max(1.0, abs(float(A[k])))
        else:
            err = np.abs(A[k] - B[k]).max() / np.abs(A[k]).max()
        print("%-10s rel err %.3e" % (k, err))
        ok = ok and err <= 1e-5
    print("TC == SIMT within 1e-5:", ok)
    return 0 if ok else 1


if __name__ == "__main__":
    if sys.argv[1] == "compare":
        sys.exit(compare(sys.argv[2], sys.argv[3]))
    run(sys.argv[2])
