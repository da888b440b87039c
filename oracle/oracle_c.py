"""ctypes wrapper of the C oracle (oracle/oracle_c.c -> oracle/_ref/liboracle_c.so).
TEST INFRASTRUCTURE / CPU BASELINE ONLY - see the header of oracle_c.c."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "oracle_c.c")
LIB = os.path.join(_HERE, "_ref", "liboracle_c.so")
_lib = None


def build():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed for oracle_c.c:\n" + r.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(LIB)
        f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        lib.orc_num_threads.restype = ctypes.c_int
        lib.orc_set_threads.restype = None
        lib.orc_set_threads.argtypes = [ctypes.c_int]
        lib.orc_nbr_list.restype = ctypes.c_int64
        lib.orc_nbr_list.argtypes = [f32p, ctypes.c_int, f32p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_int64]
        lib.orc_nbr_rows_upper.restype = None
        lib.orc_nbr_rows_upper.argtypes = [f32p, ctypes.c_int, f32p, ctypes.c_double, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.orc_pair_rows.restype = ctypes.c_double
        lib.orc_pair_rows.argtypes = [f32p, ctypes.c_int, f32p, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_int, ctypes.c_int, f32p]
        lib.orc_nhc_md.restype = ctypes.c_double
        lib.orc_nhc_md.argtypes = [f32p, f32p, f32p, f32p, ctypes.c_int, f32p, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_double, ctypes.c_int, f32p, ctypes.c_double, ctypes.c_int, f32p,
                                   ctypes.c_int, ctypes.c_int]
        _lib = lib
    return _lib


def num_threads():
    return load().orc_num_threads()


def set_threads(n):
    """use n OpenMP threads from now on (bench.py: torchrun exports OMP_NUM_THREADS=1 to every rank)"""
    load().orc_set_threads(int(n))


def nbr_list(xyz, cell3, cutoff, get_dis=True):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    cell3 = np.ascontiguousarray(cell3, dtype=np.float32)
    n = xyz.shape[0]
    lib = load()
    P = lib.orc_nbr_list(xyz, n, cell3, float(cutoff), None, None, None, 0)
    nbr = np.empty((P, 2), dtype=np.int64)
    off = np.empty((P, 3), dtype=np.float32)
    dis = np.empty((P,), dtype=np.float32)
    lib.orc_nbr_list(xyz, n, cell3, float(cutoff), nbr.ctypes.data, off.ctypes.data, dis.ctypes.data, P)
    return nbr, off, dis


def nbr_rows_upper(xyz, cell3, cutoff, sel, cap=128):
    """Reference list rows (i, j > i) of the selected atoms: returns (cnt[nsel], j[nsel, cap], off[nsel, cap, 3]);
    only the first cnt[r] slots of row r are defined."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    cell3 = np.ascontiguousarray(cell3, dtype=np.float32)
    sel = np.ascontiguousarray(sel, dtype=np.int64)
    lib = load()
    while True:
        j = np.zeros((len(sel), cap), dtype=np.int64)
        off = np.zeros((len(sel), cap, 3), dtype=np.float32)
        cnt = np.zeros((len(sel),), dtype=np.int32)
        lib.orc_nbr_rows_upper(xyz, xyz.shape[0], cell3, float(cutoff), sel.ctypes.data, len(sel), j.ctypes.data,
                               off.ctypes.data, cnt.ctypes.data, cap)
        if len(sel) == 0 or int(cnt.max()) <= cap:
            return cnt, j, off
        cap = int(cnt.max())


def lj_forces(xyz, cell3, cutoff, sigma=1.0, eps=1.0, rows=None):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    cell3 = np.ascontiguousarray(cell3, dtype=np.float32)
    n = xyz.shape[0]
    i0, i1 = (0, n) if rows is None else rows
    f = np.empty((i1 - i0, 3), dtype=np.float32)
    e = load().orc_pair_rows(xyz, n, cell3, float(cutoff), float(sigma), float(eps), i0, i1, f)
    return e, f


def nhc_md(v, q, pv, mass, cell3, cutoff, sigma, eps, Q, T, ndof, dts, rows=0):
    """In-place NH-Verlet steps on copies; returns (v, q, pv, last_energy)."""
    v = np.ascontiguousarray(v, dtype=np.float32).copy()
    q = np.ascontiguousarray(q, dtype=np.float32).copy()
    pv = np.ascontiguousarray(pv, dtype=np.float32).copy()
    mass = np.ascontiguousarray(mass, dtype=np.float32)
    cell3 = np.ascontiguousarray(cell3, dtype=np.float32)
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    dts = np.ascontiguousarray(dts, dtype=np.float32)
    e = load().orc_nhc_md(v, q, pv, mass, q.shape[0], cell3, float(cutoff), float(sigma), float(eps), len(Q), Q,
                          float(T), int(ndof), dts, len(dts), int(rows))
    return v, q, pv, e
