"""Builds libmdgrad_b200.so (sm_100a) and the C oracle in-tree with nvcc / gcc.

`python -m mdgrad_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libmdgrad_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-Xptxas", "-v"]
# per-file extra flags
EXTRA = {"engine.cu": ["-fmad=false"]}
# Build variants for A/B measurements: the same sources with extra -D feature macros, linked to
# libmdgrad_b200_<name>.so and selected at import time with MDG_LIB_VARIANT=<name> (tools/ab_variants.sh runs the
# bench on each and the GPU suite on the fastest).  The default library is "" - add entries while experimenting.
VARIANTS = {
    # int8 (VABSDIFF4 + IDP.4A) screening in the skin-list builder - build_fast.cuh; to be A/B-measured on a GPU (round 2)
    "i8": ["-DMDG_BUILD_INT8_SCREEN=1"],
    # register budget of the force kernel (default 8 CTAs/SM = 32 registers) and CTA shape of the list builder
    "mb6": ["-DMDG_FORCE_MINBLOCKS=6"],
    "mb4": ["-DMDG_FORCE_MINBLOCKS=4"],
    "u2mb6": ["-DMDG_FORCE_UNROLL=2", "-DMDG_FORCE_MINBLOCKS=6"],
    "u2mb4": ["-DMDG_FORCE_UNROLL=2", "-DMDG_FORCE_MINBLOCKS=4"],
    # software-pipelined row stream (next index block requested one iteration ahead), at three register budgets
    "pf": ["-DMDG_FORCE_PREFETCH=1"],
    "pfmb6": ["-DMDG_FORCE_PREFETCH=1", "-DMDG_FORCE_MINBLOCKS=6"],
    "pfmb4": ["-DMDG_FORCE_PREFETCH=1", "-DMDG_FORCE_MINBLOCKS=4"],
    # (8 warps per CTA would need 77 KB of STATIC shared memory - over the 48 KB limit; only the 2-warp shape is buildable)
    "fbw2": ["-DFB_WARPS=2"],
    # list builder phase 1 with two candidates per lane (half the LDS.128 per distance test)
    "fbp0": ["-DFB_PAIR=0"],
    "fbp4": ["-DFB_TRIP=4"],
    "fbp64": ["-DFB_MINBLOCKS=8"],
    "fbp4w2": ["-DFB_TRIP=4", "-DFB_WARPS=2"],
    # list builder with the static cell assignment (cell = f(blockIdx, warp)) instead of the global work counter
    "fbst": ["-DFB_DYNAMIC=0"],
    "lean": ["-DMDG_BUILD_LEAN=1"],
    # SchNet fused filter generator at 3 CTAs per SM (register cap 85 instead of the 100 it takes by itself)
    "sne3": ["-DSN_EDGE_MINBLOCKS=3"],
    "i8lean": ["-DMDG_BUILD_INT8_SCREEN=1", "-DMDG_BUILD_LEAN=1"],
}


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(path, flags):
    h = hashlib.sha1()
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    for p in [path] + headers + [os.path.join(ROOT, "include", "mdgrad_b200.h")]:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def _compile(src, verbose, variant=""):
    flags = ARCH + COMMON + EXTRA.get(src, []) + VARIANTS.get(variant, [])
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src[:-3] + (("." + variant) if variant else "") + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp(path, flags)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, False, ""
    cmd = [NVCC] + flags + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    with open(obj + ".ptxas.txt", "w") as f:
        f.write(r.stderr)
    return obj, True, r.stderr


def lib_path(variant=""):
    return os.path.join(HERE, "libmdgrad_b200%s.so" % (("_" + variant) if variant else ""))


def build(verbose=False, force=False, variant=""):
    if variant and variant not in VARIANTS:
        raise ValueError("unknown build variant %r (have %s)" % (variant, sorted(VARIANTS)))
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = _sources()
    LIB = lib_path(variant)
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose, variant), srcs))
    objs = [r[0] for r in results]
    rebuilt = any(r[1] for r in results)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        for r in results:
            if r[2]:
                print(r[2])
    return LIB


def build_oracle():
    """Compile oracle/oracle_c.c -> oracle/_ref/liboracle.so (test infrastructure)."""
    src = os.path.join(ROOT, "oracle", "oracle_c.c")
    if not os.path.exists(src):
        return None
    outdir = os.path.join(ROOT, "oracle", "_ref")
    os.makedirs(outdir, exist_ok=True)
    out = os.path.join(outdir, "liboracle_c.so")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", src, "-o", out, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed for oracle_c.c:\n%s" % r.stderr)
    return out


if __name__ == "__main__":
    lib = build(verbose="-v" in sys.argv, force="--force" in sys.argv)
    print("built", lib)
    for a in sys.argv[1:]:
        if a.startswith("--variant="):
            print("built", build(verbose="-v" in sys.argv, variant=a.split("=", 1)[1]))
    o = build_oracle()
    if o:
        print("built", o)
