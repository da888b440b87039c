"""Builds the CPU-emulated copy of libmdgrad_b200.so (TEST INFRASTRUCTURE ONLY - see cuda_runtime.h).

The product's .cu sources are used verbatim except for one mechanical rewrite: every
`kernel<targs><<<grid, block, smem, stream>>>(args)` launch becomes a call of `cuemu::launch`, and
`extern __shared__ T name[];` becomes a pointer to the emulated dynamic shared memory.  The result is compiled
with g++ against tests/cuemu/cuda_runtime.h into tests/cuemu/_build/libmdgrad_b200_emu.so.

    python tests/cuemu/build_emu.py
"""
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "mdgrad_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libmdgrad_b200_emu.so")
CXX = os.environ.get("CUEMU_CXX", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")
# CUEMU_SANITIZE=1: AddressSanitizer + UBSan build (out-of-bounds / misaligned vector accesses of the kernels); run the
# tests with LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0
SANITIZE = os.environ.get("CUEMU_SANITIZE", "") not in ("", "0")
# CUEMU_DEFINES="-DMDG_BUILD_INT8_SCREEN=1": emulate an experimental build variant (mdgrad_b200/build.py VARIANTS)
DEFINES = os.environ.get("CUEMU_DEFINES", "").split()
if SANITIZE or DEFINES:
    OUT = os.path.join(HERE, "_build_san" if SANITIZE else "_build_var")
    LIB = os.path.join(OUT, "libmdgrad_b200_emu.so")
FLAGS = (["-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"] if SANITIZE else ["-O2"]) + [
    "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-Wno-unknown-pragmas",
    "-Wno-unused-function", "-DMDG_EMU=1"] + DEFINES + ["-I", HERE, "-I", CSRC]


def _match_back_template(src, i):
    """src[i] == '>' : index of the matching '<' scanning backwards"""
    depth = 0
    while i >= 0:
        ch = src[i]
        if ch == '>':
            depth += 1
        elif ch == '<':
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template brackets before <<<")


def _match_paren(src, i):
    """src[i] == '(' : index of the matching ')'"""
    depth = 0
    n = len(src)
    while i < n:
        ch = src[i]
        if ch == '(':
            depth += 1
        elif ch == ')':
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced parentheses after >>>")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        if ch in ")]}":
            depth -= 1
        if ch == ',' and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def rewrite_launches(src):
    out = []
    pos = 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            break
        # kernel expression: identifier [+ template args] right before <<<
        j = k - 1
        while src[j].isspace():
            j -= 1
        if src[j] == '>':
            j = _match_back_template(src, j) - 1
            while src[j].isspace():
                j -= 1
        e = j
        while j >= 0 and (src[j].isalnum() or src[j] == '_' or src[j] == ':'):
            j -= 1
        start = j + 1
        assert start <= e, "no kernel name before <<< at offset %d" % k
        kernel = src[start:k].strip()
        close = src.find(">>>", k)
        cfg = _split_top(src[k + 3:close].replace("\\\n", " "))
        assert 2 <= len(cfg) <= 4, cfg
        smem = cfg[2] if len(cfg) > 2 else "0"
        p = close + 3
        while src[p].isspace() or src[p] == '\\':
            p += 1
        assert src[p] == '(', "no argument list after >>> (%s)" % kernel
        q = _match_paren(src, p)
        args = src[p + 1:q]
        out.append(src[pos:start])
        stream = cfg[3] if len(cfg) > 3 else "nullptr"
        ktext = "".join(kernel.split()).replace('"', "")
        out.append("cuemu::launch(\"%s\", dim3(%s), dim3(%s), (size_t)(%s), (cudaStream_t)(%s), [_cuemu_args = std::make_tuple(%s)]() { "
                   "std::apply([](auto... _a) { %s(_a...); }, _cuemu_args); })" % (ktext, cfg[0], cfg[1], smem, stream, args, kernel))
        pos = q + 1
    return "".join(out)


_DYN = re.compile(r"extern\s+__shared__\s+([A-Za-z_][A-Za-z0-9_ ]*?)\s+([A-Za-z_][A-Za-z0-9_]*)\s*\[\s*\]\s*;")


def translate(text):
    text = _DYN.sub(lambda m: "%s* %s = (%s*)cuemu::dyn_smem();" % (m.group(1), m.group(2), m.group(1)), text)
    return rewrite_launches(text)


def _digest(paths):
    h = hashlib.sha1()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(verbose=False):
    os.makedirs(OUT, exist_ok=True)
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps += [os.path.join(HERE, f) for f in ("cuda_runtime.h", "cuemu.cpp", "build_emu.py")]
    deps.append(os.path.join(ROOT, "include", "mdgrad_b200.h"))
    stamp = _digest(deps)
    stamp_file = os.path.join(OUT, "stamp")
    if os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    objs = []
    procs = []
    # headers with launches / dynamic smem are translated too and shadow the originals through the include order
    for f in os.listdir(CSRC):
        if f.endswith(".cuh"):
            with open(os.path.join(CSRC, f)) as fh:
                t = translate(fh.read())
            with open(os.path.join(OUT, f), "w") as fh:
                fh.write(t)
    for f in cu:
        with open(os.path.join(CSRC, f)) as fh:
            t = translate(fh.read())
        cpp = os.path.join(OUT, f[:-3] + ".emu.cpp")
        with open(cpp, "w") as fh:
            fh.write('#line 1 "%s"\n' % os.path.join(CSRC, f))
            fh.write(t)
        obj = cpp[:-4] + ".o"
        objs.append(obj)
        procs.append((f, subprocess.Popen([CXX] + FLAGS[:-2] + ["-I", OUT, "-I", CSRC, "-c", cpp, "-o", obj],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    rt = os.path.join(OUT, "cuemu.o")
    procs.append(("cuemu.cpp", subprocess.Popen([CXX] + FLAGS + ["-c", os.path.join(HERE, "cuemu.cpp"), "-o", rt],
                                                stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs.append(rt)
    for name, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("cuemu build failed for %s:\n%s" % (name, out))
        if verbose and out.strip():
            print(out)
    r = subprocess.run([CXX, "-shared", "-o", LIB] + (["-fsanitize=address,undefined"] if SANITIZE else []) + objs + ["-ldl", "-lm"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("cuemu link failed:\n" + r.stdout + r.stderr)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


def build_fake_nccl():
    """tests/cuemu/fake_nccl.cpp -> _build/libfakenccl.so (NCCL over shared memory between emulated ranks, tests/test_emu_dist.py)"""
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libfakenccl.so")
    src = os.path.join(HERE, "fake_nccl.cpp")
    if os.path.exists(lib) and os.path.getmtime(lib) >= os.path.getmtime(src):
        return lib
    r = subprocess.run([CXX, "-O2", "-g", "-std=c++17", "-fPIC", "-shared", src, "-o", lib, "-ldl", "-lrt", "-lpthread"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("fake_nccl build failed:\n" + r.stdout + r.stderr)
    return lib


def build_selftest():
    """tests/cuemu/selftest.cu -> _build/selftest (the emulator checking itself, tests/test_emu_selftest.py)"""
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "selftest")
    srcs = [os.path.join(HERE, f) for f in ("selftest.cu", "cuemu.cpp", "cuda_runtime.h", "build_emu.py")]
    if os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(f) for f in srcs):
        return exe
    cpp = os.path.join(OUT, "selftest.emu.cpp")
    with open(srcs[0]) as fh:
        t = translate(fh.read())
    with open(cpp, "w") as fh:
        fh.write('#line 1 "%s"\n' % srcs[0])
        fh.write(t)
    r = subprocess.run([CXX] + FLAGS + [cpp, srcs[1], "-o", exe, "-rdynamic", "-ldl", "-lm"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("cuemu selftest build failed:\n" + r.stdout + r.stderr)
    return exe


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
