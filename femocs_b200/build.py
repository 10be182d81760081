"""Builds femocs_b200/lib/libfemocs_b200.so with nvcc for sm_100a (in-tree, no JIT cache)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libfemocs_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-ccbin", "/usr/bin/g++",
          "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off,-Wall,-Wno-unused-function"]
# (source, extra flags).  interp_kernels.cu must not contract a*b+c into FMA: cell indices are
# compared bit-for-bit with the reference, which is built for baseline x86-64.
UNITS = [
    ("host_setup.cpp", []),
    ("partition.cpp", []),
    ("poisson_kernels.cu", []),
    ("interp_kernels.cu", ["-fmad=false"]),
    ("twolevel.cu", []),
    ("q2.cu", []),
    ("api.cu", []),
]


def nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith(".h")]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "femocs_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs = []
    procs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc()] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    relink = force or bool(procs) or not os.path.exists(LIB)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + out)
        if p.returncode:
            raise RuntimeError("nvcc failed for " + cmd[-3])
    if relink:
        cmd = [nvcc()] + ARCH + ["-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-lgomp"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
