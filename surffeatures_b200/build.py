"""Builds libsurf_b200.so (sm_100a) in-tree with nvcc.  No JIT cache: the .so travels with the repo."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsurf_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]

# (source, extra flags).  Stages 1-4 reproduce the reference's un-contracted fp32 arithmetic.
UNITS = [
    ("surf_kernels.cu", ["-fmad=false"]),
    ("surf_describe.cu", ["-fmad=false"]),
    ("surf_match.cu", []),
    ("surf_match_tc.cu", []),
    ("surf_engine.cu", []),
    ("surf_db.cu", []),
    ("surf_comm.cu", []),
    ("surf_vote.cu", ["-fmad=false"]),
    ("surf_mha.cu", ["-fmad=false"]),
    ("surf_pose.cu", ["-fmad=false"]),
]


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libsurf_b200.so cannot be built (there is no CPU fallback)")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "surf_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper without OpenMP specs; nvcc needs a plain host g++
    host = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    objs = []
    procs = []
    for src, extra in UNITS:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, "-ccbin", host, *ARCH, *COMMON, *extra, "-Xptxas", "-v" if verbose else "-warn-spills",
               "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nvcc, "-ccbin", host, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart", "-ldl"]
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
