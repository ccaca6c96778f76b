"""Build libgiwaxs_b200.so (sm_100a only) in-tree with nvcc.

    python -m giwaxsim_b200.build [--force]

The shared object is written next to this file so that it travels with the
source tree; it is git-ignored.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libgiwaxs_b200.so")
SOURCES = ["gx_api.cu", "gx_atoms.cu", "gx_project.cu", "gx_fft.cu", "gx_bin.cu", "gx_detector.cu",
           "gx_detector_affine.cu", "gx_fused.cu", "gx_slab.cu", "gx_comm.cu", "gx_compare.cu", "gx_hostcopy.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=default", "--expt-relaxed-constexpr", "-DGX_TWP=1",
         "-Xptxas", "-v"]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "giwaxs_b200.h"))
    return hdrs


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _obj_name(src):
    return os.path.splitext(src)[0] + ".o"


def _compile(src):
    obj = os.path.join(OBJ, _obj_name(src))
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + _deps()):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", path, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    return obj, p.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if force:
        for s in srcs:
            o = os.path.join(OBJ, _obj_name(s))
            if os.path.exists(o):
                os.remove(o)
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                      "-Xcompiler", "-fPIC", "-ldl"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
