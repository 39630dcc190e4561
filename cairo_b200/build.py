"""In-tree build of the native libraries (nvcc cross-compiles sm_100a without a GPU).

    python -m cairo_b200.build          # libevxgpu.so (CUDA pixel pipeline + C-ABI) and libevx1.so (C++ host API)
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
GPU_SO = os.path.join(HERE, "libevxgpu.so")
HOST_SO = os.path.join(HERE, "libevx1.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(sub, exts):
    d = os.path.join(CSRC, sub) if sub else CSRC
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_gpu(force=False, verbose=False):
    deps = _sources("", (".cu", ".cuh")) + [os.path.join(ROOT, "include", "evxgpu.h")]
    if not force and not _newer(GPU_SO, deps):
        return GPU_SO
    extra = os.environ.get("EVX_EXTRA_NVCC", "").split()      # e.g. -DEVX_K3_TRACE for profiles/trace_k3.py
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", GPU_SO, os.path.join(CSRC, "evxgpu.cu")]
    subprocess.check_call(cmd)
    lint_sass(GPU_SO)
    return GPU_SO


def lint_sass(so):
    """Guard against a ptxas 12.9 miscompile met in round 1: when it rematerialises the packed
    negation `add.u16x2 r, ~v, 0x00010001` (evx_neg16x2) in split halves it can drop the immediate and
    emit `VIADD.16x2 Rd, Ra, 0x0` -- every difference against that source word is then off by one
    (the PTX is correct; profiles/r01_summary.md has the trace).  No kernel here adds a zero
    constant on purpose, so that instruction anywhere in the library fails the build."""
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True, check=True).stdout
    bad = re.findall(r"VIADD\.16x2 R\d+, R\d+, 0x0 ;", sass)
    if bad:
        os.remove(so)
        raise RuntimeError("ptxas dropped a packed-add immediate (%d x '%s'); perturb the source and rebuild" % (len(bad), bad[0]))


def build_host(force=False):
    hdir = os.path.join(CSRC, "host")
    srcs = _sources("host", (".cpp",))
    deps = srcs + _sources("host", (".h",)) + [os.path.join(ROOT, "include", "evxgpu.h"), os.path.join(ROOT, "include", "evx1_c.h")]
    if not srcs:
        return None
    if not force and not _newer(HOST_SO, deps):
        return HOST_SO
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", hdir,
           "-o", HOST_SO] + srcs + ["-L", HERE, "-levxgpu", "-Wl,-rpath,$ORIGIN", "-lpthread"]
    subprocess.check_call(cmd)
    return HOST_SO


def build_all(force=False, verbose=False):
    build_gpu(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", GPU_SO, HOST_SO if os.path.exists(HOST_SO) else "(no host library yet)")
