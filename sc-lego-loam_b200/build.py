"""In-tree build of the native libraries (run here on CPU: nvcc cross-compiles sm_100a without a GPU)."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_SCGPU = os.path.join(PKG, "libscgpu.so")
LIB_SCANGEN = os.path.join(PKG, "libscangen.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--fmad=false",  # bit-exact stages spell out every rounding; never let ptxas contract them
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_scangen(force=False):
    src = os.path.join(CSRC, "scangen.c")
    if force or _newer(LIB_SCANGEN, [src]):
        subprocess.run(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-o", LIB_SCANGEN, src, "-lm"], check=True)
    return LIB_SCANGEN


def build_scgpu(force=False, verbose=False):
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "scgpu.h"))
    if force or _newer(LIB_SCGPU, deps):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB_SCGPU] + srcs + ["-lcudart"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout)
        if r.returncode:
            raise RuntimeError("nvcc failed")
        with open(os.path.join(PKG, "ptxas_info.txt"), "w") as f:
            f.write(r.stdout)
    return LIB_SCGPU


def build_all(force=False, verbose=False):
    build_scangen(force)
    build_scgpu(force, verbose)
