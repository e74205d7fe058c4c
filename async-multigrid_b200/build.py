"""In-tree builds: the sm_100a CUDA library (the product) and the host input-provider library.

Everything is compiled with explicit nvcc / g++ command lines into the source tree so
the resulting .so files travel to the GPU box with the repository snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")

CUDA_LIB = os.path.join(PKG, "libamg_b200.so")
HOST_LIB = os.path.join(PKG, "libamg_host.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build failed: " + " ".join(cmd[:3]))
    return r.stdout


def _nccl_dirs():
    """torch-bundled NCCL (2.28.x): headers + libnccl.so.2."""
    try:
        import nvidia.nccl as n  # type: ignore
        base = list(n.__path__)[0]
        inc, lib = os.path.join(base, "include"), os.path.join(base, "lib")
        if os.path.exists(os.path.join(inc, "nccl.h")):
            return inc, lib
    except Exception:
        pass
    return None, None


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_cuda(force=False, verbose=False):
    srcs = cuda_sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "amg_b200.h"))
    if not force and _newer(CUDA_LIB, deps):
        return CUDA_LIB
    inc, lib = _nccl_dirs()
    cmd = [NVCC, "--threads", "0", "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC,-fopenmp",
           "-rdc=false", "-I", os.path.join(ROOT, "include"), "-I", CSRC] + ARCH
    if verbose:
        cmd += ["-Xptxas", "-v"]
    if inc:
        cmd += ["-DAMG_HAVE_NCCL=1", "-I", inc]
    cmd += srcs + ["-o", CUDA_LIB, "-lcudart", "-lgomp"]
    if inc:
        cmd += ["-L", lib, "-l:libnccl.so.2", "-Xlinker", "-rpath=" + lib]
    out = _run(cmd)
    if verbose:
        print(out)
    return CUDA_LIB


def build_host(force=False):
    src = os.path.join(HOST, "amg_host.cpp")
    if not force and _newer(HOST_LIB, [src]):
        return HOST_LIB
    _run(["g++", "-O3", "-fopenmp", "-std=c++17", "-shared", "-fPIC", src, "-o", HOST_LIB])
    return HOST_LIB


def build_all(force=False):
    build_host(force)
    build_cuda(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", [p for p in (CUDA_LIB, HOST_LIB) if os.path.exists(p)])
