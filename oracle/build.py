"""TEST INFRASTRUCTURE builds: the CPU oracle (oracle/amg_oracle.c -> liboracle.so) and the
reference-object oracle (unmodified /root/reference/src translation units -> _ref/libref_smem.so).
Not imported by the product package."""
import os
import subprocess
import sys

ORACLE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(ORACLE, "liboracle.so")
REF_LIB = os.path.join(ORACLE, "_ref", "libref_smem.so")
REF_B200_LIB = os.path.join(ORACLE, "_ref", "libref_b200.so")     # the same + INTEGRATION.md's binding, linked against libamg_b200.so


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build failed: " + " ".join(cmd[:3]))
    return r.stdout


def build_oracle(force=False):
    src = os.path.join(ORACLE, "amg_oracle.c")
    if not force and _newer(ORACLE_LIB, [src]):
        return ORACLE_LIB
    # the reference's own flags: -fopenmp -O3 (/root/reference/Makefile:37)
    _run(["gcc", "-O3", "-fopenmp", "-std=gnu11", "-shared", "-fPIC", src, "-o", ORACLE_LIB, "-lm"])
    return ORACLE_LIB


def build_ref(force=False):
    """Only possible where the reference tree is mounted (the build container); the GPU box uses
    the prebuilt file that travels with the snapshot."""
    if not os.path.isdir("/root/reference/src"):
        return REF_LIB if os.path.exists(REF_LIB) else None
    script = os.path.join(ORACLE, "build_ref.sh")
    shim = os.path.join(ORACLE, "ref_shim")
    deps = [script, os.path.join(ORACLE, "ref_driver.cpp"), os.path.join(os.path.dirname(ORACLE), "integration", "SMEM_B200.hpp"),
            os.path.join(os.path.dirname(ORACLE), "integration", "DMEM_B200.hpp")]
    deps += [os.path.join(d, f) for d, _, fs in os.walk(shim) for f in fs]
    if not force and _newer(REF_LIB, deps):
        return REF_LIB
    _run(["bash", script])
    return REF_LIB


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv), build_ref(force="--force" in sys.argv))
