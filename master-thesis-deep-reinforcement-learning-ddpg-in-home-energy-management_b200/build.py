"""Builds libshems_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libshems_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-ccbin", "/usr/bin/g++"]
# env.cu restates Julia scalar arithmetic: no a*b+c contraction allowed there
UNITS = [("env.cu", ["-fmad=false"]), ("replay.cu", []), ("ddpg.cu", []), ("ddpg_fused.cu", []), ("actor_rollout.cu", ["-fmad=false"]), ("tc_gemm.cu", []), ("series.cu", [])]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines, verbose=False):
    """tools/: a differently-tuned copy of the library (extra -D macros) as libshems_b200_<name>.so"""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out = os.path.join(HERE, f"libshems_b200_{name}.so")
    objs = []
    for src, extra in UNITS:
        o = os.path.join(CSRC, src.replace(".cu", f".{name}.o"))
        subprocess.check_call([nvcc] + ARCH + COMMON + extra + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) +
                              ["-c", os.path.join(CSRC, src), "-o", o])
        objs.append(o)
    subprocess.check_call([nvcc] + ARCH + ["-shared", "-o", out] + objs + ["-ccbin", "/usr/bin/g++"])
    return out


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "shems_b200.h"))
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        if force or _newer(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            subprocess.check_call(cmd)
        objs.append(o)
    if force or _newer(LIB, objs):
        subprocess.check_call([nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
