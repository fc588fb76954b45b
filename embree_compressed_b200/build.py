"""Builds libecb200.so (CUDA kernels + C-ABI host layer) in-tree for sm_100a with nvcc.

nvcc cross-compiles without a GPU.  The resulting .so is git-ignored but travels to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libecb200.so")
SHIM = os.path.join(HERE, "libembree3_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              # arithmetic contract of the kernels (see csrc/ecb_kernels.cu): no implicit FMA contraction,
              # IEEE division and square root, denormals kept
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2,-Wall"]


def nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_variant(name, defines, only_if_stale=False):
    """libecb200_<name>.so with extra -D flags: the bounds-checking build ("check", ECB_BOUNDS_CHECK=1) and kernel A/B
    experiments (tools/kbench.py)."""
    out = os.path.join(HERE, "libecb200_%s.so" % name)
    srcs = [os.path.join(CSRC, f) for f in ("ecb_kernels.cu", "ecb_build.cu", "ecb_api.cpp", "ecb_import.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [os.path.join(HERE, "..", "include", "ecb200.h"), __file__]
    if only_if_stale and not _stale(out, deps):
        return out
    cmd = [nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-shared", "-o", out] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    return out


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("ecb_kernels.cu", "ecb_build.cu", "ecb_api.cpp", "ecb_import.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + \
        [os.path.join(HERE, "..", "include", "ecb200.h"), __file__]
    if force or _stale(LIB, deps):
        cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + srcs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building libecb200.so")
    shim_src = os.path.join(CSRC, "rtcore_shim.cpp")
    if os.path.exists(shim_src) and (force or _stale(SHIM, [shim_src, LIB] + deps)):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall",
               "-I" + os.path.join(HERE, "..", "include"), "-o", SHIM, shim_src,
               "-L" + HERE, "-lecb200", "-Wl,-rpath,$ORIGIN", "-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed building libembree3_b200.so")
    demo_src = os.path.join(HERE, "..", "tests", "cpp", "shim_demo.cpp")
    demo = os.path.join(HERE, "..", "tests", "cpp", "shim_demo")
    if os.path.exists(demo_src) and os.path.exists(SHIM) and (force or _stale(demo, [demo_src, SHIM])):
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-o", demo, demo_src, "-L" + HERE, "-lembree3_b200", "-lecb200",
               "-Wl,-rpath," + HERE]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed building tests/cpp/shim_demo")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
