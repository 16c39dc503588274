"""Builds the in-tree native libraries (sm_100a only; nvcc cross-compiles without a GPU):

  libdfsa_b200.so  CUDA kernels + runtime + transports behind the C-ABI of include/dfsa_b200.h
  libdfsa_host.so  the host C++ drop-in headers (host/*.hpp) wrapped as extern "C" for ctypes / FFI users

Usage: python distributed-full-state-algorithms_b200/build.py [--force]
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
INCLUDE = os.path.join(ROOT, "include")

CUDA_SOURCES = ["dfsa_runtime.cu", "dfsa_comm.cu", "dfsa_kernels_sv.cu", "dfsa_kernels_manytarg.cu", "dfsa_kernels_dm.cu", "dfsa_kernels_fused.cu"]
LIB_DEVICE = os.path.join(PKG, "libdfsa_b200.so")
LIB_HOST = os.path.join(PKG, "libdfsa_host.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo", "--extended-lambda",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
    "-I", INCLUDE, "-I", CSRC,
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)      # the image exports a wrapper gcc without libgomp specs
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        raise RuntimeError("build failed: %s" % cmd[0])
    return res.stdout + res.stderr


def build(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [os.path.join(INCLUDE, "dfsa_b200.h")]
    objs = []
    for src in CUDA_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        if force or _newer(o, [s] + headers):
            out = _run(["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
            if verbose:
                print(out)
        objs.append(o)
    if force or _newer(LIB_DEVICE, objs):
        _run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_DEVICE] + objs + ["-lnccl", "-lrt"])
    host_src = os.path.join(HOST, "dfsa_host_capi.cpp")
    if os.path.exists(host_src):
        host_headers = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".hpp")]
        if force or _newer(LIB_HOST, [host_src, LIB_DEVICE] + host_headers + headers):
            _run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I", INCLUDE, "-I", HOST, host_src, "-o", LIB_HOST,
                  "-L", PKG, "-ldfsa_b200", "-Wl,-rpath,$ORIGIN"])
    # the reference's timing demo against the drop-in headers: proves the C++ surface compiles and links standalone
    demo_src = os.path.join(ROOT, "examples", "main.cpp")
    demo_bin = os.path.join(ROOT, "examples", "main")
    if os.path.exists(demo_src):
        host_headers = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".hpp")]
        if force or _newer(demo_bin, [demo_src, LIB_DEVICE] + host_headers):
            _run(["g++", "-std=c++17", "-O2", "-Wall", "-I", INCLUDE, "-I", HOST, demo_src, "-o", demo_bin,
                  "-L", PKG, "-ldfsa_b200", "-Wl,-rpath," + PKG])
    return LIB_DEVICE, LIB_HOST


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB_DEVICE)
