"""In-tree build of libwpt.so with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB = os.path.join(PKG_DIR, "libwpt.so")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp", ".h", "Makefile"))]
    deps.append(os.path.join(PKG_DIR, "..", "include", "wpt.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every CUDA/C++ source of the package into wasm_pathtracer_b200/libwpt.so."""
    if force or _stale():
        cmd = ["make", "-C", CSRC] + (["-B"] if force else [])
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout)
        if r.returncode != 0:
            raise RuntimeError("building libwpt.so failed")
    return LIB


TOOLS_DIR = os.path.join(PKG_DIR, "..", "tools")
RENDER = os.path.join(TOOLS_DIR, "wpt_render")


def build_tools(force=False):
    """Compile tools/wpt_render (the worker-style progressive driver, plain C++ over include/wpt.h)."""
    src = os.path.join(TOOLS_DIR, "wpt_render.cpp")
    if force or not os.path.exists(RENDER) or os.path.getmtime(RENDER) < max(os.path.getmtime(src), os.path.getmtime(LIB)):
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", src, "-I" + os.path.join(PKG_DIR, "..", "include"), "-L" + PKG_DIR, "-lwpt",
               "-Wl,-rpath,$ORIGIN/../wasm_pathtracer_b200", "-o", RENDER]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("building tools/wpt_render failed")
    # tools/wpt_peaks: L2 / HBM / FP32 roofs measured on the box
    psrc, pexe = os.path.join(TOOLS_DIR, "wpt_peaks.cu"), os.path.join(TOOLS_DIR, "wpt_peaks")
    if force or not os.path.exists(pexe) or os.path.getmtime(pexe) < os.path.getmtime(psrc):
        r = subprocess.run(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", psrc, "-o", pexe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("building tools/wpt_peaks failed")
    return RENDER
