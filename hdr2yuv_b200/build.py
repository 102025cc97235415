#!/usr/bin/env python3
"""Build hdr2yuv_b200/libhdr2yuv_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with gpurun snapshots.
-fmad=false: the reference's numerics are SSE2 without FMA; every contraction
the kernels want is written explicitly with __fma_* intrinsics.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhdr2yuv_b200.so")
OBJDIR = os.path.join(HERE, "build")
SOURCES = ["h2y_api.cu", "h2y_stats.cu", "h2y_staged.cu", "h2y_forward.cu", "h2y_forward2.cu", "h2y_inverse.cu"]
HEADERS = [os.path.join(CSRC, "h2y_device.cuh"), os.path.join(CSRC, "h2y_internal.h"), os.path.join(CSRC, "h2y_f32x2.cuh"),
           os.path.join(ROOT, "include", "hdr2yuv_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + HEADERS):
        return obj, ""
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    out = r.stdout.decode(errors="replace")
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, out))
    return obj, out


def build(verbose=False, force=False):
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [o for o, _ in results]
    log = "".join(t for _, t in results)
    if _stale(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout.decode(errors="replace"))
    if verbose:
        sys.stdout.write(log)
    return LIB


CLI_DIR = os.path.join(HERE, "cli")
CLI_BIN = os.path.join(CLI_DIR, "bin")
CXX = os.environ.get("CXX", "g++")


def build_cli(force=False):
    """The C++ command-line hosts (hdr2yuv, yuv2tiff) and the reader test tool, linked against the in-tree
    libhdr2yuv_b200.so (rpath $ORIGIN/../.. so they run from the tree on the GPU box)."""
    build()
    os.makedirs(CLI_BIN, exist_ok=True)
    io_src = [os.path.join(CLI_DIR, "h2y_io.cpp"), os.path.join(CLI_DIR, "h2y_io.h")]
    common = [CXX, "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + CLI_DIR]
    gpu_link = ["-L" + HERE, "-lhdr2yuv_b200", "-Wl,-rpath,$ORIGIN/../..", "-lz", "-lpthread"]
    targets = {"hdr2yuv": (["hdr2yuv_main.cpp"], gpu_link), "yuv2tiff": (["yuv2tiff_main.cpp"], gpu_link),
               "h2y_iotool": (["h2y_iotool.cpp"], ["-lz"])}
    out = {}
    for name, (srcs, link) in targets.items():
        exe = os.path.join(CLI_BIN, name)
        deps = [os.path.join(CLI_DIR, s) for s in srcs] + io_src + [os.path.join(ROOT, "include", "hdr2yuv_b200.h"), LIB]
        if force or _stale(exe, deps):
            cmd = common + [os.path.join(CLI_DIR, s) for s in srcs] + [io_src[0], "-o", exe] + link
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
            if r.returncode != 0:
                raise RuntimeError("build of %s failed:\n%s" % (name, r.stdout.decode(errors="replace")))
        out[name] = exe
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
    print(build_cli(force="-f" in sys.argv))
