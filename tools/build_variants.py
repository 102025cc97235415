#!/usr/bin/env python3
"""Build experiment variants of the library: tools/build_variants.py name=-DFLAG=V,-DFLAG2=W ...

One translation unit is recompiled per variant (h2y_forward2.cu, or name@file.cu=flags for another); the other objects come from the normal build.  Output:
hdr2yuv_b200/_variants/libh2y_<name>.so (git-ignored, travels with gpurun).  bench.py / the tests pick a variant up
through H2Y_LIB=<path> (hdr2yuv_b200/_cabi.py).  Used for the A/B timings recorded in profiles/.
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hdr2yuv_b200 import build as B  # noqa: E402


def main():
    B.build()
    out_dir = os.path.join(B.HERE, "_variants")
    os.makedirs(out_dir, exist_ok=True)
    procs = []
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition("=")
        name, _, src = name.partition("@")               # name@file.cu=flags recompiles another translation unit
        src = src or "h2y_forward2.cu"
        flags = [f for f in flags.split(",") if f]
        obj = os.path.join(out_dir, "%s_%s.o" % (src.replace(".cu", ""), name))
        cmd = [B.NVCC] + B.FLAGS + flags + ["-c", os.path.join(B.CSRC, src), "-o", obj]
        procs.append((name, obj, subprocess.Popen(cmd), src))
    for name, obj, p, src in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed for variant " + name)
        objs = [os.path.join(B.OBJDIR, s.replace(".cu", ".o")) for s in B.SOURCES if s != src] + [obj]
        lib = os.path.join(out_dir, "libh2y_%s.so" % name)
        subprocess.check_call([B.NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
        print(lib)


if __name__ == "__main__":
    main()
