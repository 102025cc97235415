#!/usr/bin/env python3
"""Build the parity checkers.  TEST INFRASTRUCTURE ONLY -- nothing here ships in the product.

  oracle/libh2yoracle.so   the plain-C restatement (h2y_oracle.c); builds anywhere gcc exists.
  oracle/_ref/libh2yref.so the reference's own convert.cpp/common.cpp/tiff.cpp compiled UNMODIFIED,
                           read in place from /root/reference (never copied), plus ref_shim.cpp.
  oracle/_ref/yuv2tiff_ref the reference's whole yuv2tiff.cpp program.
  oracle/_ref/hdr2yuv_ref  the reference's whole hdr2yuv program: its own main() / parse_options() / read_file()
                           (hdr2yuv.cpp) linked with the same convert/common/tiff objects and ref_io_stubs.cpp
                           and the reference's own dpx.cpp (the EXR codec, which needs OpenEXR, stops when reached).  The CLI tests
                           byte-compare hdr2yuv_b200/cli/bin/hdr2yuv with it on .rgb / .yuv / .tiff sources.

The two reference sources that include "/usr/local/include/tiffio.h" by absolute path
(tiff.cpp:3, yuv2tiff.cpp:5) are streamed through sed into the compiler's stdin so that the
include resolves to oracle/stub/tiffio.h; no modified copy is written anywhere.

Flags follow the reference's make.sh:4-16 (g++ -O2; no -march, no FMA, no fast-math): the
reference numerics are x86-64 SSE2 scalar.  oracle/_ref/ is git-ignored but travels with gpurun.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("H2Y_REFERENCE_DIR", "/root/reference")
OUT_REF = os.path.join(HERE, "_ref")
ORACLE_SO = os.path.join(HERE, "libh2yoracle.so")
REF_SO = os.path.join(OUT_REF, "libh2yref.so")
REF_YUV2TIFF = os.path.join(OUT_REF, "yuv2tiff_ref")
REF_HDR2YUV = os.path.join(OUT_REF, "hdr2yuv_ref")

CFLAGS = ["-O2", "-fPIC", "-ffp-contract=off"]
LEGACY_INCLUDES = ["-include", "cstring", "-include", "limits", "-include", "cstdlib", "-include", "cstdio",
                   "-include", "strings.h"]


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, **kw)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout.decode(errors="replace"))
        raise RuntimeError("oracle build step failed: " + cmd[0])


def _newer(target, *sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_oracle(force=False):
    src = [os.path.join(HERE, "h2y_oracle.c"), os.path.join(HERE, "h2y_oracle.h")]
    if not force and _newer(ORACLE_SO, *src):
        return ORACLE_SO
    _run(["gcc", "-std=c11"] + CFLAGS + ["-shared", "-o", ORACLE_SO, src[0], "-lm"])
    return ORACLE_SO


def _compile_via_sed(ref_file, out_obj_or_exe, link=False):
    """Compile a reference source whose only edit is the tiffio.h include path (done in a pipe)."""
    with open(os.path.join(REF, ref_file), "rb") as f:
        text = f.read().replace(b'"/usr/local/include/tiffio.h"', b'"tiffio.h"')
    cmd = ["g++"] + CFLAGS + ["-w", "-I" + os.path.join(HERE, "stub"), "-I" + REF] + LEGACY_INCLUDES + \
          ["-x", "c++", "-"]
    cmd += ["-o", out_obj_or_exe] if link else ["-c", "-o", out_obj_or_exe]
    _run(cmd, input=text)


def build_ref(force=False):
    """Returns the .so path, or None when the reference tree is absent (e.g. on the GPU box,
    which uses the prebuilt files that travelled with the snapshot)."""
    if not os.path.isdir(REF):
        return REF_SO if os.path.exists(REF_SO) else None
    shim = os.path.join(HERE, "ref_shim.cpp")
    stub_h = os.path.join(HERE, "stub", "tiffio.h")
    io_stubs = os.path.join(HERE, "ref_io_stubs.cpp")
    if not force and _newer(REF_SO, shim, stub_h, io_stubs, os.path.join(REF, "convert.cpp")) and os.path.exists(REF_YUV2TIFF) and \
            os.path.exists(REF_HDR2YUV):
        return REF_SO
    os.makedirs(OUT_REF, exist_ok=True)
    objs = []
    for name in ("convert.cpp", "common.cpp"):
        obj = os.path.join(OUT_REF, name.replace(".cpp", ".o"))
        _run(["g++"] + CFLAGS + ["-w", "-I" + REF] + LEGACY_INCLUDES + ["-c", os.path.join(REF, name), "-o", obj])
        objs.append(obj)
    tiff_obj = os.path.join(OUT_REF, "tiff.o")
    _compile_via_sed("tiff.cpp", tiff_obj)
    objs.append(tiff_obj)
    # the reference's DPX codec, unmodified; its three OpenEXR includes resolve to oracle/stub (a stand-in for `half`)
    dpx_obj = os.path.join(OUT_REF, "dpx.o")
    _run(["g++"] + CFLAGS + ["-w", "-I" + os.path.join(HERE, "stub"), "-I" + REF] + LEGACY_INCLUDES + ["-c", os.path.join(REF, "dpx.cpp"), "-o", dpx_obj])
    objs.append(dpx_obj)
    shim_obj = os.path.join(OUT_REF, "ref_shim.o")
    _run(["g++"] + CFLAGS + ["-w", "-I" + REF] + LEGACY_INCLUDES + ["-c", shim, "-o", shim_obj])
    objs.append(shim_obj)
    _run(["g++", "-shared", "-o", REF_SO] + objs + ["-lm"])
    _compile_via_sed("yuv2tiff.cpp", REF_YUV2TIFF, link=True)
    # the reference's own main(): hdr2yuv.cpp + the three objects above + the codec stubs
    main_obj = os.path.join(OUT_REF, "hdr2yuv.o")
    stubs_obj = os.path.join(OUT_REF, "ref_io_stubs.o")
    _run(["g++"] + CFLAGS + ["-w", "-I" + REF] + LEGACY_INCLUDES + ["-c", os.path.join(REF, "hdr2yuv.cpp"), "-o", main_obj])
    _run(["g++"] + CFLAGS + ["-w", "-I" + REF] + LEGACY_INCLUDES + ["-c", io_stubs, "-o", stubs_obj])
    _run(["g++", "-o", REF_HDR2YUV, main_obj, stubs_obj] + [o for o in objs if not o.endswith("ref_shim.o")] + ["-lm"])
    objs += [main_obj, stubs_obj]
    for o in objs:
        os.remove(o)
    return REF_SO


def clean():
    if os.path.exists(ORACLE_SO):
        os.remove(ORACLE_SO)
    shutil.rmtree(OUT_REF, ignore_errors=True)


if __name__ == "__main__":
    if "clean" in sys.argv:
        clean()
    else:
        print(build_oracle(force="-f" in sys.argv))
        print(build_ref(force="-f" in sys.argv))
