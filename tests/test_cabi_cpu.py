"""CPU tests of the boundary: the library loads, exports every symbol the header declares, and
the pure-host helpers agree with the oracle.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

from hdr2yuv_b200 import _cabi as cabi
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hdr2yuv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(h2y_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "header declares %s but the library does not export it" % n
    assert sorted(cabi.SYMBOLS) == names, "ctypes binding and header disagree"
    assert cabi.lib().h2y_abi_version() == 2


def test_clip_limits_match_oracle(built_lib):
    from hdr2yuv_b200 import api
    for bd in (8, 10, 12, 14, 16):
        for fr in (0, 1):
            a = api.set_pic_clip(bd, fr)
            b = O._Clip()
            O.port_lib().orc_set_clip(bd, fr, C.byref(b))
            assert (a.minCV, a.maxCV, a.minVR, a.maxVR, a.minVRC, a.maxVRC, a.Half) == \
                   (b.minCV, b.maxCV, b.minVR, b.maxVR, b.minVRC, b.maxVRC, b.Half)
    # the values SURVEY 8(a2) quotes
    c = api.set_pic_clip(10, 0)
    assert (c.minVR, c.maxVR, c.maxVRC, c.Half) == (64, 940, 960, 512)


def test_geometry_helpers(built_lib):
    from hdr2yuv_b200 import api
    assert api.plane_dims(3840, 2160, cabi.CHROMA_420) == ([3840, 1920, 1920], [2160, 1080, 1080])
    assert api.plane_dims(1920, 1080, cabi.CHROMA_422) == ([1920, 960, 960], [1080, 1080, 1080])
    assert api.yuv_frame_bytes(1920, 1080, cabi.CHROMA_420) == 6220800
    assert api.yuv_frame_bytes(3840, 2160, cabi.CHROMA_420) == 24883200
    for chroma in (1, 2, 3):
        assert api.yuv_frame_bytes(64, 48, chroma) == 2 * O.yuv_frame_samples(64, 48, chroma)
    d = api.pic_desc(3840, 2160, layout=cabi.LAYOUT_HALF_RGBA)
    assert api.src_frame_bytes(d) == 3840 * 2160 * 8
    d = api.pic_desc(3840, 2160, layout=cabi.LAYOUT_RGB16)
    assert api.src_frame_bytes(d) == 3840 * 2160 * 6


def test_tmp_bit_depth_rule(built_lib):
    from hdr2yuv_b200 import api
    # hdr2yuv.cpp:803-808: U16 in -> keep the input depth; float in -> destination depth
    s16 = api.pic_desc(8, 8, bit_depth=16, layout=cabi.LAYOUT_RGB16)
    sh = api.pic_desc(8, 8, bit_depth=32, layout=cabi.LAYOUT_HALF_RGBA)
    d10 = api.pic_desc(8, 8, bit_depth=10)
    assert cabi.lib().h2y_tmp_bit_depth(C.byref(s16), C.byref(d10)) == 16
    assert cabi.lib().h2y_tmp_bit_depth(C.byref(sh), C.byref(d10)) == 10


def test_no_gpu_means_loud_failure(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hdr2yuv_b200 import api
    with pytest.raises(cabi.H2YError):
        api.Context(0)


def test_product_never_imports_oracle():
    # the product path must not route through the oracle or any CPU fallback
    pkg = os.path.join(ROOT, "hdr2yuv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower() or f in ("synth.py",), f
