#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref, built from /root/reference).

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
The vectors pin the C restatement (oracle/h2y_oracle.c) and the CUDA path wherever the
reference itself cannot run.  Forward vectors keep input and output; inverse vectors keep the
input frame once and a SHA-256 per parameter set plus the first 2 output rows.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from hdr2yuv_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert O.ref_available(), "needs /root/reference to build oracle/_ref"
    fwd = {}
    for name, src, dst in cases.FORWARD_CASES:
        px, _ = cases.forward_input(src, cases.GW, cases.GH)
        if src["kind"] == "tiff16":
            planes = O.load_rgb16(px, src["full_range"])
        else:
            planes = O.load_half(px)
        yuv, tmp, stats = O.forward(planes, cases.oracle_src(src), dst, backend="ref", want_tmp=True)
        if dst["chroma"] == 2:
            # the reference has no 4:2:2 branch (its chroma planes stay uninitialised):
            # luma and tmp444 come from the reference, chroma is stage 1 of its FIR on that tmp444
            n = cases.GW * cases.GH
            luma = yuv[:n].copy()
            yuv = O.forward(planes, cases.oracle_src(src), dst, backend="port")
            assert np.array_equal(yuv[:n], luma)
            top = (1 << 16) - 1
            for c in (1, 2):
                want = O.subsample_fir_h(tmp[c], 0, top).reshape(-1) >> (16 - dst["bit_depth"])
                want = np.clip(want, 64, 960)
                got = yuv[n + (c - 1) * (n // 2): n + c * (n // 2)]
                assert np.array_equal(got, want)
        fwd[name + "/in"] = px
        fwd[name + "/yuv"] = yuv
        fwd[name + "/tmp444"] = tmp
        fwd[name + "/stats"] = stats.astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "forward.npz"), **fwd)

    # inverse: one realistic 10-bit 4:2:0 frame made by the reference's own forward chain
    px = synth.exr_half_frame(cases.IW, cases.IH, seed=11, channels=4, correlated=True)
    planes = O.load_half(px)
    inv = {}
    # two input frames keep the fixture small: a Y'DzDx one, and a BT.2020 one shared by the rest
    for m_fwd, tag in ((11, "ydzdx"), (9, "ycbcr")):
        dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=m_fwd, chroma=1, resampler=1)
        inv["yuv10_" + tag] = O.forward(planes, dict(bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0),
                                        dst, backend="ref")
    for bd, m, fir, fr, al in cases.INVERSE_CASES:
        yuv = cases.widen_yuv(inv[cases.inverse_input_key(m)], bd)
        rgb, invalid = O.yuv2tiff(yuv, cases.IW, cases.IH, bd, m, fir, fr, al, backend="ref")
        key = "b%d_m%d_fir%d_fr%d_a%d" % (bd, m, fir, fr, al)
        inv[key + "/sha256"] = np.frombuffer(hashlib.sha256(rgb.tobytes()).digest(), np.uint8)
        inv[key + "/head"] = rgb[:2].copy()
        inv[key + "/invalid"] = np.array([invalid], np.int64)
    np.savez_compressed(os.path.join(HERE, "inverse.npz"), **inv)
    # matrix_inverse (convert.cpp:1320): outputs of the compiled reference; inputs are regenerated from the seed
    mi = {}
    for m, ibd, fr, obd in cases.MINV_CASES:
        out, _ = O.matrix_inverse(cases.minv_input(ibd), m, ibd, fr, obd, backend="ref")
        mi["m%d_i%d_f%d_o%d" % (m, ibd, fr, obd)] = out
    np.savez_compressed(os.path.join(HERE, "matrix_inverse.npz"), **mi)
    for f in ("forward.npz", "inverse.npz", "matrix_inverse.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
