"""CPU tests of the oracle itself: the C restatement against the golden vectors generated from
the compiled reference, against the compiled reference directly (when /root/reference is
present), and its size-independent properties."""
import hashlib

import numpy as np
import pytest

import cases
from hdr2yuv_b200 import synth
from oracle import oracle as O


def _planes(src, px):
    return O.load_rgb16(px, src["full_range"]) if src["kind"] == "tiff16" else O.load_half(px)


@pytest.mark.parametrize("name,src,dst", cases.FORWARD_CASES, ids=[c[0] for c in cases.FORWARD_CASES])
def test_port_matches_golden_forward(golden_forward, name, src, dst):
    px = golden_forward[name + "/in"]
    yuv, tmp, stats = O.forward(_planes(src, px), cases.oracle_src(src), dst, backend="port", want_tmp=True)
    assert np.array_equal(stats, golden_forward[name + "/stats"])
    assert np.array_equal(tmp, golden_forward[name + "/tmp444"])
    assert np.array_equal(yuv, golden_forward[name + "/yuv"])


def test_golden_inputs_are_reproducible(golden_forward):
    # the committed inputs are exactly what the seeded generator makes
    for name, src, _ in cases.FORWARD_CASES[::7]:
        px, _ = cases.forward_input(src, cases.GW, cases.GH)
        assert np.array_equal(px, golden_forward[name + "/in"])


@pytest.mark.parametrize("case", cases.INVERSE_CASES, ids=lambda c: "b%d_m%d_fir%d_fr%d_a%d" % c)
def test_port_matches_golden_inverse(golden_inverse, case):
    bd, m, fir, fr, al = case
    yuv = cases.widen_yuv(golden_inverse[cases.inverse_input_key(m)], bd)
    rgb, invalid = O.yuv2tiff(yuv, cases.IW, cases.IH, bd, m, fir, fr, al, backend="port")
    key = "b%d_m%d_fir%d_fr%d_a%d" % case
    assert np.array_equal(rgb[:2], golden_inverse[key + "/head"])
    assert hashlib.sha256(rgb.tobytes()).digest() == golden_inverse[key + "/sha256"].tobytes()
    assert invalid == int(golden_inverse[key + "/invalid"][0])


needs_ref = pytest.mark.skipif(not O.ref_available(), reason="compiled reference (oracle/_ref) not available")


@needs_ref
def test_port_vs_reference_random_sizes():
    rng = np.random.default_rng(5)
    for w, h in ((40, 24), (128, 66), (200, 120)):
        px = synth.tiff16_frame(w, h, seed=w)
        planes = O.load_rgb16(px, 0)
        for m in (9, 11, 13):
            for res in (0, 1):
                if res == 0 and (w % 4 or h % 4):
                    continue
                dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=m, chroma=1, resampler=res)
                src = dict(bit_depth=16, full_range=0, transfer=16, primaries=10, matrix=0)
                assert np.array_equal(O.forward(planes, src, dst, "port"), O.forward(planes, src, dst, "ref"))
        hp = O.load_half(synth.exr_half_frame(w, h, seed=h))
        src = dict(bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0)
        dst = dict(bit_depth=12, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
        assert np.array_equal(O.forward(hp, src, dst, "port"), O.forward(hp, src, dst, "ref"))
    del rng


@needs_ref
def test_transfer_functions_vs_reference():
    x = np.concatenate([np.linspace(0, 1.5, 4001), [0.0, 1.0, 1e-6, 0.018]]).astype(np.float32)
    for which in range(6):
        a, b = O.transfer(which, x, "port"), O.transfer(which, x, "ref")
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), which


@needs_ref
def test_stats_snap_cascade_vs_reference():
    # estimated ceiling snaps to YMax/CMax of the picture depth (common.cpp:94-106)
    for bd, top in ((16, 50000), (16, 60000), (16, 60500), (10, 900), (10, 950), (12, 3000)):
        pl = np.full((3, 8, 8), 100, np.uint16)
        pl[:, 0, 0] = top
        assert np.array_equal(O.pic_stats(pl, bd, "port"), O.pic_stats(pl, bd, "ref"))
    f = np.array([0.7, 3.9, 1e-3], np.float32).reshape(1, 1, 3).repeat(3, 0)
    assert np.array_equal(O.pic_stats(f, 32, "port"), O.pic_stats(f, 32, "ref"))


def test_neutral_grey_lands_on_half_minus_one():
    # R=G=B=32768 -> Y=512, Cb=Cr=511 at 10 bits
    pl = np.full((3, 16, 16), 32768, np.uint16)
    src = dict(bit_depth=16, full_range=0, transfer=16, primaries=10, matrix=0)
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    yuv = O.forward(pl, src, dst)
    assert set(yuv[:256]) == {512} and set(yuv[256:]) == {511}


def test_fir_is_integer_exact_below_16_bits():
    # Appendix A item 7: (sum + 256) >> 9 with the clamps reproduces the float FIR for <= 15-bit data
    rng = np.random.default_rng(3)
    for bits in (10, 12, 15):
        p = rng.integers(0, 1 << bits, (48, 64)).astype(np.int64)
        top = (1 << bits) - 1
        got = O.subsample_fir_h(p.astype(np.uint16), 0, top).astype(np.int64)
        idx = np.arange(0, 64, 2)
        g = lambda o: p[:, np.clip(idx + o, 0, 63)]
        acc = 21 * (g(-5) + g(5)) - 52 * (g(-3) + g(3)) + 159 * (g(-1) + g(1)) + 256 * g(0)
        want = np.clip((acc + 256) >> 9, 0, top)
        assert np.array_equal(got, want)


def test_box_and_upsample_roundtrip_constant():
    p = np.full((24, 32), 777, np.uint16)
    assert np.all(O.subsample_box(p) == 777)
    up = O.upsample_420to444(np.full((12, 16), 500, np.uint16), 1, 0, 1023)
    assert np.all(up == 500)      # FIR taps sum to 256: DC is preserved exactly


@pytest.mark.parametrize("case", cases.MINV_CASES, ids=lambda c: "m%d_i%d_f%d_o%d" % c)
def test_port_matches_golden_matrix_inverse(golden_minv, case):
    m, ibd, fr, obd = case
    out, invalid = O.matrix_inverse(cases.minv_input(ibd), m, ibd, fr, obd, backend="port")
    assert np.array_equal(out, golden_minv["m%d_i%d_f%d_o%d" % case])
    assert invalid >= 0


def test_matrix_inverse_family_quirks():
    # convert.cpp:1387 compares matrix_coeffs with booleans: only 1 takes the 709 equations, 0 is the error,
    # 10 (BT2020c) falls through to Y'DzDx exactly like 9 and 11
    pl = cases.minv_input(12)
    a, _ = O.matrix_inverse(pl, 10, 12, 0, 12)
    b, _ = O.matrix_inverse(pl, 11, 12, 0, 12)
    c, _ = O.matrix_inverse(pl, 1, 12, 0, 12)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    with pytest.raises(RuntimeError):
        O.matrix_inverse(pl, 0, 12, 0, 12)
    rgb = O.write_tiff_rows(a, 16, 12)
    assert np.array_equal(rgb[..., 0], a[2] << 4) and np.array_equal(rgb[..., 1], a[0] << 4) and np.array_equal(rgb[..., 2], a[1] << 4)


@needs_ref
def test_matrix_inverse_port_vs_reference():
    rng = np.random.default_rng(8)
    for m in (1, 9, 11):
        pl = rng.integers(0, 4096, (3, 33, 50), dtype=np.uint16)
        assert np.array_equal(O.matrix_inverse(pl, m, 12, 0, 16, "port")[0], O.matrix_inverse(pl, m, 12, 0, 16, "ref")[0])


@needs_ref
def test_ybar_mode_and_yuvprime2_keyword_vs_reference_binary():
    # yuv2tiff -X (yuv2tiff.cpp:162, 365-399) against the reference's own executable; and Y'u''v'' 4:2:0 of the
    # forward chain (convert.cpp:533-801) against the reference's convert()
    w, h = 960, 540
    rng = np.random.default_rng(11)
    for bd, full, fir in ((12, False, True), (10, False, False)):
        top = (1 << bd) - 1
        y = rng.integers(0, top + 1, w * h, dtype=np.uint16)
        y[:50] = 0
        c = rng.integers(top // 4, 3 * top // 4, w * h // 2, dtype=np.uint16)
        yuv = np.concatenate([y, c])
        a, ia = O.yuv2tiff(yuv, w, h, bit_depth=bd, fir=fir, full_range=full, backend="ref", ybar=True)
        b, ib = O.yuv2tiff(yuv, w, h, bit_depth=bd, fir=fir, full_range=full, backend="port", ybar=True)
        assert np.array_equal(a, b) and ia == ib
    planes = rng.integers(0, 65536, (3, 36, 64), dtype=np.uint16)
    src = dict(bit_depth=16, full_range=1, transfer=18, primaries=10, matrix=0)
    for res, bd, fr in ((1, 16, 1), (0, 12, 0)):
        dst = dict(bit_depth=bd, full_range=fr, transfer=18, primaries=10, matrix=15, chroma=1, resampler=res)
        assert np.array_equal(O.forward(planes, src, dst, backend="ref"), O.forward(planes, src, dst, backend="port"))


@needs_ref
def test_matrix_convert_float_output_vs_reference():
    # The F32-output twin of matrix_convert (convert.cpp:1117-1122, 1222-1304): what an .exr / .dpx destination runs on an
    # integer source (hdr2yuv.cpp:797-823, 935-962).  Restatement == compiled reference, bit for bit, with and without a
    # transfer change, for every matrix family the branch has.
    w, h = 96, 40
    planes = O.load_rgb16(synth.tiff16_frame(w, h, seed=31), 0)
    for src_tr, dst_tr in ((16, 16), (16, 8), (16, 1), (1, 16)):
        for m, prim in ((0, 10), (9, 9), (1, 1), (11, 10), (13, 10)):
            for depth in (16, 32):
                src = dict(bit_depth=16, full_range=0, transfer=src_tr, primaries=10 if src_tr == 16 else 1, matrix=0)
                # matrix 0 with other primaries has no branch in the reference ("Can't determine color difference", exit)
                dst = dict(bit_depth=depth, full_range=0, transfer=dst_tr, primaries=prim if m else src["primaries"], matrix=m)
                if depth == 32:
                    continue        # set_pic_clip shifts by bit_depth - 8 = 24 on an unsigned: defined, but Half overflows; not a route
                a = O.matrix_convert_f32(planes, src, dst, "port")
                b = O.matrix_convert_f32(planes, src, dst, "ref")
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (src_tr, dst_tr, m, int((a.view(np.uint32) != b.view(np.uint32)).sum()))
