"""-m gpu parity tests of the forward path (h2y_forward and the staged entry points) against the
golden vectors generated from the compiled reference and against the oracle."""
import os

import numpy as np
import pytest
import torch

import cases
import gpu_util as G
from hdr2yuv_b200 import _cabi as cabi
from hdr2yuv_b200 import api, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _uses_transfer(src, dst):
    # libm on the path (transfer change, or the rho-gamma EOTF inside Y'u''v''): <= 1 code, deviations counted
    return src["transfer"] != dst["transfer"] or (dst["matrix"] == 15 and dst["chroma"] == 1)


@pytest.mark.parametrize("name,src,dst", cases.FORWARD_CASES, ids=[c[0] for c in cases.FORWARD_CASES])
def test_fused_forward_matches_golden(ctx, golden_forward, name, src, dst):
    px = golden_forward[name + "/in"]
    got = G.gpu_forward(ctx, [px], src, dst)[0]
    G.compare_codes(got, golden_forward[name + "/yuv"], _uses_transfer(src, dst), name)


@pytest.mark.parametrize("name,src,dst", cases.FORWARD_CASES[::3], ids=[c[0] for c in cases.FORWARD_CASES[::3]])
def test_exact_math_and_staged_routes_agree(ctx, opt, golden_forward, name, src, dst):
    # the reciprocal fast path, the reference-order FP64 path and the staged kernels give the same codes
    px = golden_forward[name + "/in"]
    fast = G.gpu_forward(ctx, [px], src, dst)[0]
    opt("H2Y_EXACT_MATH", "1")
    exact = G.gpu_forward(ctx, [px], src, dst)[0]
    opt("H2Y_EXACT_MATH", "0")
    opt("H2Y_FORCE_STAGED", "1")
    staged = G.gpu_forward(ctx, [px], src, dst)[0]
    opt("H2Y_FORCE_STAGED", "0")
    assert np.array_equal(fast, exact)
    G.compare_codes(staged, fast, _uses_transfer(src, dst), name + " staged-vs-fused")


_TIFF = dict(kind="tiff16", bit_depth=16, full_range=0, transfer=16, primaries=10, matrix=0)
_HALF = dict(kind="half", bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0)


@pytest.mark.parametrize("w,h", [(8, 12), (240, 64), (248, 34), (488, 130), (1000, 96), (1920, 1080)])
@pytest.mark.parametrize("matrix", [9, 11])
def test_fused_forward_sizes_tiff(ctx, w, h, matrix):
    # config 1 / 3 shape: 16-bit X'Y'Z' through read_tiff's clip, no transfer change -> bit exact
    for chroma, res in ((1, 1), (1, 0), (3, 1), (2, 1)):
        if res == 0 and (w % 4 or h % 4):
            continue
        if h > 500 and (chroma, res) != (1, 1):
            continue
        dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=matrix, chroma=chroma, resampler=res)
        frames = [synth.tiff16_frame(w, h, seed=s, smooth=bool(s & 1)) for s in (1, 2)]
        got = G.gpu_forward(ctx, frames, _TIFF, dst)
        for f, g in zip(frames, got):
            G.compare_codes(g, G.oracle_forward(f, _TIFF, dst), False, "tiff %dx%d c%d r%d" % (w, h, chroma, res))


@pytest.mark.parametrize("w,h", [(16, 14), (240, 48), (496, 270), (1920, 1080)])
def test_fused_forward_sizes_half_pq(ctx, w, h):
    # config 2 / 5 shape: half RGBA linear light -> PQ; per-frame statistics differ inside one batch
    for bd, ch in ((10, 4), (12, 3)):
        dst = dict(bit_depth=bd, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
        frames = [synth.exr_half_frame(w, h, seed=0, channels=ch), synth.exr_half_frame(w, h, seed=1, channels=ch, hi=350.0),
                  synth.exr_half_frame(w, h, seed=2, channels=ch, correlated=True)]
        frames[1][..., 0] = np.minimum(frames[1][..., 0], np.float16(100.0).view(np.uint16))  # per-channel ranges differ
        got = G.gpu_forward(ctx, frames, _HALF, dst)
        for f, g in zip(frames, got):
            G.compare_codes(g, G.oracle_forward(f, _HALF, dst), True, "half %dx%d b%d" % (w, h, bd))


def test_full_4k_frame_half_pq(ctx):
    # BASELINE config 2 at full size, one frame through the oracle (about 3 s of CPU)
    w, h = 3840, 2160
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    f = synth.exr_half_frame(w, h, seed=0, channels=4)
    got = G.gpu_forward(ctx, [f, f], _HALF, dst)
    assert np.array_equal(got[0], got[1])                      # batch position does not matter
    nbad = G.compare_codes(got[0], G.oracle_forward(f, _HALF, dst), True, "4K half PQ")
    print("4K PQ frame: %d of %d samples deviate by one code" % (nbad, got[0].size))
    st = ctx.forward_last_stats(0)
    assert list(st.estimated_floor) == [0, 0, 0] and list(st.estimated_ceiling) == [4000, 4000, 4000]


def test_odd_geometry_takes_general_route(ctx):
    # widths that are not a multiple of 8 go through the staged kernels inside h2y_forward
    for w, h in ((100, 50), (36, 20)):
        dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
        f = synth.tiff16_frame(w, h, seed=3)
        G.compare_codes(G.gpu_forward(ctx, [f], _TIFF, dst)[0], G.oracle_forward(f, _TIFF, dst), False, "odd tiff")
        f = synth.exr_half_frame(w, h, seed=3, channels=4)
        G.compare_codes(G.gpu_forward(ctx, [f], _HALF, dst)[0], G.oracle_forward(f, _HALF, dst), True, "odd half")


def test_staged_entry_points_match_oracle_intermediates(ctx, golden_forward):
    # pic_stats -> matrix_convert -> convert -> write_yuv clamp, one C-ABI call per reference function
    for name, src, dst in [c for c in cases.FORWARD_CASES if c[0] in ("tiff_m9_c1_r1_b10", "half_pq_m9_b10_f0", "half_rho_m9")]:
        px = golden_forward[name + "/in"]
        h, w = px.shape[:2]
        is_u16 = src["kind"] == "tiff16"
        planes = O.load_rgb16(px, src["full_range"]) if is_u16 else O.load_half(px)
        d_in = [torch.from_numpy(planes[c].copy()).cuda() for c in range(3)]
        in_pic = api.pic_desc(w, h, cabi.CHROMA_444, src["transfer"], src["primaries"], src["matrix"], src["bit_depth"],
                              src["full_range"], cabi.PIC_TYPE_U16 if is_u16 else cabi.PIC_TYPE_F32,
                              cabi.LAYOUT_PLANAR_U16 if is_u16 else cabi.LAYOUT_PLANAR_F32)
        st = ctx.pic_stats(in_pic, d_in)
        want_stats = golden_forward[name + "/stats"]
        assert list(st.estimated_floor) + list(st.estimated_ceiling) == list(want_stats)
        tmp_depth = src["bit_depth"] if is_u16 else dst["bit_depth"]
        tmp_pic = api.pic_desc(w, h, cabi.CHROMA_444, dst["transfer"], dst["primaries"], dst["matrix"], tmp_depth,
                               dst["full_range"])
        d_tmp = [torch.zeros(h * w, dtype=torch.int16, device="cuda") for _ in range(3)]
        ctx.matrix_convert(tmp_pic, d_tmp, in_pic, d_in, st)
        torch.cuda.synchronize()
        tmp = np.stack([t.cpu().numpy().view(np.uint16).reshape(h, w) for t in d_tmp])
        G.compare_codes(tmp, golden_forward[name + "/tmp444"], _uses_transfer(src, dst), name + " tmp444")
        out_pic = api.pic_desc(w, h, dst["chroma"], dst["transfer"], dst["primaries"], dst["matrix"], dst["bit_depth"],
                               dst["full_range"])
        d_out = [torch.zeros(h * w, dtype=torch.int16, device="cuda") for _ in range(3)]
        ctx.convert(out_pic, d_out, tmp_pic, d_tmp, dst["resampler"])
        ctx.write_yuv_clamp(out_pic, d_out, tmp_depth)
        torch.cuda.synchronize()
        pw, ph = api.plane_dims(w, h, dst["chroma"])
        got = np.concatenate([d_out[c].cpu().numpy().view(np.uint16)[: pw[c] * ph[c]] for c in range(3)])
        G.compare_codes(got, golden_forward[name + "/yuv"], _uses_transfer(src, dst), name + " staged yuv")


def test_matrix_convert_float_output_twin(ctx):
    # F32-output branch (convert.cpp:1222-1304), reachable through the staged call
    w, h = 64, 16
    px = synth.exr_half_frame(w, h, seed=9, channels=4)
    planes = O.load_half(px)
    import ctypes as C
    for m in (9, 11, 13, 0):
        pic = O._Pic(w, h, 3, 8, 1, 0, 32, 1, O.PIC_F32)
        outp = O._Pic(w, h, 3, 16, 9 if m else 1, m, 12, 0, O.PIC_F32)
        want = np.zeros((3, h, w), np.float32)
        for c in range(3):
            pic.fbuf[c] = planes[c].ctypes.data
            outp.fbuf[c] = want[c].ctypes.data
        O.port_lib().orc_pic_stats(C.byref(pic), None, None)
        assert O.port_lib().orc_matrix_convert(C.byref(outp), C.byref(pic)) == 0
        in_pic = api.pic_desc(w, h, 3, 8, 1, 0, 32, 1, cabi.PIC_TYPE_F32, cabi.LAYOUT_PLANAR_F32)
        out_pic = api.pic_desc(w, h, 3, 16, 9 if m else 1, m, 12, 0, cabi.PIC_TYPE_F32, cabi.LAYOUT_PLANAR_F32)
        d_in = [torch.from_numpy(planes[c].copy()).cuda() for c in range(3)]
        d_out = [torch.zeros(h * w, dtype=torch.float32, device="cuda") for _ in range(3)]
        st = ctx.pic_stats(in_pic, d_in)
        ctx.matrix_convert(out_pic, d_out, in_pic, d_in, st)
        torch.cuda.synchronize()
        got = np.stack([t.cpu().numpy().reshape(h, w) for t in d_out])
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
        assert float(rel.max()) <= 1e-5, (m, float(rel.max()))   # north_star tolerance for float output


def test_host_pipeline_equals_device_path(ctx):
    w, h, n = 496, 270, 11
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    frames = [synth.exr_half_frame(w, h, seed=s, channels=4) for s in range(n)]
    dev = G.gpu_forward(ctx, frames, _HALF, dst)
    params = api.forward_params(w, h, cabi.LAYOUT_HALF_RGBA, _HALF, dst, resampler=1, clip_on_load=0)
    src = np.ascontiguousarray(np.stack(frames, 0))
    out = np.zeros((n, api.yuv_frame_bytes(w, h, 1) // 2), np.uint16)
    ctx.forward_host(params, src, out, n)
    for i in range(n):
        assert np.array_equal(out[i], dev[i]), i
    pin_in = api.PinnedBuffer(src.nbytes)
    pin_out = api.PinnedBuffer(out.nbytes)
    pin_in.array[:] = src.view(np.uint8).reshape(-1)
    ctx.forward_host(params, pin_in.ptr, pin_out.ptr, n)
    assert np.array_equal(pin_out.view(np.uint16).reshape(n, -1), out)
    pin_in.free()
    pin_out.free()


def test_error_convention(ctx):
    w, h = 64, 16
    d_src = torch.zeros(w * h * 8, dtype=torch.uint8, device="cuda")
    d_dst = torch.zeros(w * h * 6, dtype=torch.uint8, device="cuda")
    good_dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    p = api.forward_params(w, h, cabi.LAYOUT_HALF_RGBA, _HALF, good_dst)
    p.src.chroma_format_idc = cabi.CHROMA_420            # matrix_convert's 4:4:4 precondition (convert.cpp:886-890)
    with pytest.raises(cabi.H2YError) as e:
        ctx.forward(p, d_src, d_dst, 1)
    assert e.value.status == cabi.ERR_PRECONDITION
    p = api.forward_params(w, h, cabi.LAYOUT_HALF_RGBA, _HALF, dict(good_dst, matrix=5))
    with pytest.raises(cabi.H2YError) as e:              # "Can't determine color difference to use?" (1195-1198)
        ctx.forward(p, d_src, d_dst, 1)
    assert e.value.status == cabi.ERR_MATRIX
    p = api.forward_params(w, h, cabi.LAYOUT_RGB16, dict(_TIFF, bit_depth=10), dict(good_dst, bit_depth=12))
    with pytest.raises(cabi.H2YError) as e:              # write_yuv: dst bitdepth > src bitdepth (tiff.cpp:396-401)
        ctx.forward(p, d_src, d_dst, 1)
    assert e.value.status == cabi.ERR_BIT_DEPTH


# ---- the two EXR-route kernels (h2y_forward2.cu) --------------------------------------------------------
def _force_kernel(opt, which):
    if which:
        opt("H2Y_FORWARD_KERNEL", which)
    else:
        opt("H2Y_FORWARD_KERNEL", "")


@pytest.mark.parametrize("which", ["ring", "rows", None])
@pytest.mark.parametrize("matrix,depth,full", [(9, 10, 0), (9, 12, 0), (11, 10, 0), (1, 10, 0), (9, 10, 1), (11, 12, 1)])
def test_exr_kernels_match_oracle(ctx, opt, which, matrix, depth, full):
    # widths that are and are not multiples of the 240-px strip, several frames per batch, RGB and RGBA
    _force_kernel(opt, which)
    for (w, h, ch) in ((240, 66, 3), (488, 130, 4), (1000, 34, 3)):
        dst = dict(bit_depth=depth, full_range=full, transfer=16, primaries=9, matrix=matrix, chroma=1, resampler=1)
        frames = [synth.exr_half_frame(w, h, seed=100 + s, channels=ch) for s in range(3)]
        got = G.gpu_forward(ctx, frames, _HALF, dst)
        for f, g in zip(frames, got):
            G.compare_codes(g, G.oracle_forward(f, _HALF, dst), True, "%s %dx%d m%d b%d" % (which, w, h, matrix, depth))


def test_large_4k_batch_takes_rows_kernel_and_matches_oracle(ctx, opt):
    # 10 frames of 3840x2160: the batch size at which h2y_forward picks the warp-autonomous kernel by itself.
    # Every frame is checked through a checksum of the two oracle-verified frames' neighbours: frames repeat
    # with period 2, so frames 0 and 1 (oracle) pin all ten.
    w, h, n = 3840, 2160, 10
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    base = [synth.exr_half_frame_fast(w, h, seed=s, channels=3) for s in (0, 1)]
    frames = [base[i % 2] for i in range(n)]
    opt("H2Y_PLAN_REUSE", "0")          # the classic two-pass route (tests/test_plan_reuse_gpu.py covers the single pass)
    before = ctx.kernel_launches
    got = G.gpu_forward(ctx, frames, _HALF, dst)
    # init, stats, plan, LUT, rows kernel (two-LUT, single-LUT and the two three-table instantiations), general-kernel sweep
    assert ctx.kernel_launches - before == 9
    for i in (0, 1):
        G.compare_codes(got[i], G.oracle_forward(base[i], _HALF, dst), True, "4K frame %d" % i)
    for i in range(2, n):
        assert np.array_equal(got[i], got[i % 2]), i


def test_rows_kernel_splits_frames_by_lut_fit(ctx, opt):
    # frames whose largest sample is above ~10 800 (half code >= LUT2_CODES) cannot keep two pre-scaled LUT copies in
    # shared memory: they go to the single-LUT instantiation, the others to the two-LUT one, in the same call
    opt("H2Y_FORWARD_KERNEL", "rows")
    w, h = 480, 128
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    frames = [synth.exr_half_frame(w, h, seed=60 + s, channels=3, hi=hi) for s, hi in enumerate((4000.0, 40000.0, 900.0, 60000.0))]
    got = G.gpu_forward(ctx, frames, _HALF, dst)
    for i, f in enumerate(frames):
        G.compare_codes(got[i], G.oracle_forward(f, _HALF, dst), True, "frame %d" % i)
    opt("H2Y_NO_SPECIALISED", "1")         # and the generic (run-time constants) instantiation agrees
    again = G.gpu_forward(ctx, frames, _HALF, dst)
    for a, b in zip(got, again):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("depth", [10, 12])
def test_rows_kernel_keeps_the_clamp_when_a_frame_can_reach_it(ctx, opt, depth):
    # The two-LUT instantiation drops matrix_convert's chroma clamp only for frames whose LUT extremes prove it cannot
    # bind.  A frame with max in [1, 2) has range (int)max - (int)min = 1, so its normalised samples reach ~1.9 and the
    # PQ values ~1.07: such frames must take the single-LUT instantiation (clamp kept) inside the same call, next to
    # ordinary frames that take the clamp-free one.  Saturated primaries make the clamp actually bind.
    opt("H2Y_FORWARD_KERNEL", "rows")
    w, h = 480, 128
    dst = dict(bit_depth=depth, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    frames = [synth.exr_half_frame(w, h, seed=80 + s, channels=3, hi=hi) for s, hi in enumerate((1.9, 4000.0, 1.02, 1.5, 3.7))]
    sat = frames[0].copy()
    sat[::2, ::3, 0] = 0x3F99          # R ~ 1.9 next to G = B = 0: Cr far above, Cb far below the legal range
    sat[::2, ::3, 1:] = 0
    sat[1::2, 1::3, 2] = 0x3F99        # and the same for B
    sat[1::2, 1::3, :2] = 0
    frames.append(sat)
    got = G.gpu_forward(ctx, frames, _HALF, dst)
    for i, f in enumerate(frames):
        G.compare_codes(got[i], G.oracle_forward(f, _HALF, dst), True, "frame %d" % i)


@pytest.mark.parametrize("which", ["ring", "rows"])
def test_unclean_frames_fall_back_to_the_general_kernel(ctx, opt, which):
    # negative zero, +inf and a different (floor, ceiling) per frame: frames 1 and 2 are not "clean" and must come
    # out of the v1 kernel, bit-identical to what the oracle gives; frames 0 and 3 stay on the fast kernel
    _force_kernel(opt, which)
    w, h = 256, 64
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    frames = [synth.exr_half_frame(w, h, seed=40 + s, channels=3, hi=1000.0 * (s + 1)) for s in range(4)]
    frames[1] = frames[1].copy(); frames[1][5, 7, 1] = 0x8000                  # -0.0
    frames[2] = frames[2].copy(); frames[2][9, 100, 2] = 0x7C00                # +inf
    got = G.gpu_forward(ctx, frames, _HALF, dst)
    for i in (0, 1, 3):
        G.compare_codes(got[i], G.oracle_forward(frames[i], _HALF, dst), True, "frame %d" % i)
    # +inf makes the reference's own range (int)inf - floor overflow: only require that the fast kernel declined it
    opt("H2Y_FORCE_V1", "1")
    ref2 = G.gpu_forward(ctx, [frames[2]], _HALF, dst)[0]
    assert np.array_equal(got[2], ref2)


@pytest.mark.parametrize("which", ["ring", "rows"])
def test_wide_and_short_pictures(ctx, opt, which):
    # more strips than warps (8K: 32 strips of 240), two-row pictures, a single 8-pixel column
    _force_kernel(opt, which)
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    for (w, h, n) in ((7680, 36, 2), (8, 2, 3), (8, 130, 1), (3848, 2, 2)):
        frames = [synth.exr_half_frame(w, h, seed=200 + s, channels=3) for s in range(n)]
        got = G.gpu_forward(ctx, frames, _HALF, dst)
        for f, g in zip(frames, got):
            G.compare_codes(g, G.oracle_forward(f, _HALF, dst), True, "%s %dx%d" % (which, w, h))


@pytest.mark.parametrize("big_endian", [True, False])
def test_dpx10_layout_unpacks_on_the_gpu_and_matches_oracle(ctx, big_endian):
    # H2Y_LAYOUT_DPX10_BE / _LE: the file's packed words go to the device as stored; the unpack and the /1023.0 of
    # dpx_read (dpx.cpp:506-531) happen there and the picture continues as an F32 source (hdr2yuv.cpp:700-737).
    # Code 1023 and code 0 are present, so (int)max - (int)min = 1 (a DPX without a full-scale sample has range 0 in
    # the reference, common.cpp:135-136).
    w, h = 136, 52
    rng = np.random.default_rng(21)
    codes = rng.integers(0, 1024, (2, h, w, 3), dtype=np.uint32)
    codes[:, 3, 5] = 1023
    codes[:, 7, 9] = 0
    words = (codes[..., 0] << 22) | (codes[..., 1] << 12) | (codes[..., 2] << 2)
    stored = words.astype(">u4" if big_endian else "<u4")
    src = dict(kind="dpx", bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0)
    layout = cabi.LAYOUT_DPX10_BE if big_endian else cabi.LAYOUT_DPX10_LE
    for dst in (dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1),
                dict(bit_depth=12, full_range=1, transfer=18, primaries=9, matrix=11, chroma=3, resampler=1),
                dict(bit_depth=10, full_range=0, transfer=8, primaries=1, matrix=1, chroma=1, resampler=0)):
        params = api.forward_params(w, h, layout, src, dst, resampler=dst["resampler"])
        assert api.src_frame_bytes(params.src) == w * h * 4
        d_src = G.to_dev(stored.view(np.uint8))
        nbytes = api.yuv_frame_bytes(w, h, dst["chroma"])
        d_dst = torch.zeros(nbytes * 2, dtype=torch.uint8, device="cuda")
        ctx.forward(params, d_src, d_dst, 2)
        torch.cuda.synchronize()
        got = d_dst.cpu().numpy().view(np.uint16).reshape(2, -1)
        for i in range(2):
            want = O.forward(O.load_dpx10(stored[i], big_endian), cases.oracle_src(src), dst, backend="port")
            G.compare_codes(got[i], want, src["transfer"] != dst["transfer"], "dpx frame %d" % i)


@pytest.mark.parametrize("bits", [16, 32])
@pytest.mark.parametrize("big_endian", [True, False])
def test_dpx16_and_float_layouts_unpack_on_the_gpu_and_match_oracle(ctx, bits, big_endian):
    # H2Y_LAYOUT_DPX16_* (sample = code / 65535.0, dpx.cpp:478-494) and H2Y_LAYOUT_DPXF32_* (floats as they are,
    # dpx.cpp:412-443): the file's samples go to the device in the file's byte order.  The Python loaders used as the
    # oracle here are pinned to the reference's own dpx_read on the CPU (tests/test_cli_cpu.py).
    w, h = 136, 52
    rng = np.random.default_rng(bits + int(big_endian))
    if bits == 16:
        c = rng.integers(0, 65536, (2, h, w, 3), dtype=np.uint16)
        c[:, 3, 5] = 65535
        c[:, 7, 9] = 0
        stored = c.astype(">u2" if big_endian else "<u2")
        load, layout, bpp = O.load_dpx16, (cabi.LAYOUT_DPX16_BE if big_endian else cabi.LAYOUT_DPX16_LE), 6
    else:
        c = np.exp(rng.uniform(np.log(1e-3), np.log(300.0), (2, h, w, 3))).astype(np.float32)
        stored = c.astype(">f4" if big_endian else "<f4")
        load, layout, bpp = O.load_dpxf32, (cabi.LAYOUT_DPXF32_BE if big_endian else cabi.LAYOUT_DPXF32_LE), 12
    src = dict(kind="dpx", bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0)
    for dst in (dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1),
                dict(bit_depth=12, full_range=1, transfer=18, primaries=9, matrix=11, chroma=3, resampler=1)):
        params = api.forward_params(w, h, layout, src, dst, resampler=dst["resampler"])
        assert api.src_frame_bytes(params.src) == w * h * bpp
        d_src = G.to_dev(stored.view(np.uint8))
        nbytes = api.yuv_frame_bytes(w, h, dst["chroma"])
        d_dst = torch.zeros(nbytes * 2, dtype=torch.uint8, device="cuda")
        ctx.forward(params, d_src, d_dst, 2)
        torch.cuda.synchronize()
        got = d_dst.cpu().numpy().view(np.uint16).reshape(2, -1)
        for i in range(2):
            want = O.forward(load(stored[i], big_endian), cases.oracle_src(src), dst, backend="port")
            G.compare_codes(got[i], want, True, "dpx%d frame %d" % (bits, i))


@pytest.mark.parametrize("depth", [16, 32])
def test_float_destination_host_route(ctx, depth):
    # h2y_forward_f32_host = what an .exr / .dpx destination runs (hdr2yuv.cpp:797-823): pic_stats -> matrix_convert into
    # float planes.  Depth 16 (an .exr destination) against the compiled reference's own F32 branch when it is here
    # (else the restatement, which a CPU test pins to it bit for bit); depth 32 (what a .dpx destination is forced to)
    # is the all-zero picture the reference's program writes (tests/test_cli_cpu.py checks that against its file).
    w, h, n = 200, 64, 3
    frames = [synth.tiff16_frame(w, h, seed=60 + i) for i in range(n)]
    backend = "ref" if O.ref_available() else "port"
    for src_tr, dst_tr, m, prim in ((16, 16, 9, 9), (16, 8, 0, 10), (16, 1, 1, 1), (16, 16, 11, 10)):
        src = dict(kind="tiff16", bit_depth=16, full_range=0, transfer=src_tr, primaries=10, matrix=0)
        dst = dict(bit_depth=depth, full_range=0, transfer=dst_tr, primaries=prim, matrix=m, chroma=3, resampler=1)
        params = api.forward_params(w, h, cabi.LAYOUT_RGB16, src, dst, resampler=1, clip_on_load=1)
        h_in = np.ascontiguousarray(np.stack(frames, 0))
        h_out = np.full((n, 3, h, w), -1.0, np.float32)
        ctx.forward_f32_host(params, h_in, h_out, n)
        for i in range(n):
            if depth == 32:
                assert not h_out[i].any()
                continue
            want = O.matrix_convert_f32(O.load_rgb16(frames[i], 0), cases.oracle_src(src), dst, backend)
            if src_tr == dst_tr:
                assert np.array_equal(h_out[i].view(np.uint32), want.view(np.uint32)), (m, i)
            else:
                rel = np.abs(h_out[i].astype(np.float64) - want) / np.maximum(np.abs(want.astype(np.float64)), 1e-6)
                assert float(rel.max()) <= 1e-5, (m, i, float(rel.max()))      # the path's tolerance for float output


@pytest.mark.parametrize("which", ["ring", "rows"])
@pytest.mark.parametrize("matrix,depth,full,channels", [(9, 10, 0, 3), (11, 10, 0, 3), (9, 12, 0, 4), (9, 16, 0, 3), (11, 12, 1, 4), (9, 10, 1, 3)])
def test_tiff_420_kernels_match_oracle(ctx, opt, which, matrix, depth, full, channels):
    # the two kernels behind the integer 4:2:0 FIR route (CTA ring / warp-autonomous with a private row ring), forced in
    # turn; 16-bit tmp depth, reference operation order in both filters -> bit exact.  Widths that are and are not
    # multiples of the 240-pixel strip, several frames so that row ranges cross frame boundaries, RGB and RGBA rows.
    _force_kernel(opt, which)
    src = dict(_TIFF, full_range=full)
    for (w, h) in ((240, 66), (488, 130), (1000, 34), (8, 2), (3848, 4)):
        dst = dict(bit_depth=depth, full_range=full, transfer=16, primaries=9, matrix=matrix, chroma=1, resampler=1)
        frames = [synth.tiff16_frame(w, h, seed=300 + s, channels=channels, smooth=bool(s & 1)) for s in range(3)]
        got = G.gpu_forward(ctx, frames, src, dst)
        for f, g in zip(frames, got):
            G.compare_codes(g, G.oracle_forward(f, src, dst), False, "%s %dx%d m%d b%d" % (which, w, h, matrix, depth))


def test_large_1080p_tiff_batch_takes_rows_kernel_and_matches_oracle(ctx):
    # 40 frames of 1920x1080: the batch size at which the integer route picks the warp-autonomous kernel by itself
    w, h, n = 1920, 1080, 40
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    base = [synth.tiff16_frame(w, h, seed=s, smooth=bool(s & 1)) for s in (0, 1, 2)]
    frames = [base[i % 3] for i in range(n)]
    before = ctx.kernel_launches
    got = G.gpu_forward(ctx, frames, _TIFF, dst)
    assert ctx.kernel_launches - before == 1
    for i in range(3):
        G.compare_codes(got[i], G.oracle_forward(base[i], _TIFF, dst), False, "1080p frame %d" % i)
    for i in range(3, n):
        assert np.array_equal(got[i], got[i % 3]), i


@pytest.mark.parametrize("depth,matrix", [(10, 9), (12, 9), (10, 11), (10, 1)])
def test_rows_kernel_three_table_frames(ctx, opt, depth, matrix):
    # Real footage: the channels' (int)min / (int)max differ, so the reference normalises each channel on its own
    # (common.cpp:135-136, convert.cpp:936-940) and a frame needs three transfer tables.  The rows kernel keeps three
    # range-restricted tables in shared memory when they fit (FrameK::clean3); other frames of the same call still go
    # to the single-table instantiations or, when nothing fits, to the general kernel.
    opt("H2Y_FORWARD_KERNEL", "rows")
    w, h = 480, 130
    dst = dict(bit_depth=depth, full_range=0, transfer=16, primaries=9, matrix=matrix, chroma=1, resampler=1)
    rng = np.random.default_rng(7)

    def footage(seed, top=(900.0, 650.0, 400.0), bottom=0.02):
        v = np.exp(rng.uniform(np.log(bottom), np.log(1.0), (h, w, 3))) * np.array(top)
        v = v.astype(np.float16)
        for c in range(3):
            v[seed % h, (seed * 7 + c) % w, c] = np.float16(top[c])      # pin each channel's maximum
        return np.ascontiguousarray(v).view(np.uint16)

    frames = [footage(1), synth.exr_half_frame(w, h, seed=2, channels=3),            # three tables; one table
              footage(3, top=(4000.0, 3000.5, 1999.0)), footage(4, top=(100.0, 100.9, 100.2)),   # three; one (same ints)
              footage(5, top=(4000.0, 2500.0, 1200.0), bottom=1e-5)]                  # code ranges too wide: general kernel
    # exact zeros (black bars) in some or all channels: the tables start one entry below the smallest nonzero code
    z = footage(6).copy(); z[:9] = 0; z[-7:, :, 1] = 0; frames.append(z)
    z = footage(7, top=(3000.0, 2000.0, 1000.0)).copy(); z[40:50, 100:300, 2] = 0; z[3, 5, 0] = 1; frames.append(z)   # and a denormal
    # Three-table frames that can reach matrix_convert's chroma clamp (convert.cpp:1207-1213) must keep it: channel
    # maxima in [1, 4) give ranges of 1..3, normalised samples up to ~1.9 and PQ values above 1.  The clamp-free
    # three-table instantiation (three_nc_frame) has to decline them; saturated primaries make the clamp bind.
    frames.append(footage(8, top=(1.9, 2.5, 3.7)))
    sat = footage(9, top=(1.9, 2.9, 3.9)).copy()
    sat[::2, ::3, 0] = np.float16(1.9).view(np.uint16); sat[::2, ::3, 1:] = np.float16(0.05).view(np.uint16)
    sat[1::2, 1::3, 2] = np.float16(3.9).view(np.uint16); sat[1::2, 1::3, :2] = np.float16(0.05).view(np.uint16)
    frames.append(sat)
    got = G.gpu_forward(ctx, frames, _HALF, dst)
    for i, f in enumerate(frames):
        G.compare_codes(got[i], G.oracle_forward(f, _HALF, dst), True, "frame %d" % i)
    # Imf::Rgba rows (alpha = 1.0 rides along and must not enter any table index)
    rgba = [np.concatenate([f, np.full((h, w, 1), 0x3C00, np.uint16)], -1) for f in frames[:3]]
    got = G.gpu_forward(ctx, rgba, _HALF, dst)
    for i, f in enumerate(rgba):
        G.compare_codes(got[i], G.oracle_forward(f, _HALF, dst), True, "rgba frame %d" % i)


@pytest.mark.gpu
def test_stream_order_behind_the_overlapped_rows_launches(ctx, opt):
    # The rows-kernel instantiations of one call run on forked streams (each may start while its predecessor drains) and
    # are joined to the caller's stream by events.  Work queued behind h2y_forward on the same stream must still see every frame
    # converted: a device-to-host copy queued on that stream right after the call, with no device-wide synchronise in
    # between, has to return the same bytes as a synchronised run.  Single-table frames are converted by the FIRST
    # launch (the long one, the later launches find nothing to do), three-table frames by the last ones.
    opt("H2Y_FORWARD_KERNEL", "rows")
    w, h, n = 1920, 1080, 12
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    single = synth.exr_half_frame_fast(w, h, seed=3, channels=3)
    three = synth.exr_half_frame_smooth_fast(w, h, seed=4)
    for frames in ([single] * n, [three] * n, [single, three] * (n // 2)):
        want = np.stack(G.gpu_forward(ctx, frames, _HALF, dst))
        params = api.forward_params(w, h, cabi.LAYOUT_HALF_RGB, _HALF, dst, resampler=1, clip_on_load=0)
        d_src = G.to_dev(np.stack(frames, 0))
        nbytes = api.yuv_frame_bytes(w, h, 1)
        host = torch.empty(nbytes * n, dtype=torch.uint8).pin_memory()
        st = torch.cuda.Stream()
        for _ in range(3):
            d_dst = torch.zeros(nbytes * n, dtype=torch.uint8, device="cuda")
            host.zero_()
            torch.cuda.synchronize()
            with torch.cuda.stream(st):
                ctx.forward(params, d_src, d_dst, n, stream=st.cuda_stream)
                host.copy_(d_dst, non_blocking=True)
            st.synchronize()
            assert np.array_equal(host.numpy().view(np.uint16).reshape(n, -1), want)


@pytest.mark.gpu
def test_host_pipeline_with_the_rows_kernels(ctx, opt):
    # h2y_forward_host with the warp-autonomous kernels forced: per chunk the compute stream runs the overlapped rows
    # launches and the general-kernel sweep, records an event, and the copy stream drains the chunk behind that event.
    # Single-table and three-table frames alternate, so both the first and the last launch of a call do real work.
    opt("H2Y_FORWARD_KERNEL", "rows")
    w, h, n = 1920, 1080, 20
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    base = [synth.exr_half_frame_fast(w, h, seed=11, channels=3), synth.exr_half_frame_smooth_fast(w, h, seed=12)]
    frames = [base[(i // 3) % 2] for i in range(n)]
    want = [G.gpu_forward(ctx, [f], _HALF, dst)[0] for f in base]
    params = api.forward_params(w, h, cabi.LAYOUT_HALF_RGB, _HALF, dst, resampler=1, clip_on_load=0)
    src = np.ascontiguousarray(np.stack(frames, 0))
    out = np.zeros((n, api.yuv_frame_bytes(w, h, 1) // 2), np.uint16)
    for _ in range(2):
        out[:] = 0
        ctx.forward_host(params, src, out, n)
        for i in range(n):
            assert np.array_equal(out[i], want[(i // 3) % 2]), i
