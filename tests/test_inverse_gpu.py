"""-m gpu parity tests of the inverse path (h2y_inverse: one yuv2tiff main-loop iteration per frame)."""
import hashlib

import numpy as np
import pytest
import torch

import cases
import gpu_util as G
from hdr2yuv_b200 import _cabi as cabi
from hdr2yuv_b200 import api, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def gpu_inverse(ctx, yuvs, w, h, bd, m, fir, fr, al, ybar=0):
    p = cabi.InverseParams(w, h, bd, m, fir, fr, al, ybar)
    n = len(yuvs)
    d_yuv = G.to_dev(np.stack(yuvs, 0))
    nch = 4 if al else 3
    d_rgb = torch.zeros(n * w * h * nch * 2, dtype=torch.uint8, device="cuda")
    d_inv = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx.inverse(p, d_yuv, d_rgb, n, invalid=d_inv)
    torch.cuda.synchronize()
    rgb = d_rgb.cpu().numpy().view(np.uint16).reshape(n, h, w, nch)
    return rgb, d_inv.cpu().numpy()


@pytest.mark.parametrize("kernel", ["tile", "rows"])
@pytest.mark.parametrize("case", cases.INVERSE_CASES, ids=lambda c: "b%d_m%d_fir%d_fr%d_a%d" % c)
def test_inverse_matches_golden(ctx, golden_inverse, case, kernel, opt):
    opt("H2Y_INVERSE_KERNEL", kernel)          # both kernels behind h2y_inverse
    bd, m, fir, fr, al = case
    yuv = cases.widen_yuv(golden_inverse[cases.inverse_input_key(m)], bd)
    rgb, inv = gpu_inverse(ctx, [yuv], cases.IW, cases.IH, bd, m, fir, fr, al)
    key = "b%d_m%d_fir%d_fr%d_a%d" % case
    assert np.array_equal(rgb[0][:2], golden_inverse[key + "/head"])
    assert hashlib.sha256(rgb[0].tobytes()).digest() == golden_inverse[key + "/sha256"].tobytes()
    assert int(inv[0]) == int(golden_inverse[key + "/invalid"][0])


@pytest.mark.parametrize("kernel", ["tile", "rows"])
@pytest.mark.parametrize("bd,fir,fr,al", [(12, 1, 0, 0), (10, 0, 0, 1), (14, 1, 1, 0)])
def test_ybar_mode_matches_oracle(ctx, kernel, bd, fir, fr, al, opt):
    # yuv2tiff -X (yuv2tiff.cpp:365-399): Y'DzDx rebuilt around the 2x2 luma mean.  The oracle's restatement is pinned
    # against the reference binary in test_oracle_cpu.  Random luma next to mid-range chroma produces many negative
    # (invalid) pixels and values above Full-1.5, so every branch of the mode is taken.  A forced "rows" request still
    # runs the tile kernel (the only one with this mode); with a 709 matrix the flag is inert, as in the reference.
    opt("H2Y_INVERSE_KERNEL", kernel)
    w, h = 264, 70
    rng = np.random.default_rng(bd)
    top = (1 << bd) - 1
    yuvs = []
    for s in range(2):
        y = rng.integers(0, top + 1, w * h, dtype=np.uint16)
        y[:40] = 0
        c = rng.integers(top // 4, 3 * top // 4, w * h // 2, dtype=np.uint16)
        yuvs.append(np.concatenate([y, c]))
    rgb, inv = gpu_inverse(ctx, yuvs, w, h, bd, O.INV_YDZDX, fir, fr, al, ybar=1)
    for i, y in enumerate(yuvs):
        want, winv = O.yuv2tiff(y, w, h, bd, O.INV_YDZDX, bool(fir), bool(fr), bool(al), ybar=True)
        assert np.array_equal(rgb[i], want), (kernel, bd, i)
        assert int(inv[i]) == winv
    plain, _ = gpu_inverse(ctx, yuvs[:1], w, h, bd, O.INV_709, fir, fr, al, ybar=0)
    inert, _ = gpu_inverse(ctx, yuvs[:1], w, h, bd, O.INV_709, fir, fr, al, ybar=1)
    assert np.array_equal(plain, inert)


def _yuv_for(w, h, seed, bd, matrix_fwd):
    px = synth.exr_half_frame(w, h, seed=seed, channels=4, correlated=True)
    dst = dict(bit_depth=bd, full_range=0, transfer=16, primaries=9, matrix=matrix_fwd, chroma=1, resampler=1)
    return O.forward(O.load_half(px), dict(bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0), dst)


@pytest.mark.parametrize("kernel", ["tile", "rows"])
@pytest.mark.parametrize("w,h", [(8, 2), (128, 16), (136, 34), (1000, 250), (1920, 1080), (3840, 2160)])
def test_inverse_sizes_vs_oracle(ctx, w, h, kernel, opt):
    opt("H2Y_INVERSE_KERNEL", kernel)
    big = w * h > 1 << 20
    for bd, m, m_fwd in ((10, O.INV_2020, 9), (12, O.INV_YDZDX, 11), (10, O.INV_709, 1), (12, O.INV_Y100, 13)):
        if big and m not in (O.INV_2020, O.INV_YDZDX):
            continue
        yuvs = [_yuv_for(w, h, s, bd, m_fwd) for s in ((0,) if big else (0, 1))]
        rgb, inv = gpu_inverse(ctx, yuvs, w, h, bd, m, 1, 0, 0)
        for i, y in enumerate(yuvs):
            want, winv = O.yuv2tiff(y, w, h, bd, m, True, False, False)
            assert np.array_equal(rgb[i], want), (w, h, bd, m)
            assert int(inv[i]) == winv
    # random (unrealistic) chroma exercises the negative / invalid-pixel branches
    rng = np.random.default_rng(w)
    if not big:
        y = rng.integers(0, 1024, w * h * 3 // 2).astype(np.uint16)
        for m in range(5):
            for fir in (1, 0):
                rgb, inv = gpu_inverse(ctx, [y], w, h, 10, m, fir, 0, 0)
                want, winv = O.yuv2tiff(y, w, h, 10, m, bool(fir), False, False)
                assert np.array_equal(rgb[0], want), (m, fir)
                assert int(inv[0]) == winv


def _inverse_dev(ctx, d_yuv, n, w, h, bd, m, fir, fr):
    d_rgb = torch.zeros(n * w * h * 6, dtype=torch.uint8, device="cuda")
    d_inv = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx.inverse(cabi.InverseParams(w, h, bd, m, fir, fr, 0, 0), d_yuv, d_rgb, n, invalid=d_inv)
    torch.cuda.synchronize()
    return d_rgb, d_inv


@pytest.mark.parametrize("m", [O.INV_2020, O.INV_709])
@pytest.mark.parametrize("fr", [0, 1])
def test_rows_kernel_on_every_10bit_triplet(ctx, m, fr, opt):
    # The rows kernel evaluates the Y'CbCr inverse in fp32 behind guard bands (h2y_inverse.cu); the tile kernel runs the
    # reference's double arithmetic for every pixel and is pinned to the reference by the tests above.  With the box
    # upsampler the chroma of a pixel is the stored sample, so a 2048 x 2048 picture whose Cb plane counts columns and
    # whose Cr plane counts rows holds every (Cb, Cr) pair, and 256 frames with the luma code stepping through
    # (4 f + position in the 2 x 2 block + a per-block offset) mod 1024 hold every 10-bit (Y', Cb, Cr) triplet once:
    # 2^30 pixels, compared on the device.
    w = h = 2048
    dev = torch.device("cuda")
    i = torch.arange(1024, device=dev, dtype=torch.int32)
    cb = i.view(1, 1024).expand(1024, 1024).contiguous().view(-1)
    cr = i.view(1024, 1).expand(1024, 1024).contiguous().view(-1)
    x = torch.arange(w, device=dev, dtype=torch.int32).view(1, w)
    y = torch.arange(h, device=dev, dtype=torch.int32).view(h, 1)
    base = ((y & 1) * 2 + (x & 1) + 37 * (x >> 1) + 101 * (y >> 1))             # (h, w)
    batch = 32
    total_invalid = 0
    for f0 in range(0, 256, batch):
        frames = []
        for f in range(f0, f0 + batch):
            luma = ((base + 4 * f) & 1023).view(-1)
            frames.append(torch.cat([luma, cb, cr]).to(torch.int16))
        d_yuv = torch.stack(frames, 0).view(torch.uint8).view(-1)
        res = {}
        # "exact": the rows kernel with every pixel through its exact routine, the integer forms (inv_pixel_int10 for
        # BT.2020, inv_pixel_int with its rounding thresholds for BT.709): all 2^30 triplets reach them
        for kernel in ("tile", "rows", "exact"):
            opt("H2Y_INVERSE_KERNEL", kernel)
            res[kernel] = _inverse_dev(ctx, d_yuv, batch, w, h, 10, m, 0, fr)
        for kernel in ("rows", "exact"):
            assert torch.equal(res["tile"][0], res[kernel][0]), (kernel, m, fr, f0)
            assert torch.equal(res["tile"][1], res[kernel][1]), (kernel, m, fr, f0)
        total_invalid += int(res["tile"][1].sum().item())
        del res, d_yuv, frames
    assert total_invalid > 0                        # the lattice includes the out-of-gamut corners


@pytest.mark.parametrize("bd,m", [(10, O.INV_2020), (10, O.INV_709), (12, O.INV_2020), (12, O.INV_709), (14, O.INV_2020), (14, O.INV_709)])
def test_rows_kernel_dense_random_triplets(ctx, bd, m, opt):
    # The same comparison with the FIR upsampler (for 10-bit BT.2020 video range that is the instantiation with the
    # constants compiled in) and at the depths where the lattice is too large: 2^27 random pixels per case.
    w = h = 2048
    n = 32
    g = torch.Generator(device="cuda")
    g.manual_seed(bd * 10 + m)
    top = 1 << bd
    d_yuv = torch.randint(0, top, (n, w * h * 3 // 2), device="cuda", generator=g, dtype=torch.int32).to(torch.int16)
    # half of the frames: smooth chroma, so that the upsampled values stay near the stored ones (in-gamut pixels)
    d_yuv[n // 2:, w * h:] = (top // 2 + torch.randint(-top // 8, top // 8, (n // 2, w * h // 2), device="cuda", generator=g, dtype=torch.int32)).to(torch.int16)
    d_yuv = d_yuv.view(torch.uint8).view(-1)
    res = {}
    # "exact": every pixel through the rows kernel's exact routine, i.e. the integer forms (inv_pixel_int10 / inv_pixel_int,
    # with the float-rounding thresholds that only matter above 10 bits and for BT.709)
    for kernel in ("tile", "rows", "exact"):
        opt("H2Y_INVERSE_KERNEL", kernel)
        res[kernel] = _inverse_dev(ctx, d_yuv, n, w, h, bd, m, 1, 0)
    for kernel in ("rows", "exact"):
        assert torch.equal(res["tile"][0], res[kernel][0]), kernel
        assert torch.equal(res["tile"][1], res[kernel][1]), kernel


def test_staged_upsample_matches_oracle(ctx):
    rng = np.random.default_rng(4)
    for wh, hh, top in ((64, 24, 1023), (101, 37, 4095), (960, 540, 16383)):
        p = rng.integers(0, top + 1, (hh, wh)).astype(np.uint16)
        for fir in (0, 1):
            d_src = torch.from_numpy(p.view(np.int16)).cuda()
            d_dst = torch.zeros(4 * wh * hh, dtype=torch.int16, device="cuda")
            ctx.subsample_420_to_444(d_src, d_dst, 2 * wh, 2 * hh, fir, 0, top)
            torch.cuda.synchronize()
            got = d_dst.cpu().numpy().view(np.uint16).reshape(2 * hh, 2 * wh)
            assert np.array_equal(got, O.upsample_420to444(p, fir, 0, top))


def test_round_trip_is_close(ctx):
    # forward then inverse of a smooth 16-bit frame comes back within FIR/quantisation error:
    # a size-independent sanity property (the exact check is the oracle comparison above)
    w, h = 1920, 1080
    f = synth.tiff16_frame(w, h, seed=2, smooth=True)
    src = dict(kind="tiff16", bit_depth=16, full_range=0, transfer=16, primaries=10, matrix=0)
    dst = dict(bit_depth=12, full_range=0, transfer=16, primaries=9, matrix=11, chroma=1, resampler=1)
    yuv = G.gpu_forward(ctx, [f], src, dst)[0]
    rgb, _ = gpu_inverse(ctx, [yuv], w, h, 12, O.INV_YDZDX, 1, 0, 0)
    want = np.clip(f.astype(np.int64), 4096, 60160)
    err = np.abs(rgb[0].astype(np.int64)[8:-8, 8:-8] - want[8:-8, 8:-8])
    assert int(err[..., 1].max()) <= 15                    # Y' = G' passes through: exact to the 12-bit truncation
    assert np.median(err[..., 0]) <= 1024 and np.median(err[..., 2]) <= 1024   # chroma lost only the per-pixel noise


def test_inverse_host_pipeline(ctx):
    w, h, n = 1000, 250, 7
    yuvs = [_yuv_for(w, h, s, 10, 9) for s in range(n)]
    dev, dinv = gpu_inverse(ctx, yuvs, w, h, 10, O.INV_2020, 1, 0, 0)
    p = cabi.InverseParams(w, h, 10, O.INV_2020, 1, 0, 0)
    src = np.ascontiguousarray(np.stack(yuvs, 0))
    out = np.zeros((n, h, w, 3), np.uint16)
    inv = np.zeros(n, np.uint32)
    ctx.inverse_host(p, src, out, n, invalid=inv)
    assert np.array_equal(out, dev)
    assert np.array_equal(inv.astype(np.int64), dinv.astype(np.int64))


# ---- hdr2yuv's own 4:4:4 inverse: matrix_inverse + write_tiff (convert.cpp:1320-1867, tiff.cpp:559-652) ----
@pytest.mark.parametrize("case", cases.MINV_CASES, ids=lambda c: "m%d_i%d_f%d_o%d" % c)
def test_matrix_inverse_matches_oracle_and_golden(ctx, golden_minv, case):
    m, ibd, fr, obd = case
    pl = cases.minv_input(ibd)
    want, invalid = O.matrix_inverse(pl, m, ibd, fr, obd, backend="port")
    assert np.array_equal(want, golden_minv["m%d_i%d_f%d_o%d" % case])
    h, w = pl.shape[1:]
    d_in = torch.from_numpy(pl.view(np.int16).reshape(-1)).cuda()
    d_out = torch.zeros(3 * h * w, dtype=torch.int16, device="cuda")
    d_inv = torch.zeros(1, dtype=torch.int32, device="cuda")
    in_pic = api.pic_desc(w, h, 3, 0, 0, m, ibd, fr)
    out_pic = api.pic_desc(w, h, 3, 0, 0, 0, obd, fr)
    n = h * w
    ctx.matrix_inverse(out_pic, [d_out[i * n:(i + 1) * n] for i in range(3)], in_pic, [d_in[i * n:(i + 1) * n] for i in range(3)], invalid=d_inv)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(np.uint16).reshape(3, h, w)
    assert np.array_equal(got, want)
    assert int(d_inv.item()) == invalid
    if obd >= ibd:
        # write_tiff on top, and the fused host call that does both like main()
        d_rgb = torch.zeros(3 * n, dtype=torch.int16, device="cuda")
        tmp, _ = O.matrix_inverse(pl, m, ibd, fr, ibd, backend="port")          # tmp picture keeps the source depth
        d_tmp = torch.from_numpy(tmp.view(np.int16).reshape(-1)).cuda()
        ctx.write_tiff_rows(api.pic_desc(w, h, 3, 0, 0, 0, obd, fr), [d_tmp[i * n:(i + 1) * n] for i in range(3)], ibd, d_rgb)
        torch.cuda.synchronize()
        want_rgb = O.write_tiff_rows(tmp, obd, ibd)
        assert np.array_equal(d_rgb.cpu().numpy().view(np.uint16).reshape(h, w, 3), want_rgb)
        frames = np.ascontiguousarray(np.stack([pl, pl[:, ::-1].copy()], 0))
        out = np.zeros((2, h, w, 3), np.uint16)
        inv = np.zeros(2, np.uint32)
        ctx.inverse444_host(in_pic, obd, frames, out, 2, invalid=inv)
        assert np.array_equal(out[0], want_rgb)
        t2, i2 = O.matrix_inverse(frames[1], m, ibd, fr, ibd, backend="port")
        assert np.array_equal(out[1], O.write_tiff_rows(t2, obd, ibd)) and int(inv[1]) == i2 and int(inv[0]) == invalid


def test_matrix_inverse_errors(ctx):
    w, h = 16, 8
    d = torch.zeros(3 * w * h, dtype=torch.int16, device="cuda")
    n = w * h
    pl = [d[i * n:(i + 1) * n] for i in range(3)]
    with pytest.raises(cabi.H2YError) as e:              # "Can't determine color difference to use?" (convert.cpp:1735)
        ctx.matrix_inverse(api.pic_desc(w, h, 3, 0, 0, 0, 12, 0), pl, api.pic_desc(w, h, 3, 0, 0, 0, 12, 0), pl)
    assert e.value.status == cabi.ERR_MATRIX
    with pytest.raises(cabi.H2YError) as e:              # write_tiff would shift by a negative count
        ctx.write_tiff_rows(api.pic_desc(w, h, 3, 0, 0, 0, 10, 0), pl, 12, d)
    assert e.value.status == cabi.ERR_BIT_DEPTH


def test_empty_batches_are_no_ops(ctx):
    d = torch.zeros(64, dtype=torch.uint8, device="cuda")
    p = cabi.InverseParams(64, 16, 10, O.INV_2020, 1, 0, 0)
    ctx.inverse(p, d, d, 0)
    fp = api.forward_params(64, 16, cabi.LAYOUT_RGB16, dict(bit_depth=16, full_range=0, transfer=16, primaries=10, matrix=0),
                            dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1))
    ctx.forward(fp, d, d, 0)
    ctx.forward_host(fp, np.zeros(1, np.uint8), np.zeros(1, np.uint8), 0)
    ctx.inverse_host(p, np.zeros(1, np.uint8), np.zeros(1, np.uint8), 0)
    torch.cuda.synchronize()


@pytest.mark.parametrize("kernel", ["tile", "rows"])
def test_inverse_wide_and_short_pictures(ctx, kernel, opt):
    opt("H2Y_INVERSE_KERNEL", kernel)
    for (w, h) in ((7680, 36), (8, 2), (8, 130), (3848, 4)):
        yuvs = [_yuv_for(w, h, 300 + s, 10, 9) for s in range(2)]
        rgb, inv = gpu_inverse(ctx, yuvs, w, h, 10, O.INV_2020, 1, 0, 0)
        for i in range(2):
            want, invalid = O.yuv2tiff(yuvs[i], w, h, 10, O.INV_2020, True, False, False)
            assert np.array_equal(rgb[i], want), (kernel, w, h)
            assert int(inv[i]) == invalid


def test_linear_light_stage_behind_the_inverse(ctx, golden_inverse):
    # SURVEY.md 8a note N2: yuv2tiff stops at PQ-coded 16-bit integers; the optional linear stage is the reference's own
    # PQ10000_f (convert.cpp:43-51) on code / 65535 (the normalisation of convert.cpp:1017-1019 with floor 0, ceiling 65535).
    # Tolerance of the path for float output: 1e-5 relative.  Every 16-bit code is checked, then an inverse frame's rows.
    codes = np.arange(65536, dtype=np.uint16)
    d_codes = torch.from_numpy(codes.view(np.int16)).cuda()
    d_lin = torch.zeros(65536, dtype=torch.float32, device="cuda")
    ctx.pq_codes_to_linear(d_codes, d_lin)
    torch.cuda.synchronize()
    got = d_lin.cpu().numpy()
    # which = 0: PQ10000_f; the compiled reference itself when it is there (one C call for all codes), else the restatement
    want = O.transfer(0, codes.astype(np.float32) / np.float32(65535.0), "ref" if O.ref_available() else "port")
    rel = np.abs(got.astype(np.float64) - want.astype(np.float64)) / np.maximum(np.abs(want.astype(np.float64)), 1e-30)
    assert float(rel.max()) <= 1e-5, float(rel.max())
    assert got[0] == 0.0 and abs(float(got[65535]) - 1.0) <= 1e-6
    yuv = cases.widen_yuv(golden_inverse[cases.inverse_input_key(2)], 10)
    rgb, _ = gpu_inverse(ctx, [yuv], cases.IW, cases.IH, 10, 2, 1, 0, 0)
    d_rgb = torch.from_numpy(np.ascontiguousarray(rgb[0]).view(np.int16).reshape(-1)).cuda()
    d_out = torch.zeros(d_rgb.numel(), dtype=torch.float32, device="cuda")
    ctx.pq_codes_to_linear(d_rgb, d_out)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), got[rgb[0].reshape(-1)])
