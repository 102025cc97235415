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


def gpu_inverse(ctx, yuvs, w, h, bd, m, fir, fr, al):
    p = cabi.InverseParams(w, h, bd, m, fir, fr, al)
    n = len(yuvs)
    d_yuv = G.to_dev(np.stack(yuvs, 0))
    nch = 4 if al else 3
    d_rgb = torch.zeros(n * w * h * nch * 2, dtype=torch.uint8, device="cuda")
    d_inv = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx.inverse(p, d_yuv, d_rgb, n, invalid=d_inv)
    torch.cuda.synchronize()
    rgb = d_rgb.cpu().numpy().view(np.uint16).reshape(n, h, w, nch)
    return rgb, d_inv.cpu().numpy()


@pytest.mark.parametrize("case", cases.INVERSE_CASES, ids=lambda c: "b%d_m%d_fir%d_fr%d_a%d" % c)
def test_inverse_matches_golden(ctx, golden_inverse, case):
    bd, m, fir, fr, al = case
    yuv = cases.widen_yuv(golden_inverse[cases.inverse_input_key(m)], bd)
    rgb, inv = gpu_inverse(ctx, [yuv], cases.IW, cases.IH, bd, m, fir, fr, al)
    key = "b%d_m%d_fir%d_fr%d_a%d" % case
    assert np.array_equal(rgb[0][:2], golden_inverse[key + "/head"])
    assert hashlib.sha256(rgb[0].tobytes()).digest() == golden_inverse[key + "/sha256"].tobytes()
    assert int(inv[0]) == int(golden_inverse[key + "/invalid"][0])


def _yuv_for(w, h, seed, bd, matrix_fwd):
    px = synth.exr_half_frame(w, h, seed=seed, channels=4, correlated=True)
    dst = dict(bit_depth=bd, full_range=0, transfer=16, primaries=9, matrix=matrix_fwd, chroma=1, resampler=1)
    return O.forward(O.load_half(px), dict(bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0), dst)


@pytest.mark.parametrize("w,h", [(8, 2), (128, 16), (136, 34), (1000, 250), (1920, 1080), (3840, 2160)])
def test_inverse_sizes_vs_oracle(ctx, w, h):
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
