"""Parity cases shared by the golden-vector generator and the tests."""
import numpy as np

from hdr2yuv_b200 import synth

GW, GH = 64, 48          # geometry of the committed golden vectors (small on purpose)

# (name, src dict, dst dict): src.kind is "tiff16" (interleaved RGB16 through read_tiff's clip) or "half"
FORWARD_CASES = []


def _add(name, kind, src, dst):
    FORWARD_CASES.append((name, dict(kind=kind, **src), dst))


_TIFF = dict(bit_depth=16, full_range=0, transfer=16, primaries=10, matrix=0)
for m in (9, 1, 11, 12, 13):
    for chroma, res in ((1, 1), (1, 0), (3, 1)):
        for bd in (10, 12, 16):
            if (m in (1, 12, 13) and bd != 10) or (chroma == 3 and bd == 12):
                continue
            _add(f"tiff_m{m}_c{chroma}_r{res}_b{bd}", "tiff16", _TIFF,
                 dict(bit_depth=bd, full_range=0, transfer=16, primaries=9, matrix=m, chroma=chroma, resampler=res))
_add("tiff_full_m9", "tiff16", dict(_TIFF, full_range=1),
     dict(bit_depth=10, full_range=1, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1))
_add("tiff_pass_gbr", "tiff16", _TIFF,
     dict(bit_depth=10, full_range=0, transfer=16, primaries=10, matrix=0, chroma=3, resampler=1))
_add("tiff_422_m9", "tiff16", _TIFF,
     dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=2, resampler=1))
for tr in (1, 8, 18):
    _add(f"tiff_transfer16to{tr}", "tiff16", dict(_TIFF, primaries=9),
         dict(bit_depth=10, full_range=0, transfer=tr, primaries=9, matrix=9, chroma=1, resampler=1))
# Y'u''v'' 4:2:0 (convert.cpp:533-801): rho-gamma coded 16-bit source, libm pow on the luma plane
_add("tiff_prime2_fir_b16", "tiff16", dict(_TIFF, transfer=18, full_range=1),
     dict(bit_depth=16, full_range=1, transfer=18, primaries=10, matrix=15, chroma=1, resampler=1))
_add("tiff_prime2_box_b12", "tiff16", dict(_TIFF, transfer=18),
     dict(bit_depth=12, full_range=0, transfer=18, primaries=10, matrix=15, chroma=1, resampler=0))
_HALF = dict(bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0)
for m in (9, 11, 1, 13):
    for bd in (10, 12):
        for fr in (0, 1):
            if m in (1, 13) and (bd == 12 or fr == 1):
                continue
            _add(f"half_pq_m{m}_b{bd}_f{fr}", "half", _HALF,
                 dict(bit_depth=bd, full_range=fr, transfer=16, primaries=9, matrix=m, chroma=1, resampler=1))
_add("half_gamma_m9", "half", _HALF, dict(bit_depth=10, full_range=0, transfer=1, primaries=9, matrix=9, chroma=1, resampler=1))
_add("half_rho_m9", "half", _HALF, dict(bit_depth=10, full_range=0, transfer=18, primaries=9, matrix=9, chroma=1, resampler=1))
_add("half_pq_444_box", "half", _HALF, dict(bit_depth=12, full_range=0, transfer=16, primaries=9, matrix=9, chroma=3, resampler=0))
_add("half_pq_box", "half", _HALF, dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=0))
_add("half_same_transfer", "half", dict(_HALF, transfer=16), dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1))
_add("half_pqsrc_to_linear", "half", dict(_HALF, transfer=16), dict(bit_depth=10, full_range=0, transfer=8, primaries=9, matrix=11, chroma=1, resampler=1))


def forward_input(case_src, w, h, seed=7):
    """Returns (interleaved array as handed to the C-ABI, reference-order planes for the oracle)."""
    if case_src["kind"] == "tiff16":
        px = synth.tiff16_frame(w, h, seed=seed, channels=3)
        return px, None
    hi = 1.7 if case_src.get("transfer") == 16 else 4000.0     # PQ-coded sources live in [0, ~1]
    px = synth.exr_half_frame(w, h, seed=seed, channels=4, hi=hi)
    return px, None


def oracle_src(case_src):
    return {k: case_src[k] for k in ("bit_depth", "full_range", "transfer", "primaries", "matrix")}


# inverse (yuv2tiff) cases: (bit_depth, matrix, fir, full_range, alpha)
INVERSE_CASES = [(bd, m, fir, fr, al)
                 for bd in (10, 12, 14) for m in range(5) for fir in (1, 0) for fr in (0, 1) for al in (0, 1)
                 if not (bd != 10 and (fir == 0 or fr == 1 or al == 1)) and not (m in (3, 4) and (fr or al))]
IW, IH = 960, 540        # smallest geometry the reference program knows (yuv2tiff.cpp:188-198)


def widen_yuv(yuv10, bit_depth):
    """Deterministically extend a 10-bit frame to 12/14 bits with non-zero low bits."""
    if bit_depth == 10:
        return yuv10
    sh = bit_depth - 10
    idx = np.arange(yuv10.size, dtype=np.uint64)
    low = ((idx * np.uint64(2654435761)) >> np.uint64(13)) & np.uint64((1 << sh) - 1)
    return ((yuv10.astype(np.uint32) << sh) | low.astype(np.uint32)).astype(np.uint16)


def inverse_input_key(matrix):
    return "yuv10_ydzdx" if matrix == 0 else "yuv10_ycbcr"


# matrix_inverse (hdr2yuv .yuv 4:4:4 -> .tiff) cases: (matrix_coeffs, in_bit_depth, in_full_range, out_bit_depth)
MINV_CASES = [(m, ibd, fr, obd) for m in (1, 9, 11, 10, 12) for (ibd, fr, obd) in ((12, 0, 12), (12, 0, 16), (10, 0, 16), (10, 1, 10), (14, 0, 12))]
MW, MH = 96, 40


def minv_input(in_bit_depth, seed=21):
    """(3,H,W) u16 Y,Cb,Cr: a plausible picture (luma ramp + noise, chroma around mid-grey) plus extreme codes that
    drive components negative (invalidPixels) and above Full-1."""
    rng = np.random.default_rng(seed + in_bit_depth)
    top = (1 << in_bit_depth) - 1
    y = np.clip(np.linspace(0, top, MW)[None, :] + rng.integers(-40, 41, (MH, MW)), 0, top)
    c = np.clip((top + 1) // 2 + rng.integers(-(top // 6), top // 6 + 1, (2, MH, MW)), 0, top)
    pl = np.concatenate([y[None], c], 0).astype(np.uint16)
    spr = rng.random(pl.shape) < 0.03
    pl[spr] = rng.choice(np.array([0, 1, top // 2, top - 1, top], np.uint16), int(spr.sum()))
    return np.ascontiguousarray(pl)
