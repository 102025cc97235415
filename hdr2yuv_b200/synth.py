"""Synthetic frames for tests and bench.py: seeded, numpy only.

There is no network and the reference ships no clips, so every workload is synthetic:
  tiff16_frame  -- what a 16-bit X'Y'Z' PQ TIFF strip row looks like to read_tiff: interleaved
                   R,G,B u16, uniform over the video range plus a 1 % sprinkle of out-of-range codes
                   that exercise the on-read clip (tiff.cpp:296-304)
  exr_half_frame -- linear-light half-float RGB(A) "in nits": log-uniform in [0.005, 4000], exact
                   zeros present, all >= 0 and frame max >= 1 (else the reference divides by a zero
                   range, common.cpp:135-136 / convert.cpp:939)
"""
import numpy as np


def tiff16_frame(width, height, seed=1, channels=3, smooth=False):
    rng = np.random.default_rng(seed)
    if smooth:
        y, x = np.mgrid[0:height, 0:width]
        base = 4096 + (56064 * (0.5 + 0.5 * np.sin(x / 97.0) * np.cos(y / 61.0)))
        px = np.stack([base, np.roll(base, 13, 1), np.roll(base, 29, 0)], -1)
        px = px + rng.integers(-64, 65, px.shape)
        px = np.clip(px, 0, 65535).astype(np.uint16)
    else:
        px = rng.integers(4096, 60161, (height, width, 3), dtype=np.uint16)
    spr = rng.random((height, width, 3)) < 0.01
    px[spr] = rng.choice(np.array([0, 4095, 60161, 65535], np.uint16), int(spr.sum()))
    if channels == 4:
        a = np.full((height, width, 1), 65535, np.uint16)
        px = np.concatenate([px, a], -1)
    return np.ascontiguousarray(px)


def exr_half_frame(width, height, seed=0, channels=4, lo=0.005, hi=4000.0, correlated=False):
    """Returns (H, W, channels) uint16 half bit patterns, r,g,b(,a) order."""
    rng = np.random.default_rng(seed)
    if correlated:
        base = np.exp(rng.uniform(np.log(lo), np.log(hi), (height, width, 1)))
        v = base * rng.uniform(0.7, 1.3, (height, width, 3))
        v = np.clip(v, 0.0, hi)
    else:
        v = np.exp(rng.uniform(np.log(lo), np.log(hi), (height, width, 3)))
    h = v.astype(np.float16)
    flat = h.reshape(-1)
    flat[rng.integers(0, flat.size, max(3, flat.size // 1000))] = np.float16(0.0)
    flat[rng.integers(0, flat.size)] = np.float16(hi)       # pin the frame maximum
    if channels == 4:
        a = np.full((height, width, 1), np.float16(1.0))
        h = np.concatenate([h, a], -1)
    return np.ascontiguousarray(h).view(np.uint16)


def planes_from_interleaved(px):
    """(H,W,C) r,g,b(,a) -> (3,H,W) in the reference's G,B,R plane order (tiff.cpp:309-311)."""
    return np.ascontiguousarray(np.stack([px[..., 1], px[..., 2], px[..., 0]], 0))


_HALF_LO, _HALF_HI = 0x1D1F, 0x6BD0        # bit patterns of half(0.005) and half(4000.0)


def exr_half_frame_fast(width, height, seed=0, channels=3):
    """bench.py's generator for the same workload as exr_half_frame, ~10x faster: half bit patterns
    drawn uniformly between half(0.005) and half(4000) are log-uniform to within one binade's
    linear spacing.  Exact zeros are sprinkled in and the frame maximum is pinned to 4000, so every
    value is >= 0 and (int)max - (int)min > 0 (common.cpp:135-136, convert.cpp:939).
    Returns (H, W, channels) uint16 half bit patterns, r,g,b(,a)."""
    n = width * height * 3
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 1 << 64, (n + 3) // 4, dtype=np.uint64).view(np.uint16)[:n].astype(np.uint32)
    raw *= np.uint32(_HALF_HI - _HALF_LO + 1)
    raw >>= np.uint32(16)
    raw += np.uint32(_HALF_LO)
    bits = raw.astype(np.uint16)
    bits[rng.integers(0, n, max(3, n // 1000))] = 0
    bits[int(rng.integers(0, n))] = _HALF_HI
    px = bits.reshape(height, width, 3)
    if channels == 4:
        out = np.empty((height, width, 4), np.uint16)
        out[..., :3] = px
        out[..., 3] = 0x3C00                    # alpha = 1.0
        return out
    return px


def exr_half_frame_smooth_fast(width, height, seed=0, block=16, peak_white=False):
    """Spatially correlated linear-light frame (what decoded video looks like to the inverse path): a log-uniform
    luminance field at 1/block resolution, bilinearly smooth through a separable box blur, with a mild colour cast
    per channel and 2 % pixel noise.  (H, W, 3) uint16 half bit patterns, all >= 0, max pinned to 4000.
    peak_white: one peak-white (4000, 4000, 4000) and one near-black pixel per frame, as in a graded master: the three
    channels then share pic_stats' floor and ceiling (common.cpp:135-136) and the frame takes the single-table kernels."""
    rng = np.random.default_rng(seed)
    bh, bw = height // block + 2, width // block + 2
    low = np.exp(rng.uniform(np.log(0.05), np.log(2000.0), (bh, bw))).astype(np.float32)
    lum = np.repeat(np.repeat(low, block, 0), block, 1)
    k = block
    c = np.cumsum(lum, 0); lum = (c[k:] - c[:-k]) / k
    c = np.cumsum(lum, 1); lum = (c[:, k:] - c[:, :-k]) / k
    lum = lum[:height, :width]
    cast = rng.uniform(0.6, 1.4, (bh, bw, 3)).astype(np.float32)
    cast = np.repeat(np.repeat(cast, block, 0), block, 1)[:height, :width]
    noise = rng.uniform(0.98, 1.02, (height, width, 1)).astype(np.float32)
    v = (lum[..., None] * cast * noise).astype(np.float16)
    flat = v.reshape(-1)
    flat[int(rng.integers(0, flat.size))] = np.float16(4000.0)
    if peak_white:
        v[int(rng.integers(0, height)), int(rng.integers(0, width))] = np.float16(4000.0)
        v[int(rng.integers(0, height)), int(rng.integers(0, width))] = np.float16(0.01)
    return np.ascontiguousarray(v).view(np.uint16)


def exr_half_frame_flat(width, height, seed=0):
    """Diagnostic content: 64 x 64 tiles of constant colour (so the 32 lanes of a warp, 8 pixels apart, gather at most a
    handful of distinct table entries and shared memory serves them by broadcast), plus one peak-white and one near-black
    pixel so that the channels share floor and ceiling.  (H, W, 3) uint16 half bit patterns."""
    rng = np.random.default_rng(seed)
    t = 64
    low = np.exp(rng.uniform(np.log(0.05), np.log(2000.0), (height // t + 1, width // t + 1, 3))).astype(np.float16)
    v = np.repeat(np.repeat(low, t, 0), t, 1)[:height, :width].copy()
    v[int(rng.integers(0, height)), int(rng.integers(0, width))] = np.float16(4000.0)
    v[int(rng.integers(0, height)), int(rng.integers(0, width))] = np.float16(0.01)
    return np.ascontiguousarray(v).view(np.uint16)
