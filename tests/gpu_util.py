"""Helpers for the -m gpu parity tests: everything goes through the C-ABI (hdr2yuv_b200.api)."""
import numpy as np
import torch

import cases
from hdr2yuv_b200 import _cabi as cabi
from hdr2yuv_b200 import api
from oracle import oracle as O


def layout_of(src_kind, channels):
    if src_kind == "tiff16":
        return cabi.LAYOUT_RGB16 if channels == 3 else cabi.LAYOUT_RGBA16
    return cabi.LAYOUT_HALF_RGB if channels == 3 else cabi.LAYOUT_HALF_RGBA


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()


def gpu_forward(ctx, frames, src, dst, layout=None):
    """frames: list of (H,W,C) uint16 arrays (or one (3,H,W) planar array list).  Returns list of yuv arrays."""
    h, w, ch = frames[0].shape
    layout = layout_of(src["kind"], ch) if layout is None else layout
    params = api.forward_params(w, h, layout, src, dst, resampler=dst["resampler"],
                                clip_on_load=1 if src["kind"] == "tiff16" else 0)
    d_src = to_dev(np.stack(frames, 0))
    nbytes = api.yuv_frame_bytes(w, h, dst["chroma"])
    d_dst = torch.zeros(nbytes * len(frames), dtype=torch.uint8, device="cuda")
    ctx.forward(params, d_src, d_dst, len(frames))
    torch.cuda.synchronize()
    out = d_dst.cpu().numpy().view(np.uint16).reshape(len(frames), -1)
    return [out[i] for i in range(len(frames))]


def oracle_forward(px, src, dst):
    planes = O.load_rgb16(px, src["full_range"]) if src["kind"] == "tiff16" else O.load_half(px)
    return O.forward(planes, cases.oracle_src(src), dst, backend="port")


def compare_codes(got, want, uses_transfer, what=""):
    """Bit-exact unless the transfer LUT (CUDA pow vs glibc pow) is involved: then <= 1 code,
    deviations counted (north_star tolerance)."""
    got = np.asarray(got).astype(np.int64)
    want = np.asarray(want).astype(np.int64)
    assert got.shape == want.shape, (got.shape, want.shape)
    d = np.abs(got - want)
    nbad = int((d != 0).sum())
    if not uses_transfer:
        assert nbad == 0, "%s: %d of %d samples differ (max %d)" % (what, nbad, d.size, int(d.max()))
    else:
        assert int(d.max()) <= 1, "%s: max deviation %d codes" % (what, int(d.max()))
        assert nbad <= max(2, d.size // 2000), "%s: %d of %d samples deviate by 1 code" % (what, nbad, d.size)
    return nbad
