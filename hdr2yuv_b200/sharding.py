"""Frame-range sharding across GPUs.

Every frame is independent (no temporal state anywhere in the reference: n_frames is unused,
hdr2yuv.cpp:157-160; pic_stats is per frame, common.cpp:66), so N GPUs take contiguous frame
ranges and write at byte offset frame * frame_bytes of the .yuv -- the same bytes the reference's
sequential append produces (tiff.cpp:440).  No collective sits on the data path; the helpers
below are bookkeeping for timing only.
"""


def frame_range(rank, world, nframes):
    """Contiguous range [lo, hi) of rank `rank`; sizes differ by at most one frame.  The formula lives in the C-ABI
    (h2y_frame_range), which the CLI's --devices workers use as well."""
    import ctypes as C
    from ._cabi import check, lib
    lo, hi = C.c_int(0), C.c_int(0)
    check(lib().h2y_frame_range(int(rank), int(world), int(nframes), C.byref(lo), C.byref(hi)), "h2y_frame_range")
    return lo.value, hi.value


def output_offset(frame, frame_bytes):
    return frame * frame_bytes


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def max_over_ranks(value, device=None):
    d = _dist()
    if d is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    d = _dist()
    if d is None:
        return value
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return type(value)(t.item())
