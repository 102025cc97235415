"""hdr2yuv_b200 -- B200-native (sm_100a) implementation of hdr2yuv's per-pixel conversion hot path.

The product is the C-ABI shared library built from csrc/ (include/hdr2yuv_b200.h); this package
is the thin host-side mirror of the reference's function boundary on top of it.  There is no
CPU fallback: importing `hdr2yuv_b200.api` without the built library raises.
"""
__version__ = "0.1.0"
