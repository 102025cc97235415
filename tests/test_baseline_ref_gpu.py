"""-m gpu: the BASELINE.json 4K configurations at FULL size against the compiled REFERENCE itself (oracle backend "ref" =
the reference's convert.cpp / common.cpp / tiff.cpp and its yuv2tiff program, built unmodified into oracle/_ref), not the
C restatement.  The chain restatement == reference is pinned on the CPU at small sizes (tests/test_oracle_cpu.py); these
tests close the gap at 3840x2160, one frame per configuration (a few seconds of CPU each).  oracle/_ref travels to the
GPU box prebuilt; without it the tests skip."""
import numpy as np
import pytest
import torch

import cases
import gpu_util as G
from hdr2yuv_b200 import _cabi as cabi
from hdr2yuv_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not O.ref_available(), reason="compiled reference (oracle/_ref) not available")

_HALF = dict(kind="half", bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0)
W, H = 3840, 2160


@needs_ref
@pytest.mark.parametrize("depth", [10, 12], ids=["configs1_pq10", "configs4_pq12"])
def test_exr_4k_forward_against_the_compiled_reference(ctx, opt, depth):
    # configs[1] / configs[4]: 3840x2160 half RGB (linear, BT.709) -> PQ 10- / 12-bit BT.2020nc 4:2:0 FIR.  Ten frames so
    # that h2y_forward takes the rows kernel (the bench's route), two-pass and then single-pass with plan reuse.
    dst = dict(bit_depth=depth, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    base = synth.exr_half_frame_fast(W, H, seed=40 + depth, channels=3)
    want = O.forward(O.load_half(base), cases.oracle_src(_HALF), dst, backend="ref")
    frames = [base] * 10
    for reuse in ("0", "1"):
        opt("H2Y_PLAN_REUSE", reuse)
        got = G.gpu_forward(ctx, frames, _HALF, dst)
        if reuse == "1":
            got = G.gpu_forward(ctx, frames, _HALF, dst)       # the first call with plan reuse switched on only leaves the seed
            attempted, nframes, nredone = ctx.forward_last_plan_reuse()
            assert attempted and nredone == 0
        nbad = G.compare_codes(got[0], want, True, "4K half PQ%d vs compiled reference (plan reuse %s)" % (depth, reuse))
        print("4K PQ%d, plan reuse %s: %d of %d samples deviate by one code from the compiled reference" % (depth, reuse, nbad, want.size))
        for i in range(1, 10):
            assert np.array_equal(got[i], got[0]), i


@needs_ref
def test_inverse_4k_against_the_reference_program(ctx):
    # configs[3]: 3840x2160 10-bit BT.2020 4:2:0 .yuv -> FIR upsample -> 16-bit RGB rows, against the reference's yuv2tiff
    # program run file to file.  The input is a forward conversion of synthetic content (what an encoder would be fed).
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    base = synth.exr_half_frame_fast(W, H, seed=77, channels=3)
    yuv = G.gpu_forward(ctx, [base], _HALF, dst)[0]
    want, invalid = O.yuv2tiff(yuv, W, H, 10, O.INV_2020, True, False, False, backend="ref")
    n = 10                                              # enough rows per worker for the rows kernel
    p = cabi.InverseParams(W, H, 10, O.INV_2020, 1, 0, 0, 0)
    d_yuv = G.to_dev(np.stack([yuv] * n, 0))
    d_rgb = torch.zeros(n * W * H * 3 * 2, dtype=torch.uint8, device="cuda")
    d_inv = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx.inverse(p, d_yuv, d_rgb, n, invalid=d_inv)
    torch.cuda.synchronize()
    rgb = d_rgb.cpu().numpy().view(np.uint16).reshape(n, H, W, 3)
    for i in range(n):
        assert np.array_equal(rgb[i], want), i
    assert all(int(v) == invalid for v in d_inv.cpu().numpy())
