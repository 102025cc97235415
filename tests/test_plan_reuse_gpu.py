"""-m gpu tests of the single-pass forward route (plan reuse, DESIGN.md 4 "K1s"): the rows kernel converts a call's frames
with the plan of the previous call's last frame while it gathers their extrema; the verify step hands every frame whose
own plan differs back to the classic kernels of the same call.  Whatever the prediction, the output must be what the
classic route gives (which the other tests pin to the oracle) -- and the oracle itself is consulted directly here too."""
import numpy as np
import pytest

import gpu_util as G
from hdr2yuv_b200 import synth

pytestmark = pytest.mark.gpu

_HALF = dict(kind="half", bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0)


def _dst(depth=10, transfer=16):
    return dict(bit_depth=depth, full_range=0, transfer=transfer, primaries=9, matrix=9, chroma=1, resampler=1)


def _frame(w, h, seed, hi=4000.0, channels=3, lo=0.005):
    """A single-table frame: every channel contains an exact 0 and the value `hi`, so (int)min = 0 and (int)max = (int)hi."""
    f = synth.exr_half_frame(w, h, seed=seed, channels=channels, lo=lo, hi=hi)
    f[0, 0, :3] = np.float16(hi).view(np.uint16)
    f[0, 1, :3] = 0
    return f


def _classic(ctx, opt, frames, dst):
    opt("H2Y_PLAN_REUSE", "0")
    out = G.gpu_forward(ctx, frames, _HALF, dst)
    assert ctx.forward_last_plan_reuse()[0] is False
    return out


def _seed(ctx, opt, frame, dst):
    """A classic call whose last frame becomes the seed of the next call."""
    opt("H2Y_PLAN_REUSE", "1")
    G.gpu_forward(ctx, [frame] * 2, _HALF, dst)


def _reuse(ctx, opt, frames, dst, expect_redone):
    opt("H2Y_PLAN_REUSE", "1")
    out = G.gpu_forward(ctx, frames, _HALF, dst)
    attempted, n, redone = ctx.forward_last_plan_reuse()
    assert attempted and n == len(frames)
    if expect_redone is not None:
        assert redone == expect_redone, (redone, expect_redone)
    return out


@pytest.mark.parametrize("depth,channels", [(10, 3), (12, 3), (10, 4)])
def test_uniform_sequence_is_converted_in_one_pass(ctx, opt, depth, channels):
    opt("H2Y_FORWARD_KERNEL", "rows")
    w, h = 480, 130
    dst = _dst(depth)
    frames = [_frame(w, h, 300 + s, channels=channels) for s in range(6)]
    want = _classic(ctx, opt, frames, dst)
    _seed(ctx, opt, frames[0], dst)
    got = _reuse(ctx, opt, frames, dst, expect_redone=0)
    for i, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), i
    G.compare_codes(got[3], G.oracle_forward(frames[3], _HALF, dst), True, "plan reuse vs oracle")
    # the statistics the pass gathered are the frame's own (pic_stats, common.cpp:135-136)
    st = ctx.forward_last_stats(2)
    assert list(st.estimated_floor) == [0, 0, 0] and list(st.estimated_ceiling) == [4000, 4000, 4000]
    f = frames[2][..., :3].view(np.float16).astype(np.float32)
    assert np.allclose(list(st.f_max), [f[..., 1].max(), f[..., 2].max(), f[..., 0].max()])     # G, B, R


def test_frames_with_another_plan_are_converted_again(ctx, opt):
    # frames 2 and 7 have another ceiling, frame 4 another floor: exactly those are handed back (general kernel: few)
    opt("H2Y_FORWARD_KERNEL", "rows")
    w, h = 480, 130
    dst = _dst()
    frames = [_frame(w, h, 320 + s) for s in range(12)]
    frames[2] = _frame(w, h, 340, hi=900.0)
    frames[7] = _frame(w, h, 341, hi=4002.0)
    frames[4] = frames[4].copy()
    frames[4][frames[4] < 0x3E00] = 0x3E00              # every sample >= 1.5: floor 1
    want = _classic(ctx, opt, frames, dst)
    _seed(ctx, opt, frames[0], dst)
    got = _reuse(ctx, opt, frames, dst, expect_redone=3)
    for i, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), i
    for i in (2, 4, 7):
        G.compare_codes(got[i], G.oracle_forward(frames[i], _HALF, dst), True, "redone frame %d" % i)


def test_a_sequence_that_changed_goes_back_to_the_rows_kernels(ctx, opt):
    # the seed is from other content: every frame fails and the general kernel converts the whole call
    opt("H2Y_FORWARD_KERNEL", "rows")
    w, h = 480, 130
    dst = _dst()
    frames = [_frame(w, h, 360 + s, hi=1000.0) for s in range(12)]
    want = _classic(ctx, opt, frames, dst)
    _seed(ctx, opt, _frame(w, h, 399, hi=4000.0), dst)
    got = _reuse(ctx, opt, frames, dst, expect_redone=12)
    for i, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), i
    # the call left the new content's plan as the seed: the next call passes without a redo
    got = _reuse(ctx, opt, frames, dst, expect_redone=0)
    for i, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), i


def test_codes_outside_the_window_are_never_used_as_indices(ctx, opt):
    # negative zero, +inf, NaN, a value far above the predicted ceiling (beyond the two-copy LUT), three-table frames
    opt("H2Y_FORWARD_KERNEL", "rows")
    w, h = 480, 130
    dst = _dst()
    frames = [_frame(w, h, 380 + s) for s in range(8)]
    frames[1] = frames[1].copy(); frames[1][5, 7, 1] = 0x8000                  # -0.0
    frames[2] = frames[2].copy(); frames[2][9, 100, 2] = 0x7C00                # +inf
    frames[3] = frames[3].copy(); frames[3][129, 479, 0] = 0x7E00              # NaN
    frames[5] = frames[5].copy(); frames[5][64, 200, 1] = np.float16(60000.0).view(np.uint16)
    frames[6] = frames[6].copy(); frames[6][..., 0] = np.minimum(frames[6][..., 0], np.float16(100.0).view(np.uint16))
    want = _classic(ctx, opt, frames, dst)
    _seed(ctx, opt, frames[0], dst)
    got = _reuse(ctx, opt, frames, dst, expect_redone=5)
    for i, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), i
    for i in (0, 1, 5, 6):
        G.compare_codes(got[i], G.oracle_forward(frames[i], _HALF, dst), True, "frame %d" % i)


def test_a_seed_belongs_to_its_transfer_functions(ctx, opt):
    # same extrema, another destination transfer: the seed LUT is for another function and must not be used
    opt("H2Y_FORWARD_KERNEL", "rows")
    w, h = 480, 130
    frames = [_frame(w, h, 400 + s) for s in range(4)]
    _seed(ctx, opt, frames[0], _dst(10, transfer=16))
    dst = _dst(10, transfer=18)                            # rho-gamma
    opt("H2Y_PLAN_REUSE", "1")
    got = G.gpu_forward(ctx, frames, _HALF, dst)
    assert ctx.forward_last_plan_reuse()[0] is False       # no seed for these parameters yet: classic call
    got2 = _reuse(ctx, opt, frames, dst, expect_redone=0)  # now there is one
    for a, b in zip(got, got2):
        assert np.array_equal(a, b)
    G.compare_codes(got2[1], G.oracle_forward(frames[1], _HALF, dst), True, "rho-gamma, plan reuse")


def test_the_policy_follows_the_content(ctx, opt):
    # automatic mode: a uniform sequence switches plan reuse on after its first (classic) call has reported back;
    # content whose plans differ frame by frame switches it off again
    opt("H2Y_FORWARD_KERNEL", "rows")
    opt("H2Y_PLAN_REUSE", "")
    w, h = 480, 130
    dst = _dst()
    uniform = [_frame(w, h, 420 + s) for s in range(8)]
    varied = [_frame(w, h, 440 + s, hi=500.0 + 37.0 * s) for s in range(8)]
    want_u = _classic(ctx, opt, uniform, dst)
    want_v = _classic(ctx, opt, varied, dst)
    opt("H2Y_PLAN_REUSE", "")
    seen = []
    for frames, want in ((uniform, want_u),) * 3 + ((varied, want_v),) * 3 + ((uniform, want_u),) * 3:
        got = G.gpu_forward(ctx, frames, _HALF, dst)       # gpu_forward synchronises: the feedback has landed
        seen.append(ctx.forward_last_plan_reuse()[0])
        for i, (a, b) in enumerate(zip(got, want)):
            assert np.array_equal(a, b), i
    assert seen[1] and seen[2], seen                       # on for the uniform sequence
    assert not seen[5], seen                               # off once the varied content has been seen
    assert seen[8], seen                                   # and on again


def test_full_size_batch_takes_the_single_pass_by_itself(ctx, opt):
    # BASELINE configs[1] geometry, 10 frames: large enough for the rows kernel without forcing anything
    opt("H2Y_PLAN_REUSE", "")
    w, h, n = 3840, 2160, 10
    dst = _dst()
    base = [synth.exr_half_frame_fast(w, h, seed=s, channels=3) for s in (0, 1)]
    frames = [base[i % 2] for i in range(n)]
    first = G.gpu_forward(ctx, frames, _HALF, dst)         # classic, or a reuse attempt with whatever seed an earlier test left
    G.gpu_forward(ctx, frames, _HALF, dst)                 # by now the policy has seen a uniform call with this content
    second = G.gpu_forward(ctx, frames, _HALF, dst)
    attempted, nf, redone = ctx.forward_last_plan_reuse()
    assert attempted and nf == n and redone == 0
    for a, b in zip(first, second):
        assert np.array_equal(a, b)
    G.compare_codes(second[1], G.oracle_forward(base[1], _HALF, dst), True, "4K frame, single pass")
