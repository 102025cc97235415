"""-m gpu end-to-end tests of the C++ command-line hosts: files in, files out, compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import cases
from hdr2yuv_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cli():
    import torch
    assert torch.cuda.is_available()
    from hdr2yuv_b200 import build
    return build.build_cli()


def run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    text = r.stdout.decode(errors="replace")
    assert r.returncode == 0, text
    return text


def test_tiff_to_yuv_and_append(cli, tmp_path):
    w, h = 256, 96
    px = synth.tiff16_frame(w, h, seed=4)
    px.tofile(tmp_path / "t.raw")
    run([cli["h2y_iotool"], "write-tiff", str(tmp_path / "in.tiff"), str(w), str(h), "3", str(tmp_path / "t.raw")])
    out = tmp_path / "out.yuv"
    cmd = [cli["hdr2yuv"], "--src_filename", str(tmp_path / "in.tiff"), "--dst_filename", str(out), "--src_pic_width", str(w),
           "--src_pic_height", str(h), "--src_bit_depth", "16", "--dst_bit_depth", "10", "--src_chroma_format_idc", "3",
           "--dst_chroma_format_idc", "1", "--src_matrix_coeffs", "0", "--dst_matrix_coeffs", "9",
           "--src_transfer_characteristics", "16", "--dst_transfer_characteristics", "16", "--src_colour_primaries", "10",
           "--dst_colour_primaries", "9", "--chroma_resampler_type", "1"]
    run(cmd)
    src = dict(bit_depth=16, full_range=0, transfer=16, primaries=10, matrix=0)
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    want = O.forward(O.load_rgb16(px, 0), src, dst, backend="port")
    got = np.fromfile(out, np.uint16)
    assert np.array_equal(got, want)
    run(cmd)                                   # the reference appends (tiff.cpp:440): a second run doubles the file
    got = np.fromfile(out, np.uint16)
    assert got.size == 2 * want.size and np.array_equal(got[want.size:], want)


def test_exr_sequence_to_yuv(cli, tmp_path):
    w, h, n = 256, 96, 5
    frames = [synth.exr_half_frame(w, h, seed=10 + i, channels=3) for i in range(n)]
    for i, f in enumerate(frames):
        f.tofile(tmp_path / "f.raw")
        run([cli["h2y_iotool"], "write-exr", str(tmp_path / ("shot.%04d.exr" % i)), str(w), str(h), "3", "3" if i % 2 else "0",
             str(tmp_path / "f.raw")])
    out = tmp_path / "out.yuv"
    run([cli["hdr2yuv"], "--src_filename", str(tmp_path / "shot.0000.exr"), "--dst_filename", str(out), "--src_pic_width", str(w),
         "--src_pic_height", str(h), "--src_bit_depth", "16", "--dst_bit_depth", "12", "--src_chroma_format_idc", "3",
         "--dst_chroma_format_idc", "1", "--src_matrix_coeffs", "0", "--dst_matrix_coeffs", "9",
         "--src_transfer_characteristics", "LINEAR", "--dst_transfer_characteristics", "PQ", "--src_colour_primaries", "1",
         "--dst_colour_primaries", "9", "--chroma_resampler_type", "1", "--n_frames", str(n), "--batch_frames", "2",
         "--dst_video_full_range_flag", "0"])
    src = dict(bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0)
    dst = dict(bit_depth=12, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    got = np.fromfile(out, np.uint16).reshape(n, -1)
    for i in range(n):
        want = O.forward(O.load_half(frames[i]), src, dst, backend="port")
        d = np.abs(got[i].astype(np.int32) - want.astype(np.int32))
        assert d.max() <= 1 and (d != 0).sum() <= max(2, d.size // 2000)


def test_rgb_to_444_yuv_box_default(cli, tmp_path):
    # .rgb planar input is NOT clipped on read (only read_tiff clips); chroma_resampler_type defaults to 0 (box)
    w, h = 128, 64
    px = synth.tiff16_frame(w, h, seed=6)
    np.ascontiguousarray(px.transpose(2, 0, 1)).tofile(tmp_path / "in.rgb")
    out = tmp_path / "out.yuv"
    run([cli["hdr2yuv"], "--src_filename", str(tmp_path / "in.rgb"), "--dst_filename", str(out), "--src_pic_width", str(w),
         "--src_pic_height", str(h), "--src_bit_depth", "16", "--dst_bit_depth", "10", "--src_chroma_format_idc", "3",
         "--dst_chroma_format_idc", "1", "--src_matrix_coeffs", "0", "--dst_matrix_coeffs", "11",
         "--src_transfer_characteristics", "16", "--src_colour_primaries", "10"])
    planes = np.ascontiguousarray(np.stack([px[..., 1], px[..., 2], px[..., 0]], 0))
    src = dict(bit_depth=16, full_range=0, transfer=16, primaries=10, matrix=0)
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=10, matrix=11, chroma=1, resampler=0)
    assert np.array_equal(np.fromfile(out, np.uint16), O.forward(planes, src, dst, backend="port"))


def test_yuv2tiff_matches_oracle(cli, tmp_path, golden_inverse):
    yuv = golden_inverse[cases.inverse_input_key(2)]           # 960x540 10-bit Y'CbCr frame
    two = np.concatenate([yuv, yuv[::-1].copy()]) if False else np.concatenate([yuv, yuv])
    two.tofile(tmp_path / "in.yuv")
    text = run([cli["yuv2tiff"], str(tmp_path / "in.yuv"), "B10", "2020", "HD960", "-I", "-o", str(tmp_path / "tifXYZ")])
    want, invalid = O.yuv2tiff(yuv, cases.IW, cases.IH, 10, O.INV_2020, True, False, False, backend="port")
    for i in range(2):
        p = tmp_path / "tifXYZ" / ("XpYpZp%05d.tif" % i)
        run([cli["h2y_iotool"], "read-tiff", str(p), str(tmp_path / "o.raw")])
        got = np.fromfile(tmp_path / "o.raw", np.uint16).reshape(cases.IH, cases.IW, 3)
        assert np.array_equal(got, want)
    assert "Invalid Pixels:  %d" % invalid in text


def test_yuv444_to_tiff_uses_matrix_inverse(cli, tmp_path):
    # the reference's YCbCr 4:4:4 -> RGB tiff case of test.sh:70-80
    pl = cases.minv_input(12)
    pl.tofile(tmp_path / "in.yuv")
    out = tmp_path / "out.tiff"
    text = run([cli["hdr2yuv"], "--src_filename", str(tmp_path / "in.yuv"), "--dst_filename", str(out), "--src_pic_width", str(cases.MW),
                "--src_pic_height", str(cases.MH), "--src_bit_depth", "12", "--dst_bit_depth", "16", "--src_chroma_format_idc", "3",
                "--dst_chroma_format_idc", "3", "--src_matrix_coeffs", "1", "--dst_matrix_coeffs", "0",
                "--src_transfer_characteristics", "1", "--dst_transfer_characteristics", "1", "--src_colour_primaries", "1",
                "--dst_colour_primaries", "1", "--verbose_level", "1"])
    tmp, invalid = O.matrix_inverse(pl, 1, 12, 0, 12, backend="port")
    want = O.write_tiff_rows(tmp, 16, 12)
    run([cli["h2y_iotool"], "read-tiff", str(out), str(tmp_path / "o.raw")])
    got = np.fromfile(tmp_path / "o.raw", np.uint16).reshape(cases.MH, cases.MW, 3)
    assert np.array_equal(got, want)
    assert "invalid pixels %d" % invalid in text


def test_dpx_sequence_to_yuv(cli, tmp_path):
    # 10-bit packed DPX frames (both byte orders in one sequence is refused; here big-endian, as film scanners write):
    # reader on the host, unpack + /1023.0 + the whole chain on the GPU
    w, h, n = 256, 96, 3
    rng = np.random.default_rng(5)
    stored = []
    for i in range(n):
        c = rng.integers(0, 1024, (h, w, 3), dtype=np.uint16)
        c[2, 3] = 1023
        c[4, 5] = 0
        c.tofile(tmp_path / "c.raw")
        run([cli["h2y_iotool"], "write-dpx", str(tmp_path / ("scan_%04d.dpx" % i)), str(w), str(h), "1", str(tmp_path / "c.raw")])
        c = c.astype(np.uint32)
        stored.append(((c[..., 0] << 22) | (c[..., 1] << 12) | (c[..., 2] << 2)).astype(">u4"))
    out = tmp_path / "out.yuv"
    run([cli["hdr2yuv"], "--src_filename", str(tmp_path / "scan_0000.dpx"), "--dst_filename", str(out), "--src_pic_width", str(w),
         "--src_pic_height", str(h), "--src_bit_depth", "10", "--dst_bit_depth", "10", "--src_chroma_format_idc", "3",
         "--dst_chroma_format_idc", "1", "--src_matrix_coeffs", "0", "--dst_matrix_coeffs", "9",
         "--src_transfer_characteristics", "8", "--dst_transfer_characteristics", "16", "--src_colour_primaries", "1",
         "--dst_colour_primaries", "9", "--src_video_full_range_flag", "1", "--dst_video_full_range_flag", "0",       # dst would inherit 1
         "--chroma_resampler_type", "1", "--n_frames", str(n)])
    src = dict(bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0)
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    got = np.fromfile(out, np.uint16).reshape(n, -1)
    for i in range(n):
        want = O.forward(O.load_dpx10(stored[i], True), src, dst, backend="port")
        d = np.abs(got[i].astype(int) - want.astype(int))
        assert d.max() <= 1 and (d != 0).sum() <= max(2, d.size // 2000), (i, int(d.max()), int((d != 0).sum()), got[i][:8], want[:8])


def test_devices_2_shards_the_frame_range(cli, tmp_path):
    # --devices N: one worker thread per GPU, contiguous frame ranges (sharding.frame_range), every worker writes its frames
    # at their own offsets of the one .yuv file.  Needs two GPUs (gpurun --gpus 2); the result must equal the one-GPU file.
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    w, h, n = 256, 96, 7
    for i in range(n):
        synth.exr_half_frame(w, h, seed=30 + i, channels=3).tofile(tmp_path / "f.raw")
        run([cli["h2y_iotool"], "write-exr", str(tmp_path / ("shot.%04d.exr" % i)), str(w), str(h), "3", "0", str(tmp_path / "f.raw")])
    outs = []
    for devices in (1, 2):
        out = tmp_path / ("out%d.yuv" % devices)
        text = run([cli["hdr2yuv"], "--src_filename", str(tmp_path / "shot.0000.exr"), "--dst_filename", str(out), "--src_pic_width", str(w),
                    "--src_pic_height", str(h), "--src_bit_depth", "16", "--dst_bit_depth", "10", "--src_chroma_format_idc", "3",
                    "--dst_chroma_format_idc", "1", "--src_matrix_coeffs", "0", "--dst_matrix_coeffs", "9",
                    "--src_transfer_characteristics", "LINEAR", "--dst_transfer_characteristics", "PQ", "--src_colour_primaries", "1",
                    "--dst_colour_primaries", "9", "--chroma_resampler_type", "1", "--n_frames", str(n), "--batch_frames", "2",
                    "--dst_video_full_range_flag", "0", "--devices", str(devices)])
        outs.append(np.fromfile(out, np.uint16))
    assert outs[0].size == outs[1].size == n * (w * h * 3 // 2)
    assert np.array_equal(outs[0], outs[1])


def test_rgb_to_exr_float_destination(cli, tmp_path):
    # An .exr destination (hdr2yuv.cpp:935-957): matrix_convert's F32 twin at tmp depth 16, then write_exr_file's RGBA
    # halfs with alpha 0 (exr.cpp:99-133).  The reference's writer needs OpenEXR, so the planes are checked against the
    # reference's own matrix_convert (or its pinned restatement) rounded to half, read back through this repo's reader.
    w, h = 128, 64
    px = synth.tiff16_frame(w, h, seed=71)
    np.ascontiguousarray(px.transpose(2, 0, 1)).tofile(tmp_path / "in.rgb")
    out = tmp_path / "out.exr"
    run([cli["hdr2yuv"], "--src_filename", str(tmp_path / "in.rgb"), "--dst_filename", str(out), "--src_transfer_characteristics", "16",
         "--dst_transfer_characteristics", "16", "--src_pic_width", str(w), "--src_pic_height", str(h), "--src_bit_depth", "16",
         "--dst_bit_depth", "16", "--src_chroma_format_idc", "3", "--src_matrix_coeffs", "0", "--dst_matrix_coeffs", "9",
         "--src_colour_primaries", "9", "--dst_colour_primaries", "9"])
    run([cli["h2y_iotool"], "read-exr", str(out), str(tmp_path / "o.raw"), "4"])
    got = np.fromfile(tmp_path / "o.raw", np.uint16).reshape(h, w, 4)
    planes = np.ascontiguousarray(np.stack([px[..., 1], px[..., 2], px[..., 0]], 0))        # .rgb is read unclipped
    src = dict(bit_depth=16, full_range=0, transfer=16, primaries=9, matrix=0)
    dst = dict(bit_depth=16, full_range=0, transfer=16, primaries=9, matrix=9)
    want = O.matrix_convert_f32(planes, src, dst, "ref" if O.ref_available() else "port")   # G, B, R = Y, Cb, Cr planes
    want_half = np.stack([want[2], want[0], want[1]], -1).astype(np.float16).view(np.uint16)
    assert np.array_equal(got[..., :3], want_half)
    assert not got[..., 3].any()
