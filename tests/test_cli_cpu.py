"""CPU tests of the C++ command-line hosts' host side: the native TIFF / EXR / raw readers and writers
(checked against OpenCV's libtiff / OpenEXR codecs where available), frame-sequence naming and the
reference's option handling.  No GPU: hdr2yuv --dump_input stops before any CUDA call."""
import os
import subprocess

import numpy as np
import pytest

from hdr2yuv_b200 import synth


@pytest.fixture(scope="module")
def cli():
    from hdr2yuv_b200 import build
    return build.build_cli()


def run(cmd, cwd=None, check=True):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    text = r.stdout.decode(errors="replace")
    if check:
        assert r.returncode == 0, text
    return r.returncode, text


def _halfs(rng, h, w, c):
    v = np.exp(rng.uniform(np.log(1e-3), np.log(4000.0), (h, w, c))).astype(np.float16)
    v.reshape(-1)[::17] = 0
    return v.view(np.uint16)


@pytest.mark.parametrize("comp", [0, 2, 3])
@pytest.mark.parametrize("channels", [3, 4])
def test_exr_round_trip(cli, tmp_path, comp, channels):
    rng = np.random.default_rng(comp * 10 + channels)
    w, h = 37, 35                       # 35 rows: two full 16-line ZIP blocks and a partial one
    px = _halfs(rng, h, w, channels)
    raw, exr, back = tmp_path / "in.raw", tmp_path / "a.exr", tmp_path / "out.raw"
    px.tofile(raw)
    run([cli["h2y_iotool"], "write-exr", str(exr), str(w), str(h), str(channels), str(comp), str(raw)])
    run([cli["h2y_iotool"], "read-exr", str(exr), str(back), str(channels)])
    assert np.array_equal(np.fromfile(back, np.uint16).reshape(h, w, channels), px)
    if channels == 4:                   # reading 3 of 4 channels drops alpha; reading 4 of 3 fills alpha with 1.0
        run([cli["h2y_iotool"], "read-exr", str(exr), str(back), "3"])
        assert np.array_equal(np.fromfile(back, np.uint16).reshape(h, w, 3), px[..., :3])
    else:
        run([cli["h2y_iotool"], "read-exr", str(exr), str(back), "4"])
        got = np.fromfile(back, np.uint16).reshape(h, w, 4)
        assert np.array_equal(got[..., :3], px) and (got[..., 3] == 0x3C00).all()


def test_readers_against_opencv_codecs(cli, tmp_path):
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    w, h = 53, 41
    f = np.exp(rng.uniform(np.log(1e-3), np.log(4000.0), (h, w, 3))).astype(np.float32)
    for name, comp in (("none", cv2.IMWRITE_EXR_COMPRESSION_NO), ("zip", cv2.IMWRITE_EXR_COMPRESSION_ZIP),
                       ("zips", cv2.IMWRITE_EXR_COMPRESSION_ZIPS)):
        p = str(tmp_path / ("cv_%s.exr" % name))
        try:
            ok = cv2.imwrite(p, f, [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION, comp])
        except cv2.error:
            pytest.skip("this OpenCV build has no OpenEXR codec")
        assert ok
        run([cli["h2y_iotool"], "read-exr", p, str(tmp_path / "o.raw")])
        got = np.fromfile(tmp_path / "o.raw", np.uint16).reshape(h, w, 3).view(np.float16)
        assert np.array_equal(got, f[..., ::-1].astype(np.float16))          # OpenCV is B,G,R
    # a FLOAT-channel file is narrowed to half like Imf::RgbaInputFile does
    p = str(tmp_path / "cv_f32.exr")
    assert cv2.imwrite(p, f, [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_FLOAT])
    run([cli["h2y_iotool"], "read-exr", p, str(tmp_path / "o.raw")])
    got = np.fromfile(tmp_path / "o.raw", np.uint16).reshape(h, w, 3).view(np.float16)
    assert np.array_equal(got, f[..., ::-1].astype(np.float16))
    # TIFF: libtiff-written, and our writer read back by libtiff
    t = rng.integers(0, 65536, (h, w, 3), dtype=np.uint16)
    p = str(tmp_path / "cv.tiff")
    assert cv2.imwrite(p, t, [cv2.IMWRITE_TIFF_COMPRESSION, 1])
    run([cli["h2y_iotool"], "read-tiff", p, str(tmp_path / "o.raw")])
    assert np.array_equal(np.fromfile(tmp_path / "o.raw", np.uint16).reshape(h, w, 3), t[..., ::-1])
    t.tofile(tmp_path / "t.raw")
    p2 = str(tmp_path / "ours.tiff")
    run([cli["h2y_iotool"], "write-tiff", p2, str(w), str(h), "3", str(tmp_path / "t.raw")])
    back = cv2.imread(p2, cv2.IMREAD_UNCHANGED)
    assert back is not None and back.dtype == np.uint16 and np.array_equal(back[..., ::-1], t)


def test_tiff_round_trip_and_centre_cutout(cli, tmp_path):
    rng = np.random.default_rng(9)
    for channels in (3, 4):
        w, h = 100, 60
        t = rng.integers(0, 65536, (h, w, channels), dtype=np.uint16)
        t.tofile(tmp_path / "t.raw")
        p = str(tmp_path / "a.tiff")
        run([cli["h2y_iotool"], "write-tiff", p, str(w), str(h), str(channels), str(tmp_path / "t.raw")])
        run([cli["h2y_iotool"], "read-tiff", p, str(tmp_path / "o.raw")])
        assert np.array_equal(np.fromfile(tmp_path / "o.raw", np.uint16).reshape(h, w, channels), t)
        run([cli["h2y_iotool"], "read-tiff", p, str(tmp_path / "o.raw"), "40", "20"])      # tiff.cpp:191-220
        assert np.array_equal(np.fromfile(tmp_path / "o.raw", np.uint16).reshape(20, 40, channels), t[20:40, 30:70])


BASE = ["--src_chroma_format_idc", "3", "--dst_chroma_format_idc", "1", "--dst_bit_depth", "10",
        "--src_matrix_coeffs", "0", "--dst_matrix_coeffs", "9", "--src_colour_primaries", "1", "--dst_colour_primaries", "9"]


def test_dump_input_layouts_and_sequences(cli, tmp_path):
    w, h = 64, 48
    # numbered EXR sequence: frame i = seed i
    frames = [synth.exr_half_frame(w, h, seed=i, channels=3) for i in range(3)]
    for i, f in enumerate(frames):
        f.tofile(tmp_path / "f.raw")
        run([cli["h2y_iotool"], "write-exr", str(tmp_path / ("clip_%05d.exr" % (7 + i))), str(w), str(h), "3", "3", str(tmp_path / "f.raw")])
    dump = tmp_path / "dump.raw"
    _, text = run([cli["hdr2yuv"], "--src_filename", str(tmp_path / "clip_00007.exr"), "--dst_filename", str(tmp_path / "o.yuv"),
                   "--src_pic_width", str(w), "--src_pic_height", str(h), "--src_bit_depth", "16", "--n_frames", "3",
                   "--src_transfer_characteristics", "8", "--dst_transfer_characteristics", "PQ", "--dump_input", str(dump)] + BASE)
    got = np.fromfile(dump, np.uint16).reshape(3, h, w, 3)
    assert all(np.array_equal(got[i], frames[i]) for i in range(3)), text
    assert "dst_transfer_characteristics: 16 (type: PQ)" in text          # names resolve like the reference's table
    # printf-style pattern with a start frame
    _, _ = run([cli["hdr2yuv"], "--src_filename", str(tmp_path / "clip_%05d.exr"), "--dst_filename", str(tmp_path / "o.yuv"),
                "--src_pic_width", str(w), "--src_pic_height", str(h), "--src_bit_depth", "16", "--n_frames", "2",
                "--src_start_frame", "8", "--src_transfer_characteristics", "8", "--dst_transfer_characteristics", "16",
                "--dump_input", str(dump)] + BASE)
    got = np.fromfile(dump, np.uint16).reshape(2, h, w, 3)
    assert np.array_equal(got[0], frames[1]) and np.array_equal(got[1], frames[2])
    # planar .rgb, second frame of the file: R,G,B planes -> interleaved R,G,B
    rgb = np.stack([synth.tiff16_frame(w, h, seed=s) for s in (1, 2)], 0)
    np.ascontiguousarray(rgb.transpose(0, 3, 1, 2)).tofile(tmp_path / "clip.rgb")
    run([cli["hdr2yuv"], "--src_filename", str(tmp_path / "clip.rgb"), "--dst_filename", str(tmp_path / "o.yuv"),
         "--src_pic_width", str(w), "--src_pic_height", str(h), "--src_bit_depth", "16", "--src_start_frame", "1",
         "--src_transfer_characteristics", "16", "--dump_input", str(dump)] + BASE)
    assert np.array_equal(np.fromfile(dump, np.uint16).reshape(h, w, 3), rgb[1])


def test_option_errors_follow_the_reference(cli, tmp_path):
    # 4:2:0 input is refused with the reference's message and exit(0) (hdr2yuv.cpp:536-574)
    rc, text = run([cli["hdr2yuv"], "--src_filename", "a.yuv", "--dst_filename", "b.yuv", "--src_pic_width", "64",
                    "--src_pic_height", "48", "--src_bit_depth", "10", "--src_chroma_format_idc", "1"], check=False)
    assert rc == 0 and "Only 4:4:4 input supported" in text and "TOO MANY ARGUMENT ERRORS" in text
    # unknown extension
    rc, text = run([cli["hdr2yuv"], "--src_filename", "a.png", "--dst_filename", "b.yuv", "--src_pic_width", "64",
                    "--src_pic_height", "48", "--src_bit_depth", "10", "--src_chroma_format_idc", "3"], check=False)
    assert rc == 0 and "not recongized or not supported" in text
    # unknown flags only warn
    rc, text = run([cli["hdr2yuv"], "--bogus", "1", "--help"], check=False)
    assert "WARNING: argument (--bogus) unrecongized" in text and "Transfer chracteristics options" in text
    # a short raw file is reported, not read past its end
    np.zeros(10, np.uint16).tofile(tmp_path / "short.rgb")
    rc, text = run([cli["hdr2yuv"], "--src_filename", str(tmp_path / "short.rgb"), "--dst_filename", str(tmp_path / "o.yuv"),
                    "--src_pic_width", "64", "--src_pic_height", "48", "--src_bit_depth", "16", "--src_chroma_format_idc", "3",
                    "--dst_chroma_format_idc", "1", "--dst_bit_depth", "10"], check=False)
    assert rc == 1 and "holds 0 frames" in text


@pytest.mark.parametrize("big_endian", [1, 0])
def test_dpx_10bit_reader_and_dump(cli, tmp_path, big_endian):
    # 10-bit packed DPX (dpx.cpp:210-360, 506-531): the header fields dpx_read uses, both byte orders; the host hands
    # the stored words to the GPU untouched, so --dump_input must return exactly the file's pixel words
    w, h = 72, 40
    rng = np.random.default_rng(3 + big_endian)
    rgb10 = rng.integers(0, 1024, (h, w, 3), dtype=np.uint16)
    rgb10.tofile(tmp_path / "c.raw")
    dpx = tmp_path / "in.dpx"
    run([cli["h2y_iotool"], "write-dpx", str(dpx), str(w), str(h), str(big_endian), str(tmp_path / "c.raw")])
    raw = np.fromfile(dpx, np.uint8)
    assert raw.size == 2048 + w * h * 4 and bytes(raw[:4]) == (b"SDPX" if big_endian else b"XPDS")
    _, text = run([cli["h2y_iotool"], "read-dpx", str(dpx), str(tmp_path / "w.raw")])
    assert text.split() == [str(w), str(h), "10", str(big_endian)]
    words = np.fromfile(tmp_path / "w.raw", np.dtype(">u4") if big_endian else np.dtype("<u4")).reshape(h, w)
    assert np.array_equal(words >> 22, rgb10[..., 0]) and np.array_equal((words >> 12) & 1023, rgb10[..., 1])
    assert np.array_equal((words >> 2) & 1023, rgb10[..., 2])
    dump = tmp_path / "dump.raw"
    size = ["--src_pic_width", str(w), "--src_pic_height", str(h), "--src_bit_depth", "10"]
    _, text = run([cli["hdr2yuv"], "--src_filename", str(dpx), "--dst_filename", str(tmp_path / "o.yuv"),
                   "--src_transfer_characteristics", "8", "--dst_transfer_characteristics", "16", "--dump_input", str(dump)] + size + BASE)
    assert np.array_equal(np.fromfile(dump, np.uint8), raw[2048:]), text
    assert "layout %d" % (6 if big_endian else 7) in text
    # a 12-bit DPX is refused with dpx_read's message, like every packing it does not know (dpx.cpp:333-338)
    bad = raw.copy(); bad[803] = 12
    bad.tofile(tmp_path / "bad.dpx")
    rc, text = run([cli["hdr2yuv"], "--src_filename", str(tmp_path / "bad.dpx"), "--dst_filename", str(tmp_path / "o.yuv"),
                    "--dump_input", str(dump)] + size + BASE, check=False)
    assert rc != 0 and "12-bit" in text


def test_readers_reject_damaged_files(cli, tmp_path):
    # Damaged inputs must come back as an error message and exit status 1, never as a crash or an out-of-bounds read:
    # a raw EXR chunk that claims fewer bytes than its lines need, a chunk size beyond the block, a truncated file,
    # an inverted data window, and a TIFF whose strip table claims 2^31 entries.
    rng = np.random.default_rng(5)
    w, h = 24, 18
    px = _halfs(rng, h, w, 3)
    raw, exr, out = tmp_path / "in.raw", tmp_path / "a.exr", tmp_path / "o.raw"
    px.tofile(raw)
    run([cli["h2y_iotool"], "write-exr", str(exr), str(w), str(h), "3", "0", str(raw)])
    good = bytearray(exr.read_bytes())
    line = w * 3 * 2
    first_chunk = len(good) - h * (8 + line)            # uncompressed: h chunks of (y, size, line)
    assert int.from_bytes(good[first_chunk + 4:first_chunk + 8], "little") == line

    def damaged(name, edit):
        b = bytearray(good)
        b = edit(b) or b
        p = tmp_path / name
        p.write_bytes(bytes(b))
        rc, text = run([cli["h2y_iotool"], "read-exr", str(p), str(out)], check=False)
        assert rc == 1, (name, rc, text)
        return text

    def chunk_size(v):
        def edit(b):
            b[first_chunk + 4:first_chunk + 8] = int(v).to_bytes(4, "little", signed=True)
        return edit

    assert "chunk" in damaged("short_chunk.exr", chunk_size(line - 2))
    assert "chunk" in damaged("huge_chunk.exr", chunk_size(0x7fffffff))
    assert "chunk" in damaged("negative_chunk.exr", chunk_size(-4))
    damaged("truncated.exr", lambda b: b[:len(b) - line // 2])
    i = bytes(good).index(b"dataWindow\0box2i\0") + len(b"dataWindow\0box2i\0") + 4
    def flip_window(b):
        b[i + 8:i + 12] = (-5).to_bytes(4, "little", signed=True)          # xmax < xmin
    assert "data window" in damaged("window.exr", flip_window)

    tif = tmp_path / "a.tiff"
    synth.tiff16_frame(w, h, seed=2).tofile(raw)
    run([cli["h2y_iotool"], "write-tiff", str(tif), str(w), str(h), "3", str(raw)])
    b = bytearray(tif.read_bytes())
    ifd = int.from_bytes(b[4:8], "little")
    n = int.from_bytes(b[ifd:ifd + 2], "little")
    for k in range(n):
        e = ifd + 2 + 12 * k
        if int.from_bytes(b[e:e + 2], "little") == 273:                     # StripOffsets: count = 2^31
            b[e + 4:e + 8] = (1 << 31).to_bytes(4, "little")
    bad = tmp_path / "bad.tiff"
    bad.write_bytes(bytes(b))
    rc, text = run([cli["h2y_iotool"], "read-tiff", str(bad), str(out)], check=False)
    assert rc == 1 and "strips" in text, text
    (tmp_path / "cut.tiff").write_bytes(bytes(tif.read_bytes()[:200]))
    rc, text = run([cli["h2y_iotool"], "read-tiff", str(tmp_path / "cut.tiff"), str(out)], check=False)
    assert rc == 1, text

    # a strip table that does not cover the picture, RowsPerStrip = 0, byte counts smaller than the rows they stand for:
    # no row of the destination may be left as it was (the CLI's destination is uninitialised pinned memory)
    def tiff_tag_edit(name, tag, count=None, value=None, entry_of_array=None):
        b = bytearray(tif.read_bytes())
        for k in range(n):
            e = ifd + 2 + 12 * k
            if int.from_bytes(b[e:e + 2], "little") == tag:
                if count is not None:
                    b[e + 4:e + 8] = int(count).to_bytes(4, "little")
                if value is not None:
                    b[e + 8:e + 12] = int(value).to_bytes(4, "little")
                if entry_of_array is not None:
                    arr = int.from_bytes(b[e + 8:e + 12], "little")
                    b[arr + 4 * entry_of_array[0]:arr + 4 * entry_of_array[0] + 4] = int(entry_of_array[1]).to_bytes(4, "little")
        p = tmp_path / name
        p.write_bytes(bytes(b))
        rc, text = run([cli["h2y_iotool"], "read-tiff", str(p), str(out)], check=False)
        assert rc == 1, (name, rc, text)
        return text

    assert "cover" in tiff_tag_edit("short_table.tiff", 273, count=h - 3)
    assert "RowsPerStrip" in tiff_tag_edit("rps0.tiff", 278, value=0)
    assert "byte count" in tiff_tag_edit("small_count.tiff", 279, entry_of_array=(5, 10))


# ---- the reference's own program (oracle/_ref/hdr2yuv_ref = hdr2yuv.cpp's main() built unmodified) ------------------------
# The GPU CLI tests hand the oracle the source / destination parameters the test author derived from an argv.  Here that
# derivation (option inheritance hdr2yuv.cpp:265-318, defaults, .rgb read path, write_yuv append) is checked against the
# reference program itself, on the CPU: same argv -> the reference's output file must equal the oracle's forward() result.
REF_BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "hdr2yuv_ref")
needs_ref_bin = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/hdr2yuv_ref not built (no /root/reference here)")

REF_PROGRAM_CASES = {
    # argv tail (after the file names and geometry) -> (src bits, src dict, dst dict) as the GPU tests would derive them
    "rgb12_to_420_bt709_box_default": (
        12, ["--src_transfer_characteristics", "1", "--dst_transfer_characteristics", "1", "--dst_bit_depth", "10",
             "--dst_chroma_format_idc", "1", "--src_matrix_coeffs", "0", "--dst_matrix_coeffs", "1", "--src_colour_primaries", "1",
             "--dst_colour_primaries", "1"],
        dict(bit_depth=12, full_range=0, transfer=1, primaries=1, matrix=0),
        dict(bit_depth=10, full_range=0, transfer=1, primaries=1, matrix=1, chroma=1, resampler=0)),
    "rgb16_everything_inherited": (
        16, ["--src_transfer_characteristics", "16", "--dst_chroma_format_idc", "1", "--src_matrix_coeffs", "0",
             "--dst_matrix_coeffs", "9", "--src_colour_primaries", "9"],
        dict(bit_depth=16, full_range=0, transfer=16, primaries=9, matrix=0),
        dict(bit_depth=16, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=0)),
    "rgb16_ydzdx_fir_10": (
        16, ["--src_transfer_characteristics", "16", "--dst_bit_depth", "10", "--dst_chroma_format_idc", "1", "--src_matrix_coeffs", "0",
             "--dst_matrix_coeffs", "11", "--src_colour_primaries", "10", "--chroma_resampler_type", "1"],
        dict(bit_depth=16, full_range=0, transfer=16, primaries=10, matrix=0),
        dict(bit_depth=10, full_range=0, transfer=16, primaries=10, matrix=11, chroma=1, resampler=1)),
    # (no 4:2:2 case: the reference's convert() has no 4:2:2 branch, its program writes constant chroma planes; this repo
    # defines 4:2:2 as stage 1 of the reference's two-stage FIR, pinned through ref_subsample_fir's first stage instead)
    "rgb16_full_range_flags": (
        16, ["--src_transfer_characteristics", "16", "--dst_bit_depth", "12", "--dst_chroma_format_idc", "3", "--src_matrix_coeffs", "0",
             "--dst_matrix_coeffs", "9", "--src_colour_primaries", "9", "--src_video_full_range_flag", "1"],
        dict(bit_depth=16, full_range=1, transfer=16, primaries=9, matrix=0),
        dict(bit_depth=12, full_range=1, transfer=16, primaries=9, matrix=9, chroma=3, resampler=0)),
}


@needs_ref_bin
@pytest.mark.parametrize("name", sorted(REF_PROGRAM_CASES))
def test_reference_program_agrees_with_the_oracle_parameters(tmp_path, name):
    from oracle import oracle as O
    bits, tail, src, dst = REF_PROGRAM_CASES[name]
    w, h = 128, 64
    px = synth.tiff16_frame(w, h, seed=len(name)) >> (16 - bits)
    planes_rgb = np.ascontiguousarray(px.transpose(2, 0, 1)).astype(np.uint16)      # .rgb files are planar R, G, B
    planes_rgb.tofile(tmp_path / "in.rgb")
    # transfer options first: the reference indexes its transfer table with the argv position (hdr2yuv.cpp:200)
    argv = [REF_BIN, "--src_filename", "in.rgb", "--dst_filename", "out.yuv"] + tail[:2] + \
           (tail[2:4] if tail[2] == "--dst_transfer_characteristics" else []) + \
           ["--src_pic_width", str(w), "--src_pic_height", str(h), "--src_bit_depth", str(bits), "--src_chroma_format_idc", "3"] + \
           (tail[4:] if tail[2] == "--dst_transfer_characteristics" else tail[2:])
    rc, text = run(argv, cwd=tmp_path)
    got = np.fromfile(tmp_path / "out.yuv", np.uint16)
    planes = np.ascontiguousarray(np.stack([planes_rgb[1], planes_rgb[2], planes_rgb[0]], 0))    # G, B, R
    want = O.forward(planes, src, dst, backend="port")
    assert got.size == want.size, (got.size, want.size, text[-800:])
    assert np.array_equal(got, want), (name, int((got != want).sum()))
    rc, text = run(argv, cwd=tmp_path)                 # write_yuv appends (tiff.cpp:440)
    assert np.fromfile(tmp_path / "out.yuv", np.uint16).size == 2 * want.size


@needs_ref_bin
@pytest.mark.parametrize("declared_bits", [16, 12])
def test_reference_program_tiff_route_agrees_with_the_oracle(cli, tmp_path, declared_bits):
    # test.sh:5-16 on the CPU: the reference's own read_tiff (de-interleave, clip on read, header overriding the declared
    # depth: tiff.cpp:296-304, 322-338) through the oracle's libtiff stand-in, against the restatement's load_rgb16 + forward.
    # The file comes from this repo's TIFF writer; read_tiff only accepts 1080 or 2160 strips (tiff.cpp:178-180).
    from oracle import oracle as O
    w, h = 256, 1080
    px = synth.tiff16_frame(w, h, seed=21)
    px.tofile(tmp_path / "t.raw")
    run([cli["h2y_iotool"], "write-tiff", str(tmp_path / "in.tiff"), str(w), str(h), "3", str(tmp_path / "t.raw")])
    run([REF_BIN, "--src_filename", "in.tiff", "--dst_filename", "out.yuv", "--src_transfer_characteristics", "16",
         "--dst_transfer_characteristics", "16", "--src_pic_width", str(w), "--src_pic_height", str(h), "--src_bit_depth", str(declared_bits),
         "--dst_bit_depth", "10", "--src_chroma_format_idc", "3", "--dst_chroma_format_idc", "1", "--src_matrix_coeffs", "0",
         "--dst_matrix_coeffs", "9", "--src_colour_primaries", "10", "--dst_colour_primaries", "9", "--chroma_resampler_type", "1"], cwd=tmp_path)
    got = np.fromfile(tmp_path / "out.yuv", np.uint16)
    src = dict(bit_depth=16, full_range=0, transfer=16, primaries=10, matrix=0)
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    assert np.array_equal(got, O.forward(O.load_rgb16(px, 0), src, dst, backend="port"))


@needs_ref_bin
def test_dpx_float_writer_equals_the_reference_programs_file(cli, tmp_path):
    # An .rgb -> .dpx run of the reference program (its own dpx_write_float, dpx.cpp:719-920) leaves a float DPX whose
    # picture is all zero: parse_options forces the destination depth to 32 (hdr2yuv.cpp:436-446) and set_pic_clip then
    # shifts by 32 and 24 (common.cpp:303-311), so every clip limit is 0 on x86-64 and matrix_convert's clamp
    # (convert.cpp:1286-1293) zeroes the planes.  This host's writer must produce the same bytes from the same planes.
    w, h = 64, 32
    px = synth.tiff16_frame(w, h, seed=4)
    np.ascontiguousarray(px.transpose(2, 0, 1)).tofile(tmp_path / "in.rgb")
    run([REF_BIN, "--src_filename", "in.rgb", "--dst_filename", "ref.dpx", "--src_transfer_characteristics", "16", "--src_pic_width", str(w),
         "--src_pic_height", str(h), "--src_bit_depth", "16", "--src_chroma_format_idc", "3", "--src_matrix_coeffs", "0",
         "--src_colour_primaries", "9"], cwd=tmp_path)
    ref = (tmp_path / "ref.dpx").read_bytes()
    assert len(ref) == 2048 + w * h * 12 and not any(ref[2048:])
    np.zeros((3, h, w), np.float32).tofile(tmp_path / "planes.bin")
    run([cli["h2y_iotool"], "write-dpx-float", str(tmp_path / "ours.dpx"), str(w), str(h), str(tmp_path / "planes.bin")])
    assert (tmp_path / "ours.dpx").read_bytes() == ref
    # and with real content the pixel section is R, G, B interleaved little-endian floats from the G, B, R planes
    planes = np.random.default_rng(2).random((3, h, w), dtype=np.float32)
    planes.tofile(tmp_path / "planes.bin")
    run([cli["h2y_iotool"], "write-dpx-float", str(tmp_path / "ours2.dpx"), str(w), str(h), str(tmp_path / "planes.bin")])
    b = (tmp_path / "ours2.dpx").read_bytes()
    assert b[:2048] == ref[:2048]
    got = np.frombuffer(b[2048:], "<f4").reshape(h, w, 3)
    assert np.array_equal(got, np.stack([planes[2], planes[0], planes[1]], -1))


@needs_ref_bin
@pytest.mark.parametrize("bits,big_endian", [(16, 0), (16, 1), (32, 0), (32, 1), (10, 1)])
def test_reference_program_reads_our_dpx_files_like_the_oracle(cli, tmp_path, bits, big_endian):
    # The reference's own dpx_read (dpx.cpp:208-547, built unmodified with a stand-in for OpenEXR's half) reads the DPX
    # files this repo's tools write, and its whole program turns them into the .yuv the restatement predicts from the
    # Python loaders (oracle.load_dpx10 / load_dpx16 / load_dpxf32): pins those loaders, which the GPU layout tests use.
    from oracle import oracle as O
    w, h = 64, 32
    rng = np.random.default_rng(bits + big_endian)
    if bits == 10:
        c = rng.integers(0, 1024, (h, w, 3), dtype=np.uint16)
        c.tofile(tmp_path / "c.raw")
        run([cli["h2y_iotool"], "write-dpx", str(tmp_path / "in.dpx"), str(w), str(h), str(big_endian), str(tmp_path / "c.raw")])
        c32 = c.astype(np.uint32)
        planes = O.load_dpx10(((c32[..., 0] << 22) | (c32[..., 1] << 12) | (c32[..., 2] << 2)).astype(">u4" if big_endian else "<u4"), bool(big_endian))
    elif bits == 16:
        c = rng.integers(0, 65536, (h, w, 3), dtype=np.uint16)
        c.tofile(tmp_path / "c.raw")
        run([cli["h2y_iotool"], "write-dpx-raw", str(tmp_path / "in.dpx"), str(w), str(h), str(big_endian), "16", str(tmp_path / "c.raw")])
        planes = O.load_dpx16(c.astype(">u2" if big_endian else "<u2"), bool(big_endian))
    else:
        c = np.exp(rng.uniform(np.log(1e-3), np.log(100.0), (h, w, 3))).astype(np.float32)
        c.tofile(tmp_path / "c.raw")
        run([cli["h2y_iotool"], "write-dpx-raw", str(tmp_path / "in.dpx"), str(w), str(h), str(big_endian), "32", str(tmp_path / "c.raw")])
        planes = O.load_dpxf32(c.astype(">f4" if big_endian else "<f4"), bool(big_endian))
    run([REF_BIN, "--src_filename", "in.dpx", "--dst_filename", "out.yuv", "--src_transfer_characteristics", "8",
         "--dst_transfer_characteristics", "16", "--src_pic_width", str(w), "--src_pic_height", str(h), "--src_bit_depth", "32",
         "--dst_bit_depth", "10", "--src_chroma_format_idc", "3", "--dst_chroma_format_idc", "1", "--src_matrix_coeffs", "0",
         "--dst_matrix_coeffs", "9", "--src_colour_primaries", "1", "--dst_colour_primaries", "9", "--src_video_full_range_flag", "1",
         "--dst_video_full_range_flag", "0", "--chroma_resampler_type", "1"], cwd=tmp_path)
    got = np.fromfile(tmp_path / "out.yuv", np.uint16)
    src = dict(bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0)
    dst = dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1)
    want = O.forward(planes, src, dst, backend="port")
    assert got.size == want.size and np.array_equal(got, want), (bits, big_endian, int((got != want).sum()))
