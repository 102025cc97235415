"""numpy-facing loader for the parity checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package hdr2yuv_b200 never does.

Two backends with the same call shapes:
  port -- oracle/libh2yoracle.so, the plain-C restatement (h2y_oracle.c)
  ref  -- oracle/_ref/libh2yref.so + yuv2tiff_ref, the reference's own sources compiled unmodified
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

CHROMA_420, CHROMA_422, CHROMA_444 = 1, 2, 3
PIC_U16, PIC_F32 = 1, 2
TRANSFER_BT709, TRANSFER_BT601, TRANSFER_LINEAR, TRANSFER_PQ, TRANSFER_RHO_GAMMA = 1, 6, 8, 16, 18
MATRIX_GBR, MATRIX_BT709, MATRIX_BT2020NC, MATRIX_YDZDX, MATRIX_Y500, MATRIX_Y100 = 0, 1, 9, 11, 12, 13
INV_YDZDX, INV_709, INV_2020, INV_Y100, INV_Y500 = 0, 1, 2, 3, 4
_INV_KEYWORD = {INV_YDZDX: None, INV_709: "709", INV_2020: "2020", INV_Y100: "Y100", INV_Y500: "Y500"}


class _Clip(C.Structure):
    _fields_ = [("minCV", C.c_ulong), ("maxCV", C.c_ulong), ("minVR", C.c_ushort), ("maxVR", C.c_ushort),
                ("minVRC", C.c_ushort), ("maxVRC", C.c_ushort), ("Half", C.c_ushort)]


class _Pic(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("chroma_format_idc", C.c_int),
                ("transfer_characteristics", C.c_int), ("colour_primaries", C.c_int), ("matrix_coeffs", C.c_int),
                ("bit_depth", C.c_int), ("video_full_range_flag", C.c_int), ("pic_buffer_type", C.c_int),
                ("est_floor", C.c_int * 3), ("est_ceiling", C.c_int * 3),
                ("buf", C.c_void_p * 3), ("fbuf", C.c_void_p * 3)]


class _Dst(C.Structure):
    _fields_ = [("dst_bit_depth", C.c_int), ("dst_full_range", C.c_int), ("dst_transfer", C.c_int),
                ("dst_primaries", C.c_int), ("dst_matrix", C.c_int), ("dst_chroma", C.c_int),
                ("resampler", C.c_int)]


class _Inv(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("bit_depth", C.c_int), ("matrix", C.c_int),
                ("fir", C.c_int), ("full_range", C.c_int), ("alpha", C.c_int), ("ybar", C.c_int)]


def _build():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_h2y_oracle_build", os.path.join(HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_port = None
_ref = None


def port_lib():
    global _port
    if _port is None:
        _port = C.CDLL(_build().build_oracle())
        _port.orc_half_to_float.restype = C.c_float
        _port.orc_yuv2tiff_frame.restype = C.c_long
        for n in ("orc_pq_eotf", "orc_pq_oetf", "orc_bt1886_eotf", "orc_bt1886_oetf", "orc_rho_gamma_eotf",
                  "orc_rho_gamma_oetf"):
            getattr(_port, n).restype = C.c_float
            getattr(_port, n).argtypes = [C.c_float]
    return _port


def ref_available():
    b = _build()
    return b.build_ref() is not None and os.path.exists(b.REF_YUV2TIFF)


def ref_lib():
    global _ref
    if _ref is None:
        path = _build().build_ref()
        if path is None:
            raise RuntimeError("compiled reference (oracle/_ref) is not available here")
        _ref = C.CDLL(path)
    return _ref


def chroma_dims(w, h, chroma):
    cw = w if chroma == CHROMA_444 else w >> 1
    ch = h >> 1 if chroma == CHROMA_420 else h
    return cw, ch


def yuv_frame_samples(w, h, chroma):
    cw, ch = chroma_dims(w, h, chroma)
    return w * h + 2 * cw * ch


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------
def forward(planes, src, dst, backend="port", want_tmp=False, timing=None):
    """planes: (3,H,W) uint16 or float32 in G,B,R order.  src/dst: dicts with the pic_t fields
    (bit_depth, full_range, transfer, primaries, matrix) + dst chroma, resampler.
    Returns the .yuv frame (flat uint16: Y, Cb, Cr) [, tmp444 (3,H,W), stats(6)]."""
    planes = np.ascontiguousarray(planes)
    _, h, w = planes.shape
    is_f32 = planes.dtype == np.float32
    assert is_f32 or planes.dtype == np.uint16
    out = np.zeros(yuv_frame_samples(w, h, dst["chroma"]), np.uint16)
    if backend == "ref":
        cfg = (C.c_int * 15)(w, h, int(is_f32), src["bit_depth"], src["full_range"], src["transfer"],
                             src["primaries"], src["matrix"], dst["bit_depth"], dst["full_range"], dst["transfer"],
                             dst["primaries"], dst["matrix"], dst["chroma"], dst["resampler"])
        pl = (C.c_void_p * 3)(*[planes[c].ctypes.data for c in range(3)])
        tmp = np.zeros((3, h, w), np.uint16)
        stats = (C.c_int * 6)()
        secs = C.c_double(0)
        rc = ref_lib().ref_forward_frame(cfg, pl, _ptr(out), _ptr(tmp), stats, C.byref(secs))
        if rc:
            raise RuntimeError("reference forward chain returned %d" % rc)
        if timing is not None:
            timing.append(secs.value)
        return (out, tmp, np.array(list(stats))) if want_tmp else out
    lib = port_lib()
    pic = _Pic(w, h, CHROMA_444, src["transfer"], src["primaries"], src["matrix"], src["bit_depth"],
               src["full_range"], PIC_F32 if is_f32 else PIC_U16)
    for c in range(3):
        (pic.fbuf if is_f32 else pic.buf)[c] = planes[c].ctypes.data
    d = _Dst(dst["bit_depth"], dst["full_range"], dst["transfer"], dst["primaries"], dst["matrix"], dst["chroma"],
             dst["resampler"])
    if want_tmp:
        lib.orc_pic_stats(C.byref(pic), None, None)
        tmp = np.zeros((3, h, w), np.uint16)
        t = _Pic(w, h, CHROMA_444, dst["transfer"], dst["primaries"], dst["matrix"],
                 dst["bit_depth"] if is_f32 else src["bit_depth"], dst["full_range"], PIC_U16)
        for c in range(3):
            t.buf[c] = tmp[c].ctypes.data
        rc = lib.orc_matrix_convert(C.byref(t), C.byref(pic))
        if rc:
            raise RuntimeError("oracle matrix_convert returned %d" % rc)
        stats = np.array(list(pic.est_floor) + list(pic.est_ceiling))
    rc = lib.orc_forward_frame(C.byref(pic), C.byref(d), _ptr(out))
    if rc:
        raise RuntimeError("oracle forward chain returned %d" % rc)
    return (out, tmp, stats) if want_tmp else out


def matrix_convert_f32(planes, src, dst, backend="port"):
    """matrix_convert into an F32 picture (convert.cpp:1222-1304; what an .exr / .dpx destination runs, hdr2yuv.cpp:797-823).
    planes: (3,H,W) uint16 or float32, G,B,R.  Returns (3,H,W) float32.  The tmp depth is the reference's
    (hdr2yuv.cpp:805-808): the source depth for an F32 source, the destination depth for an integer source."""
    planes = np.ascontiguousarray(planes)
    _, h, w = planes.shape
    is_f32 = planes.dtype == np.float32
    out = np.zeros((3, h, w), np.float32)
    if backend == "ref":
        cfg = (C.c_int * 15)(w, h, int(is_f32), src["bit_depth"], src["full_range"], src["transfer"], src["primaries"],
                             src["matrix"], dst["bit_depth"], dst["full_range"], dst["transfer"], dst["primaries"],
                             dst["matrix"], 3, 0)
        pl = (C.c_void_p * 3)(*[planes[c].ctypes.data for c in range(3)])
        rc = ref_lib().ref_matrix_convert_f32out(cfg, pl, _ptr(out))
        if rc:
            raise RuntimeError("reference matrix_convert returned %d" % rc)
        return out
    lib = port_lib()
    pic = _Pic(w, h, CHROMA_444, src["transfer"], src["primaries"], src["matrix"], src["bit_depth"], src["full_range"],
               PIC_F32 if is_f32 else PIC_U16)
    t = _Pic(w, h, CHROMA_444, dst["transfer"], dst["primaries"], dst["matrix"], src["bit_depth"] if is_f32 else dst["bit_depth"],
             dst["full_range"], PIC_F32)
    for c in range(3):
        (pic.fbuf if is_f32 else pic.buf)[c] = planes[c].ctypes.data
        t.fbuf[c] = out[c].ctypes.data
    lib.orc_pic_stats(C.byref(pic), None, None)
    rc = lib.orc_matrix_convert(C.byref(t), C.byref(pic))
    if rc:
        raise RuntimeError("oracle matrix_convert returned %d" % rc)
    return out


def pic_stats(planes, bit_depth, backend="port"):
    planes = np.ascontiguousarray(planes)
    _, h, w = planes.shape
    is_f32 = planes.dtype == np.float32
    if backend == "ref":
        pl = (C.c_void_p * 3)(*[planes[c].ctypes.data for c in range(3)])
        fc = (C.c_int * 6)()
        ref_lib().ref_pic_stats(int(is_f32), bit_depth, w, h, pl, fc)
        return np.array(list(fc))
    pic = _Pic(w, h, CHROMA_444, 0, 0, 0, bit_depth, 1, PIC_F32 if is_f32 else PIC_U16)
    for c in range(3):
        (pic.fbuf if is_f32 else pic.buf)[c] = planes[c].ctypes.data
    port_lib().orc_pic_stats(C.byref(pic), None, None)
    return np.array(list(pic.est_floor) + list(pic.est_ceiling))


def subsample_fir(plane, lo, hi, backend="port"):
    plane = np.ascontiguousarray(plane, np.uint16)
    h, w = plane.shape
    out = np.zeros((h // 2, w // 2), np.uint16)
    fn = ref_lib().ref_subsample_fir if backend == "ref" else port_lib().orc_subsample_fir
    fn(_ptr(out), _ptr(plane), w, h, C.c_ulong(lo), C.c_ulong(hi))
    return out


def subsample_fir_h(plane, lo, hi):
    plane = np.ascontiguousarray(plane, np.uint16)
    h, w = plane.shape
    out = np.zeros((h, w // 2), np.uint16)
    port_lib().orc_subsample_fir_h(_ptr(out), _ptr(plane), w, h, C.c_ulong(lo), C.c_ulong(hi))
    return out


def subsample_box(plane, backend="port"):
    plane = np.ascontiguousarray(plane, np.uint16)
    h, w = plane.shape
    out = np.zeros((h // 2, w // 2), np.uint16)
    fn = ref_lib().ref_subsample_box if backend == "ref" else port_lib().orc_subsample_box
    fn(_ptr(out), _ptr(plane), w, h)
    return out


def upsample_420to444(plane, fir, lo, hi, backend="port"):
    plane = np.ascontiguousarray(plane, np.uint16)
    hh, wh = plane.shape
    out = np.zeros((2 * hh, 2 * wh), np.uint16)
    fn = ref_lib().ref_upsample_420to444 if backend == "ref" else port_lib().orc_upsample_420to444
    fn(_ptr(plane), _ptr(out), 2 * wh, 2 * hh, int(fir), C.c_ushort(lo), C.c_ushort(hi))
    return out


_TRANSFER_NAMES = ["orc_pq_eotf", "orc_pq_oetf", "orc_bt1886_eotf", "orc_bt1886_oetf", "orc_rho_gamma_eotf",
                   "orc_rho_gamma_oetf"]


def transfer(which, x, backend="port"):
    """which: 0 PQ_f 1 PQ_r 2 bt1886_f 3 bt1886_r 4 rho_f 5 rho_r (float32 in -> float32 out)."""
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    if backend == "ref":
        ref_lib().ref_transfer(which, _ptr(x), _ptr(out), C.c_long(x.size))
        return out
    fn = getattr(port_lib(), _TRANSFER_NAMES[which])
    flat_in, flat_out = x.ravel(), out.ravel()
    for i in range(flat_in.size):
        flat_out[i] = fn(float(flat_in[i]))
    return out


def load_rgb16(rgb, full_range):
    """(H,W,C) interleaved R,G,B(,A) u16 -> (3,H,W) G,B,R planes with read_tiff's on-read clip."""
    rgb = np.ascontiguousarray(rgb, np.uint16)
    h, w, nch = rgb.shape
    planes = np.zeros((3, h, w), np.uint16)
    pl = (C.c_void_p * 3)(*[planes[c].ctypes.data for c in range(3)])
    port_lib().orc_load_rgb16(_ptr(rgb), h * w, nch, int(full_range), pl)
    return planes


def load_half(px):
    """(H,W,C) interleaved r,g,b(,a) half bit patterns (uint16) -> (3,H,W) float32 G,B,R planes."""
    px = np.ascontiguousarray(px, np.uint16)
    h, w, nch = px.shape
    planes = np.zeros((3, h, w), np.float32)
    pl = (C.c_void_p * 3)(*[planes[c].ctypes.data for c in range(3)])
    port_lib().orc_load_half(_ptr(px), h * w, nch, pl)
    return planes


def load_dpx10(words, big_endian):
    """What dpx_read + muxed_dpx_to_planar_float_buf leave in in_pic->fbuf for a 10-bit packed DPX (dpx.cpp:506-531,
    common.cpp:12-28): (H,W) stored 32-bit words -> (3,H,W) float32 planes G,B,R with sample = code / 1023.0."""
    w = np.ascontiguousarray(words).view(np.dtype(">u4") if big_endian else np.dtype("<u4")).astype(np.uint32)
    rr, gg, bb = w >> 22, (w >> 12) & 1023, (w >> 2) & 1023
    return np.stack([(c.astype(np.float64) / 1023.0).astype(np.float32) for c in (gg, bb, rr)], 0)


def load_dpx16(samples, big_endian):
    """16-bit DPX (dpx.cpp:478-494): (H,W,3) stored u16 R,G,B samples in file byte order -> (3,H,W) float32 planes G,B,R with
    sample = code / 65535.0 (double division rounded to float by the assignment)."""
    v = np.ascontiguousarray(samples).view(np.dtype(">u2") if big_endian else np.dtype("<u2")).astype(np.float64)
    f = (v / 65535.0).astype(np.float32)
    return np.ascontiguousarray(np.stack([f[..., 1], f[..., 2], f[..., 0]], 0))


def load_dpxf32(samples, big_endian):
    """32-bit float DPX (dpx.cpp:412-443): (H,W,3) stored floats R,G,B in file byte order, taken as they are."""
    f = np.ascontiguousarray(samples).view(np.dtype(">f4") if big_endian else np.dtype("<f4")).astype(np.float32)
    return np.ascontiguousarray(np.stack([f[..., 1], f[..., 2], f[..., 0]], 0))


def yuv2tiff(yuv, w, h, bit_depth=12, matrix=INV_YDZDX, fir=True, full_range=False, alpha=False, backend="port", ybar=False):
    """One 4:2:0 frame (flat u16 Y,Cb,Cr) -> (H,W,3|4) interleaved RGB16.  Returns (rgb, invalid_pixels)."""
    yuv = np.ascontiguousarray(yuv, np.uint16)
    nch = 4 if alpha else 3
    if backend == "ref":
        # the reference program only knows 3840x2160 / 1920x1080 / 960x540 (yuv2tiff.cpp:188-198)
        size_kw = {(3840, 2160): None, (1920, 1080): "HD1920", (960, 540): "HD960"}[(w, h)]
        args = [k for k in (size_kw, {10: "B10", 12: None, 14: "B14"}[bit_depth], _INV_KEYWORD[matrix],
                            None if fir else "BOX", "FULL" if full_range else None, "ALPHA" if alpha else None,
                            "-X" if ybar else None) if k]
        exe = os.path.join(HERE, "_ref", "yuv2tiff_ref")
        with tempfile.TemporaryDirectory() as d:
            os.mkdir(os.path.join(d, "tifXYZ"))
            yuv.tofile(os.path.join(d, "in.yuv"))
            r = subprocess.run([exe, "in.yuv"] + args + ["-f", "1"], cwd=d, stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT)
            text = r.stdout.decode(errors="replace")
            rgb = np.fromfile(os.path.join(d, "tifXYZ", "XpYpZp00000.tif"), np.uint16).reshape(h, w, nch)
        invalid = [int(l.split(":")[1]) for l in text.splitlines() if l.startswith("Invalid Pixels:")]
        return rgb, (invalid[0] if invalid else -1)
    p = _Inv(w, h, bit_depth, matrix, int(fir), int(full_range), int(alpha), int(ybar))
    rgb = np.zeros((h, w, nch), np.uint16)
    inv = port_lib().orc_yuv2tiff_frame(C.byref(p), _ptr(yuv), _ptr(rgb))
    return rgb, int(inv)


def matrix_inverse(planes, matrix_coeffs, in_bit_depth, in_full_range, out_bit_depth, backend="port"):
    """hdr2yuv's 4:4:4 inverse (convert.cpp:1320-1867): (3,H,W) uint16 Y,Cb,Cr -> (3,H,W) uint16 G,B,R.
    Returns (planes, invalid_pixels); invalid_pixels is None for the compiled reference (it only prints it)."""
    planes = np.ascontiguousarray(planes, np.uint16)
    _, h, w = planes.shape
    out = np.zeros((3, h, w), np.uint16)
    pi = (C.c_void_p * 3)(*[planes[c].ctypes.data for c in range(3)])
    po = (C.c_void_p * 3)(*[out[c].ctypes.data for c in range(3)])
    if backend == "ref":
        cfg = (C.c_int * 6)(w, h, matrix_coeffs, in_bit_depth, in_full_range, out_bit_depth)
        rc = ref_lib().ref_matrix_inverse(cfg, pi, po)
        if rc:
            raise RuntimeError("reference matrix_inverse returned %d" % rc)
        return out, None
    lib = port_lib()
    lib.orc_matrix_inverse.restype = C.c_long
    inv = lib.orc_matrix_inverse(po, pi, w, h, matrix_coeffs, in_bit_depth, in_full_range, out_bit_depth)
    if inv == -2:
        raise RuntimeError("Can't determine color difference to use?")
    return out, int(inv)


def write_tiff_rows(planes, pic_bit_depth, src_bit_depth):
    """write_tiff's compute (tiff.cpp:605-628): (3,H,W) G,B,R -> (H,W,3) interleaved R,G,B, << depth difference."""
    planes = np.ascontiguousarray(planes, np.uint16)
    _, h, w = planes.shape
    rgb = np.zeros((h, w, 3), np.uint16)
    pi = (C.c_void_p * 3)(*[planes[c].ctypes.data for c in range(3)])
    port_lib().orc_write_tiff_rows(_ptr(rgb), pi, C.c_long(h * w), pic_bit_depth, src_bit_depth)
    return rgb
