"""ctypes binding of include/hdr2yuv_b200.h -- the same stub a reference maintainer would add
(see INTEGRATION.md).  No CPU fallback: if the library is missing or was not built, importing
this module raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# H2Y_LIB selects an experiment build of the same library (tools/build_variants.py); the default is the in-tree product
LIB_PATH = os.environ.get("H2Y_LIB") or os.path.join(HERE, "libhdr2yuv_b200.so")

# status codes (include/hdr2yuv_b200.h)
OK, ERR_PRECONDITION, ERR_MATRIX, ERR_BIT_DEPTH, ERR_UNSUPPORTED, ERR_ARG, ERR_CUDA, ERR_NOMEM = range(8)

CHROMA_400, CHROMA_420, CHROMA_422, CHROMA_444 = 0, 1, 2, 3
PIC_TYPE_U16, PIC_TYPE_F32 = 1, 2
TRANSFER_BT709, TRANSFER_BT601, TRANSFER_LINEAR, TRANSFER_BT2020_10bit, TRANSFER_BT2020_12bit = 1, 6, 8, 14, 15
TRANSFER_PQ, TRANSFER_RHO_GAMMA = 16, 18
MATRIX_GBR, MATRIX_BT709, MATRIX_BT2020nc, MATRIX_BT2020c = 0, 1, 9, 10
MATRIX_YDzDx, MATRIX_YDzDx_Y500, MATRIX_YDzDx_Y100, MATRIX_YUVPRIME1, MATRIX_YUVPRIME2 = 11, 12, 13, 14, 15
LAYOUT_PLANAR_U16, LAYOUT_PLANAR_F32, LAYOUT_RGB16, LAYOUT_RGBA16, LAYOUT_HALF_RGB, LAYOUT_HALF_RGBA = range(6)
LAYOUT_DPX10_BE, LAYOUT_DPX10_LE = 6, 7
LAYOUT_DPX16_BE, LAYOUT_DPX16_LE, LAYOUT_DPXF32_BE, LAYOUT_DPXF32_LE = 8, 9, 10, 11
INV_YDzDx, INV_709, INV_2020, INV_Y100, INV_Y500 = range(5)


class PicDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("chroma_format_idc", C.c_int32),
                ("transfer_characteristics", C.c_int32), ("colour_primaries", C.c_int32),
                ("matrix_coeffs", C.c_int32), ("bit_depth", C.c_int32), ("video_full_range_flag", C.c_int32),
                ("pic_buffer_type", C.c_int32), ("layout", C.c_int32)]


class ClipLimits(C.Structure):
    _fields_ = [("minCV", C.c_uint32), ("maxCV", C.c_uint32), ("minVR", C.c_uint16), ("maxVR", C.c_uint16),
                ("minVRC", C.c_uint16), ("maxVRC", C.c_uint16), ("Half", C.c_uint16), ("_pad", C.c_uint16)]


class PicStats(C.Structure):
    _fields_ = [("f_min", C.c_float * 3), ("f_max", C.c_float * 3), ("i_min", C.c_uint16 * 3),
                ("i_max", C.c_uint16 * 3), ("estimated_ceiling", C.c_int32 * 3), ("estimated_floor", C.c_int32 * 3)]


class ForwardParams(C.Structure):
    _fields_ = [("src", PicDesc), ("dst", PicDesc), ("chroma_resampler_type", C.c_int32),
                ("clip_on_load", C.c_int32)]


class InverseParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bit_depth", C.c_int32), ("matrix", C.c_int32),
                ("fir", C.c_int32), ("full_range", C.c_int32), ("alpha", C.c_int32), ("ybar", C.c_int32)]


_P3 = C.c_void_p * 3

# every symbol include/hdr2yuv_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "h2y_abi_version": (C.c_int, []),
    "h2y_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "h2y_ctx_destroy": (C.c_int, [C.c_void_p]),
    "h2y_ctx_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "h2y_status_string": (C.c_char_p, [C.c_int]),
    "h2y_last_cuda_error": (C.c_int, [C.c_void_p]),
    "h2y_host_alloc": (C.c_void_p, [C.c_size_t]),
    "h2y_host_free": (None, [C.c_void_p]),
    "h2y_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "h2y_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "h2y_profile_last_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "h2y_set_pic_clip": (C.c_int, [C.c_int, C.c_int, C.POINTER(ClipLimits)]),
    "h2y_plane_dims": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int * 3), C.POINTER(C.c_int * 3)]),
    "h2y_frame_range": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "h2y_src_frame_bytes": (C.c_size_t, [C.POINTER(PicDesc)]),
    "h2y_yuv_frame_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "h2y_tmp_bit_depth": (C.c_int, [C.POINTER(PicDesc), C.POINTER(PicDesc)]),
    "h2y_pic_stats": (C.c_int, [C.c_void_p, C.POINTER(PicDesc), C.POINTER(_P3), C.POINTER(PicStats), C.c_void_p]),
    "h2y_matrix_convert": (C.c_int, [C.c_void_p, C.POINTER(PicDesc), C.POINTER(_P3), C.POINTER(PicDesc),
                                     C.POINTER(_P3), C.POINTER(PicStats), C.c_void_p]),
    "h2y_convert": (C.c_int, [C.c_void_p, C.POINTER(PicDesc), C.POINTER(_P3), C.POINTER(PicDesc), C.POINTER(_P3),
                              C.c_int, C.c_void_p]),
    "h2y_write_yuv_clamp": (C.c_int, [C.c_void_p, C.POINTER(PicDesc), C.POINTER(_P3), C.c_int, C.c_void_p]),
    "h2y_subsample_420_to_444": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.c_uint16, C.c_uint16, C.c_void_p]),
    "h2y_matrix_inverse": (C.c_int, [C.c_void_p, C.POINTER(PicDesc), C.POINTER(_P3), C.POINTER(PicDesc), C.POINTER(_P3),
                                     C.c_void_p, C.c_void_p]),
    "h2y_write_tiff_rows": (C.c_int, [C.c_void_p, C.POINTER(PicDesc), C.POINTER(_P3), C.c_int, C.c_void_p, C.c_void_p]),
    "h2y_inverse444_host": (C.c_int, [C.c_void_p, C.POINTER(PicDesc), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                      C.c_size_t, C.c_int, C.c_void_p]),
    "h2y_forward": (C.c_int, [C.c_void_p, C.POINTER(ForwardParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                              C.c_int, C.c_void_p]),
    "h2y_forward_host": (C.c_int, [C.c_void_p, C.POINTER(ForwardParams), C.c_void_p, C.c_size_t, C.c_void_p,
                                   C.c_size_t, C.c_int]),
    "h2y_forward_f32_host": (C.c_int, [C.c_void_p, C.POINTER(ForwardParams), C.c_void_p, C.c_size_t, C.c_void_p,
                                   C.c_size_t, C.c_int]),
    "h2y_forward_last_stats": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(PicStats)]),
    "h2y_forward_last_plan_reuse": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "h2y_pq_codes_to_linear": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "h2y_rgb_frame_bytes": (C.c_size_t, [C.POINTER(InverseParams)]),
    "h2y_inverse": (C.c_int, [C.c_void_p, C.POINTER(InverseParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                              C.c_int, C.c_void_p, C.c_void_p]),
    "h2y_inverse_host": (C.c_int, [C.c_void_p, C.POINTER(InverseParams), C.c_void_p, C.c_size_t, C.c_void_p,
                                   C.c_size_t, C.c_int, C.c_void_p]),
}

_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("hdr2yuv_b200: %s is missing -- run `python -m hdr2yuv_b200.build` "
                              "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)          # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        if l.h2y_abi_version() != 2:
            raise ImportError("hdr2yuv_b200: ABI version mismatch")
        _lib = l
    return _lib


class H2YError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        msg = lib().h2y_status_string(status).decode()
        super().__init__("%s: %s (h2y_status %d)" % (where, msg, status))


def check(status, where):
    if status != OK:
        raise H2YError(status, where)
