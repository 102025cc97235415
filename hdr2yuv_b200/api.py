"""Host-side mirror of the reference's function boundary (hdr.h:420-423, 439-443) over the C-ABI.

Same names and argument meaning as the reference: pic_stats, matrix_convert, convert,
write_yuv (compute half) plus the fused `forward` / `inverse` calls a drop-in host makes.
Device buffers are torch CUDA tensors (PyTorch is plumbing: device memory and streams only);
host buffers are numpy arrays.  Every call goes through libhdr2yuv_b200.so; nothing here
computes pixels on the CPU.
"""
import ctypes as C

import numpy as np

from . import _cabi as cabi
from ._cabi import (ForwardParams, InverseParams, PicDesc, PicStats, check, lib)


def pic_desc(width, height, chroma_format_idc=cabi.CHROMA_444, transfer=0, primaries=0, matrix=0, bit_depth=16,
             full_range=0, pic_buffer_type=cabi.PIC_TYPE_U16, layout=cabi.LAYOUT_PLANAR_U16):
    return PicDesc(width, height, chroma_format_idc, transfer, primaries, matrix, bit_depth, full_range,
                   pic_buffer_type, layout)


def forward_params(width, height, layout, src, dst, resampler=1, clip_on_load=0):
    """src/dst: dicts with bit_depth, full_range, transfer, primaries, matrix (+ dst chroma)."""
    is_u16 = layout in (cabi.LAYOUT_PLANAR_U16, cabi.LAYOUT_RGB16, cabi.LAYOUT_RGBA16)
    s = pic_desc(width, height, cabi.CHROMA_444, src["transfer"], src["primaries"], src["matrix"], src["bit_depth"],
                 src["full_range"], cabi.PIC_TYPE_U16 if is_u16 else cabi.PIC_TYPE_F32, layout)
    d = pic_desc(width, height, dst["chroma"], dst["transfer"], dst["primaries"], dst["matrix"], dst["bit_depth"],
                 dst["full_range"], cabi.PIC_TYPE_U16, cabi.LAYOUT_PLANAR_U16)
    return ForwardParams(s, d, resampler, clip_on_load)


def set_pic_clip(bit_depth, full_range):
    """set_pic_clip (common.cpp:300-327)."""
    out = cabi.ClipLimits()
    check(lib().h2y_set_pic_clip(bit_depth, full_range, C.byref(out)), "h2y_set_pic_clip")
    return out


def plane_dims(width, height, chroma):
    pw, ph = (C.c_int * 3)(), (C.c_int * 3)()
    check(lib().h2y_plane_dims(width, height, chroma, C.byref(pw), C.byref(ph)), "h2y_plane_dims")
    return list(pw), list(ph)


def src_frame_bytes(desc):
    return lib().h2y_src_frame_bytes(C.byref(desc))


def yuv_frame_bytes(width, height, chroma):
    return lib().h2y_yuv_frame_bytes(width, height, chroma)


def _stream_ptr(stream):
    if stream is None:
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return C.c_void_p(int(stream))


def _planes(tensors):
    return cabi._P3(*[t.data_ptr() for t in tensors])


class Context:
    """One per GPU (h2y_ctx): scratch device memory, LUTs, pipeline streams."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().h2y_ctx_create(int(device), C.byref(self._h)), "h2y_ctx_create")
        self.device = int(device)

    def close(self):
        if self._h:
            lib().h2y_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name=None, value=None):
        """h2y_ctx_set_option: a test / experiment switch of this context (name=None restores the defaults)."""
        check(lib().h2y_ctx_set_option(self._h, None if name is None else name.encode(),
                                       None if value is None else str(value).encode()), "h2y_ctx_set_option")

    @property
    def kernel_launches(self):
        return int(lib().h2y_kernel_launches(self._h))

    def profile_enable(self, on=True):
        check(lib().h2y_profile_enable(self._h, int(on)), "h2y_profile_enable")

    def profile_last_ms(self):
        """(dominant kernel ms, prologue kernels ms) of the last bracketed forward/inverse call."""
        a, b = C.c_float(0), C.c_float(0)
        check(lib().h2y_profile_last_ms(self._h, C.byref(a), C.byref(b)), "h2y_profile_last_ms")
        return a.value, b.value

    # ---- staged: one call per reference function ---------------------------------------------
    def pic_stats(self, pic, planes, stream=None):
        """pic_stats (common.cpp:66-168).  planes: three CUDA tensors."""
        out = PicStats()
        pl = _planes(planes)
        check(lib().h2y_pic_stats(self._h, C.byref(pic), C.byref(pl), C.byref(out), _stream_ptr(stream)),
              "h2y_pic_stats")
        return out

    def matrix_convert(self, out_pic, out_planes, in_pic, in_planes, in_stats=None, stream=None):
        """matrix_convert (convert.cpp:879-1315)."""
        po, pi = _planes(out_planes), _planes(in_planes)
        st = C.byref(in_stats) if in_stats is not None else None
        check(lib().h2y_matrix_convert(self._h, C.byref(out_pic), C.byref(po), C.byref(in_pic), C.byref(pi), st,
                                       _stream_ptr(stream)), "h2y_matrix_convert")

    def convert(self, out_pic, out_planes, in_pic, in_planes, chroma_resampler_type, stream=None):
        """convert (convert.cpp:513-874)."""
        po, pi = _planes(out_planes), _planes(in_planes)
        check(lib().h2y_convert(self._h, C.byref(out_pic), C.byref(po), C.byref(in_pic), C.byref(pi),
                                int(chroma_resampler_type), _stream_ptr(stream)), "h2y_convert")

    def write_yuv_clamp(self, pic, planes, src_bit_depth, stream=None):
        """compute half of write_yuv (tiff.cpp:457-550), in place."""
        pl = _planes(planes)
        check(lib().h2y_write_yuv_clamp(self._h, C.byref(pic), C.byref(pl), int(src_bit_depth), _stream_ptr(stream)),
              "h2y_write_yuv_clamp")

    def subsample_420_to_444(self, src, dst, width, height, algorithm, minCV, maxCV, stream=None):
        """Subsample420to444 (yuv2tiff.cpp:575-692), row-major planes."""
        check(lib().h2y_subsample_420_to_444(self._h, src.data_ptr(), dst.data_ptr(), width, height, int(algorithm),
                                             minCV, maxCV, _stream_ptr(stream)), "h2y_subsample_420_to_444")

    def matrix_inverse(self, out_pic, out_planes, in_pic, in_planes, invalid=None, stream=None):
        """matrix_inverse (convert.cpp:1320-1867): contiguous U16 planes Y,Cb,Cr -> G,B,R."""
        po, pi = _planes(out_planes), _planes(in_planes)
        check(lib().h2y_matrix_inverse(self._h, C.byref(out_pic), C.byref(po), C.byref(in_pic), C.byref(pi),
                                       invalid.data_ptr() if invalid is not None else None, _stream_ptr(stream)),
              "h2y_matrix_inverse")

    def write_tiff_rows(self, pic, planes, src_bit_depth, rgb, stream=None):
        """compute half of write_tiff (tiff.cpp:559-652): planes G,B,R -> interleaved R,G,B."""
        pl = _planes(planes)
        check(lib().h2y_write_tiff_rows(self._h, C.byref(pic), C.byref(pl), int(src_bit_depth), rgb.data_ptr(),
                                        _stream_ptr(stream)), "h2y_write_tiff_rows")

    def inverse444_host(self, in_pic, out_bit_depth, yuv, rgb, nframes, invalid=None):
        """.yuv 4:4:4 frames -> RGB16 rows as main() + write_tiff do (hdr2yuv.cpp:803-821, 898-933)."""
        n = in_pic.width * in_pic.height * 6
        check(lib().h2y_inverse444_host(self._h, C.byref(in_pic), int(out_bit_depth), _addr(yuv), n, _addr(rgb), n,
                                        int(nframes), _addr(invalid) if invalid is not None else None),
              "h2y_inverse444_host")

    # ---- fused ----------------------------------------------------------------------------------
    def forward(self, params, src, dst, nframes, src_stride=None, dst_stride=None, stream=None):
        """pic_stats -> matrix_convert -> convert -> write_yuv clamp on frames resident in HBM."""
        ss = src_stride or src_frame_bytes(params.src)
        ds = dst_stride or yuv_frame_bytes(params.src.width, params.src.height, params.dst.chroma_format_idc)
        check(lib().h2y_forward(self._h, C.byref(params), src.data_ptr(), ss, dst.data_ptr(), ds, int(nframes),
                                _stream_ptr(stream)), "h2y_forward")

    def forward_host(self, params, src, dst, nframes, src_stride=None, dst_stride=None):
        """Same through host buffers (numpy arrays or raw addresses); H2D/compute/D2H pipelined."""
        ss = src_stride or src_frame_bytes(params.src)
        ds = dst_stride or yuv_frame_bytes(params.src.width, params.src.height, params.dst.chroma_format_idc)
        check(lib().h2y_forward_host(self._h, C.byref(params), _addr(src), ss, _addr(dst), ds, int(nframes)),
              "h2y_forward_host")

    def forward_f32_host(self, params, src, dst, nframes, src_stride=None, dst_stride=None):
        """F32 destinations (.exr / .dpx outputs): pic_stats -> matrix_convert into float planes G, B, R (h2y_forward_f32_host)."""
        ss = src_stride or src_frame_bytes(params.src)
        ds = dst_stride or params.src.width * params.src.height * 12
        check(lib().h2y_forward_f32_host(self._h, C.byref(params), _addr(src), ss, _addr(dst), ds, int(nframes)), "h2y_forward_f32_host")

    def forward_last_plan_reuse(self):
        """(attempted, nframes, nredone) of the last forward call's last group (h2y_forward_last_plan_reuse)."""
        a, n, r = C.c_int(0), C.c_int(0), C.c_int(0)
        check(lib().h2y_forward_last_plan_reuse(self._h, C.byref(a), C.byref(n), C.byref(r)), "h2y_forward_last_plan_reuse")
        return bool(a.value), n.value, r.value

    def forward_last_stats(self, frame=0):
        out = PicStats()
        check(lib().h2y_forward_last_stats(self._h, frame, C.byref(out)), "h2y_forward_last_stats")
        return out

    def pq_codes_to_linear(self, codes, linear, stream=None):
        """optional linear-light stage behind the inverse path: PQ10000_f(code / 65535) per u16 sample (h2y_pq_codes_to_linear)."""
        check(lib().h2y_pq_codes_to_linear(self._h, codes.data_ptr(), int(codes.numel()), linear.data_ptr(), _stream_ptr(stream)),
              "h2y_pq_codes_to_linear")

    def inverse(self, params, yuv, rgb, nframes, invalid=None, yuv_stride=None, rgb_stride=None, stream=None):
        """one yuv2tiff main-loop iteration per frame (yuv2tiff.cpp:278-552), device resident."""
        ys = yuv_stride or yuv_frame_bytes(params.width, params.height, cabi.CHROMA_420)
        rs = rgb_stride or lib().h2y_rgb_frame_bytes(C.byref(params))
        check(lib().h2y_inverse(self._h, C.byref(params), yuv.data_ptr(), ys, rgb.data_ptr(), rs, int(nframes),
                                invalid.data_ptr() if invalid is not None else None, _stream_ptr(stream)),
              "h2y_inverse")

    def inverse_host(self, params, yuv, rgb, nframes, invalid=None, yuv_stride=None, rgb_stride=None):
        ys = yuv_stride or yuv_frame_bytes(params.width, params.height, cabi.CHROMA_420)
        rs = rgb_stride or lib().h2y_rgb_frame_bytes(C.byref(params))
        check(lib().h2y_inverse_host(self._h, C.byref(params), _addr(yuv), ys, _addr(rgb), rs, int(nframes),
                                     _addr(invalid) if invalid is not None else None), "h2y_inverse_host")


def _addr(x):
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return C.c_void_p(x.ctypes.data)
    return C.c_void_p(int(x))


class PinnedBuffer:
    """numpy view over h2y_host_alloc memory (cudaHostAlloc), for asynchronous PCIe copies."""

    def __init__(self, nbytes):
        self.ptr = lib().h2y_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError("h2y_host_alloc(%d) failed" % nbytes)
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def view(self, dtype):
        return self.array.view(dtype)

    def free(self):
        if self.ptr:
            self.array = None
            lib().h2y_host_free(self.ptr)
            self.ptr = None
