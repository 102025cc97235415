/*
 * hdr2yuv_b200.h -- C-ABI of the B200-native (sm_100a) hdr2yuv / yuv2tiff conversion hot path.
 *
 * This is the drop-in boundary.  The reference has no plugin/FFI layer: its
 * boundary is four C++ free functions over POD structs (hdr.h:420-423, 439-443).  Each entry
 * point below names the reference interface it replaces.  Plain pointers and sizes only; no
 * C++ or torch types.  All `d_*` pointers are CUDA device pointers, all `h_*` pointers host
 * memory (pinned memory from h2y_host_alloc gives asynchronous copies).  `stream` is a
 * cudaStream_t passed as void* (NULL = the legacy default stream).
 *
 * Error convention (reference: `return 1` + printf for precondition failures,
 * convert.cpp:523-528, 886-890; printf + exit(0) for fatal ones, convert.cpp:1196-1197,
 * tiff.cpp:399-400): every function returns an h2y_status; the library never exits, never
 * prints, and never falls back to a CPU implementation.
 *
 * Plane order everywhere is index 0/1/2 = G/B/R = Y/Cb/Cr (tiff.cpp:309-311, exr.cpp:233-235).
 */
#ifndef HDR2YUV_B200_H
#define HDR2YUV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H2Y_ABI_VERSION 2

typedef enum h2y_status {
    H2Y_OK = 0,
    H2Y_ERR_PRECONDITION = 1, /* the reference's `return(1)`: non-4:4:4 input to matrix_convert
                                 (convert.cpp:886-890), non-U16 pictures to convert (523-528) */
    H2Y_ERR_MATRIX = 2,       /* "Can't determine color difference to use?" (convert.cpp:1195-1198) */
    H2Y_ERR_BIT_DEPTH = 3,    /* dst bit depth > src bit depth in write_yuv (tiff.cpp:396-401) */
    H2Y_ERR_UNSUPPORTED = 4,  /* a combination the reference accepts that this path does not serve */
    H2Y_ERR_ARG = 5,          /* NULL pointer, bad geometry, unknown enum value */
    H2Y_ERR_CUDA = 6,         /* a CUDA runtime call failed; see h2y_last_cuda_error */
    H2Y_ERR_NOMEM = 7
} h2y_status;

/* hdr.h:14-17 */
enum { H2Y_CHROMA_400 = 0, H2Y_CHROMA_420 = 1, H2Y_CHROMA_422 = 2, H2Y_CHROMA_444 = 3 };
/* hdr.h:296-297 */
enum { H2Y_PIC_TYPE_U16 = 1, H2Y_PIC_TYPE_F32 = 2 };
/* hdr.h:104-134, values the hot path tests */
enum {
    H2Y_TRANSFER_BT709 = 1, H2Y_TRANSFER_BT601 = 6, H2Y_TRANSFER_LINEAR = 8, H2Y_TRANSFER_BT2020_10bit = 14,
    H2Y_TRANSFER_BT2020_12bit = 15, H2Y_TRANSFER_PQ = 16, H2Y_TRANSFER_RHO_GAMMA = 18
};
/* hdr.h:168-191 */
enum {
    H2Y_MATRIX_GBR = 0, H2Y_MATRIX_BT709 = 1, H2Y_MATRIX_BT2020nc = 9, H2Y_MATRIX_BT2020c = 10,
    H2Y_MATRIX_YDzDx = 11, H2Y_MATRIX_YDzDx_Y500 = 12, H2Y_MATRIX_YDzDx_Y100 = 13,
    H2Y_MATRIX_YUVPRIME1 = 14, H2Y_MATRIX_YUVPRIME2 = 15
};

/* How the samples of one source frame lie in memory.  The reference's readers de-interleave on
 * the host (tiff.cpp:265-315, exr.cpp:209-235, hdr2yuv.cpp:633-643); here the raw decoded
 * samples go to the GPU and the de-interleave is part of the kernel's load. */
typedef enum h2y_layout {
    H2Y_LAYOUT_PLANAR_U16 = 0, /* three W*H planes G,B,R back to back (pic_t.buf order, hdr.h:382) */
    H2Y_LAYOUT_PLANAR_F32 = 1, /* three W*H float planes G,B,R (pic_t.fbuf, hdr.h:383) */
    H2Y_LAYOUT_RGB16 = 2,      /* interleaved R,G,B u16: a TIFF strip row (tiff.cpp:276-278) */
    H2Y_LAYOUT_RGBA16 = 3,     /* interleaved R,G,B,A u16 (--alpha_channel, tiff.cpp:281-282) */
    H2Y_LAYOUT_HALF_RGB = 4,   /* interleaved r,g,b IEEE half */
    H2Y_LAYOUT_HALF_RGBA = 5,  /* interleaved Imf::Rgba {half r,g,b,a} (exr.cpp:143-192) */
    /* one 32-bit word per pixel as a 10-bit DPX file stores it, R[31:22] G[21:12] B[11:2] (dpx.cpp:506-531), in the
     * file's byte order (big-endian "SDPX", little-endian "XPDS").  The picture is F32: sample = code / 1023.0 */
    H2Y_LAYOUT_DPX10_BE = 6,
    H2Y_LAYOUT_DPX10_LE = 7,
    /* 16-bit DPX: interleaved R,G,B u16 in the file's byte order; the picture is F32, sample = code / 65535.0
     * (dpx.cpp:478-494).  (A 16-bit DPX read with the half-float flag, dpx.cpp:454-477, is H2Y_LAYOUT_HALF_RGB after a
     * byte swap on the host.) */
    H2Y_LAYOUT_DPX16_BE = 8,
    H2Y_LAYOUT_DPX16_LE = 9,
    /* 32-bit float DPX: interleaved R,G,B IEEE floats in the file's byte order, taken as they are (dpx.cpp:412-443) */
    H2Y_LAYOUT_DPXF32_BE = 10,
    H2Y_LAYOUT_DPXF32_LE = 11
} h2y_layout;

/* The fields of pic_t (hdr.h:359-392) the hot path reads. */
typedef struct h2y_pic_desc {
    int32_t width;
    int32_t height;
    int32_t chroma_format_idc;
    int32_t transfer_characteristics;
    int32_t colour_primaries;
    int32_t matrix_coeffs;
    int32_t bit_depth;
    int32_t video_full_range_flag;
    int32_t pic_buffer_type; /* H2Y_PIC_TYPE_* */
    int32_t layout;          /* h2y_layout; PLANAR_* for every staged call */
} h2y_pic_desc;

/* clip_limits_t (hdr.h:345-356) as set_pic_clip derives it (common.cpp:300-327). */
typedef struct h2y_clip_limits {
    uint32_t minCV, maxCV;
    uint16_t minVR, maxVR, minVRC, maxVRC, Half, _pad;
} h2y_clip_limits;

/* pic_stats_t (hdr.h:302-318) without the log-only averages. */
typedef struct h2y_pic_stats_s {
    float f_min[3], f_max[3];
    uint16_t i_min[3], i_max[3];
    int32_t estimated_ceiling[3];
    int32_t estimated_floor[3];
} h2y_pic_stats_t;

typedef struct h2y_ctx h2y_ctx; /* one per GPU; owns scratch device memory, LUTs, streams */

/* ---- context ------------------------------------------------------------------------------ */
int h2y_abi_version(void);
h2y_status h2y_ctx_create(int device, h2y_ctx **out);
h2y_status h2y_ctx_destroy(h2y_ctx *ctx);
const char *h2y_status_string(h2y_status s);
int h2y_last_cuda_error(const h2y_ctx *ctx);             /* cudaError_t of the last H2Y_ERR_CUDA */
void *h2y_host_alloc(size_t bytes);                       /* pinned host memory (NULL on failure) */
void h2y_host_free(void *p);
uint64_t h2y_kernel_launches(const h2y_ctx *ctx);         /* kernels this context has launched so far */

/* Threading and streams.  A context owns one set of scratch buffers (statistics, per-frame plans, LUTs, intermediate
 * planes), so it serves ONE call at a time: use it from one host thread at a time, and one context per thread that
 * wants to work concurrently.  Calls are asynchronous on the caller's stream; a call that arrives on a different stream
 * than the previous one first waits, on the device, for that previous call's work (an event), so streams may be mixed
 * safely but do not overlap within a context.  Growing a scratch buffer synchronises the device.  Every entry point
 * switches to the context's device and restores the caller's current device before it returns. */

/* Test and experiment switches.  The library reads the environment exactly once, in h2y_ctx_create; afterwards the
 * switches belong to the context and are changed with this call (tests force each kernel of a route this way).
 * Names: "H2Y_FORWARD_KERNEL" = "ring" | "rows" | "", "H2Y_INVERSE_KERNEL" = "tile" | "rows" | "exact" (the rows
 * kernel with every pixel through its exact integer routine) | "", "H2Y_FORCE_STAGED",
 * "H2Y_FORCE_V1", "H2Y_NO_SPECIALISED", "H2Y_EXACT_MATH" = "1" | "0", "H2Y_STATS_GX" = n, "H2Y_PLAN_REUSE" = "0" never,
 * "1" whenever a previous plan exists, "" automatic (see h2y_forward).  name == NULL restores the defaults of
 * h2y_ctx_create.  Unknown names return H2Y_ERR_ARG. */
h2y_status h2y_ctx_set_option(h2y_ctx *ctx, const char *name, const char *value);

/* Profiling hooks for bench.py: when enabled, h2y_forward / h2y_inverse bracket their kernels with
 * CUDA events on the caller's stream.  h2y_profile_last_ms waits for the bracketed calls made since
 * h2y_profile_enable (the 16 most recent) and returns the AVERAGE duration of their dominant kernel
 * (the fused forward / inverse kernel) and of the kernels before it (statistics + LUT build; 0 when
 * there are none).  A forward call of more than 256 frames brackets each 256-frame group. */
h2y_status h2y_profile_enable(h2y_ctx *ctx, int on);
h2y_status h2y_profile_last_ms(h2y_ctx *ctx, float *main_kernel_ms, float *prologue_ms);

/* ---- geometry / limits (pure host helpers, no GPU needed) -------------------------------- */
/* set_pic_clip, common.cpp:300-327 */
h2y_status h2y_set_pic_clip(int bit_depth, int video_full_range_flag, h2y_clip_limits *out);
/* init_pic plane geometry, common.cpp:191-198 */
h2y_status h2y_plane_dims(int width, int height, int chroma_format_idc, int plane_w[3], int plane_h[3]);
size_t h2y_src_frame_bytes(const h2y_pic_desc *src);      /* bytes of one source frame in its layout */
size_t h2y_yuv_frame_bytes(int width, int height, int chroma_format_idc); /* one .yuv frame: Y,Cb,Cr u16 */
/* Frame-range sharding (frames are independent: pic_stats is per frame, common.cpp:66; n_frames is unused,
 * hdr2yuv.cpp:157-160): worker `rank` of `world` converts frames [*lo, *hi) of `nframes`, contiguous, sizes differing
 * by at most one.  The one partition formula of the repo: the CLI's --devices threads, bench.py's ranks and
 * hdr2yuv_b200/sharding.py all call it. */
h2y_status h2y_frame_range(int rank, int world, int nframes, int *lo, int *hi);
/* tmp picture bit depth rule of main(), hdr2yuv.cpp:803-808 (destination of a .yuv is U16) */
int h2y_tmp_bit_depth(const h2y_pic_desc *src, const h2y_pic_desc *dst);

/* ---- staged entry points: one per reference function; planar device buffers --------------- */

/* pic_stats (common.cpp:66-168): per-plane extrema and the estimated floor/ceiling that feed
 * matrix_convert's normalisation.  `d_planes` are the three planes of a PLANAR_U16/F32 picture. */
h2y_status h2y_pic_stats(h2y_ctx *ctx, const h2y_pic_desc *pic, const void *const d_planes[3],
                         h2y_pic_stats_t *h_stats_out, void *stream);

/* matrix_convert (convert.cpp:879-1315): normalise -> transfer change -> range scale ->
 * colour-difference matrix -> offset -> clamp, 4:4:4 in, 4:4:4 out.  `in_stats` may be NULL
 * when in->transfer_characteristics == out->transfer_characteristics.  out->bit_depth is the
 * tmp picture's depth.  out->pic_buffer_type selects the U16 (1146-1220) or F32 (1222-1304) twin. */
h2y_status h2y_matrix_convert(h2y_ctx *ctx, const h2y_pic_desc *out, void *const d_out_planes[3],
                              const h2y_pic_desc *in, const void *const d_in_planes[3],
                              const h2y_pic_stats_t *in_stats, void *stream);

/* convert (convert.cpp:513-874): chroma resample of planes 1,2 + copy of plane 0.
 * chroma_resampler_type as user_args_t (hdr.h:276): 0 = box, otherwise FIR.  4:2:2 output is
 * FIR stage 1 only (convert.cpp:290-321); the reference itself has no 4:2:2 branch. */
h2y_status h2y_convert(h2y_ctx *ctx, const h2y_pic_desc *out, void *const d_out_planes[3],
                       const h2y_pic_desc *in, const void *const d_in_planes[3], int chroma_resampler_type,
                       void *stream);

/* the compute half of write_yuv (tiff.cpp:457-550): >> (src_bit_depth - pic->bit_depth) then the
 * range clamp, in place on the logical plane sizes.  The file append stays with the host. */
h2y_status h2y_write_yuv_clamp(h2y_ctx *ctx, const h2y_pic_desc *pic, void *const d_planes[3], int src_bit_depth,
                               void *stream);

/* Subsample420to444 (yuv2tiff.cpp:575-692 == convert.cpp:1869-1986) on one row-major plane:
 * src (w/2)x(h/2) -> dst w x h; algorithm 0 = box replicate, otherwise FIR. */
h2y_status h2y_subsample_420_to_444(h2y_ctx *ctx, const void *d_src, void *d_dst, int width, int height,
                                    int algorithm, uint16_t minCV, uint16_t maxCV, void *stream);

/* matrix_inverse (convert.cpp:1320-1867): hdr2yuv's own 4:4:4 inverse, used when a .yuv is converted to a .tiff
 * (dispatch at hdr2yuv.cpp:818-819).  U16 4:4:4 planes Y,Cb,Cr -> planes G,B,R at out->bit_depth.  The family is
 * chosen from in->matrix_coeffs exactly as the reference does, quirks included: 1 takes the BT.709 equations,
 * 0 is "Can't determine color difference to use?" (H2Y_ERR_MATRIX), every other value the Y'DzDx ones.  Half/Full
 * are the reference's hard-coded 2048/4096.  d_invalid_pixels (optional, one uint32) receives invalidPixels. */
h2y_status h2y_matrix_inverse(h2y_ctx *ctx, const h2y_pic_desc *out, void *const d_out_planes[3],
                              const h2y_pic_desc *in, const void *const d_in_planes[3], uint32_t *d_invalid_pixels,
                              void *stream);

/* the compute half of write_tiff (tiff.cpp:559-652): planes G,B,R -> interleaved R,G,B u16 rows, every sample
 * << (pic->bit_depth - src_bit_depth).  H2Y_ERR_BIT_DEPTH when that difference is negative. */
h2y_status h2y_write_tiff_rows(h2y_ctx *ctx, const h2y_pic_desc *pic, const void *const d_planes[3], int src_bit_depth,
                               void *d_rgb, void *stream);

/* .yuv (4:4:4) -> .tiff as main() does it (hdr2yuv.cpp:803-821, 898-933): matrix_inverse into a tmp picture of the
 * source's depth, then write_tiff at out_bit_depth.  Frames of three u16 planes in host memory -> interleaved RGB16
 * rows in host memory; h_invalid_pixels (optional) gets one count per frame. */
h2y_status h2y_inverse444_host(h2y_ctx *ctx, const h2y_pic_desc *in, int out_bit_depth, const void *h_yuv444,
                               size_t yuv_stride_bytes, void *h_rgb, size_t rgb_stride_bytes, int nframes,
                               uint32_t *h_invalid_pixels);

/* ---- fused forward path: what main() does between read_file and the file append ----------- */
typedef struct h2y_forward_params {
    h2y_pic_desc src;            /* in_pic after the reader: geometry, layout, depth, range, VUI codes */
    h2y_pic_desc dst;            /* out_pic after parse_options: depth, range, VUI codes, chroma format */
    int32_t chroma_resampler_type; /* user_args_t.chroma_resampler_type (hdr.h:276) */
    int32_t clip_on_load;        /* 1: read_tiff's clip of 16-bit samples to [minVR,maxVR] when
                                    src.video_full_range_flag == 0 (tiff.cpp:296-304) */
} h2y_forward_params;

/* pic_stats -> matrix_convert -> convert -> write_yuv clamp (hdr2yuv.cpp:797-928) for `nframes`
 * independent frames resident in HBM.  Frame i is read at d_src + i*src_stride_bytes in
 * params->src.layout and written at d_dst + i*dst_stride_bytes as one .yuv frame (Y, Cb, Cr
 * planes of u16, h2y_yuv_frame_bytes).  Asynchronous on `stream`.
 *
 * Plan reuse (large batches of half-float frames on the compiled-in BT.2020 10/12-bit configurations).  pic_stats
 * needs a frame's extrema before its first pixel is converted (common.cpp:66-168 -> convert.cpp:936-940), i.e. a
 * second pass over the input.  Frames of a sequence usually truncate to the same (int) floor / ceiling, so a call may
 * convert every frame with the plan of the previous call's last frame while gathering the frame's own extrema from the
 * samples it loads anyway; a verify step on the device then compares plans and the frames that differ are converted
 * again the classic way inside the same call.  The output is the reference's in every case; only the time differs.  The
 * library decides per call from the outcome of earlier calls (never waiting for the GPU); "H2Y_PLAN_REUSE" = "0" / "1"
 * (h2y_ctx_set_option) switches it off / forces it. */
h2y_status h2y_forward(h2y_ctx *ctx, const h2y_forward_params *params, const void *d_src, size_t src_stride_bytes,
                       void *d_dst, size_t dst_stride_bytes, int nframes, void *stream);

/* Same, from and to host memory: the library runs its own pinned-ring H2D / compute / D2H
 * pipeline on private streams and returns when the last frame has landed in h_dst. */
h2y_status h2y_forward_host(h2y_ctx *ctx, const h2y_forward_params *params, const void *h_src,
                            size_t src_stride_bytes, void *h_dst, size_t dst_stride_bytes, int nframes);

/* The per-frame statistics the last h2y_forward call derived (frame index in that call).  Synchronises the stream
 * used.  Only meaningful when the transfer changed.  The plans live in scratch that every 256-frame group of a call
 * and every chunk of h2y_forward_host's pipeline overwrites: after a call that spanned several groups or chunks this
 * returns H2Y_ERR_UNSUPPORTED.  i_min / i_max are filled for integer sources only (0 for half / float sources). */
h2y_status h2y_forward_last_stats(h2y_ctx *ctx, int frame, h2y_pic_stats_t *out);

/* What the last 256-frame group of the last h2y_forward call did about plan reuse: *attempted = 1 when it took the
 * single-pass route, *nframes = frames in the group, *nredone = frames the verify step handed back to the classic
 * kernels (0 when every prediction held).  Synchronises the stream used.  Any pointer may be NULL. */
h2y_status h2y_forward_last_plan_reuse(h2y_ctx *ctx, int *attempted, int *nframes, int *nredone);

/* ---- inverse path: one loop iteration of yuv2tiff's main() (yuv2tiff.cpp:278-552) --------- */
enum { H2Y_INV_YDzDx = 0, H2Y_INV_709 = 1, H2Y_INV_2020 = 2, H2Y_INV_Y100 = 3, H2Y_INV_Y500 = 4 };

typedef struct h2y_inverse_params {
    int32_t width, height;  /* the reference hard-codes 3840x2160 / HD1920 / HD960 (yuv2tiff.cpp:188-198) */
    int32_t bit_depth;      /* 10 (B10), 12 (default), 14 (B14): yuv2tiff.cpp:137-157 */
    int32_t matrix;         /* H2Y_INV_*: keywords 709 / 2020 / Y100 / Y500, default Y'DzDx (120-132) */
    int32_t fir;            /* 0 when BOX is given (117) */
    int32_t full_range;     /* FULL (118) */
    int32_t alpha;          /* ALPHA: 4 samples per pixel, A = 65535 (119-123, 544) */
    int32_t ybar;           /* -X: Y'DzDx rebuilt around the 2x2 mean of Y' (162, 365-399); ignored by the other matrices */
} h2y_inverse_params;

size_t h2y_rgb_frame_bytes(const h2y_inverse_params *p);   /* W*H*(3|4)*2 */

/* read-clamp -> 4:2:0->4:4:4 upsample -> inverse colour difference -> negatives to 0 -> range
 * clamp -> << (16-bit_depth), interleaved R,G,B(,A) u16 rows.  d_invalid_pixels (optional, device,
 * one uint32 per frame) receives the reference's invalidPixels count (yuv2tiff.cpp:478-511). */
h2y_status h2y_inverse(h2y_ctx *ctx, const h2y_inverse_params *params, const void *d_yuv, size_t yuv_stride_bytes,
                       void *d_rgb, size_t rgb_stride_bytes, int nframes, uint32_t *d_invalid_pixels, void *stream);

h2y_status h2y_inverse_host(h2y_ctx *ctx, const h2y_inverse_params *params, const void *h_yuv,
                            size_t yuv_stride_bytes, void *h_rgb, size_t rgb_stride_bytes, int nframes,
                            uint32_t *h_invalid_pixels);

/* F32 destinations (.exr / .dpx outputs, hdr2yuv.cpp:797-823, 935-962): pic_stats -> matrix_convert into an F32 picture
 * (convert.cpp:1117-1122, 1222-1304), frame after frame through the pinned H2D / compute / D2H pipeline.  h_dst receives
 * three float planes G, B, R per frame (pic_t.fbuf order), 12 bytes per pixel.  dst->bit_depth is the tmp picture's depth
 * for an integer source (hdr2yuv.cpp:805-808): 16 is the one value for which the reference's result is defined (an .exr
 * destination); 32 (what parse_options forces for a .dpx destination, hdr2yuv.cpp:436-446) makes set_pic_clip shift by
 * 32 and 24 (common.cpp:303-311), which on x86-64 leaves maxCV = 0 and every clip limit 0, so matrix_convert's clamp
 * (convert.cpp:1286-1293) writes an all-zero picture: that is reproduced.  F32 sources return H2Y_ERR_UNSUPPORTED (their
 * tmp depth is the source's 32 with a U16-style clip that has no defined value). */
h2y_status h2y_forward_f32_host(h2y_ctx *ctx, const h2y_forward_params *params, const void *h_src, size_t src_stride_bytes,
                                void *h_dst, size_t dst_stride_bytes, int nframes);

/* Optional linear-light stage behind the inverse path (SURVEY.md 8a note N2).  yuv2tiff stops at PQ-coded 16-bit
 * integers (yuv2tiff.cpp:535-552; the reference leaves linearisation to an external CTL step, README.md:113).  The in-repo
 * definition of that step is PQ10000_f (convert.cpp:43-51) on the code normalised as convert.cpp:1017-1019 does with
 * floor 0 and ceiling 65535:  linear = PQ10000_f((float)code / 65535.0f), in units of 10 000 cd/m2.  d_codes: n u16
 * samples in any layout (e.g. the interleaved RGB rows h2y_inverse wrote); d_linear: n floats in the same order.
 * The function is tabulated per 16-bit code in FP64 (once per context) and gathered: within 1 float ulp of the double
 * evaluation, far inside the 1e-5 relative tolerance the path states. */
h2y_status h2y_pq_codes_to_linear(h2y_ctx *ctx, const void *d_codes, size_t n, void *d_linear, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HDR2YUV_B200_H */
