// h2y_internal.h -- host-side declarations shared by the .cu translation units.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "h2y_device.cuh"
#include "hdr2yuv_b200.h"

namespace h2y {

// Per-frame statistics + normalisation constants, produced on the device by the stats/plan
// kernels and consumed by the LUT build and the fused forward kernel (no host round trip).
struct FrameK {
    float fmin[3], fmax[3];     // extrema as pic_stats sees them (common.cpp:66-168)
    int est_floor[3], est_ceiling[3];
    float offset[3], range[3];  // convert.cpp:936-940
    int lut_slot[3];            // index (frame*3+channel) of the LUT this channel reads
    int same_lut;               // 1 when the three channels share one LUT
    // half-float sources: every sample finite and non-negative (sign bit clear), normalisation range > 0 and one
    // shared LUT -- the precondition of the v2 kernel (h2y_forward2.cu); code_lo..code_hi = raw code range
    int clean;
    unsigned code_lo, code_hi;
    int lut2_ok;                // clean and code_hi < LUT2_CODES: two pre-scaled LUT copies fit in shared memory
    // same preconditions but a table per channel (the channels' floor / ceiling differ, as in real footage): the rows
    // kernel keeps three range-restricted tables, channel c indexed by code - ch_lo[c]; clean3 also says that they fit
    int clean3;
    unsigned ch_lo[3], ch_hi[3];
    int zero_entry[3];          // table entry ch_lo[c] holds the value of code 0 (see k_plan)
};
constexpr unsigned LUT3_FLOATS = 58000;  // three tables together: 232 000 B
constexpr unsigned LUT2_CODES = 29000;   // 2 x 29000 floats = 232 000 B of the 232 448 B a CTA may use (half code 0x7148 ~ 10 800)

struct h2y_ctx_impl {
    int device;
    int sm_count;
    int last_cuda_error;
    unsigned long long launches;
    // scratch, grown on demand
    void *scratch[8];
    size_t scratch_bytes[8];
    // host pipeline
    cudaStream_t s_h2d, s_compute, s_d2h;
    int pipeline_ready;
    FrameK *h_framek;           // pinned copy for h2y_forward_last_stats
    int last_nframes;
    cudaStream_t last_stream;
    // profiling hooks
    int profile_on, profile_count;      // bracketed calls since h2y_profile_enable (ring of PROFILE_RING)
    cudaEvent_t ev[16][3];
};

constexpr int PROFILE_RING = 16;
enum ScratchSlot { SCR_STATS = 0, SCR_FRAMEK = 1, SCR_LUT = 2, SCR_TMP444 = 3, SCR_UNPACK = 4, SCR_OUT = 5,
                   SCR_RING_IN = 6, SCR_RING_OUT = 7 };

h2y_status scratch_reserve(h2y_ctx_impl *c, int slot, size_t bytes, void **out);
h2y_status cuda_fail(h2y_ctx_impl *c, cudaError_t e);

#define H2Y_CUDA(ctx, call)                                                  \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return ::h2y::cuda_fail((ctx), e__);         \
    } while (0)

// host-side derivation of the per-launch constants from the two picture descriptors
h2y_status make_pixk(const h2y_pic_desc &in, const h2y_pic_desc &tmp_out, int out_bit_depth, int out_full_range,
                     int clip_on_load, PixK *k);
void clip_of(int bit_depth, int full_range, h2y_clip_limits *c);

inline bool layout_is_planar(int l) { return l == H2Y_LAYOUT_PLANAR_U16 || l == H2Y_LAYOUT_PLANAR_F32; }
inline bool layout_is_half(int l) { return l == H2Y_LAYOUT_HALF_RGB || l == H2Y_LAYOUT_HALF_RGBA; }
inline bool layout_is_dpx(int l) { return l == H2Y_LAYOUT_DPX10_BE || l == H2Y_LAYOUT_DPX10_LE; }
inline int layout_channels(int l) { return (l == H2Y_LAYOUT_RGBA16 || l == H2Y_LAYOUT_HALF_RGBA) ? 4 : 3; }

// ---- launchers (h2y_stats.cu) ----------------------------------------------------------------
h2y_status launch_stats_and_luts(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, const void *d_src,
                                 size_t src_stride, int nframes, FrameK **d_framek, float **d_luts,
                                 cudaStream_t st);
h2y_status launch_stats_planar(h2y_ctx_impl *c, const h2y_pic_desc &pic, const void *const d_planes[3],
                               FrameK **d_framek, cudaStream_t st);

// ---- launchers (h2y_staged.cu) ---------------------------------------------------------------
struct NormK { float offset[3], range[3]; };
h2y_status launch_matrix_convert(h2y_ctx_impl *c, const PixK &k, const NormK &nk, int w, int h, int in_is_f32,
                                 const void *const d_in[3], int out_is_f32, void *const d_out[3], cudaStream_t st);
h2y_status launch_unpack(h2y_ctx_impl *c, int layout, int w, int h, const void *d_src, void *const d_planes[3],
                         int clip_on_load, unsigned lo, unsigned hi, cudaStream_t st);
h2y_status launch_fir_420(h2y_ctx_impl *c, const uint16_t *d_src, uint16_t *d_dst, uint16_t *d_mid, int w, int h,
                          unsigned maxCV, cudaStream_t st);
h2y_status launch_fir_422(h2y_ctx_impl *c, const uint16_t *d_src, uint16_t *d_dst, int w, int h, unsigned maxCV,
                          cudaStream_t st);
h2y_status launch_box_420(h2y_ctx_impl *c, const uint16_t *d_src, uint16_t *d_dst, int w, int h, cudaStream_t st);
h2y_status launch_yuvprime2_420(h2y_ctx_impl *c, const uint16_t *const d_in[3], uint16_t *d_u, uint16_t *d_v, uint16_t *scratch,
                                int w, int h, int resampler, unsigned maxCV, cudaStream_t st);
h2y_status launch_out_clamp(h2y_ctx_impl *c, uint16_t *d_plane, size_t n, int shift, unsigned lo, unsigned hi,
                            cudaStream_t st);
h2y_status launch_upsample(h2y_ctx_impl *c, const uint16_t *d_src, uint16_t *d_dst, uint16_t *d_mid, int w, int h,
                           int fir, unsigned minCV, unsigned maxCV, cudaStream_t st);

h2y_status launch_matrix_inverse(h2y_ctx_impl *c, int family, int minVR, int maxVR, int shift_right, int shift_left,
                                 long npix, const uint16_t *d_in, size_t in_stride_elems, uint16_t *d_out,
                                 size_t out_stride_elems, int nframes, int interleave, uint32_t *d_invalid, cudaStream_t st);
h2y_status launch_write_tiff_rows(h2y_ctx_impl *c, long npix, const uint16_t *g, const uint16_t *b, const uint16_t *r,
                                  uint16_t *rgb, int sr, cudaStream_t st);

// ---- launchers (h2y_forward.cu / h2y_inverse.cu) ---------------------------------------------
bool fused_forward_supported(const h2y_forward_params &p);
bool forward_u16_420_supported(const h2y_forward_params &p, const PixK &k);
h2y_status launch_forward_u16_420(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, const void *d_src,
                                  size_t src_stride, void *d_dst, size_t dst_stride, int nframes, cudaStream_t st);
bool forward_exr420_supported(const h2y_forward_params &p, const PixK &k, int tmp_bit_depth);
h2y_status launch_forward_exr420(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, int tmp_bit_depth,
                                 const void *d_src, size_t src_stride, void *d_dst, size_t dst_stride, int nframes,
                                 const FrameK *d_framek, const float *d_luts, cudaStream_t st, int *took_three_table_frames = nullptr);
h2y_status launch_forward_fused(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, const void *d_src,
                                size_t src_stride, void *d_dst, size_t dst_stride, int nframes,
                                const FrameK *d_framek, const float *d_luts, cudaStream_t st, int skip_clean = 0);

struct InvK {
    int w, h, bit_depth, matrix, fir, full_range, alpha, ybar;
    int SR;
    unsigned Half, Full, maxCV;
    unsigned minVR, maxVR, minVRC, maxVRC;
    double kb, kr, wb, wr, wg;
    float T, U, V, W;
};
h2y_status make_invk(const h2y_inverse_params &p, InvK *k);
h2y_status launch_inverse(h2y_ctx_impl *c, const InvK &k, const void *d_yuv, size_t yuv_stride, void *d_rgb,
                          size_t rgb_stride, int nframes, uint32_t *d_invalid, cudaStream_t st);

}   // namespace h2y
