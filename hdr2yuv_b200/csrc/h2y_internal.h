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
    // 1: the single-pass (plan-reuse) kernel converted this frame with the previous plan and the statistics it gathered
    // on the way confirmed that plan; every later kernel of the call leaves the frame alone (h2y_forward2.cu, "SPEC")
    int spec_done;
};

// ---- single-pass forward with plan reuse ("K1s") ---------------------------------------------------
// The reference needs a frame's extrema before its first pixel (common.cpp:66-168 -> convert.cpp:936-940), which costs
// a second read of the input.  Frames of a sequence usually share their (int) floor / ceiling, so the rows kernel can
// convert with the PREVIOUS plan while it gathers this frame's extrema from the samples it loads anyway; a verify step
// then compares plans and only the frames that differ are converted again the classic way.
struct SpecSeed {               // device-resident, one per context: the plan predicted for the next call's frames
    int valid;
    int est_floor, est_ceiling; // shared by G, B, R
    unsigned code_lo, code_hi;  // every half code whose value truncates into [floor, ceiling]: the LUT window
    int lut2_ok;
};
struct SpecCtl {                // device-resident, one per call
    int skip;                   // no usable seed: the speculative kernels return at once, every frame is flagged
    int nflag;                  // frames the classic kernels still have to convert
    int mode;                   // 0 nothing to do, 1 the general kernel converts the flagged frames
    int nframes;
    int seq;                    // call counter (the host reads a copy of this struct, late, as feedback for its policy)
    int uniform;                // frames of this call whose final plan equals the last frame's (classic calls)
    int pad[2];
};

// test / experiment switches: read from the environment once, at h2y_ctx_create, and changed per context through
// h2y_ctx_set_option (the product path never calls getenv afterwards)
struct Switches {
    int force_staged, force_v1, no_specialised, exact_math;
    int fwd_kernel;             // 0 automatic, 1 "ring", 2 "rows"
    int inv_kernel;             // 0 automatic, 1 "tile", 2 "rows", 3 "exact" = rows with every pixel through the exact routine
    int stats_gx;               // 0 automatic
    int guard_log2;             // < 0: the derived guard band; otherwise 2^-n (timing experiments, breaks parity)
    int spec;                   // plan reuse: -1 automatic (host policy), 0 never, 1 whenever a seed exists
};
void switches_from_env(Switches *s);
int switches_set(Switches *s, const char *name, const char *value);
constexpr unsigned LUT3_FLOATS = 58000;  // three tables together: 232 000 B
constexpr unsigned LUT2_CODES = 29000;   // 2 x 29000 floats = 232 000 B of the 232 448 B a CTA may use (half code 0x7148 ~ 10 800)

struct h2y_ctx_impl {
    int device;
    int sm_count;
    int last_cuda_error;
    unsigned long long launches;
    // scratch, grown on demand
    void *scratch[9];
    size_t scratch_bytes[9];
    // host pipeline
    cudaStream_t s_h2d, s_compute, s_d2h;
    int pipeline_ready;
    FrameK *h_framek;           // pinned copy for h2y_forward_last_stats
    int last_nframes;
    cudaStream_t last_stream;
    // profiling hooks
    int profile_on, profile_count;      // bracketed calls since h2y_profile_enable (ring of PROFILE_RING)
    cudaEvent_t ev[16][3];
    Switches sw;
    // instantiations of one call that convert disjoint frames run on forked streams joined by events
    cudaStream_t s_aux[4];
    cudaEvent_t ev_fork, ev_join[4];
    int aux_ready, aux_forked, aux_used;
    // one stream in flight per context: a call on another stream first waits for the previous call's work
    cudaEvent_t ev_busy;
    int busy_valid;
    cudaStream_t busy_stream;
    // plan reuse (SpecSeed / SpecCtl above)
    void *spec_dev;                     // SpecSeed + SpecCtl + seed LUT (65536 floats) + one predicted FrameK
    SpecCtl *spec_fb;                   // pinned: copy of the last SpecCtl, read late by the host policy
    int spec_seen_seq, spec_hint;       // policy state: feedback already acted upon; 1 = try plan reuse in the next call
    int spec_seq;                       // calls issued
    unsigned long long spec_key;        // parameters the seed belongs to (0: none yet)
    int last_is_float;
    int last_spec;                      // the last group of the last h2y_forward call took the single-pass route
    int last_groups;                    // 256-frame groups of the last h2y_forward call (h2y_forward_last_stats)
    float *pq_linear_lut;               // PQ10000_f per 16-bit code (h2y_pq_codes_to_linear), built on first use
};
// fork / join of the context's auxiliary streams around kernels that convert disjoint frames (h2y_api.cu)
h2y_status aux_fork(h2y_ctx_impl *c, cudaStream_t st);     // record the fork point on st (idempotent until aux_join)
cudaStream_t aux_stream(h2y_ctx_impl *c, int i);            // auxiliary stream i, made to wait for the fork point
h2y_status aux_join(h2y_ctx_impl *c, cudaStream_t st);     // st waits for every auxiliary stream used since aux_fork

// what a launch behind (or as) a plan-reuse pass needs to know
struct SpecLaunch {
    const SpecCtl *ctl;
    const int *flag;            // per frame: 1 = still to be converted by the classic kernels
    int *bail;                  // per frame: raised by the SPEC kernel when a code left the predicted LUT window
    unsigned *slots;            // statistics slots the SPEC kernel fills
    int spec;                   // 1: launch the SPEC instantiations of the rows kernel; 0: the general kernel, restricted to `flag`
};

constexpr int PROFILE_RING = 16;
enum ScratchSlot { SCR_STATS = 0, SCR_FRAMEK = 1, SCR_LUT = 2, SCR_TMP444 = 3, SCR_UNPACK = 4, SCR_OUT = 5,
                   SCR_RING_IN = 6, SCR_RING_OUT = 7, SCR_SPEC = 8, SCR_COUNT = 9 };

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
inline bool layout_is_dpx(int l) { return l >= H2Y_LAYOUT_DPX10_BE && l <= H2Y_LAYOUT_DPXF32_LE; }
inline int layout_channels(int l) { return (l == H2Y_LAYOUT_RGBA16 || l == H2Y_LAYOUT_HALF_RGBA) ? 4 : 3; }

// ---- launchers (h2y_stats.cu) ----------------------------------------------------------------
h2y_status launch_stats_and_luts(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, const void *d_src,
                                 size_t src_stride, int nframes, FrameK **d_framek, float **d_luts,
                                 cudaStream_t st);
// plan reuse: device state of the context, the passes around the SPEC kernel, and the seed update after any call
struct SpecDev { SpecSeed *seed; SpecCtl *ctl; FrameK *pred; float *seed_lut; };
h2y_status spec_dev(h2y_ctx_impl *c, SpecDev *out);
h2y_status launch_spec_prepare(h2y_ctx_impl *c, int nframes, int seq, FrameK **d_framek, unsigned **d_slots, int **d_bail,
                               int **d_flag, cudaStream_t st);
h2y_status launch_spec_verify_and_redo_prologue(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, const void *d_src,
                                                size_t src_stride, int nframes, float **d_luts, cudaStream_t st);
h2y_status launch_seed_update(h2y_ctx_impl *c, const FrameK *d_framek, const float *d_luts, int nframes, int seq,
                              int after_spec, int force, cudaStream_t st);
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
h2y_status launch_pq_codes_to_linear(h2y_ctx_impl *c, const uint16_t *d_codes, size_t n, float *d_linear, cudaStream_t st);
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
bool forward_exr420_spec_supported(const h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, int tmp_bit_depth, int nframes);
h2y_status launch_forward_exr420(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, int tmp_bit_depth,
                                 const void *d_src, size_t src_stride, void *d_dst, size_t dst_stride, int nframes,
                                 const FrameK *d_framek, const float *d_luts, cudaStream_t st, int *took_three_table_frames = nullptr,
                                 const SpecLaunch *sl = nullptr);
h2y_status launch_forward_fused(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, const void *d_src,
                                size_t src_stride, void *d_dst, size_t dst_stride, int nframes,
                                const FrameK *d_framek, const float *d_luts, cudaStream_t st, int skip_clean = 0,
                                const SpecLaunch *sl = nullptr, cudaStream_t sweep_stream = nullptr);

struct InvK {
    int w, h, bit_depth, matrix, fir, full_range, alpha, ybar;
    int int10;                  // `B10 2020`: the exact routine has an integer form (inv_pixel_int10)
    // integer form for the other Y'CbCr depths / family (inv_pixel_int): B' = ((2 Cb - (Full-1)) ikb + 10000 Y') / 10000,
    // R' likewise with ikr, G' = (igy Y' + igb B' + igr R' + igc) / igd; igm = floor(2^32 / igd)
    int ikb, ikr, igy, igb, igr, igc, igd;
    unsigned igm;
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
