// h2y_forward.cu -- K1+K2: the fused forward kernel.
//
// One persistent kernel does, for every pixel of a batch of frames resident in HBM:
//   reader de-interleave + on-read clip (tiff.cpp:265-315 / exr.cpp:209-235)
//   -> normalise + transfer change through the exact per-frame LUT (convert.cpp:1017-1109)
//   -> range scale (1116-1144) -> colour-difference matrix, offset, clamp (1146-1220)
//   -> 4:4:4 -> 4:2:0 / 4:2:2 chroma FIR or box (convert.cpp:261-383, 91-172)
//   -> write_yuv's shift + range clamp (tiff.cpp:457-550)
// and writes the planar .yuv frame.  4:4:4 chroma and the 4:2:2 intermediate never touch HBM.
// Algorithmic traffic: 6 B/px in (3 x 16-bit; 8 with alpha) + 3 B/px out (4:2:0) = 9 B/px.
//
// Decomposition.  A work item is (frame, row segment, column strip).  A warp owns one image row
// of a strip at a time; a lane owns 8 consecutive pixels (16-byte vector loads/stores).  In the
// FIR modes lanes 0 and 31 are halo lanes (they compute chroma for the 8 pixels left/right of
// the strip so the 7-tap horizontal filter can fetch its +-5 neighbours with warp shuffles), so
// a strip is 240 pixels wide and 3840 = 16 strips exactly.  The horizontally filtered rows go
// into a 48-row shared-memory ring (the reference's u16 `dst422` intermediate, kept as exact
// integer-valued floats); after every 16 rows the CTA runs the 12-tap vertical filter on the
// rows that became complete.  Only the 11 halo rows at a segment boundary are recomputed.
// Picture edges replicate by index clamp exactly as convert.cpp:295-300, 337-347.
//
// The CTA is persistent (grid = SMs x CTAs/SM) and walks items in frame order so that the
// transfer LUT staged in shared memory is reloaded only when the frame's (floor, ceiling)
// pair changes.
#include "h2y_internal.h"

namespace h2y {

enum ChromaMode : int { CM_444 = 0, CM_420_FIR = 1, CM_420_BOX = 2, CM_422_FIR = 3 };

namespace {
constexpr int THREADS = 512, WARPS = THREADS / 32;
constexpr int RING_ROWS = 48, RING_W = 128;          // 120 used columns per plane, padded
constexpr int LUT_SMEM_CODES = 0x7C01;               // non-negative halfs up to +inf
}   // namespace

struct FwdArgs {
    const uint8_t *src;
    size_t src_stride;       // bytes between frames
    uint8_t *dst;
    size_t dst_stride;
    int w, h, nframes;
    int layout;
    int use_lut;             // transfer changes: per-frame LUT gather
    int exact_math;          // force the reference-order FP64 path for every pixel (debug / tests)
    int skip_clean;          // 1: frames flagged clean were converted by the fast kernels (h2y_forward2.cu); 2: clean3 frames too
    int strip_w, nstrips, seg_rows, nsegs, nitems;
    PixK k;
    const FrameK *framek;
    const float *luts;       // [nframes*3][65536]
    unsigned long long *fallback_count;   // pixels that took the exact fallback (diagnostic)
    // behind a plan-reuse pass (h2y_internal.h): only frames with flag[frame] != 0 are left, this kernel converts all of them
    const SpecCtl *ctl;
    const int *flag;
};

// ---- 8-pixel loads -------------------------------------------------------------------------------
struct Px8 { unsigned g[8], b[8], r[8]; };   // 16-bit codes

__device__ __forceinline__ void unpack_planar(const uint4 &v, unsigned o[8])
{
    o[0] = v.x & 0xffffu; o[1] = v.x >> 16; o[2] = v.y & 0xffffu; o[3] = v.y >> 16;
    o[4] = v.z & 0xffffu; o[5] = v.z >> 16; o[6] = v.w & 0xffffu; o[7] = v.w >> 16;
}

struct Raw8 { uint4 v[4]; };

__device__ __forceinline__ void issue_loads(const FwdArgs &a, const uint8_t *frame, int row, int x, Raw8 &raw)
{
    const size_t px = (size_t)row * a.w + x;
    if (a.layout == H2Y_LAYOUT_PLANAR_U16) {
        const size_t plane = (size_t)a.w * a.h * 2;
        raw.v[0] = __ldg(reinterpret_cast<const uint4 *>(frame + px * 2));
        raw.v[1] = __ldg(reinterpret_cast<const uint4 *>(frame + plane + px * 2));
        raw.v[2] = __ldg(reinterpret_cast<const uint4 *>(frame + 2 * plane + px * 2));
    } else if (a.layout == H2Y_LAYOUT_RGB16 || a.layout == H2Y_LAYOUT_HALF_RGB) {
        const uint4 *p = reinterpret_cast<const uint4 *>(frame + px * 6);
        raw.v[0] = __ldg(p); raw.v[1] = __ldg(p + 1); raw.v[2] = __ldg(p + 2);
    } else {
        const uint4 *p = reinterpret_cast<const uint4 *>(frame + px * 8);
        raw.v[0] = __ldg(p); raw.v[1] = __ldg(p + 1); raw.v[2] = __ldg(p + 2); raw.v[3] = __ldg(p + 3);
    }
}

__device__ __forceinline__ void decode_loads(const FwdArgs &a, const Raw8 &raw, Px8 &p)
{
    if (a.layout == H2Y_LAYOUT_PLANAR_U16) {
        unpack_planar(raw.v[0], p.g); unpack_planar(raw.v[1], p.b); unpack_planar(raw.v[2], p.r);
    } else if (a.layout == H2Y_LAYOUT_RGB16 || a.layout == H2Y_LAYOUT_HALF_RGB) {
        unsigned s[24];
        unpack_planar(raw.v[0], s); unpack_planar(raw.v[1], s + 8); unpack_planar(raw.v[2], s + 16);
#pragma unroll
        for (int q = 0; q < 8; q++) { p.r[q] = s[3 * q]; p.g[q] = s[3 * q + 1]; p.b[q] = s[3 * q + 2]; }
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint4 &v = raw.v[q];      // two RGBA pixels per 16 bytes
            p.r[2 * q] = v.x & 0xffffu; p.g[2 * q] = v.x >> 16; p.b[2 * q] = v.y & 0xffffu;
            p.r[2 * q + 1] = v.z & 0xffffu; p.g[2 * q + 1] = v.z >> 16; p.b[2 * q + 1] = v.w & 0xffffu;
        }
    }
}

// ---- code -> float sample in the destination transfer domain --------------------------------------
struct LutView {
    const float *g0, *g1, *g2;   // global LUTs per channel
    const float *s;              // shared LUT (all channels) or nullptr
    unsigned s_lo, s_hi;         // code range resident in shared memory
};

__device__ __forceinline__ float lut_fetch(const LutView &lv, const float *g, unsigned code)
{
    if (lv.s) {
        unsigned o = code - lv.s_lo;
        if (o <= lv.s_hi - lv.s_lo) return lv.s[o];
    }
    return __ldg(g + code);
}

template <int MK>
__device__ __forceinline__ void pixel8(const FwdArgs &a, const LutView &lv, const Px8 &p, unsigned Y[8],
                                       unsigned Cb[8], unsigned Cr[8], unsigned &fallbacks)
{
    const PixK &k = a.k;
    const bool half_in = a.layout == H2Y_LAYOUT_HALF_RGB || a.layout == H2Y_LAYOUT_HALF_RGBA;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        float G, B, R;
        if (a.use_lut) {
            G = lut_fetch(lv, lv.g0, p.g[q]);
            B = lut_fetch(lv, lv.g1, p.b[q]);
            R = lut_fetch(lv, lv.g2, p.r[q]);
            scale_to_codes(G, B, R, k);
        } else if (half_in) {
            G = half_bits_to_float(p.g[q]); B = half_bits_to_float(p.b[q]); R = half_bits_to_float(p.r[q]);
        } else {
            unsigned g = p.g[q], b = p.b[q], r = p.r[q];
            if (k.clip_on_load) {
                g = min(max(g, k.loadLo), k.loadHi); b = min(max(b, k.loadLo), k.loadHi); r = min(max(r, k.loadLo), k.loadHi);
            }
            G = (float)g; B = (float)b; R = (float)r;
        }
        bool ok = !a.exact_math && px_matrix_fast<MK>(G, B, R, k, Y[q], Cb[q], Cr[q]);
        if (!ok) {
            px_matrix_exact<MK>(G, B, R, k, Y[q], Cb[q], Cr[q]);
            fallbacks++;
        }
    }
}

__device__ __forceinline__ uint4 pack8_clamped(const unsigned v[8], int shift, unsigned lo, unsigned hi)
{
    unsigned c[8];
#pragma unroll
    for (int q = 0; q < 8; q++) c[q] = out_clamp(v[q], shift, lo, hi);
    return make_uint4(c[0] | (c[1] << 16), c[2] | (c[3] << 16), c[4] | (c[5] << 16), c[6] | (c[7] << 16));
}

// ---- the kernel --------------------------------------------------------------------------------------
template <int MK, int CM>
__global__ void __launch_bounds__(THREADS, 1) k_forward_fused(const FwdArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);                   // [RING_ROWS][2][RING_W]
    float *lut_s = ring + ((CM == CM_420_FIR) ? RING_ROWS * 2 * RING_W : 0);
    constexpr int HALO = (CM == CM_420_FIR || CM == CM_422_FIR) ? 1 : 0;
    constexpr int ROWS_PER_STEP = (CM == CM_420_BOX) ? 2 * WARPS : WARPS;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PixK &k = a.k;
    const int w = a.w, h = a.h, wh = w >> 1;
    const float maxCVf = (float)k.maxCV;
    const bool half_in = a.layout == H2Y_LAYOUT_HALF_RGB || a.layout == H2Y_LAYOUT_HALF_RGBA;
    int cur_slot = -1;
    unsigned cur_lo = 0, cur_hi = 0;
    unsigned fallbacks = 0;
    if (a.ctl && a.ctl->nflag == 0) return;                     // the plan-reuse pass left nothing to do

    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int strip = item % a.nstrips;
        const int seg = (item / a.nstrips) % a.nsegs;
        const int frame = item / (a.nstrips * a.nsegs);
        if (a.flag && !a.flag[frame]) continue;                 // converted and confirmed by the plan-reuse pass
        if (a.skip_clean && !a.ctl &&
            (a.framek[frame].clean || (a.skip_clean > 1 && a.framek[frame].clean3))) continue;   // uniform per CTA
        const uint8_t *fsrc = a.src + (size_t)frame * a.src_stride;
        uint16_t *fY = reinterpret_cast<uint16_t *>(a.dst + (size_t)frame * a.dst_stride);
        uint16_t *fCb = fY + (size_t)w * h;
        const int cw = (CM == CM_444) ? w : wh;
        const int chh = (CM == CM_420_FIR || CM == CM_420_BOX) ? (h >> 1) : h;
        uint16_t *fCr = fCb + (size_t)cw * chh;

        // ---- LUT view for this frame ----
        LutView lv;
        lv.s = nullptr; lv.s_lo = 0; lv.s_hi = 0; lv.g0 = lv.g1 = lv.g2 = nullptr;
        if (a.use_lut) {
            const FrameK &fk = a.framek[frame];
            lv.g0 = a.luts + (size_t)fk.lut_slot[0] * 65536;
            lv.g1 = a.luts + (size_t)fk.lut_slot[1] * 65536;
            lv.g2 = a.luts + (size_t)fk.lut_slot[2] * 65536;
            if (half_in && fk.same_lut) {
                // resident code range: [half(min) .. half(max)] of the non-negative values
                float mn = fminf(fminf(fk.fmin[0], fk.fmin[1]), fk.fmin[2]);
                float mx = fmaxf(fmaxf(fk.fmax[0], fk.fmax[1]), fk.fmax[2]);
                unsigned lo = mn > 0.0f ? (unsigned)__half_as_ushort(__float2half_rd(mn)) : 0u;
                unsigned hi = mx < 65504.0f ? (unsigned)__half_as_ushort(__float2half_ru(mx)) : 0x7C00u;
                lo = min(lo, 0x7C00u); hi = min(max(hi, lo), 0x7C00u);
                __syncthreads();        // the previous item's readers are done
                if (fk.lut_slot[0] != cur_slot || lo < cur_lo || hi > cur_hi) {
                    for (unsigned c = lo + threadIdx.x; c <= hi; c += THREADS) lut_s[c - lo] = __ldg(lv.g0 + c);
                    cur_slot = fk.lut_slot[0]; cur_lo = lo; cur_hi = hi;
                    __syncthreads();
                }
                lo = cur_lo; hi = cur_hi;
                lv.s = lut_s; lv.s_lo = lo; lv.s_hi = hi;
            } else {
                __syncthreads();
            }
        } else {
            __syncthreads();
        }

        const int x0 = strip * a.strip_w;
        const int ys = seg * a.seg_rows, ye = min(ys + a.seg_rows, h);
        const int xl = x0 + 8 * (lane - HALO);                 // first pixel of this lane
        const bool lane_in_pic = xl >= 0 && xl < w;
        const bool lane_interior = lane >= HALO && lane < 32 - HALO && xl < min(x0 + a.strip_w, w);

        if (CM == CM_420_FIR) {
            const int r0 = ys - 6;
            const int nsteps = (ye - ys + 12 + 15) / 16;
            Raw8 raw;
            {   // prefetch step 0
                const int row = r0 + warp;
                if (row >= 0 && row < h && lane_in_pic) issue_loads(a, fsrc, row, xl, raw);
            }
            for (int s = 0; s < nsteps; s++) {
                const int row = r0 + 16 * s + warp;
                const bool row_ok = row >= 0 && row < h && row < ye + 6;
                unsigned Y[8], Cb[8], Cr[8];
#pragma unroll
                for (int q = 0; q < 8; q++) { Y[q] = 0; Cb[q] = 0; Cr[q] = 0; }
                if (row_ok && lane_in_pic) {
                    Px8 p;
                    decode_loads(a, raw, p);
                    pixel8<MK>(a, lv, p, Y, Cb, Cr, fallbacks);
                    if (lane_interior && row >= ys && row < ye)
                        *reinterpret_cast<uint4 *>(fY + (size_t)row * w + xl) = pack8_clamped(Y, k.down_shift, k.loY, k.hiY);
                }
                {   // prefetch the next step's row while this one is filtered
                    const int nrow = row + 16;
                    if (s + 1 < nsteps && nrow >= 0 && nrow < h && nrow < ye + 6 && lane_in_pic)
                        issue_loads(a, fsrc, nrow, xl, raw);
                }
                if (row_ok) {   // warp-uniform
                    // horizontal 7-tap at even x (convert.cpp:290-321): neighbours via shuffles
                    const unsigned c3 = Cb[3] | (Cr[3] << 16), c5 = Cb[5] | (Cr[5] << 16), c7 = Cb[7] | (Cr[7] << 16);
                    const unsigned c1 = Cb[1] | (Cr[1] << 16);
                    unsigned l3 = __shfl_up_sync(0xffffffffu, c3, 1), l5 = __shfl_up_sync(0xffffffffu, c5, 1),
                             l7 = __shfl_up_sync(0xffffffffu, c7, 1);
                    unsigned n1 = __shfl_down_sync(0xffffffffu, c1, 1), n3 = __shfl_down_sync(0xffffffffu, c3, 1);
                    if (xl == 0) { const unsigned e = Cb[0] | (Cr[0] << 16); l3 = l5 = l7 = e; }          // replicate s[0]
                    if (xl + 8 >= w) { const unsigned e = Cb[7] | (Cr[7] << 16); n1 = n3 = e; }            // replicate s[W-1]
                    if (lane_interior) {
                        float o[2][4];
#pragma unroll
                        for (int pl = 0; pl < 2; pl++) {
                            const unsigned *C = pl ? Cr : Cb;
                            const int sh = pl * 16;
                            const float L3 = (float)((l3 >> sh) & 0xffffu), L5 = (float)((l5 >> sh) & 0xffffu),
                                        L7 = (float)((l7 >> sh) & 0xffffu), N1 = (float)((n1 >> sh) & 0xffffu),
                                        N3 = (float)((n3 >> sh) & 0xffffu);
                            float f[8];
#pragma unroll
                            for (int q = 0; q < 8; q++) f[q] = (float)C[q];
                            o[pl][0] = (float)fir_h7(L3, L5, L7, f[0], f[1], f[3], f[5], maxCVf);
                            o[pl][1] = (float)fir_h7(L5, L7, f[1], f[2], f[3], f[5], f[7], maxCVf);
                            o[pl][2] = (float)fir_h7(L7, f[1], f[3], f[4], f[5], f[7], N1, maxCVf);
                            o[pl][3] = (float)fir_h7(f[1], f[3], f[5], f[6], f[7], N1, N3, maxCVf);
                        }
                        float *rr = ring + (size_t)((row + RING_ROWS) % RING_ROWS) * (2 * RING_W) + (lane - 1) * 4;
                        *reinterpret_cast<float4 *>(rr) = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
                        *reinterpret_cast<float4 *>(rr + RING_W) = make_float4(o[1][0], o[1][1], o[1][2], o[1][3]);
                    }
                }
                __syncthreads();
                // vertical 12-tap at even y (convert.cpp:333-377) on the rows that are now complete
                if (threadIdx.x < 480) {
                    const int jj = threadIdx.x / 60, rem = threadIdx.x % 60, pl = rem / 30, cg = rem % 30;
                    const int j = (ys >> 1) + 8 * s - 6 + jj;
                    const int col = (x0 >> 1) + cg * 4;
                    if (j >= (ys >> 1) && j < (ye >> 1) && col < min((x0 + a.strip_w) >> 1, wh)) {
                        float4 t[12];
#pragma unroll
                        for (int tt = 0; tt < 12; tt++) {
                            int rr = 2 * j - 5 + tt;
                            rr = rr < 0 ? 0 : (rr > h - 1 ? h - 1 : rr);
                            t[tt] = *reinterpret_cast<const float4 *>(ring + (size_t)(rr % RING_ROWS) * (2 * RING_W) +
                                                                      pl * RING_W + cg * 4);
                        }
                        unsigned o[4];
                        float r12[12];
#pragma unroll
                        for (int tt = 0; tt < 12; tt++) r12[tt] = t[tt].x;
                        o[0] = fir_v12(r12, maxCVf);
#pragma unroll
                        for (int tt = 0; tt < 12; tt++) r12[tt] = t[tt].y;
                        o[1] = fir_v12(r12, maxCVf);
#pragma unroll
                        for (int tt = 0; tt < 12; tt++) r12[tt] = t[tt].z;
                        o[2] = fir_v12(r12, maxCVf);
#pragma unroll
                        for (int tt = 0; tt < 12; tt++) r12[tt] = t[tt].w;
                        o[3] = fir_v12(r12, maxCVf);
#pragma unroll
                        for (int q = 0; q < 4; q++) o[q] = out_clamp(o[q], k.down_shift, k.loC, k.hiC);
                        uint16_t *dstp = (pl ? fCr : fCb) + (size_t)j * wh + col;
                        *reinterpret_cast<uint2 *>(dstp) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
                    }
                }
            }
        } else {
            // 4:4:4, 4:2:2 (FIR stage 1) and 4:2:0 box: no vertical halo, no ring
            for (int rbase = ys; rbase < ye; rbase += ROWS_PER_STEP) {
                const int nr = (CM == CM_420_BOX) ? 2 : 1;
                unsigned keepCb[4], keepCr[4];
#pragma unroll
                for (int rr = 0; rr < nr; rr++) {
                    const int row = rbase + warp * nr + rr;
                    const bool row_ok = row < ye;
                    unsigned Y[8], Cb[8], Cr[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) { Y[q] = 0; Cb[q] = 0; Cr[q] = 0; }
                    if (row_ok && lane_in_pic) {
                        Raw8 raw;
                        Px8 p;
                        issue_loads(a, fsrc, row, xl, raw);
                        decode_loads(a, raw, p);
                        pixel8<MK>(a, lv, p, Y, Cb, Cr, fallbacks);
                        if (lane_interior) {
                            *reinterpret_cast<uint4 *>(fY + (size_t)row * w + xl) = pack8_clamped(Y, k.down_shift, k.loY, k.hiY);
                            if (CM == CM_444) {
                                *reinterpret_cast<uint4 *>(fCb + (size_t)row * w + xl) = pack8_clamped(Cb, k.down_shift, k.loC, k.hiC);
                                *reinterpret_cast<uint4 *>(fCr + (size_t)row * w + xl) = pack8_clamped(Cr, k.down_shift, k.loC, k.hiC);
                            }
                        }
                    }
                    if (CM == CM_422_FIR && row_ok) {
                        const unsigned c3 = Cb[3] | (Cr[3] << 16), c5 = Cb[5] | (Cr[5] << 16), c7 = Cb[7] | (Cr[7] << 16);
                        const unsigned c1 = Cb[1] | (Cr[1] << 16);
                        unsigned l3 = __shfl_up_sync(0xffffffffu, c3, 1), l5 = __shfl_up_sync(0xffffffffu, c5, 1),
                                 l7 = __shfl_up_sync(0xffffffffu, c7, 1);
                        unsigned n1 = __shfl_down_sync(0xffffffffu, c1, 1), n3 = __shfl_down_sync(0xffffffffu, c3, 1);
                        if (xl == 0) { const unsigned e = Cb[0] | (Cr[0] << 16); l3 = l5 = l7 = e; }
                        if (xl + 8 >= w) { const unsigned e = Cb[7] | (Cr[7] << 16); n1 = n3 = e; }
                        if (lane_interior) {
#pragma unroll
                            for (int pl = 0; pl < 2; pl++) {
                                const unsigned *C = pl ? Cr : Cb;
                                const int sh = pl * 16;
                                const float L3 = (float)((l3 >> sh) & 0xffffu), L5 = (float)((l5 >> sh) & 0xffffu),
                                            L7 = (float)((l7 >> sh) & 0xffffu), N1 = (float)((n1 >> sh) & 0xffffu),
                                            N3 = (float)((n3 >> sh) & 0xffffu);
                                float f[8];
#pragma unroll
                                for (int q = 0; q < 8; q++) f[q] = (float)C[q];
                                unsigned o[4];
                                o[0] = fir_h7(L3, L5, L7, f[0], f[1], f[3], f[5], maxCVf);
                                o[1] = fir_h7(L5, L7, f[1], f[2], f[3], f[5], f[7], maxCVf);
                                o[2] = fir_h7(L7, f[1], f[3], f[4], f[5], f[7], N1, maxCVf);
                                o[3] = fir_h7(f[1], f[3], f[5], f[6], f[7], N1, N3, maxCVf);
#pragma unroll
                                for (int q = 0; q < 4; q++) o[q] = out_clamp(o[q], k.down_shift, k.loC, k.hiC);
                                uint16_t *dstp = (pl ? fCr : fCb) + (size_t)row * wh + (xl >> 1);
                                *reinterpret_cast<uint2 *>(dstp) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
                            }
                        }
                    }
                    if (CM == CM_420_BOX) {
                        // truncating mean of each 2x2 (convert.cpp:157-160)
                        if (rr == 0) {
#pragma unroll
                            for (int q = 0; q < 4; q++) { keepCb[q] = Cb[2 * q] + Cb[2 * q + 1]; keepCr[q] = Cr[2 * q] + Cr[2 * q + 1]; }
                        } else if (row_ok && lane_in_pic && lane_interior) {
                            unsigned ob[4], orr[4];
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                ob[q] = out_clamp((keepCb[q] + Cb[2 * q] + Cb[2 * q + 1]) >> 2, k.down_shift, k.loC, k.hiC);
                                orr[q] = out_clamp((keepCr[q] + Cr[2 * q] + Cr[2 * q + 1]) >> 2, k.down_shift, k.loC, k.hiC);
                            }
                            const size_t o = (size_t)(row >> 1) * wh + (xl >> 1);
                            *reinterpret_cast<uint2 *>(fCb + o) = make_uint2(ob[0] | (ob[1] << 16), ob[2] | (ob[3] << 16));
                            *reinterpret_cast<uint2 *>(fCr + o) = make_uint2(orr[0] | (orr[1] << 16), orr[2] | (orr[3] << 16));
                        }
                    }
                }
            }
        }
    }
    if (a.fallback_count && fallbacks) atomicAdd(a.fallback_count, (unsigned long long)fallbacks);
}

// ---- host side -------------------------------------------------------------------------------------------

bool fused_forward_supported(const h2y_forward_params &p)
{
    const int w = p.src.width, h = p.src.height;
    if (p.src.layout == H2Y_LAYOUT_PLANAR_F32 || layout_is_dpx(p.src.layout)) return false;   // staged route
    if (w < 8 || (w & 7) || h < 2) return false;
    if ((p.dst.chroma_format_idc == H2Y_CHROMA_420) && (h & 1)) return false;
    if (p.dst.chroma_format_idc == H2Y_CHROMA_420 && p.chroma_resampler_type == 0 && ((w & 3) || (h & 3))) return false;
    return true;
}

template <int MK, int CM>
static h2y_status launch_one(h2y_ctx_impl *c, const FwdArgs &a, int grid, size_t smem, cudaStream_t st)
{
    H2Y_CUDA(c, cudaFuncSetAttribute(k_forward_fused<MK, CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // The sweep behind the fast kernels converts only the frames they left (disjoint output, inputs older than both):
    // the caller hands it an auxiliary stream forked at the same point as the fast kernels (h2y_api.cu) and joins later.
    k_forward_fused<MK, CM><<<grid, THREADS, smem, st>>>(a);
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

template <int MK>
static h2y_status launch_mk(h2y_ctx_impl *c, const FwdArgs &a, int cm, int grid, size_t smem, cudaStream_t st)
{
    switch (cm) {
    case CM_444: return launch_one<MK, CM_444>(c, a, grid, smem, st);
    case CM_420_FIR: return launch_one<MK, CM_420_FIR>(c, a, grid, smem, st);
    case CM_420_BOX: return launch_one<MK, CM_420_BOX>(c, a, grid, smem, st);
    default: return launch_one<MK, CM_422_FIR>(c, a, grid, smem, st);
    }
}

h2y_status launch_forward_fused(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, const void *d_src,
                                size_t src_stride, void *d_dst, size_t dst_stride, int nframes,
                                const FrameK *d_framek, const float *d_luts, cudaStream_t st, int skip_clean,
                                const SpecLaunch *sl, cudaStream_t sweep_stream)
{
    FwdArgs a;
    if (sweep_stream) st = sweep_stream;
    a.ctl = sl ? sl->ctl : nullptr;
    a.flag = sl ? sl->flag : nullptr;
    a.skip_clean = skip_clean;
    a.src = (const uint8_t *)d_src; a.src_stride = src_stride;
    a.dst = (uint8_t *)d_dst; a.dst_stride = dst_stride;
    a.w = p.src.width; a.h = p.src.height; a.nframes = nframes;
    a.layout = p.src.layout;
    a.use_lut = k.convert_transfer;
    a.exact_math = c->sw.exact_math;
    a.k = k; a.framek = d_framek; a.luts = d_luts;
    a.fallback_count = nullptr;

    int cm;
    if (p.dst.chroma_format_idc == H2Y_CHROMA_444) cm = CM_444;
    else if (p.dst.chroma_format_idc == H2Y_CHROMA_422) cm = CM_422_FIR;
    else cm = p.chroma_resampler_type == 0 ? CM_420_BOX : CM_420_FIR;
    const int halo = (cm == CM_420_FIR || cm == CM_422_FIR) ? 1 : 0;
    a.strip_w = (32 - 2 * halo) * 8;
    a.nstrips = (a.w + a.strip_w - 1) / a.strip_w;

    // segment height: enough items to balance the persistent grid, tall enough to amortise the
    // 11 halo rows of the vertical filter
    const int grid_max = c->sm_count;
    const int align = cm == CM_420_BOX ? 32 : 16;
    long want_items = 6L * grid_max;
    int nsegs = (int)((want_items + (long)nframes * a.nstrips - 1) / ((long)nframes * a.nstrips));
    int max_segs = a.h / 96 > 0 ? a.h / 96 : 1;
    if (nsegs > max_segs) nsegs = max_segs;
    if (nsegs < 1) nsegs = 1;
    int seg_rows = (a.h + nsegs - 1) / nsegs;
    seg_rows = (seg_rows + align - 1) / align * align;
    a.seg_rows = seg_rows;
    a.nsegs = (a.h + seg_rows - 1) / seg_rows;
    a.nitems = nframes * a.nsegs * a.nstrips;

    size_t smem = 0;
    if (cm == CM_420_FIR) smem += (size_t)RING_ROWS * 2 * RING_W * sizeof(float);
    const bool half_in = layout_is_half(a.layout);
    if (a.use_lut && half_in) smem += (size_t)LUT_SMEM_CODES * sizeof(float);
    if (smem < 16) smem = 16;
    const int grid = a.nitems < grid_max ? a.nitems : grid_max;

    switch (k.mat_kind) {
    case MK_PASS: return launch_mk<MK_PASS>(c, a, cm, grid, smem, st);
    case MK_YDZDX: return launch_mk<MK_YDZDX>(c, a, cm, grid, smem, st);
    case MK_YCBCR: return launch_mk<MK_YCBCR>(c, a, cm, grid, smem, st);
    case MK_Y100: return launch_mk<MK_Y100>(c, a, cm, grid, smem, st);
    default: return H2Y_ERR_UNSUPPORTED;
    }
}

}   // namespace h2y
