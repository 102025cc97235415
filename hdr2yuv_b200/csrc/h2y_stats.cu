// h2y_stats.cu -- K0: frame statistics (pic_stats, common.cpp:66-168), the per-frame
// normalisation plan, and the exact transfer LUTs.
//
// Why a LUT: steps load -> normalise -> transfer change of matrix_convert
// (convert.cpp:978-1109) are a pure function of the 16-bit input code (u16 sample or half bit
// pattern) and two per-frame, per-channel integers (estimated floor / ceiling).  PQ10000_r is
// 2-3 double pow() per component (convert.cpp:61), far beyond what a memory-bound kernel can
// afford per pixel, so a tiny FP64 kernel evaluates the function once per code (65 536 codes)
// and the fused forward kernel gathers from it.  Bit-exact by construction up to CUDA-vs-glibc
// pow() last-ulp differences before the rounding to float.
#include "h2y_internal.h"

namespace h2y {

// order-preserving float <-> uint key for atomicMin / atomicMax
__device__ __forceinline__ unsigned fkey(float f)
{
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// stats slots: [frame][channel][2] = {min key, max key}
__global__ void k_stats_init(unsigned *slots, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slots[i] = (i & 1) ? 0u : 0xffffffffu;
}

// One block-strided pass over a frame; 16-bit code layouts.  grid = (blocks, nframes).
template <bool HALF>
__global__ void __launch_bounds__(256)
k_stats_codes(const uint16_t *__restrict__ src, size_t frame_stride_elems, int layout, long npix, int clip_on,
              unsigned lo, unsigned hi, unsigned *slots)
{
    const uint16_t *f = src + (size_t)blockIdx.y * frame_stride_elems;
    unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    const bool planar = layout == H2Y_LAYOUT_PLANAR_U16;
    const int nch = (layout == H2Y_LAYOUT_RGBA16 || layout == H2Y_LAYOUT_HALF_RGBA) ? 4 : 3;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long)gridDim.x * blockDim.x) {
        unsigned v[3];
        if (planar) {
            v[0] = f[i]; v[1] = f[npix + i]; v[2] = f[2 * npix + i];
        } else {
            const uint16_t *p = f + i * nch;   // R,G,B(,A) -> G,B,R
            v[0] = p[1]; v[1] = p[2]; v[2] = p[0];
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            unsigned key;
            if (HALF) {
                float x = half_bits_to_float(v[c]);
                if (x != x) continue;            // NaN never wins a '<' or '>' in the reference loop
                key = fkey(x);
            } else {
                unsigned s = v[c];
                if (clip_on) { s = s < lo ? lo : s; s = s > hi ? hi : s; }
                key = s;
            }
            mn[c] = min(mn[c], key);
            mx[c] = max(mx[c], key);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = min(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = max(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&slots[(blockIdx.y * 3 + c) * 2 + 0], mn[c]);
            atomicMax(&slots[(blockIdx.y * 3 + c) * 2 + 1], mx[c]);
        }
    }
}

// planar float pictures (staged API)
__global__ void __launch_bounds__(256)
k_stats_f32(const float *__restrict__ p0, const float *__restrict__ p1, const float *__restrict__ p2, long npix,
            unsigned *slots)
{
    const float *pl[3] = {p0, p1, p2};
    for (int c = 0; c < 3; c++) {
        unsigned mn = 0xffffffffu, mx = 0u;
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long)gridDim.x * blockDim.x) {
            float x = pl[c][i];
            if (x != x) continue;
            unsigned key = fkey(x);
            mn = min(mn, key);
            mx = max(mx, key);
        }
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&slots[c * 2 + 0], mn);
            atomicMax(&slots[c * 2 + 1], mx);
        }
    }
}

__global__ void __launch_bounds__(256)
k_stats_u16_planes(const uint16_t *__restrict__ p0, const uint16_t *__restrict__ p1, const uint16_t *__restrict__ p2,
                   long n0, long n12, unsigned *slots)
{
    const uint16_t *pl[3] = {p0, p1, p2};
    for (int c = 0; c < 3; c++) {
        unsigned mn = 0xffffffffu, mx = 0u;
        long n = c == 0 ? n0 : n12;
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
            unsigned key = pl[c][i];
            mn = min(mn, key);
            mx = max(mx, key);
        }
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&slots[c * 2 + 0], mn);
            atomicMax(&slots[c * 2 + 1], mx);
        }
    }
}

// Turn extrema into estimated floor / ceiling, offset / range and LUT slots.  One block.
__global__ void k_plan(const unsigned *slots, FrameK *fk, int nframes, int is_float, int bit_depth)
{
    const int n = nframes * 3;
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        FrameK &f = fk[p / 3];
        const int c = p % 3;
        unsigned kmin = slots[p * 2], kmax = slots[p * 2 + 1];
        int fl, ce;
        if (is_float) {
            // seeds FLT_MAX / FLT_MIN(smallest positive) of common.cpp:118-119, then (int) truncation 135-136
            float lo = kmin == 0xffffffffu ? 3.402823466e+38f : fminf(fkey_inv(kmin), 3.402823466e+38f);
            float hi = kmax == 0u ? 1.175494351e-38f : fmaxf(fkey_inv(kmax), 1.175494351e-38f);
            f.fmin[c] = lo;
            f.fmax[c] = hi;
            fl = f2i_x86(lo);
            ce = f2i_x86(hi);
        } else {
            unsigned lo = min(kmin, 65535u), hi = kmax;
            f.fmin[c] = (float)lo;
            f.fmax[c] = (float)hi;
            const int D = 1 << (bit_depth - 8);                 // snap cascade, common.cpp:94-106
            const int ymax = 219 * D + 16 * D, cmax = 224 * D + 16 * D;
            fl = (int)lo;
            ce = (int)hi;
            if (ce < ymax && ce > (ymax * 3) / 4) ce = ymax;
            if (ce < cmax && ce > (cmax * 3) / 4) ce = cmax;
        }
        f.est_floor[c] = fl;
        f.est_ceiling[c] = ce;
        f.range[c] = (float)(ce - fl);                           // convert.cpp:939-940
        f.offset[c] = (float)fl;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        FrameK &f = fk[p / 3];
        const int c = p % 3;
        int slot = p;
        for (int q = 0; q < p; q++) {
            const FrameK &g = fk[q / 3];
            if (g.est_floor[q % 3] == f.est_floor[c] && g.est_ceiling[q % 3] == f.est_ceiling[c]) { slot = q; break; }
        }
        f.lut_slot[c] = slot;
    }
    __syncthreads();
    for (int fi = threadIdx.x; fi < nframes; fi += blockDim.x)
        fk[fi].same_lut = fk[fi].lut_slot[0] == fk[fi].lut_slot[1] && fk[fi].lut_slot[0] == fk[fi].lut_slot[2];
}

// grid = (256, nframes*3): LUT p is built only by its owner (lut_slot == p).
__global__ void __launch_bounds__(256)
k_build_lut(const FrameK *fk, float *luts, int is_half, int tf_lin, int tf_enc, int clip_on, unsigned lo, unsigned hi)
{
    const int p = blockIdx.y;
    const FrameK &f = fk[p / 3];
    const int c = p % 3;
    if (f.lut_slot[c] != p) return;
    const unsigned code = blockIdx.x * blockDim.x + threadIdx.x;
    float x;
    if (is_half) x = half_bits_to_float(code);
    else {
        unsigned s = code;
        if (clip_on) { s = s < lo ? lo : s; s = s > hi ? hi : s; }
        x = (float)s;
    }
    x = __fdiv_rn(__fsub_rn(x, f.offset[c]), f.range[c]);        // convert.cpp:1017-1019
    luts[(size_t)p * 65536 + code] = change_transfer(x, tf_lin, tf_enc);
}

// ---- launchers ---------------------------------------------------------------------------------

h2y_status launch_stats_and_luts(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, const void *d_src,
                                 size_t src_stride, int nframes, FrameK **d_framek, float **d_luts, cudaStream_t st)
{
    void *slots, *fk, *luts;
    h2y_status s;
    if ((s = scratch_reserve(c, SCR_STATS, (size_t)nframes * 6 * sizeof(unsigned), &slots)) != H2Y_OK) return s;
    if ((s = scratch_reserve(c, SCR_FRAMEK, (size_t)nframes * sizeof(FrameK), &fk)) != H2Y_OK) return s;
    if ((s = scratch_reserve(c, SCR_LUT, (size_t)nframes * 3 * 65536 * sizeof(float), &luts)) != H2Y_OK) return s;
    const long npix = (long)p.src.width * p.src.height;
    const int nslots = nframes * 6;
    k_stats_init<<<(nslots + 255) / 256, 256, 0, st>>>((unsigned *)slots, nslots);
    const int blocks = (int)(((npix + 2047) / 2048) < 4L * c->sm_count ? ((npix + 2047) / 2048) : 4L * c->sm_count);
    dim3 grid(blocks, nframes);
    const bool half = layout_is_half(p.src.layout);
    if (half)
        k_stats_codes<true><<<grid, 256, 0, st>>>((const uint16_t *)d_src, src_stride / 2, p.src.layout, npix, 0, 0, 0,
                                                  (unsigned *)slots);
    else
        k_stats_codes<false><<<grid, 256, 0, st>>>((const uint16_t *)d_src, src_stride / 2, p.src.layout, npix,
                                                   k.clip_on_load, k.loadLo, k.loadHi, (unsigned *)slots);
    k_plan<<<1, 256, 0, st>>>((const unsigned *)slots, (FrameK *)fk, nframes, half ? 1 : 0, p.src.bit_depth);
    k_build_lut<<<dim3(256, nframes * 3), 256, 0, st>>>((const FrameK *)fk, (float *)luts, half ? 1 : 0, k.tf_linearise,
                                                         k.tf_encode, k.clip_on_load, k.loadLo, k.loadHi);
    c->launches += 4;
    H2Y_CUDA(c, cudaGetLastError());
    *d_framek = (FrameK *)fk;
    *d_luts = (float *)luts;
    return H2Y_OK;
}

h2y_status launch_stats_planar(h2y_ctx_impl *c, const h2y_pic_desc &pic, const void *const d_planes[3],
                               FrameK **d_framek, cudaStream_t st)
{
    void *slots, *fk;
    h2y_status s;
    if ((s = scratch_reserve(c, SCR_STATS, 6 * sizeof(unsigned), &slots)) != H2Y_OK) return s;
    if ((s = scratch_reserve(c, SCR_FRAMEK, sizeof(FrameK), &fk)) != H2Y_OK) return s;
    int pw[3], ph[3];
    h2y_plane_dims(pic.width, pic.height, pic.chroma_format_idc, pw, ph);
    const long n0 = (long)pw[0] * ph[0], n12 = (long)pw[1] * ph[1];
    k_stats_init<<<1, 256, 0, st>>>((unsigned *)slots, 6);
    const int blocks = (int)(((n0 + 2047) / 2048) < 4L * c->sm_count ? ((n0 + 2047) / 2048) : 4L * c->sm_count);
    const bool isf = pic.pic_buffer_type == H2Y_PIC_TYPE_F32;
    if (isf)
        k_stats_f32<<<blocks, 256, 0, st>>>((const float *)d_planes[0], (const float *)d_planes[1],
                                            (const float *)d_planes[2], n0, (unsigned *)slots);
    else
        k_stats_u16_planes<<<blocks, 256, 0, st>>>((const uint16_t *)d_planes[0], (const uint16_t *)d_planes[1],
                                                   (const uint16_t *)d_planes[2], n0, n12, (unsigned *)slots);
    k_plan<<<1, 256, 0, st>>>((const unsigned *)slots, (FrameK *)fk, 1, isf ? 1 : 0, pic.bit_depth);
    c->launches += 3;
    H2Y_CUDA(c, cudaGetLastError());
    *d_framek = (FrameK *)fk;
    return H2Y_OK;
}

}   // namespace h2y
