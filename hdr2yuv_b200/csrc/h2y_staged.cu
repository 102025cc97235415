// h2y_staged.cu -- one simple kernel per reference function, planar buffers, any geometry.
//
// These back the staged C-ABI entry points (h2y_matrix_convert, h2y_convert,
// h2y_write_yuv_clamp, h2y_subsample_420_to_444) that exist so parity of INTERMEDIATES can be
// tested, and they are the general-geometry route of h2y_forward when the fused kernel's
// vector path does not apply (width not a multiple of 8, planar-float input, ...).  They
// evaluate the transfer functions directly in FP64 per pixel (no LUT), which also makes them an
// independent on-GPU check of the LUT route.  One thread per sample; not the fast path.
#include "h2y_internal.h"

namespace h2y {

// ---- matrix_convert (convert.cpp:879-1315) -----------------------------------------------------

__device__ __forceinline__ void px_matrix_f32out(int mk, float G, float B, float R, const PixK &k, float half,
                                                 float &Y, float &Cb, float &Cr)
{
    if (mk == MK_PASS) { Y = G; Cb = B; Cr = R; }
    else {
        if (mk == MK_YDZDX) {
            Y = G;
            double hg = __dmul_rn((double)(-G), 0.5);
            Cb = __double2float_rn(__dadd_rn(__dadd_rn(hg, __dmul_rn((double)B, 0.5)), 0.5));
            Cr = __double2float_rn(__dadd_rn(__dadd_rn(hg, __dmul_rn((double)R, 0.5)), 0.5));
        } else if (mk == MK_YCBCR) {
            double s = __dadd_rn(__dadd_rn(__dmul_rn(k.wr, (double)R), __dmul_rn(k.wg, (double)G)),
                                 __dmul_rn(k.wb, (double)B));
            float tmpF = __double2float_rn(__dadd_rn(s, 0.5));
            Y = tmpF;
            Cb = __double2float_rn(__dadd_rn(__ddiv_rn((double)__fsub_rn(B, tmpF), k.db), 0.5));
            Cr = __double2float_rn(__dadd_rn(__ddiv_rn((double)__fsub_rn(R, tmpF), k.dr), 0.5));
        } else if (mk == MK_Y100) {
            Y = G;
            Cb = __double2float_rn(__dadd_rn((double)__fadd_rn(__fmul_rn(k.P, G), __fmul_rn(k.Q, B)), 0.5));
            Cr = __double2float_rn(__dadd_rn((double)__fadd_rn(__fmul_rn(k.RR, R), __fmul_rn(k.S, G)), 0.5));
        } else { Y = G; Cb = B; Cr = R; }
        Cb = __fsub_rn(__fadd_rn(Cb, half), 1.0f);     // Cb + clip->Half - 1, convert.cpp:1280-1281
        Cr = __fsub_rn(__fadd_rn(Cr, half), 1.0f);
    }
    const float hi = (float)k.maxCV;
    if (Y > hi) Y = hi;
    if (Y < 0.0f) Y = 0.0f;
    if (Cb > hi) Cb = hi;
    if (Cb < 0.0f) Cb = 0.0f;
    if (Cr > hi) Cr = hi;
    if (Cr < 0.0f) Cr = 0.0f;
}

template <bool IN_F32, bool OUT_F32>
__global__ void __launch_bounds__(256)
k_matrix_convert(PixK k, NormK nk, long npix, const void *in0, const void *in1, const void *in2, void *out0,
                 void *out1, void *out2)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float G, B, R;
    if (IN_F32) {
        G = ((const float *)in0)[i]; B = ((const float *)in1)[i]; R = ((const float *)in2)[i];
    } else {
        G = (float)((const uint16_t *)in0)[i]; B = (float)((const uint16_t *)in1)[i]; R = (float)((const uint16_t *)in2)[i];
    }
    if (k.convert_transfer) {
        G = change_transfer(__fdiv_rn(__fsub_rn(G, nk.offset[0]), nk.range[0]), k.tf_linearise, k.tf_encode);
        B = change_transfer(__fdiv_rn(__fsub_rn(B, nk.offset[1]), nk.range[1]), k.tf_linearise, k.tf_encode);
        R = change_transfer(__fdiv_rn(__fsub_rn(R, nk.offset[2]), nk.range[2]), k.tf_linearise, k.tf_encode);
        if (OUT_F32) {                                  // convert.cpp:1116-1122
            G = __fadd_rn(__fmul_rn(G, nk.range[0]), nk.offset[0]);
            B = __fadd_rn(__fmul_rn(B, nk.range[1]), nk.offset[1]);
            R = __fadd_rn(__fmul_rn(R, nk.range[2]), nk.offset[2]);
        } else scale_to_codes(G, B, R, k);
    }
    if (OUT_F32) {
        float Y, Cb, Cr;
        px_matrix_f32out(k.mat_kind, G, B, R, k, (float)(k.half_m1 + 1), Y, Cb, Cr);
        ((float *)out0)[i] = Y; ((float *)out1)[i] = Cb; ((float *)out2)[i] = Cr;
    } else {
        unsigned Y, Cb, Cr;
        switch (k.mat_kind) {
        case MK_PASS: px_matrix_exact<MK_PASS>(G, B, R, k, Y, Cb, Cr); break;
        case MK_YDZDX: px_matrix_exact<MK_YDZDX>(G, B, R, k, Y, Cb, Cr); break;
        case MK_YCBCR: px_matrix_exact<MK_YCBCR>(G, B, R, k, Y, Cb, Cr); break;
        case MK_Y100: px_matrix_exact<MK_Y100>(G, B, R, k, Y, Cb, Cr); break;
        default: px_matrix_exact<MK_PRIME2>(G, B, R, k, Y, Cb, Cr); break;
        }
        ((uint16_t *)out0)[i] = (uint16_t)Y; ((uint16_t *)out1)[i] = (uint16_t)Cb; ((uint16_t *)out2)[i] = (uint16_t)Cr;
    }
}

h2y_status launch_matrix_convert(h2y_ctx_impl *c, const PixK &k, const NormK &nk, int w, int h, int in_is_f32,
                                 const void *const d_in[3], int out_is_f32, void *const d_out[3], cudaStream_t st)
{
    const long npix = (long)w * h;
    const int blocks = (int)((npix + 255) / 256);
#define MC(A, B) k_matrix_convert<A, B><<<blocks, 256, 0, st>>>(k, nk, npix, d_in[0], d_in[1], d_in[2], d_out[0], d_out[1], d_out[2])
    if (in_is_f32 && out_is_f32) MC(true, true);
    else if (in_is_f32) MC(true, false);
    else if (out_is_f32) MC(false, true);
    else MC(false, false);
#undef MC
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

// ---- reader de-interleave (tiff.cpp:265-315, exr.cpp:209-235) -----------------------------------
// u16 layouts -> three u16 planes (with the on-read clip); half layouts -> three float planes.
__global__ void __launch_bounds__(256)
k_unpack(int layout, long npix, const uint16_t *__restrict__ src, void *o0, void *o1, void *o2, int clip_on,
         unsigned lo, unsigned hi)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    unsigned v[3];
    if (layout == H2Y_LAYOUT_PLANAR_U16) { v[0] = src[i]; v[1] = src[npix + i]; v[2] = src[2 * npix + i]; }
    else {
        const int nch = (layout == H2Y_LAYOUT_RGBA16 || layout == H2Y_LAYOUT_HALF_RGBA) ? 4 : 3;
        const uint16_t *p = src + i * nch;
        v[0] = p[1]; v[1] = p[2]; v[2] = p[0];
    }
    if (layout == H2Y_LAYOUT_HALF_RGB || layout == H2Y_LAYOUT_HALF_RGBA) {
        ((float *)o0)[i] = half_bits_to_float(v[0]);
        ((float *)o1)[i] = half_bits_to_float(v[1]);
        ((float *)o2)[i] = half_bits_to_float(v[2]);
    } else {
        if (clip_on)
            for (int c = 0; c < 3; c++) { v[c] = v[c] < lo ? lo : v[c]; v[c] = v[c] > hi ? hi : v[c]; }
        ((uint16_t *)o0)[i] = (uint16_t)v[0];
        ((uint16_t *)o1)[i] = (uint16_t)v[1];
        ((uint16_t *)o2)[i] = (uint16_t)v[2];
    }
}

// 10-bit packed DPX words -> three float planes G,B,R with sample = code / 1023.0, the double division rounded to
// float as dpx_read's assignment does (dpx.cpp:506-531; muxed_dpx_to_planar_float_buf, common.cpp:12-28)
__global__ void __launch_bounds__(256)
k_unpack_dpx10(long npix, const unsigned *__restrict__ src, int big_endian, float *__restrict__ g, float *__restrict__ b,
               float *__restrict__ r)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    unsigned w = __ldg(src + i);
    if (big_endian) w = __byte_perm(w, 0, 0x0123);
    const unsigned rr = w >> 22, gg = (w >> 12) & 1023u, bb = (w >> 2) & 1023u;
    r[i] = __double2float_rn(__ddiv_rn((double)rr, 1023.0));
    g[i] = __double2float_rn(__ddiv_rn((double)gg, 1023.0));
    b[i] = __double2float_rn(__ddiv_rn((double)bb, 1023.0));
}

// 16-bit DPX: interleaved R,G,B u16 in file byte order -> float planes, sample = code / 65535.0 with the double division
// rounded to float as dpx_read's assignment does (dpx.cpp:478-494)
__global__ void __launch_bounds__(256)
k_unpack_dpx16(long npix, const uint16_t *__restrict__ src, int big_endian, float *__restrict__ g, float *__restrict__ b,
               float *__restrict__ r)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    unsigned v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        v[c] = src[3 * i + c];
        if (big_endian) v[c] = ((v[c] & 0xffu) << 8) | (v[c] >> 8);
    }
    r[i] = __double2float_rn(__ddiv_rn((double)v[0], 65535.0));
    g[i] = __double2float_rn(__ddiv_rn((double)v[1], 65535.0));
    b[i] = __double2float_rn(__ddiv_rn((double)v[2], 65535.0));
}

// 32-bit float DPX: interleaved R,G,B floats in file byte order, taken as they are (dpx.cpp:412-443)
__global__ void __launch_bounds__(256)
k_unpack_dpxf32(long npix, const unsigned *__restrict__ src, int big_endian, float *__restrict__ g, float *__restrict__ b,
                float *__restrict__ r)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    unsigned v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        v[c] = __ldg(src + 3 * i + c);
        if (big_endian) v[c] = __byte_perm(v[c], 0, 0x0123);
    }
    r[i] = __uint_as_float(v[0]); g[i] = __uint_as_float(v[1]); b[i] = __uint_as_float(v[2]);
}

h2y_status launch_unpack(h2y_ctx_impl *c, int layout, int w, int h, const void *d_src, void *const d_planes[3],
                         int clip_on_load, unsigned lo, unsigned hi, cudaStream_t st)
{
    const long npix = (long)w * h;
    if (layout == H2Y_LAYOUT_DPX16_BE || layout == H2Y_LAYOUT_DPX16_LE || layout == H2Y_LAYOUT_DPXF32_BE || layout == H2Y_LAYOUT_DPXF32_LE) {
        const int grid = (int)((npix + 255) / 256);
        if (layout == H2Y_LAYOUT_DPX16_BE || layout == H2Y_LAYOUT_DPX16_LE)
            k_unpack_dpx16<<<grid, 256, 0, st>>>(npix, (const uint16_t *)d_src, layout == H2Y_LAYOUT_DPX16_BE, (float *)d_planes[0],
                                                 (float *)d_planes[1], (float *)d_planes[2]);
        else
            k_unpack_dpxf32<<<grid, 256, 0, st>>>(npix, (const unsigned *)d_src, layout == H2Y_LAYOUT_DPXF32_BE, (float *)d_planes[0],
                                                  (float *)d_planes[1], (float *)d_planes[2]);
        c->launches++;
        H2Y_CUDA(c, cudaGetLastError());
        return H2Y_OK;
    }
    if (layout_is_dpx(layout)) {
        k_unpack_dpx10<<<(int)((npix + 255) / 256), 256, 0, st>>>(npix, (const unsigned *)d_src, layout == H2Y_LAYOUT_DPX10_BE,
                                                                   (float *)d_planes[0], (float *)d_planes[1], (float *)d_planes[2]);
        c->launches++;
        H2Y_CUDA(c, cudaGetLastError());
        return H2Y_OK;
    }
    k_unpack<<<(int)((npix + 255) / 256), 256, 0, st>>>(layout, npix, (const uint16_t *)d_src, d_planes[0], d_planes[1],
                                                         d_planes[2], clip_on_load, lo, hi);
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

// ---- Subsample444to420_FIR (convert.cpp:261-383), two passes with the u16 intermediate ---------
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(256)
k_fir_h(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int w, int h, float maxCV)
{
    const int wh = w >> 1;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)wh * h) return;
    const int y = (int)(i / wh), x = 2 * (int)(i % wh);
    const uint16_t *s = src + (long)y * w;
    dst[i] = (uint16_t)fir_h7((float)s[clampi(x - 5, 0, w - 1)], (float)s[clampi(x - 3, 0, w - 1)],
                              (float)s[clampi(x - 1, 0, w - 1)], (float)s[x], (float)s[clampi(x + 1, 0, w - 1)],
                              (float)s[clampi(x + 3, 0, w - 1)], (float)s[clampi(x + 5, 0, w - 1)], maxCV);
}

__global__ void __launch_bounds__(256)
k_fir_v(const uint16_t *__restrict__ mid, uint16_t *__restrict__ dst, int wh, int h, float maxCV)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int hh = h >> 1;
    if (i >= (long)wh * hh) return;
    const int j = (int)(i / wh), x = (int)(i % wh), y = 2 * j;
    float r[12];
#pragma unroll
    for (int t = 0; t < 12; t++) r[t] = (float)mid[(long)clampi(y - 5 + t, 0, h - 1) * wh + x];
    dst[i] = (uint16_t)fir_v12(r, maxCV);
}

h2y_status launch_fir_420(h2y_ctx_impl *c, const uint16_t *d_src, uint16_t *d_dst, uint16_t *d_mid, int w, int h,
                          unsigned maxCV, cudaStream_t st)
{
    const int wh = w >> 1;
    const long n1 = (long)wh * h, n2 = (long)wh * (h >> 1);
    k_fir_h<<<(int)((n1 + 255) / 256), 256, 0, st>>>(d_src, d_mid, w, h, (float)maxCV);
    k_fir_v<<<(int)((n2 + 255) / 256), 256, 0, st>>>(d_mid, d_dst, wh, h, (float)maxCV);
    c->launches += 2;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

h2y_status launch_fir_422(h2y_ctx_impl *c, const uint16_t *d_src, uint16_t *d_dst, int w, int h, unsigned maxCV,
                          cudaStream_t st)
{
    const long n1 = (long)(w >> 1) * h;
    k_fir_h<<<(int)((n1 + 255) / 256), 256, 0, st>>>(d_src, d_dst, w, h, (float)maxCV);
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

// ---- Subsample444to420_box (convert.cpp:91-172) --------------------------------------------------
// The reference walks 4x4 blocks and reads out of bounds unless w and h are multiples of 4; the
// API layer rejects other sizes.
__global__ void __launch_bounds__(256)
k_box(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int w, int h)
{
    const int wh = w / 2, hh = h / 2;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)wh * hh) return;
    const int j = (int)(i / wh), x = (int)(i % wh);
    const uint16_t *p = src + (long)(2 * j) * w + 2 * x;
    unsigned sum = (unsigned)p[0] + p[1] + p[w] + p[w + 1];
    dst[i] = (uint16_t)(sum / 4);
}

h2y_status launch_box_420(h2y_ctx_impl *c, const uint16_t *d_src, uint16_t *d_dst, int w, int h, cudaStream_t st)
{
    const long n = (long)(w / 2) * (h / 2);
    k_box<<<(int)((n + 255) / 256), 256, 0, st>>>(d_src, d_dst, w, h);
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

// ---- Y'u''v'' 4:2:0 (convert.cpp:533-801) ------------------------------------------------------------
// Stage 1: linear Y from the rho-gamma coded 16-bit luma (561-596).  One table entry per code, built on the fly:
// the function of a single 16-bit value is evaluated once per code (FP64 pow), then the plane is a gather.
__global__ void __launch_bounds__(256) k_prime2_lut(uint16_t *lut)
{
    const unsigned code = blockIdx.x * blockDim.x + threadIdx.x;
    const float gamma_Y = __double2float_rn(__ddiv_rn((double)(float)code, 65535.0));
    const float Y_f = tf_rho_eotf(gamma_Y);
    lut[code] = (uint16_t)d2i_x86(__dmul_rn((double)Y_f, 65535.0));
}

__global__ void __launch_bounds__(256) k_prime2_linear_y(const uint16_t *__restrict__ y, const uint16_t *__restrict__ lut,
                                                         uint16_t *__restrict__ lin, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lin[i] = __ldg(lut + y[i]);
}

// Stage 3: u' = 4X/(X+15Y+3Z), v' = 9Y/(X+15Y+3Z) on the subsampled planes, clipped to [0,1], 16-bit (662-753).
// The reference then replaces u'',v'' by u',v' ("HACK", 732-734), so the coded luma and the 0.25 floor drop out.
__global__ void __launch_bounds__(256) k_prime2_uv(const uint16_t *__restrict__ sX, const uint16_t *__restrict__ sY,
                                                   const uint16_t *__restrict__ sZ, uint16_t *__restrict__ u,
                                                   uint16_t *__restrict__ v, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double X = __ddiv_rn((double)sX[i], 65535.0), Z = __ddiv_rn((double)sZ[i], 65535.0), Y = __ddiv_rn((double)sY[i], 65535.0);
    const double sum = __dadd_rn(__dadd_rn(X, __dmul_rn(15.0, Y)), __dmul_rn(3.0, Z));
    double up = 0.0, vp = 0.0;
    if (sum > 0.0) { up = __ddiv_rn(__dmul_rn(4.0, X), sum); vp = __ddiv_rn(__dmul_rn(9.0, Y), sum); }
    up = up < 0.0 ? 0.0 : up; vp = vp < 0.0 ? 0.0 : vp;
    up = up > 1.0 ? 1.0 : up; vp = vp > 1.0 ? 1.0 : vp;
    u[i] = (uint16_t)d2i_x86(__dmul_rn(up, 65535.0));
    v[i] = (uint16_t)d2i_x86(__dmul_rn(vp, 65535.0));
}

// d_in: the 4:4:4 planes Y', Z, X (plane 1 is Z, plane 2 is X: convert.cpp:593-594); d_u / d_v: (w/2)x(h/2).
// scratch: 4 planes of w*h u16 (linear Y, three subsampled planes sharing one, FIR intermediate) + the 128 KiB table.
h2y_status launch_yuvprime2_420(h2y_ctx_impl *c, const uint16_t *const d_in[3], uint16_t *d_u, uint16_t *d_v, uint16_t *scratch,
                                int w, int h, int resampler, unsigned maxCV, cudaStream_t st)
{
    const long n = (long)w * h, nc = (long)(w / 2) * (h / 2);
    uint16_t *lin = scratch, *sub = scratch + n, *mid = scratch + 2 * n, *lut = scratch + 3 * n;
    uint16_t *sY = sub, *sZ = sub + nc, *sX = sub + 2 * nc;
    k_prime2_lut<<<256, 256, 0, st>>>(lut);
    k_prime2_linear_y<<<(int)((n + 255) / 256), 256, 0, st>>>(d_in[0], lut, lin, n);
    c->launches += 2;
    const uint16_t *src[3] = {lin, d_in[1], d_in[2]};
    uint16_t *dst[3] = {sY, sZ, sX};
    for (int p = 0; p < 3; p++) {
        h2y_status s = resampler == 0 ? launch_box_420(c, src[p], dst[p], w, h, st)
                                      : launch_fir_420(c, src[p], dst[p], mid, w, h, maxCV, st);
        if (s != H2Y_OK) return s;
    }
    k_prime2_uv<<<(int)((nc + 255) / 256), 256, 0, st>>>(sX, sY, sZ, d_u, d_v, nc);
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

// ---- write_yuv compute stage (tiff.cpp:457-550) ---------------------------------------------------
__global__ void __launch_bounds__(256)
k_out_clamp(uint16_t *p, size_t n, int shift, unsigned lo, unsigned hi)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint16_t)out_clamp(p[i], shift, lo, hi);
}

h2y_status launch_out_clamp(h2y_ctx_impl *c, uint16_t *d_plane, size_t n, int shift, unsigned lo, unsigned hi,
                            cudaStream_t st)
{
    if (n == 0) return H2Y_OK;
    k_out_clamp<<<(int)((n + 255) / 256), 256, 0, st>>>(d_plane, n, shift, lo, hi);
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

// ---- optional linear-light stage behind the inverse path: PQ10000_f per 16-bit code --------------
__global__ void __launch_bounds__(256) k_build_pq_linear_lut(float *lut)
{
    const unsigned code = blockIdx.x * blockDim.x + threadIdx.x;
    lut[code] = tf_pq_eotf(__fdiv_rn((float)code, 65535.0f));                 // convert.cpp:1017-1019 with floor 0, ceiling 65535
}

// eight codes per thread: one 16-byte load, two 16-byte stores; the 256 KB table stays in L1 / L2
__global__ void __launch_bounds__(256) k_pq_codes_to_linear(const uint16_t *__restrict__ codes, size_t n, const float *__restrict__ lut,
                                                            float *__restrict__ out)
{
    const size_t nvec = n / 8;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        const uint4 c = __ldg(reinterpret_cast<const uint4 *>(codes) + v);
        const unsigned w[4] = {c.x, c.y, c.z, c.w};
        float f[8];
#pragma unroll
        for (int i = 0; i < 4; i++) { f[2 * i] = __ldg(lut + (w[i] & 0xffffu)); f[2 * i + 1] = __ldg(lut + (w[i] >> 16)); }
        float4 *o = reinterpret_cast<float4 *>(out) + 2 * v;
        o[0] = make_float4(f[0], f[1], f[2], f[3]);
        o[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (unsigned)(n - nvec * 8)) out[nvec * 8 + threadIdx.x] = __ldg(lut + codes[nvec * 8 + threadIdx.x]);
}

h2y_status launch_pq_codes_to_linear(h2y_ctx_impl *c, const uint16_t *d_codes, size_t n, float *d_linear, cudaStream_t st)
{
    if (n == 0) return H2Y_OK;
    if (!c->pq_linear_lut) {
        H2Y_CUDA(c, cudaMalloc(&c->pq_linear_lut, 65536 * sizeof(float)));
        k_build_pq_linear_lut<<<256, 256, 0, st>>>(c->pq_linear_lut);
        c->launches++;
    }
    const bool vec = (((uintptr_t)d_codes | (uintptr_t)d_linear) & 15) == 0;
    if (!vec) return H2Y_ERR_ARG;                                        // both buffers 16-byte aligned
    const size_t want = (n / 8 + 255) / 256;
    const int blocks = (int)(want < (size_t)(16 * c->sm_count) ? (want ? want : 1) : (size_t)(16 * c->sm_count));
    k_pq_codes_to_linear<<<blocks, 256, 0, st>>>(d_codes, n, c->pq_linear_lut, d_linear);
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

// ---- Subsample420to444 (yuv2tiff.cpp:575-692), row-major ------------------------------------------
__device__ __forceinline__ unsigned up_finish(float t, float lo, float hi)
{
    t = __fadd_rn(t, 0.5f);
    if (t > hi) t = hi;
    if (t < lo) t = lo;
    return (unsigned)__float2int_rz(t);
}

// vertical 2-phase 6-tap: src wh x hh -> mid wh x h
__global__ void __launch_bounds__(256)
k_up_v(const uint16_t *__restrict__ src, uint16_t *__restrict__ mid, int wh, int hh, float lo, float hi)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)wh * hh) return;
    const int j = (int)(i / wh), x = (int)(i % wh);
    const float a3 = 3.0f / 256.0f, a16 = 16.0f / 256.0f, a67 = 67.0f / 256.0f, a227 = 227.0f / 256.0f,
                a32 = 32.0f / 256.0f, a7 = 7.0f / 256.0f;
    float m3 = (float)src[(long)clampi(j - 3, 0, hh - 1) * wh + x], m2 = (float)src[(long)clampi(j - 2, 0, hh - 1) * wh + x],
          m1 = (float)src[(long)clampi(j - 1, 0, hh - 1) * wh + x], c0 = (float)src[(long)j * wh + x],
          p1 = (float)src[(long)clampi(j + 1, 0, hh - 1) * wh + x], p2 = (float)src[(long)clampi(j + 2, 0, hh - 1) * wh + x],
          p3 = (float)src[(long)clampi(j + 3, 0, hh - 1) * wh + x];
    float t = __fmul_rn(a3, m3);
    t = __fsub_rn(t, __fmul_rn(a16, m2));
    t = __fadd_rn(t, __fmul_rn(a67, m1));
    t = __fadd_rn(t, __fmul_rn(a227, c0));
    t = __fsub_rn(t, __fmul_rn(a32, p1));
    t = __fadd_rn(t, __fmul_rn(a7, p2));
    mid[(long)(2 * j) * wh + x] = (uint16_t)up_finish(t, lo, hi);
    t = __fmul_rn(a3, p3);
    t = __fsub_rn(t, __fmul_rn(a16, p2));
    t = __fadd_rn(t, __fmul_rn(a67, p1));
    t = __fadd_rn(t, __fmul_rn(a227, c0));
    t = __fsub_rn(t, __fmul_rn(a32, m1));
    t = __fadd_rn(t, __fmul_rn(a7, m2));
    mid[(long)(2 * j + 1) * wh + x] = (uint16_t)up_finish(t, lo, hi);
}

// horizontal: mid wh x h -> dst w x h
__global__ void __launch_bounds__(256)
k_up_h(const uint16_t *__restrict__ mid, uint16_t *__restrict__ dst, int wh, int h, float lo, float hi)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)wh * h) return;
    const int y = (int)(i / wh), x = (int)(i % wh);
    const uint16_t *r = mid + (long)y * wh;
    const float b21 = 21.0f / 256.0f, b52 = 52.0f / 256.0f, b159 = 159.0f / 256.0f;
    float l2 = (float)r[clampi(x - 2, 0, wh - 1)], l1 = (float)r[clampi(x - 1, 0, wh - 1)], c0 = (float)r[x],
          r1 = (float)r[clampi(x + 1, 0, wh - 1)], r2 = (float)r[clampi(x + 2, 0, wh - 1)],
          r3 = (float)r[clampi(x + 3, 0, wh - 1)];
    float t = __fmul_rn(b21, __fadd_rn(l2, r3));
    t = __fsub_rn(t, __fmul_rn(b52, __fadd_rn(l1, r2)));
    t = __fadd_rn(t, __fmul_rn(b159, __fadd_rn(c0, r1)));
    uint16_t *o = dst + (long)y * (2 * wh) + 2 * x;
    o[0] = r[x];
    o[1] = (uint16_t)up_finish(t, lo, hi);
}

__global__ void __launch_bounds__(256)
k_up_box(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int wh, int hh)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)wh * hh) return;
    const int j = (int)(i / wh), x = (int)(i % wh), w = 2 * wh;
    const uint16_t v = src[i];
    dst[(long)(2 * j) * w + 2 * x] = v;
    dst[(long)(2 * j) * w + 2 * x + 1] = v;
    dst[(long)(2 * j + 1) * w + 2 * x] = v;
    dst[(long)(2 * j + 1) * w + 2 * x + 1] = v;
}

h2y_status launch_upsample(h2y_ctx_impl *c, const uint16_t *d_src, uint16_t *d_dst, uint16_t *d_mid, int w, int h,
                           int fir, unsigned minCV, unsigned maxCV, cudaStream_t st)
{
    const int wh = w / 2, hh = h / 2;
    const long n = (long)wh * hh;
    if (!fir) {
        k_up_box<<<(int)((n + 255) / 256), 256, 0, st>>>(d_src, d_dst, wh, hh);
        c->launches++;
    } else {
        k_up_v<<<(int)((n + 255) / 256), 256, 0, st>>>(d_src, d_mid, wh, hh, (float)minCV, (float)maxCV);
        const long n2 = (long)wh * h;
        k_up_h<<<(int)((n2 + 255) / 256), 256, 0, st>>>(d_mid, d_dst, wh, h, (float)minCV, (float)maxCV);
        c->launches += 2;
    }
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

// ---- matrix_inverse (convert.cpp:1320-1867) and write_tiff's compute (tiff.cpp:605-628) ----------------
// hdr2yuv's 4:4:4-only inverse, float variables with double sub-expressions exactly as written, including the
// hard-coded Half = 2048 / Full = 4096 (1337-1338) and the family selection quirk (1387): matrix_coeffs == 1 takes
// the BT.709 equations, everything else except 0 the Y'DzDx ones.  family: 0 = Y'DzDx, 1 = BT.709.
struct MinvK {
    int family;
    int minVR, maxVR;        // in_pic->clip (FULLRANGE is hard-coded 0, 1343)
    int shift_right, shift_left;   // 1796-1809, then write_tiff's << (pic depth - src depth) when interleaving
};

__device__ __forceinline__ int minv_pixel(const MinvK &k, unsigned y, unsigned cb, unsigned cr, int &R, int &G, int &B)
{
    float Yav = (float)y, Cb = (float)cb, Cr = (float)cr, Rp, Bp, tmpF;
    if (k.family == 0) {
        Rp = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(2.0, (double)Cr), -4095.0), (double)Yav));
        Bp = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(2.0, (double)Cb), -4095.0), (double)Yav));
    } else {
        tmpF = __double2float_rn(__dadd_rn(__dmul_rn(__dadd_rn((double)Cb, -2047.5), 1.8556), (double)Yav));
        if ((double)tmpF > 4095.0) tmpF = 4095.0f;
        Bp = tmpF;
        tmpF = __double2float_rn(__dadd_rn(__dmul_rn(__dadd_rn((double)Cr, -2047.5), 1.5748), (double)Yav));
        if ((double)tmpF > 4095.0) tmpF = 4095.0f;
        Rp = tmpF;
        const double g = __dadd_rn(__dadd_rn((double)Yav, -__dmul_rn(0.07222, (double)Bp)), -__dmul_rn(0.2126, (double)Rp));
        tmpF = __double2float_rn(__dadd_rn(__ddiv_rn(g, 0.7152), 0.5));
        if ((double)tmpF > 4095.0) tmpF = 4095.0f;
        Yav = tmpF;
    }
    G = f2i_x86(Yav); B = f2i_x86(Bp); R = f2i_x86(Rp);
    int invalid = 0;
    if (G < 0) { G = 0; invalid++; }
    if (R < 0) { R = 0; invalid++; }
    if (B < 0) { B = 0; invalid++; }
    R = R < k.minVR ? k.minVR : R; G = G < k.minVR ? k.minVR : G; B = B < k.minVR ? k.minVR : B;
    R = R > k.maxVR ? k.maxVR : R; G = G > k.maxVR ? k.maxVR : G; B = B > k.maxVR ? k.maxVR : B;
    R = (R >> k.shift_right) << k.shift_left; G = (G >> k.shift_right) << k.shift_left; B = (B >> k.shift_right) << k.shift_left;
    return invalid;
}

// planar Y,Cb,Cr -> planar G,B,R (INTERLEAVE = false, the staged matrix_inverse) or interleaved R,G,B rows
// (INTERLEAVE = true: matrix_inverse + write_tiff fused); grid.y = frame
template <bool INTERLEAVE>
__global__ void __launch_bounds__(256)
k_matrix_inverse(MinvK k, long npix, const uint16_t *__restrict__ in, size_t in_stride, uint16_t *__restrict__ out,
                 size_t out_stride, uint32_t *invalid_out)
{
    const uint16_t *fi = in + (size_t)blockIdx.y * in_stride;
    uint16_t *fo = out + (size_t)blockIdx.y * out_stride;
    int invalid = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long)gridDim.x * blockDim.x) {
        int R, G, B;
        invalid += minv_pixel(k, fi[i], fi[npix + i], fi[2 * npix + i], R, G, B);
        if (INTERLEAVE) { fo[3 * i] = (uint16_t)R; fo[3 * i + 1] = (uint16_t)G; fo[3 * i + 2] = (uint16_t)B; }
        else { fo[i] = (uint16_t)G; fo[npix + i] = (uint16_t)B; fo[2 * npix + i] = (uint16_t)R; }
    }
    if (invalid_out) {
        for (int o = 16; o > 0; o >>= 1) invalid += __shfl_xor_sync(0xffffffffu, invalid, o);
        if ((threadIdx.x & 31) == 0 && invalid) atomicAdd(&invalid_out[blockIdx.y], (uint32_t)invalid);
    }
}

// write_tiff's compute on its own: planes G,B,R -> interleaved R,G,B, << sr
__global__ void __launch_bounds__(256)
k_write_tiff_rows(long npix, const uint16_t *__restrict__ g, const uint16_t *__restrict__ b, const uint16_t *__restrict__ r,
                  uint16_t *__restrict__ rgb, int sr)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long)gridDim.x * blockDim.x) {
        rgb[3 * i] = (uint16_t)((int)r[i] << sr); rgb[3 * i + 1] = (uint16_t)((int)g[i] << sr); rgb[3 * i + 2] = (uint16_t)((int)b[i] << sr);
    }
}

h2y_status launch_matrix_inverse(h2y_ctx_impl *c, int family, int minVR, int maxVR, int shift_right, int shift_left,
                                 long npix, const uint16_t *d_in, size_t in_stride_elems, uint16_t *d_out,
                                 size_t out_stride_elems, int nframes, int interleave, uint32_t *d_invalid, cudaStream_t st)
{
    MinvK k = {family, minVR, maxVR, shift_right, shift_left};
    long want = (npix + 255) / 256;
    const int blocks = (int)(want < 8L * c->sm_count ? want : 8L * c->sm_count);
    dim3 grid(blocks, nframes);
    if (interleave) k_matrix_inverse<true><<<grid, 256, 0, st>>>(k, npix, d_in, in_stride_elems, d_out, out_stride_elems, d_invalid);
    else k_matrix_inverse<false><<<grid, 256, 0, st>>>(k, npix, d_in, in_stride_elems, d_out, out_stride_elems, d_invalid);
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

h2y_status launch_write_tiff_rows(h2y_ctx_impl *c, long npix, const uint16_t *g, const uint16_t *b, const uint16_t *r,
                                  uint16_t *rgb, int sr, cudaStream_t st)
{
    long want = (npix + 255) / 256;
    const int blocks = (int)(want < 8L * c->sm_count ? want : 8L * c->sm_count);
    k_write_tiff_rows<<<blocks, 256, 0, st>>>(npix, g, b, r, rgb, sr);
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

}   // namespace h2y
