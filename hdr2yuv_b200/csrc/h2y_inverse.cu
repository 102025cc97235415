// h2y_inverse.cu -- K3: one loop iteration of yuv2tiff's main() (yuv2tiff.cpp:278-552) as one
// fused kernel: read-clamp -> 4:2:0->4:4:4 FIR/box upsample (yuv2tiff.cpp:575-692) -> inverse
// colour difference -> negatives to 0 -> range clamp -> << SR -> interleaved R,G,B(,A) u16.
//
// Tiling: a CTA of 256 threads produces a 128 x 16 tile of output pixels.  The 4:2:0 chroma
// window it needs (64+5 columns x 8+6 rows, indices clamped = edge replicate) is staged in
// shared memory, the vertical 2-phase 6-tap result (the reference's u16 `dst422`
// intermediate, clamped and truncated exactly as there) is kept in shared memory too, and
// every thread then finishes 8 consecutive pixels of one row: 16-byte luma load, three
// 16-byte interleaved stores.  4:4:4 chroma never touches HBM.
// Algorithmic traffic: 3 B/px in (4:2:0 u16) + 6 B/px out (RGB16) = 9 B/px.
#include "h2y_internal.h"

namespace h2y {

namespace {
constexpr int TILE_W = 128, TILE_H = 16;
constexpr int CW = TILE_W / 2, CH = TILE_H / 2;     // chroma samples per tile
constexpr int SRC_W = CW + 5, SRC_H = CH + 6;       // window: cols i-2..i+3, rows j-3..j+3
constexpr int SRC_P = 72, MID_P = 72;               // padded row pitches
}   // namespace

__device__ __forceinline__ int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ float up_fin(float t, float hi)
{
    t = __fadd_rn(t, 0.5f);
    if (t > hi) t = hi;
    if (t < 0.0f) t = 0.0f;
    return (float)__float2int_rz(t);
}

// per-pixel inverse colour difference (yuv2tiff.cpp:399-430, 478-544); returns invalid count
__device__ __forceinline__ int inv_pixel(const InvK &k, int Y, float fcb, float fcr, unsigned &Ro, unsigned &Go,
                                         unsigned &Bo)
{
    int Yav = Y, Rp, Bp;
    const int Ysave = Y;
    const double top = (double)k.Full - 1.0;
    const double halfm = (double)k.Half - 0.5;
    if (k.matrix == H2Y_INV_YDzDx) {
        Rp = 2 * (int)fcr - (int)(k.Full - 1) + Yav;
        Bp = 2 * (int)fcb - (int)(k.Full - 1) + Yav;
    } else if (k.matrix == H2Y_INV_2020 || k.matrix == H2Y_INV_709) {
        float t = __double2float_rn(__dadd_rn(__dmul_rn(__dadd_rn((double)fcb, -halfm), k.kb), (double)Yav));
        if ((double)t > top) t = (float)top;
        Bp = f2i_x86(t);
        t = __double2float_rn(__dadd_rn(__dmul_rn(__dadd_rn((double)fcr, -halfm), k.kr), (double)Yav));
        if ((double)t > top) t = (float)top;
        Rp = f2i_x86(t);
        double g = __dadd_rn(__dadd_rn((double)(float)Yav, -__dmul_rn(k.wb, (double)(float)Bp)),
                             -__dmul_rn(k.wr, (double)(float)Rp));
        t = __double2float_rn(__dadd_rn(__ddiv_rn(g, k.wg), 0.5));
        if ((double)t > top) t = (float)top;
        Yav = f2i_x86(t);
    } else {   // Y100 / Y500 (yuv2tiff.cpp:424-430)
        float t = __double2float_rn(__dadd_rn(__dmul_rn(__dadd_rn((double)fcb, -halfm), (double)k.W),
                                              (double)__fmul_rn((float)Yav, k.V)));
        if ((double)t > top) t = (float)top;
        Bp = f2i_x86(t);
        t = __double2float_rn(__dadd_rn(__dmul_rn(__dadd_rn((double)fcr, -halfm), (double)k.U),
                                        (double)__fmul_rn((float)Yav, k.T)));
        if ((double)t > top) t = (float)top;
        Rp = f2i_x86(t);
    }
    int invalid = 0;
    if (Yav < 0) { Yav = 0; invalid++; }
    if (Rp < 0) { Rp = 0; if (Ysave != 0) invalid++; }
    if (Bp < 0) { Bp = 0; if (Ysave != 0) invalid++; }
    if (!k.full_range) {
        Rp = iclamp(Rp, (int)k.minVR, (int)k.maxVR);
        Yav = iclamp(Yav, (int)k.minVR, (int)k.maxVR);
        Bp = iclamp(Bp, (int)k.minVR, (int)k.maxVR);
    }
    Ro = ((unsigned)(Rp & 0xffff) << k.SR) & 0xffffu;
    Go = ((unsigned)(Yav & 0xffff) << k.SR) & 0xffffu;
    Bo = ((unsigned)(Bp & 0xffff) << k.SR) & 0xffffu;
    return invalid;
}

template <bool FIR, bool ALPHA>
__global__ void __launch_bounds__(256)
k_inverse_fused(InvK k, const uint16_t *__restrict__ yuv, size_t yuv_stride_elems, uint16_t *__restrict__ rgb,
                size_t rgb_stride_elems, uint32_t *invalid_out)
{
    __shared__ float s_src[2][SRC_H][SRC_P];
    __shared__ float s_mid[2][TILE_H][MID_P];
    const int w = k.w, h = k.h, wh = w >> 1, hh = h >> 1;
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    const int i0 = x0 >> 1, j0 = y0 >> 1;
    const uint16_t *fy = yuv + (size_t)blockIdx.z * yuv_stride_elems;
    const uint16_t *fcbp = fy + (size_t)w * h;
    const uint16_t *fcrp = fcbp + (size_t)wh * hh;
    const float hiCV = (float)k.maxCV;

    // 1. chroma window with the read clamp (yuv2tiff.cpp:297-320); indices clamped = replicate
    for (int t = threadIdx.x; t < 2 * SRC_H * SRC_W; t += 256) {
        const int pl = t / (SRC_H * SRC_W), r = (t / SRC_W) % SRC_H, cidx = t % SRC_W;
        const int jj = iclamp(j0 - 3 + r, 0, hh - 1), ii = iclamp(i0 - 2 + cidx, 0, wh - 1);
        unsigned v = (pl ? fcrp : fcbp)[(size_t)jj * wh + ii];
        if (!k.full_range) { v = v < k.minVRC ? k.minVRC : v; v = v > k.maxVRC ? k.maxVRC : v; }
        s_src[pl][r][cidx] = (float)v;
    }
    __syncthreads();

    // 2. vertical pass -> the reference's dst422 rows y0 .. y0+15 (yuv2tiff.cpp:615-650)
    if (FIR) {
        const float a3 = 3.0f / 256.0f, a16 = 16.0f / 256.0f, a67 = 67.0f / 256.0f, a227 = 227.0f / 256.0f,
                    a32 = 32.0f / 256.0f, a7 = 7.0f / 256.0f;
        for (int t = threadIdx.x; t < 2 * CH * SRC_W; t += 256) {
            const int pl = t / (CH * SRC_W), jl = (t / SRC_W) % CH, cidx = t % SRC_W;
            // window row of chroma row j is jl+3; rows beyond the picture replicate via hh clamp:
            // the window was filled with clamped indices, but clamping must follow the PICTURE
            // edge, which the fill already did (row r <-> picture row clamp(j0-3+r)).
            const float m3 = s_src[pl][jl][cidx], m2 = s_src[pl][jl + 1][cidx], m1 = s_src[pl][jl + 2][cidx],
                        c0 = s_src[pl][jl + 3][cidx], p1 = s_src[pl][jl + 4][cidx], p2 = s_src[pl][jl + 5][cidx],
                        p3 = s_src[pl][jl + 6][cidx];
            float e = __fmul_rn(a3, m3);
            e = __fsub_rn(e, __fmul_rn(a16, m2));
            e = __fadd_rn(e, __fmul_rn(a67, m1));
            e = __fadd_rn(e, __fmul_rn(a227, c0));
            e = __fsub_rn(e, __fmul_rn(a32, p1));
            e = __fadd_rn(e, __fmul_rn(a7, p2));
            float o = __fmul_rn(a3, p3);
            o = __fsub_rn(o, __fmul_rn(a16, p2));
            o = __fadd_rn(o, __fmul_rn(a67, p1));
            o = __fadd_rn(o, __fmul_rn(a227, c0));
            o = __fsub_rn(o, __fmul_rn(a32, m1));
            o = __fadd_rn(o, __fmul_rn(a7, m2));
            s_mid[pl][2 * jl][cidx] = up_fin(e, hiCV);
            s_mid[pl][2 * jl + 1][cidx] = up_fin(o, hiCV);
        }
        __syncthreads();
    }

    // 3. eight consecutive pixels of one row per thread
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int x = x0 + tx * 8, y = y0 + ty;
    int invalid = 0;
    if (x < w && y < h) {
        const uint4 yv = *reinterpret_cast<const uint4 *>(fy + (size_t)y * w + x);
        const unsigned yw[4] = {yv.x, yv.y, yv.z, yv.w};
        unsigned outw[16];   // up to 8 px * 4 samples, packed 2 per word
        constexpr int nch = ALPHA ? 4 : 3;
        unsigned short samples[32];
        const float b21 = 21.0f / 256.0f, b52 = 52.0f / 256.0f, b159 = 159.0f / 256.0f;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            unsigned Y = (yw[q >> 1] >> ((q & 1) * 16)) & 0xffffu;
            if (!k.full_range) { Y = Y < k.minVR ? k.minVR : Y; Y = Y > k.maxVR ? k.maxVR : Y; }   // 283-294
            const int ci = tx * 4 + (q >> 1) + 2;      // window column of chroma sample i = (x+q)/2
            float cb, cr;
            if (FIR) {
                if ((q & 1) == 0) {
                    cb = s_mid[0][ty][ci];
                    cr = s_mid[1][ty][ci];
                } else {                                  // yuv2tiff.cpp:678-683
                    float t = __fmul_rn(b21, __fadd_rn(s_mid[0][ty][ci - 2], s_mid[0][ty][ci + 3]));
                    t = __fsub_rn(t, __fmul_rn(b52, __fadd_rn(s_mid[0][ty][ci - 1], s_mid[0][ty][ci + 2])));
                    t = __fadd_rn(t, __fmul_rn(b159, __fadd_rn(s_mid[0][ty][ci], s_mid[0][ty][ci + 1])));
                    cb = up_fin(t, hiCV);
                    t = __fmul_rn(b21, __fadd_rn(s_mid[1][ty][ci - 2], s_mid[1][ty][ci + 3]));
                    t = __fsub_rn(t, __fmul_rn(b52, __fadd_rn(s_mid[1][ty][ci - 1], s_mid[1][ty][ci + 2])));
                    t = __fadd_rn(t, __fmul_rn(b159, __fadd_rn(s_mid[1][ty][ci], s_mid[1][ty][ci + 1])));
                    cr = up_fin(t, hiCV);
                }
            } else {                                      // box replicate, yuv2tiff.cpp:577-588
                cb = s_src[0][(ty >> 1) + 3][ci];
                cr = s_src[1][(ty >> 1) + 3][ci];
            }
            unsigned R, G, B;
            invalid += inv_pixel(k, (int)Y, cb, cr, R, G, B);
            samples[q * nch + 0] = (unsigned short)R;
            samples[q * nch + 1] = (unsigned short)G;
            samples[q * nch + 2] = (unsigned short)B;
            if (nch == 4) samples[q * nch + 3] = 65535;
        }
        const int nwords = 4 * nch;
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (i < nwords) outw[i] = (unsigned)samples[2 * i] | ((unsigned)samples[2 * i + 1] << 16);
        uint4 *o = reinterpret_cast<uint4 *>(rgb + (size_t)blockIdx.z * rgb_stride_elems + ((size_t)y * w + x) * nch);
        o[0] = make_uint4(outw[0], outw[1], outw[2], outw[3]);
        o[1] = make_uint4(outw[4], outw[5], outw[6], outw[7]);
        o[2] = make_uint4(outw[8], outw[9], outw[10], outw[11]);
        if (nch == 4) o[3] = make_uint4(outw[12], outw[13], outw[14], outw[15]);
    }
    if (invalid_out) {
        for (int o = 16; o > 0; o >>= 1) invalid += __shfl_xor_sync(0xffffffffu, invalid, o);
        if ((threadIdx.x & 31) == 0 && invalid) atomicAdd(&invalid_out[blockIdx.z], (uint32_t)invalid);
    }
}

h2y_status make_invk(const h2y_inverse_params &p, InvK *k)
{
    if (p.width < 8 || p.height < 2 || (p.width & 7) || (p.height & 1)) return H2Y_ERR_ARG;
    if (p.bit_depth != 10 && p.bit_depth != 12 && p.bit_depth != 14) return H2Y_ERR_ARG;
    if (p.matrix < H2Y_INV_YDzDx || p.matrix > H2Y_INV_Y500) return H2Y_ERR_ARG;
    k->w = p.width; k->h = p.height; k->bit_depth = p.bit_depth; k->matrix = p.matrix;
    k->fir = p.fir != 0; k->full_range = p.full_range != 0; k->alpha = p.alpha != 0;
    k->SR = 16 - p.bit_depth;                                   // yuv2tiff.cpp:89, 139, 150
    k->Half = 1u << (p.bit_depth - 1);
    k->Full = 1u << p.bit_depth;
    k->maxCV = k->Full - 1;
    const unsigned D = 1u << (p.bit_depth - 10);                // yuv2tiff.cpp:99, 144, 155
    k->minVR = 64 * D; k->maxVR = 876 * D + k->minVR;           // yuv2tiff.cpp:178-186
    k->minVRC = k->minVR; k->maxVRC = 896 * D + k->minVRC;
    const bool is2020 = p.matrix == H2Y_INV_2020;
    k->kb = is2020 ? 1.8814 : 1.8556; k->kr = is2020 ? 1.4746 : 1.5748;       // yuv2tiff.cpp:404-423
    k->wb = is2020 ? 0.0593 : 0.07222; k->wr = is2020 ? 0.2627 : 0.2126; k->wg = is2020 ? 0.6780 : 0.7152;
    k->T = k->U = k->V = k->W = 0.0f;
    if (p.matrix == H2Y_INV_Y100) { k->T = 0.98989899; k->U = 2.0; k->V = 1.016835017; k->W = 2.03367; }
    else if (p.matrix == H2Y_INV_Y500) { k->T = 0.99203764; k->U = 2.0; k->V = 1.013391241; k->W = 2.026782; }
    return H2Y_OK;
}

h2y_status launch_inverse(h2y_ctx_impl *c, const InvK &k, const void *d_yuv, size_t yuv_stride, void *d_rgb,
                          size_t rgb_stride, int nframes, uint32_t *d_invalid, cudaStream_t st)
{
    dim3 grid((k.w + TILE_W - 1) / TILE_W, (k.h + TILE_H - 1) / TILE_H, nframes);
#define INV(F, A) k_inverse_fused<F, A><<<grid, 256, 0, st>>>(k, (const uint16_t *)d_yuv, yuv_stride / 2, \
                                                              (uint16_t *)d_rgb, rgb_stride / 2, d_invalid)
    if (k.fir && k.alpha) INV(true, true);
    else if (k.fir) INV(true, false);
    else if (k.alpha) INV(false, true);
    else INV(false, false);
#undef INV
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

}   // namespace h2y
