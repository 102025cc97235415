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
#include <cstdlib>

#include "h2y_f32x2.cuh"
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
// MKIND: -1 = decide at run time from k.matrix; H2Y_INV_* = compiled for that family only (smaller code)
// Ybar > 0: the -X mode, Y'DzDx rebuilt around the 2x2 luma mean and rescaled by Y'/Ybar (yuv2tiff.cpp:390-398)
template <int MKIND = -1>
__device__ __forceinline__ int inv_pixel(const InvK &k, int Y, float fcb, float fcr, unsigned &Ro, unsigned &Go,
                                         unsigned &Bo, int Ybar = 0)
{
    const int matrix = MKIND < 0 ? k.matrix : MKIND;
    int Yav = Y, Rp, Bp;
    const int Ysave = Y;
    const double top = (double)k.Full - 1.0;
    const double halfm = (double)k.Half - 0.5;
    if (matrix == H2Y_INV_YDzDx && Ybar > 0) {
        // all operands are small integers in double: the sums and the product are exact, the division rounds once,
        // then the reference stores to float and truncates
        const double fm1 = (double)(float)((double)k.Full - 1.0), yb = (double)(float)Ybar, ya = (double)(float)Yav;
        float RED = __double2float_rn(__ddiv_rn(__dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(2.0, (double)fcr), -fm1), yb), ya), yb));
        float BLUE = __double2float_rn(__ddiv_rn(__dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(2.0, (double)fcb), -fm1), yb), ya), yb));
        if ((double)RED > (double)k.Full - 1.5) RED = (float)((double)k.Full - 1.0);
        if ((double)BLUE > (double)k.Full - 1.5) BLUE = (float)((double)k.Full - 1.0);
        Rp = f2i_x86(RED);
        Bp = f2i_x86(BLUE);
    } else if (matrix == H2Y_INV_YDzDx) {
        Rp = 2 * (int)fcr - (int)(k.Full - 1) + Yav;
        Bp = 2 * (int)fcb - (int)(k.Full - 1) + Yav;
    } else if (matrix == H2Y_INV_2020 || matrix == H2Y_INV_709) {
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

// The same pixel for `B10 2020` in integer arithmetic (InvK::int10; used by the rows kernel for the pixels its fp32
// guard bands hand back).  With C a chroma code and Y', B', R' integers,
//     (C - 511.5) * 1.8814 + Y'                        = ((2 C - 1023) * 9407 + 10000 Y') / 10000
//     (Y' - 0.0593 B' - 0.2627 R') / 0.6780 + 0.5      = (2 (10000 Y' - 593 B' - 2627 R') + 6780) / 13560
// exactly (and 1.4746 / 2 = 0.7373 for R').  The reference evaluates the left sides in double (a handful of roundings,
// 1e-12 absolute at these magnitudes), rounds to float (at most 2^-15 = 3.05e-5 below 1024, 6.1e-5 below 2048) and
// truncates (yuv2tiff.cpp:404-413).  A quotient that is not an integer is at least 1/13560 = 7.4e-5 away from the
// next integer, further than those roundings can move it, and an integer quotient is rounded back onto itself: the
// truncated float is the integer quotient rounded toward zero, which is what C's division gives.  Values above
// Full-1 are set to Full-1 before the truncation there, after it here.  Every 10-bit triplet is compared with
// inv_pixel() by tests/test_inverse_gpu.py (H2Y_INVERSE_KERNEL=exact sends every pixel here).
__device__ __forceinline__ int inv_pixel_int10(const InvK &k, int Y, int cb, int cr, unsigned &Ro, unsigned &Go, unsigned &Bo)
{
    const int y4 = Y * 10000;
    int Bp = min(((2 * cb - 1023) * 9407 + y4) / 10000, 1023);
    int Rp = min(((2 * cr - 1023) * 7373 + y4) / 10000, 1023);
    int Gp = min((2 * (y4 - 593 * Bp - 2627 * Rp) + 6780) / 13560, 1023);
    int invalid = 0;
    if (Gp < 0) { Gp = 0; invalid++; }
    if (Rp < 0) { Rp = 0; if (Y != 0) invalid++; }
    if (Bp < 0) { Bp = 0; if (Y != 0) invalid++; }
    if (!k.full_range) {
        Rp = iclamp(Rp, (int)k.minVR, (int)k.maxVR);
        Gp = iclamp(Gp, (int)k.minVR, (int)k.maxVR);
        Bp = iclamp(Bp, (int)k.minVR, (int)k.maxVR);
    }
    Ro = (unsigned)Rp << 6; Go = (unsigned)Gp << 6; Bo = (unsigned)Bp << 6;
    return invalid;
}

// The integer form at any depth (10 / 12 / 14 bits) and for both Y'CbCr families.  The quotients are the same exact
// rationals (BT.709: 1.8556 / 2 = 0.9278, 1.5748 / 2 = 0.7874, and with the reference's literals 0.07222, 0.2126, 0.7152
// G' = (50000 Y' - 3611 B' - 10630 R' + 17880) / 35760), but above 10 bits, and for BT.709's denominator already from
// 256 up, the float rounding can reach the next integer: the reference's value is RN32(x + d), x = N / D the exact
// quotient, |d| < 2e-11 from its double arithmetic.  With q = floor(|x|), r = |N| mod D and k = q + 1, the floats just
// below k are 2^(e-23) apart, e = floor(log2(2k - 1)) - 1, so RN32 lands on k exactly when (D - r) / D < 2^(e-24), i.e.
// when D - r <= D >> (24 - e): D 2^(e-24) is never an integer here, the nearest quotient to that boundary stays 0.09 / D
// > 2.5e-6 away for every e up to 13 and each of the three denominators (10000, 13560, 35760), far more than d, and an
// exact tie would need x to be a multiple of 2^(e-24) with e <= 13, which a quotient by 625, 1695 or 2235 times a power of
// two cannot be unless it is an integer.  Negative quotients mirror this (RN32 and the truncation are odd functions).
// Checked against inv_pixel(): every 10-bit BT.709 triplet and 2^27 random 12- and 14-bit pixels per family, each sent
// here through H2Y_INVERSE_KERNEL=exact (tests/test_inverse_gpu.py).
__device__ __forceinline__ int trunc_rn32_quot(int N, unsigned q, unsigned r, unsigned D)
{
    const int e = 30 - __clz((int)(2u * q + 1u));                // binade below k = q + 1: 2k - 1 = 2q + 1
    q += (D - r) <= (D >> (24 - e)) ? 1u : 0u;                   // r = 0: D <= D >> n never holds (n >= 11)
    return N < 0 ? -(int)q : (int)q;
}
__device__ __forceinline__ int inv_pixel_int(const InvK &k, int Y, int cb, int cr, unsigned &Ro, unsigned &Go, unsigned &Bo)
{
    const int top = (int)k.Full - 1, y4 = Y * 10000;
    const int nb = (2 * cb - top) * k.ikb + y4, nr = (2 * cr - top) * k.ikr + y4;
    const unsigned ab = (unsigned)abs(nb), ar = (unsigned)abs(nr);
    int Bp = min(trunc_rn32_quot(nb, ab / 10000u, ab % 10000u, 10000u), top);
    int Rp = min(trunc_rn32_quot(nr, ar / 10000u, ar % 10000u, 10000u), top);
    const int ng = Y * k.igy + Bp * k.igb + Rp * k.igr + k.igc;
    const unsigned ag = (unsigned)abs(ng), D = (unsigned)k.igd;
    unsigned qg = __umulhi(ag, k.igm), rg = ag - qg * D;         // floor(2^32 / D) as multiplier: short by at most one
    if (rg >= D) { qg++; rg -= D; }
    int Gp = min(trunc_rn32_quot(ng, qg, rg, D), top);
    int invalid = 0;
    if (Gp < 0) { Gp = 0; invalid++; }
    if (Rp < 0) { Rp = 0; if (Y != 0) invalid++; }
    if (Bp < 0) { Bp = 0; if (Y != 0) invalid++; }
    if (!k.full_range) {
        Rp = iclamp(Rp, (int)k.minVR, (int)k.maxVR);
        Gp = iclamp(Gp, (int)k.minVR, (int)k.maxVR);
        Bp = iclamp(Bp, (int)k.minVR, (int)k.maxVR);
    }
    Ro = (unsigned)Rp << k.SR; Go = (unsigned)Gp << k.SR; Bo = (unsigned)Bp << k.SR;
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
        unsigned yo[4] = {0, 0, 0, 0};             // the other row of the 2x2 luma blocks (-X only)
        if (k.ybar) {
            const uint4 t = *reinterpret_cast<const uint4 *>(fy + (size_t)(y ^ 1) * w + x);
            yo[0] = t.x; yo[1] = t.y; yo[2] = t.z; yo[3] = t.w;
        }
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
            int Ybar = 0;
            if (k.ybar) {                                 // yuv2tiff.cpp:365-387: mean of the clamped 2x2 block, in [1, Full-1]
                unsigned b4[4] = {yw[q >> 1] & 0xffffu, yw[q >> 1] >> 16, yo[q >> 1] & 0xffffu, yo[q >> 1] >> 16};
                int sum = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    unsigned v = b4[i];
                    if (!k.full_range) { v = v < k.minVR ? k.minVR : v; v = v > k.maxVR ? k.maxVR : v; }
                    sum += (int)v;
                }
                Ybar = min(max(sum / 4, 1), (int)k.Full - 1);
            }
            unsigned R, G, B;
            invalid += inv_pixel(k, (int)Y, cb, cr, R, G, B, Ybar);
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

// =================================================================================================
// Rows kernel (everything but tiny inputs): warp-autonomous, same idea as k_forward_exr420_rows (h2y_forward2.cu).
// A warp owns a 240-pixel column strip (lane = 8 luma pixels = 4 chroma columns, lanes 0/31 are halo lanes of
// the horizontal interpolation) and a long run of chroma rows.  The 7-row vertical window lives in a private
// 8-slot shared-memory ring that only its own lanes touch (no barrier of any kind); the luma rows arrive by
// cp.async in per-lane stages one trip ahead; both upsampling stages are FFMA2 chains on {Cb,Cr} pairs (all
// terms are integers/256 below 2^23, so the order is free) with the reference's clamp + truncation taken as a
// round-down add and an integer clamp.  The Y'CbCr inverse is evaluated in fp32 with guard bands: each
// truncation is taken at x-G and x+G, and an exact routine (the integer forms inv_pixel_int10 / inv_pixel_int
// of the reference's double arithmetic; inv_pixel() itself for Y100 / Y500) runs only for the rare pixel
// where the two disagree or a component is negative (which also carries the invalidPixels bookkeeping).
//
// Guard bands, in units of u = 2^(d-25) (half an fp32 ulp of values in [2^(d-1), 2^d)).  x = the exact real value of a
// chain, t_ref = the reference's float (double arithmetic rounded once to float: |t_ref - x| <= 1u below 2^d; above
// Full-1 both sides clamp).  The kernel takes floor(lo) and floor(RN(lo + 2G)) with lo ~ x - G and accepts the pixel
// when they agree, i.e. when no integer lies in (lo, RN(lo + 2G)], so it needs lo < t_ref <= RN(lo + 2G).
//  * B', R': lo = fma(C - (Half-0.5), k32, Y - G).  C - (Half-0.5) and Y - G are exact; |k32 - k| <= 2^-24 (k in [1,2))
//    times 2^(d-1) is 1u, the fma's rounding 1u: |lo - (x - G)| <= 2u.  t_ref > lo needs G > 3u, t_ref <= RN(lo + 2G)
//    (another 1u of rounding) needs G >= 4u.  G = 6u (6u keeps Y - G exact: 24 bits).
//  * G': g2 = fma(-wr32, R', fma(-wb32, B', Y)) with integer B', R' (the accepted values): weight images u/8 + u/2, two
//    roundings 2u: 2.625u; times 1/wg <= 1.475 is 3.87u; the image of 1/wg 2u; the fma's rounding 1u:
//    |lo - (x - G)| <= 6.87u, so G >= 8.87u.  G = 10u.
// (One band of 2^(d-21) = 16u for all three chains sent 2.2 times as many pixels to the exact routine.)
struct Inv2Args {
    InvK k;
    const uint8_t *yuv;
    size_t yuv_stride;
    uint8_t *rgb;
    size_t rgb_stride;
    uint32_t *invalid;
    int nstrips, sub, wps;
    long total_crows;        // nframes * (h / 2)
    float hm, kb, kr, nwb, nwr, rwg;          // Half-0.5, chroma gains, -wb, -wr, 1/wg as fp32
    float guard_c, guard_g;                   // guard bands of the B' / R' chains and of the G' chain (see k_inverse_rows)
};

namespace {
#ifndef H2Y_INV_THREADS
#define H2Y_INV_THREADS 512
#endif
#ifndef H2Y_INV_MINB
#define H2Y_INV_MINB 1
#endif
#ifndef H2Y_INV_GUARD_C
#define H2Y_INV_GUARD_C 6.0f         // guard bands in units of u = 2^(d-25); 16 / 16 was the single band of rounds 1-2
#define H2Y_INV_GUARD_G 10.0f
#endif
#ifndef H2Y_INV_UNROLL_HALF
#define H2Y_INV_UNROLL_HALF 0
#endif
#ifndef H2Y_INV_LUMA_STAGES
#define H2Y_INV_LUMA_STAGES 2        // cp.async stages of the luma rows per warp: 2 = one trip ahead, 4 = three
#endif
constexpr int RTHREADS = H2Y_INV_THREADS, RWARPS = RTHREADS / 32, RSLOTS = 8, RMINB = H2Y_INV_MINB;
constexpr int LSTAGES = H2Y_INV_LUMA_STAGES;          // luma trips in flight + 1 (power of two)
constexpr float TWO23 = 8388608.0f;
constexpr size_t INV_ROWS_SMEM = 8192 + (size_t)RWARPS * 8192 + (size_t)RWARPS * LSTAGES * 1024;
}

// shared-memory accesses by 32-bit shared-window address (the ring and the luma stages are private to a lane)
__device__ __forceinline__ float4 lds128(unsigned a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128u(unsigned a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(unsigned a, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(a), "f"(x), "f"(y), "f"(z), "f"(w));
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N)); }

// Per-launch constants: CFG 0 from the launch arguments, CFG 10 = `B10 2020`, video range (yuv2tiff.cpp:137-147,
// 178-186, 404-413) with every constant an immediate (no constant-bank reloads in the pixel loop).
template <int CFG> struct IC {
    static constexpr float GC = H2Y_INV_GUARD_C / 32768.0f, GG = H2Y_INV_GUARD_G / 32768.0f;     // 6u and 10u, u = 2^(10-25)
#define ICF(name, rt, ct) __device__ __forceinline__ static float name(const Inv2Args &A) { return CFG ? (ct) : (rt); }
#define ICU(name, rt, ct) __device__ __forceinline__ static unsigned name(const Inv2Args &A) { return CFG ? (unsigned)(ct) : (unsigned)(rt); }
    ICF(hm, A.hm, 511.5f) ICF(kb, A.kb, (float)1.8814) ICF(kr, A.kr, (float)1.4746)
    ICF(nwb, A.nwb, -(float)0.0593) ICF(nwr, A.nwr, -(float)0.2627) ICF(rwg, A.rwg, (float)(1.0 / 0.6780)) ICF(guard_c, A.guard_c, GC) ICF(guard_g, A.guard_g, GG)
    ICU(full_range, A.k.full_range, 0) ICU(minVR, A.k.minVR, 64) ICU(maxVR, A.k.maxVR, 940) ICU(minVRC, A.k.minVRC, 64)
    ICU(maxVRC, A.k.maxVRC, 960) ICU(maxCV, A.k.maxCV, 1023) ICU(Full, A.k.Full, 1024) ICU(SR, A.k.SR, 6)
#undef ICF
#undef ICU
};

// mode: 0 = every pixel through inv_pixel (Y100 / Y500), 1 = guarded fp32 (709 / 2020), 2 = integer Y'DzDx
// ALLSLOW: every pixel takes the exact routine (tests: H2Y_INVERSE_KERNEL=exact)
template <int MODE, bool FIR, bool ALPHA, int CFG = 0, bool ALLSLOW = false>
__global__ void __launch_bounds__(RTHREADS, RMINB) k_inverse_rows(const Inv2Args A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const InvK &k = A.k;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Shared memory, all of it private to a lane (no barrier anywhere):
    //  * the chroma ring: per warp 8 slots x {first, second pair of columns} x 32 lanes x float4 = 8 KB, each warp's ring
    //    aligned to 8 KB so that the slot number is address bits 10-12 and a slot's address is one LOP3 away
    //    (ring_lane | ((row << 10) & 0x1c00));
    //  * the luma stages: rows 2c and 2c+1 of the next trips arrive there with cp.async (LDGSTS: no registers, and a
    //    whole trip or two of distance; as register loads they were ~300 instructions ahead of their use and the warps
    //    spent 17 % of their time waiting for them, ncu of round 2's kernel).
    const unsigned smem_sa = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned ring_lane = ((smem_sa + 8191u) & ~8191u) + (unsigned)warp * 8192u + (unsigned)lane * 16u;
    const unsigned luma_lane = ((smem_sa + 8191u) & ~8191u) + RWARPS * 8192u + (unsigned)warp * (LSTAGES * 1024u) + (unsigned)lane * 16u;
    const int w = k.w, h = k.h, wh = w >> 1, hh = h >> 1;
    constexpr int nch = ALPHA ? 4 : 3;
    typedef IC<CFG> C;
    const int hi_bits = MAGIC_BITS + (int)C::maxCV(A);
    const int wk = warp / A.wps, sfirst = warp - wk * A.wps;
    if (wk >= A.sub) return;
    const long K = (long)gridDim.x * A.sub, kid = (long)blockIdx.x * A.sub + wk;
    const long g0 = (kid * A.total_crows) / K, g1 = ((kid + 1) * A.total_crows) / K;
    if (g1 <= g0) return;
    const float guardC = C::guard_c(A), guardG = C::guard_g(A);
    int invalid_frame = -1;
    unsigned invalid = 0;

    for (int frame = (int)(g0 / hh); frame <= (int)((g1 - 1) / hh); frame++) {
        const long fbase = (long)frame * hh;
        const int cs = (int)(max(g0, fbase) - fbase), ce = (int)(min(g1, fbase + hh) - fbase);
        const uint16_t *fy = reinterpret_cast<const uint16_t *>(A.yuv + (size_t)frame * A.yuv_stride);
        const uint16_t *fcb = fy + (size_t)w * h;
        const unsigned crd = (unsigned)wh * (unsigned)hh * 2u;                  // Cr plane behind Cb, in bytes (far below 2^31)
        uint16_t *frgb = reinterpret_cast<uint16_t *>(A.rgb + (size_t)frame * A.rgb_stride);
        if (A.invalid && invalid_frame != frame) {
            if (invalid && invalid_frame >= 0) atomicAdd(&A.invalid[invalid_frame], invalid);
            invalid = 0; invalid_frame = frame;
        }
        for (int strip = sfirst; strip < A.nstrips; strip += A.wps) {
            const int x0 = strip * 240, xl = x0 + 8 * (lane - 1);
            const bool lane_in_pic = xl >= 0 && xl < w;
            const bool lane_interior = lane >= 1 && lane < 31 && xl < min(x0 + 240, w);
            const int cl = (lane_in_pic ? xl : (xl < 0 ? 0 : w - 8)) >> 1;       // first of this lane's 4 chroma columns

            // Columns clamp to [0, w/2-1] (yuv2tiff.cpp:655-683): the halo lane left of the picture holds column 0 four
            // times, the lanes right of it the last column, so the neighbours' shuffles deliver the replicated samples.
            const unsigned sel0 = xl < 0 ? 0x1010u : (xl >= w ? 0x7676u : 0x3210u), sel1 = xl < 0 ? 0x1010u : (xl >= w ? 0x7676u : 0x7654u);
            // chroma row -> two float4 {cb0,cr0,cb1,cr1},{cb2,cr2,cb3,cr3} in ring slot (slot10 >> 10) & 7
            auto stash = [&](unsigned slot10, uint2 vb, uint2 vr) {
                unsigned wb[2] = {__byte_perm(vb.x, vb.y, sel0), __byte_perm(vb.x, vb.y, sel1)};
                unsigned wr[2] = {__byte_perm(vr.x, vr.y, sel0), __byte_perm(vr.x, vr.y, sel1)};
                if (!C::full_range(A)) {                                          // read clamp, yuv2tiff.cpp:297-320
                    const unsigned lo2 = C::minVRC(A) * 0x10001u, hi2 = C::maxVRC(A) * 0x10001u;
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        asm("max.u16x2 %0, %0, %1;" : "+r"(wb[i]) : "r"(lo2)); asm("min.u16x2 %0, %0, %1;" : "+r"(wb[i]) : "r"(hi2));
                        asm("max.u16x2 %0, %0, %1;" : "+r"(wr[i]) : "r"(lo2)); asm("min.u16x2 %0, %0, %1;" : "+r"(wr[i]) : "r"(hi2));
                    }
                }
                const u64 m2 = pk(-TWO23, -TWO23);
                u64 p[4];
#pragma unroll
                for (int i = 0; i < 2; i++) {        // 0x4B000000 | v is the float 2^23 + v
                    p[2 * i] = fadd2(pk(__uint_as_float(__byte_perm(wb[i], 0x4B000000u, 0x7610)), __uint_as_float(__byte_perm(wr[i], 0x4B000000u, 0x7610))), m2);
                    p[2 * i + 1] = fadd2(pk(__uint_as_float(__byte_perm(wb[i], 0x4B000000u, 0x7632)), __uint_as_float(__byte_perm(wr[i], 0x4B000000u, 0x7632))), m2);
                }
                const unsigned sa = ring_lane | (slot10 & 0x1c00u);
                sts128(sa, plo(p[0]), phi(p[0]), plo(p[1]), phi(p[1]));
                sts128(sa + 512u, plo(p[2]), phi(p[2]), plo(p[3]), phi(p[3]));
            };
            // Running pointers.  The ring always holds rows clamp(c-3) .. clamp(c+3) in slots (c-3)&7 .. (c+3)&7: at the
            // picture's top the first row is stashed for rows -3..-1 as well, at its bottom the last fetched row (hh-1)
            // is stashed again, so the window read needs no row clamps (yuv2tiff.cpp:617-622 replicates the edge rows).
            const uint16_t *pcb = fcb + (size_t)min(max(cs - 3, 0), hh - 1) * wh + cl;       // next chroma row to fetch
            int rnext = cs - 3;                                                              // its (unclamped) row number
            uint2 nb, nr;
            auto fetch_next = [&]() {
                nb = __ldg(reinterpret_cast<const uint2 *>(pcb));
                nr = __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(pcb) + crd));
                pcb += ((unsigned)rnext < (unsigned)(hh - 1)) ? wh : 0;         // rows above the picture and the last row: stay
                rnext++;
            };
            for (int r = cs - 3; r <= cs + 3; r++) {
                fetch_next();
                stash((unsigned)(r + 8) << 10, nb, nr);
            }
            fetch_next();                                                                    // row clamp(cs + 4)
            const uint16_t *pyl = fy + (size_t)(2 * cs) * w + (cl << 1);                     // luma rows of the next trip to request
            uint16_t *po = frgb + ((size_t)(2 * cs) * w + xl) * nch;
            auto luma_request = [&](int cc) {
                const unsigned sa = luma_lane + (unsigned)(cc & (LSTAGES - 1)) * 1024u;
                cp_async16(sa, pyl);
                cp_async16(sa + 512u, pyl + w);
                pyl += 2 * w;
            };
#pragma unroll
            for (int i = 0; i < LSTAGES - 1; i++) {
                if (cs + i < ce) luma_request(cs + i);
                cp_async_commit();
            }
            unsigned c10 = (unsigned)(cs + 5) << 10;                                         // slot of row c-3, times 1024

            for (int c = cs; c < ce; c++, c10 += 1024u) {
                if (c + LSTAGES - 1 < ce) luma_request(c + LSTAGES - 1);
                cp_async_commit();
                __syncwarp();
                // ---- vertical 2-phase 6-tap (yuv2tiff.cpp:615-650) on the 7-row window ----
                u64 win[7][4];
#pragma unroll
                for (int t = 0; t < 7; t++) {
                    const unsigned sa = ring_lane | ((c10 + 1024u * t) & 0x1c00u);
                    const float4 a = lds128(sa), b = lds128(sa + 512u);
                    win[t][0] = pk(a.x, a.y); win[t][1] = pk(a.z, a.w); win[t][2] = pk(b.x, b.y); win[t][3] = pk(b.z, b.w);
                }
                u64 dv[2][4];          // the reference's dst422 rows 2c and 2c+1, this lane's 4 columns
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (FIR) {
                        const u64 a3 = pk(3.0f / 256.0f, 3.0f / 256.0f), a16n = pk(-16.0f / 256.0f, -16.0f / 256.0f),
                                  a67 = pk(67.0f / 256.0f, 67.0f / 256.0f), a227 = pk(227.0f / 256.0f, 227.0f / 256.0f),
                                  a32n = pk(-32.0f / 256.0f, -32.0f / 256.0f), a7 = pk(7.0f / 256.0f, 7.0f / 256.0f), hf = pk(0.5f, 0.5f);
                        u64 e = ffma2(a3, win[0][i], hf);
                        e = ffma2(a16n, win[1][i], e); e = ffma2(a67, win[2][i], e); e = ffma2(a227, win[3][i], e);
                        e = ffma2(a32n, win[4][i], e); e = ffma2(a7, win[5][i], e);
                        u64 o = ffma2(a3, win[6][i], hf);
                        o = ffma2(a16n, win[5][i], o); o = ffma2(a67, win[4][i], o); o = ffma2(a227, win[3][i], o);
                        o = ffma2(a32n, win[2][i], o); o = ffma2(a7, win[1][i], o);
                        int e0, e1, o0, o1;
                        unpk(fadd2_rm(e, pk(MAGIC, MAGIC)), e0, e1);
                        unpk(fadd2_rm(o, pk(MAGIC, MAGIC)), o0, o1);
                        dv[0][i] = fadd2(pk(__int_as_float(clamp3(e0, MAGIC_BITS, hi_bits)), __int_as_float(clamp3(e1, MAGIC_BITS, hi_bits))), pk(-MAGIC, -MAGIC));
                        dv[1][i] = fadd2(pk(__int_as_float(clamp3(o0, MAGIC_BITS, hi_bits)), __int_as_float(clamp3(o1, MAGIC_BITS, hi_bits))), pk(-MAGIC, -MAGIC));
                    } else {
                        dv[0][i] = dv[1][i] = win[3][i];        // box: replicate (yuv2tiff.cpp:577-588)
                    }
                }
                // the next row enters the ring while the pixels are finished; then prefetch the one after
                stash(c10 + 7u * 1024u, nb, nr);                         // row clamp(c + 4)
                if (c + 1 < ce) fetch_next();                            // row clamp(c + 5)
                cp_async_wait<LSTAGES - 1>();                            // this trip's luma rows have landed

#if H2Y_INV_UNROLL_HALF
#pragma unroll
#else
#pragma unroll 1
#endif
                for (int half = 0; half < 2; half++) {                  // not unrolled: the instruction cache is the limit
                    // both passes read dv[0]; the second row moves down at the end of the first pass
                    // (12 moves per chroma row instead of 12 selects per luma row, and no run-time array index)
                    u64 d[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) d[i] = dv[0][i];
                    // ---- horizontal: even x copies, odd x 6-tap over columns i-2 .. i+3 (yuv2tiff.cpp:655-683) ----
                    u64 cpx[8];
#pragma unroll
                    for (int i = 0; i < 4; i++) cpx[2 * i] = d[i];
                    if (FIR) {
                        u64 L2 = pk(__shfl_up_sync(0xffffffffu, plo(d[2]), 1), __shfl_up_sync(0xffffffffu, phi(d[2]), 1));
                        u64 L3 = pk(__shfl_up_sync(0xffffffffu, plo(d[3]), 1), __shfl_up_sync(0xffffffffu, phi(d[3]), 1));
                        u64 R0 = pk(__shfl_down_sync(0xffffffffu, plo(d[0]), 1), __shfl_down_sync(0xffffffffu, phi(d[0]), 1));
                        u64 R1 = pk(__shfl_down_sync(0xffffffffu, plo(d[1]), 1), __shfl_down_sync(0xffffffffu, phi(d[1]), 1));
                        u64 R2 = pk(__shfl_down_sync(0xffffffffu, plo(d[2]), 1), __shfl_down_sync(0xffffffffu, phi(d[2]), 1));
                        // (the picture's edge columns are replicated by the halo lanes, see sel0 / sel1)
                        const u64 nb8[9] = {L2, L3, d[0], d[1], d[2], d[3], R0, R1, R2};
                        const u64 b21 = pk(21.0f / 256.0f, 21.0f / 256.0f), b52n = pk(-52.0f / 256.0f, -52.0f / 256.0f),
                                  b159 = pk(159.0f / 256.0f, 159.0f / 256.0f), hf = pk(0.5f, 0.5f);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            u64 t = ffma2(b21, nb8[i], hf);
                            t = ffma2(b21, nb8[i + 5], t); t = ffma2(b52n, nb8[i + 1], t); t = ffma2(b52n, nb8[i + 4], t);
                            t = ffma2(b159, nb8[i + 2], t); t = ffma2(b159, nb8[i + 3], t);
                            int t0, t1;
                            unpk(fadd2_rm(t, pk(MAGIC, MAGIC)), t0, t1);
                            cpx[2 * i + 1] = fadd2(pk(__int_as_float(clamp3(t0, MAGIC_BITS, hi_bits)), __int_as_float(clamp3(t1, MAGIC_BITS, hi_bits))), pk(-MAGIC, -MAGIC));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; i++) cpx[2 * i + 1] = d[i];
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++) dv[0][i] = dv[1][i];
                    uint16_t *const orow = po;
                    po += (size_t)w * nch;
                    if (!lane_interior) continue;
                    const uint4 yr = lds128u(luma_lane + (unsigned)(c & (LSTAGES - 1)) * 1024u + (unsigned)half * 512u);
                    // ---- eight pixels of luma row 2c + half ----
                    // Phase 1, branch-free: the colour difference inverse as integers Rp/Gp/Bp (before the output clamp
                    // and shift) plus one bit per pixel that needs the reference-exact routine.
                    unsigned yw[4] = {yr.x, yr.y, yr.z, yr.w};
                    if (!C::full_range(A)) {                                                  // yuv2tiff.cpp:283-294
                        const unsigned lo2 = C::minVR(A) * 0x10001u, hi2 = C::maxVR(A) * 0x10001u;
#pragma unroll
                        for (int i = 0; i < 4; i++) yw[i] = clamp_u16x2(yw[i], lo2, hi2);
                    }
                    int Rv[8], Gv8[8], Bv[8];
                    unsigned slow_mask = MODE == 0 || ALLSLOW ? 0xffu : 0u;
#pragma unroll
                    for (int q = 0; q < 8; q += 2) {                     // two pixels at a time: one luma word
                        if (MODE == 1) {
                            // B', R' on the {Cb,Cr} pair of each pixel, G' on the pixel pair: all packed.
                            // One PRMT per pixel builds the float 2^23 + Y from the luma word.
                            const u64 Yf2 = fadd2(pk(__uint_as_float(__byte_perm(yw[q >> 1], 0x4B000000u, 0x7610)),
                                                     __uint_as_float(__byte_perm(yw[q >> 1], 0x4B000000u, 0x7632))), pk(-TWO23, -TWO23));
                            const u64 Yg2 = fadd2(Yf2, pk(-guardC, -guardC));
                            int B1[2], R1[2];
                            u64 dBR[2];                      // floor(upper end) - floor(lower end) per chain: 0.0f or 1.0f
#pragma unroll
                            for (int e = 0; e < 2; e++) {
                                const float Yg = e ? phi(Yg2) : plo(Yg2);
                                const u64 tlo = ffma2(fadd2(cpx[q + e], pk(-C::hm(A), -C::hm(A))), pk(C::kb(A), C::kr(A)), pk(Yg, Yg));
                                const u64 t1 = fadd2_rm(tlo, pk(MAGIC, MAGIC));
                                dBR[e] = fsub2(fadd2_rm(fadd2(tlo, pk(2.0f * guardC, 2.0f * guardC)), pk(MAGIC, MAGIC)), t1);
                                unpk(t1, B1[e], R1[e]);
                            }
                            // t > Full-1 -> Full-1 before B' and R' enter G' (yuv2tiff.cpp:406-412).  The instantiation with
                            // compiled-in constants takes the minimum here (its video-range content has such pixels, and the
                            // output clamp makes them ordinary); the others leave them to the exact routine like negative
                            // ones, which drops the minima and the sign test for one bit test (measured both ways for both:
                            // B10 2020 1.105 / 1.132 ms, B12 2020 1.312 / 1.256 ms)
                            constexpr bool TOPMIN = CFG != 0;
                            int Bc[2] = {B1[0], B1[1]}, Rc[2] = {R1[0], R1[1]};
                            if (TOPMIN) {
#pragma unroll
                                for (int e = 0; e < 2; e++) { Bc[e] = min(B1[e], hi_bits); Rc[e] = min(R1[e], hi_bits); }
                            }
                            const u64 bf = fadd2(pk(__int_as_float(Bc[0]), __int_as_float(Bc[1])), pk(-MAGIC, -MAGIC));
                            const u64 rf = fadd2(pk(__int_as_float(Rc[0]), __int_as_float(Rc[1])), pk(-MAGIC, -MAGIC));
                            const u64 g2 = ffma2(pk(C::nwr(A), C::nwr(A)), rf, ffma2(pk(C::nwb(A), C::nwb(A)), bf, Yf2));
                            const u64 glo = ffma2(g2, pk(C::rwg(A), C::rwg(A)), pk(0.5f - guardG, 0.5f - guardG));
                            int G1[2];
                            const u64 g1 = fadd2_rm(glo, pk(MAGIC, MAGIC));
                            const u64 dG = fsub2(fadd2_rm(fadd2(glo, pk(2.0f * guardG, 2.0f * guardG)), pk(MAGIC, MAGIC)), g1);
                            unpk(g1, G1[0], G1[1]);
#pragma unroll
                            for (int e = 0; e < 2; e++) {
                                // guarded (a floor changed inside the band: the exact difference of the two floors is 1.0f,
                                // taken on the packed-float pipe), negative, or above Full-1.  Every component n lies in
                                // (-Full, 2 Full): |C - Half + 0.5| k < 0.95 Full and Y' < Full for B' and R', and
                                // -0.33 Full / wg < G' < Full / wg.  So bit d = log2(Full) of MAGIC_BITS + n is clear exactly
                                // for 0 <= n < Full (a negative n leaves bits d..21 set): one OR and one AND with Full replace
                                // the sign test and the three `t > Full-1 -> Full-1` minima (!TOPMIN).  The mask is built
                                // without predicates, pixel q in bit 7 - q
                                unsigned wq = (unsigned)ilo(dBR[e]) | (unsigned)ihi(dBR[e]) | (unsigned)(e ? ihi(dG) : ilo(dG));
                                if (TOPMIN) wq |= ~(unsigned)(B1[e] & R1[e] & G1[e]) & 0x00400000u;      // negative: bit 22 clear
                                else wq |= (unsigned)(B1[e] | R1[e] | G1[e]) & C::Full(A);
                                slow_mask = slow_mask * 2u + min(wq, 1u);
                                // still biased by MAGIC_BITS, whose low half is zero: phase 3 packs the low halves
                                Bv[q + e] = Bc[e]; Rv[q + e] = Rc[e];
                                // TOPMIN: G > Full-1 -> Full-1 (yuv2tiff.cpp:412); the video-range output clamp is tighter
                                Gv8[q + e] = TOPMIN && C::full_range(A) ? min(G1[e], hi_bits) : G1[e];
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 2; e++) {
                                const int Y = (int)(e ? yw[q >> 1] >> 16 : yw[q >> 1] & 0xffffu);
                                int Rp = 0, Gp = 0, Bp = 0;
                                if (MODE == 2) {                                             // yuv2tiff.cpp:401-402
                                    const int off = Y - (int)(C::Full(A) - 1);
                                    Rp = 2 * (int)phi(cpx[q + e]) + off; Bp = 2 * (int)plo(cpx[q + e]) + off; Gp = Y;
                                    if ((Rp | Bp) < 0) slow_mask |= 0x80u >> (q + e);
                                }
                                Rv[q + e] = Rp; Gv8[q + e] = Gp; Bv[q + e] = Bp;
                            }
                        }
                    }
                    // Phase 2, rare and divergent: the reference's own arithmetic (double, true division, invalid-pixel
                    // rules).  Its result is final (clamped, shifted); shifted back it passes the common tail unchanged.
                    if (slow_mask) {
#pragma unroll
                        for (int q = 0; q < 8; q++)
                            if (slow_mask & (0x80u >> q)) {
                                unsigned R, Gg, B;
                                const int Yq = (int)((q & 1) ? yw[q >> 1] >> 16 : yw[q >> 1] & 0xffffu);
                                if (MODE == 1 && (CFG == 10 || k.int10))
                                    invalid += (unsigned)inv_pixel_int10(k, Yq, (int)plo(cpx[q]), (int)phi(cpx[q]), R, Gg, B);
                                else if (MODE == 1)
                                    invalid += (unsigned)inv_pixel_int(k, Yq, (int)plo(cpx[q]), (int)phi(cpx[q]), R, Gg, B);
                                else
                                    invalid += (unsigned)inv_pixel<MODE == 1 ? H2Y_INV_2020 : (MODE == 2 ? H2Y_INV_YDzDx : -1)>(k, Yq, plo(cpx[q]), phi(cpx[q]), R, Gg, B);
                                Rv[q] = (int)(R >> C::SR(A)); Gv8[q] = (int)(Gg >> C::SR(A)); Bv[q] = (int)(B >> C::SR(A));
                            }
                    }
                    // Phase 3: interleave, then output clamp (yuv2tiff.cpp:515-524) and << SR on two samples at a time
                    // (samples are below 2^(16-SR), so a 32-bit shift cannot carry between the halves)
                    unsigned ow[16];
                    if (!ALPHA) {
                        int flat[24];
#pragma unroll
                        for (int q = 0; q < 8; q++) { flat[3 * q] = Rv[q]; flat[3 * q + 1] = Gv8[q]; flat[3 * q + 2] = Bv[q]; }
#pragma unroll
                        for (int i = 0; i < 12; i++) ow[i] = __byte_perm((unsigned)flat[2 * i], (unsigned)flat[2 * i + 1], 0x5410);
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; q++) { ow[2 * q] = __byte_perm((unsigned)Rv[q], (unsigned)Gv8[q], 0x5410); ow[2 * q + 1] = (unsigned)Bv[q] & 0xffffu; }
                    }
                    {
                        const unsigned lo2 = C::minVR(A) * 0x10001u, hi2 = C::maxVR(A) * 0x10001u;
#pragma unroll
                        for (int i = 0; i < (ALPHA ? 16 : 12); i++) {
                            if (!C::full_range(A)) ow[i] = clamp_u16x2(ow[i], ALPHA && (i & 1) ? (lo2 & 0xffffu) : lo2, ALPHA && (i & 1) ? (hi2 & 0xffffu) : hi2);
                            else if (MODE == 2) ow[i] &= (0xffffu >> C::SR(A)) * 0x10001u;    // unclamped Y'DzDx sums wrap like the reference's u16 store
                            ow[i] <<= C::SR(A);
                            if (ALPHA && (i & 1)) ow[i] |= 0xffff0000u;
                        }
                    }
                    uint4 *o = reinterpret_cast<uint4 *>(orow);
#pragma unroll
                    for (int i = 0; i < nch; i++) o[i] = make_uint4(ow[4 * i], ow[4 * i + 1], ow[4 * i + 2], ow[4 * i + 3]);
                }
            }
            __syncwarp();
        }
    }
    if (A.invalid && invalid && invalid_frame >= 0) atomicAdd(&A.invalid[invalid_frame], invalid);
}

h2y_status make_invk(const h2y_inverse_params &p, InvK *k)
{
    if (p.width < 8 || p.height < 2 || (p.width & 7) || (p.height & 1)) return H2Y_ERR_ARG;
    if (p.bit_depth != 10 && p.bit_depth != 12 && p.bit_depth != 14) return H2Y_ERR_ARG;
    if (p.matrix < H2Y_INV_YDzDx || p.matrix > H2Y_INV_Y500) return H2Y_ERR_ARG;
    k->w = p.width; k->h = p.height; k->bit_depth = p.bit_depth; k->matrix = p.matrix;
    k->fir = p.fir != 0; k->full_range = p.full_range != 0; k->alpha = p.alpha != 0;
    k->int10 = p.matrix == H2Y_INV_2020 && p.bit_depth == 10;    // inv_pixel_int10
    k->ikb = k->ikr = k->igy = k->igb = k->igr = k->igc = 0; k->igd = 1; k->igm = 0;
    if (p.matrix == H2Y_INV_2020) {          // inv_pixel_int: 0.9407, 0.7373; (2 (10000 Y - 593 B - 2627 R) + 6780) / 13560
        k->ikb = 9407; k->ikr = 7373; k->igy = 20000; k->igb = -1186; k->igr = -5254; k->igc = 6780; k->igd = 13560;
    } else if (p.matrix == H2Y_INV_709) {    // 0.9278, 0.7874; (50000 Y - 3611 B - 10630 R + 17880) / 35760
        k->ikb = 9278; k->ikr = 7874; k->igy = 50000; k->igb = -3611; k->igr = -10630; k->igc = 17880; k->igd = 35760;
    }
    k->igm = (unsigned)(4294967296.0 / (double)k->igd);
    k->ybar = p.ybar != 0 && p.matrix == H2Y_INV_YDzDx;          // -X only acts inside the Y'DzDx branch (yuv2tiff.cpp:389-399)
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
    // large batches: warp-autonomous rows kernel
    {
        Inv2Args A;
        A.k = k; A.yuv = (const uint8_t *)d_yuv; A.yuv_stride = yuv_stride; A.rgb = (uint8_t *)d_rgb; A.rgb_stride = rgb_stride;
        A.invalid = d_invalid;
        A.nstrips = (k.w + 239) / 240;
        A.wps = 1;
        for (int d = 1; d <= RWARPS; d++) if (RWARPS % d == 0 && A.nstrips % d == 0) A.wps = d;
        A.sub = RWARPS / A.wps;
        A.total_crows = (long)nframes * (k.h / 2);
        A.hm = (float)k.Half - 0.5f; A.kb = (float)k.kb; A.kr = (float)k.kr;
        A.nwb = -(float)k.wb; A.nwr = -(float)k.wr; A.rwg = (float)(1.0 / k.wg);
        A.guard_c = H2Y_INV_GUARD_C / (float)(1 << (25 - k.bit_depth));        // 6u and 10u with u = 2^(d-25): bounds in k_inverse_rows
        A.guard_g = H2Y_INV_GUARD_G / (float)(1 << (25 - k.bit_depth));
        const long rows_per_worker = A.total_crows / ((long)c->sm_count * RMINB * A.sub);
        // forced by tests and experiments (h2y_ctx_set_option); -X: tile kernel only
        // (a worker's halo is seven chroma rows that are only fetched and stashed, so short runs cost little: one 4K frame,
        // 7 chroma rows per worker, 0.036 ms against the tile kernel's 0.079 ms; eight frames 0.165 against 0.598 ms)
        const bool want_rows = !k.ybar && (c->sw.inv_kernel ? c->sw.inv_kernel >= 2 : rows_per_worker >= 4);
        if (want_rows) {
            int grid = c->sm_count * RMINB;
            while (grid > 1 && A.total_crows / ((long)grid * A.sub) < 4) grid >>= 1;
            const size_t smem = INV_ROWS_SMEM;
            const int mode = (k.matrix == H2Y_INV_2020 || k.matrix == H2Y_INV_709) ? 1 : (k.matrix == H2Y_INV_YDzDx ? 2 : 0);
#define LR(M, F, AL)                                                                                                       \
    do {                                                                                                                   \
        H2Y_CUDA(c, cudaFuncSetAttribute(k_inverse_rows<M, F, AL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_inverse_rows<M, F, AL><<<grid, RTHREADS, smem, st>>>(A);                                                        \
    } while (0)
#define LRM(M)                                                                                                             \
    do {                                                                                                                   \
        if (k.fir && k.alpha) LR(M, true, true);                                                                           \
        else if (k.fir) LR(M, true, false);                                                                                \
        else if (k.alpha) LR(M, false, true);                                                                              \
        else LR(M, false, false);                                                                                          \
    } while (0)
            const bool cfg10 = k.matrix == H2Y_INV_2020 && k.bit_depth == 10 && !k.full_range && k.fir && !k.alpha &&
                               !c->sw.no_specialised;
            if (c->sw.inv_kernel == 3 && mode == 1 && !k.alpha) {
                if (k.fir) {
                    H2Y_CUDA(c, cudaFuncSetAttribute(k_inverse_rows<1, true, false, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    k_inverse_rows<1, true, false, 0, true><<<grid, RTHREADS, smem, st>>>(A);
                } else {
                    H2Y_CUDA(c, cudaFuncSetAttribute(k_inverse_rows<1, false, false, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    k_inverse_rows<1, false, false, 0, true><<<grid, RTHREADS, smem, st>>>(A);
                }
            } else if (cfg10) {
                H2Y_CUDA(c, cudaFuncSetAttribute(k_inverse_rows<1, true, false, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_inverse_rows<1, true, false, 10><<<grid, RTHREADS, smem, st>>>(A);
            } else if (mode == 1) LRM(1); else if (mode == 2) LRM(2); else LRM(0);
#undef LRM
#undef LR
            c->launches++;
            H2Y_CUDA(c, cudaGetLastError());
            return H2Y_OK;
        }
    }
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
