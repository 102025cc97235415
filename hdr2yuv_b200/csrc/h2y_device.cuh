// h2y_device.cuh -- device-side building blocks shared by every kernel of the hot path.
//
// Everything here reproduces the reference's C++ arithmetic bit for bit:
// float vs double per sub-expression, no FMA contraction (this TU set is compiled with
// -fmad=false and the parity-critical expressions use the explicit _rn intrinsics anyway),
// truncation toward zero, the signed/unsigned clamp quirk and the Half-1 chroma offset.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "hdr2yuv_b200.h"

namespace h2y {

// ---- per-launch constants ------------------------------------------------------------------

// colour-difference family selected by out->matrix_coeffs (convert.cpp:1159-1198)
enum MatKind : int { MK_PASS = 0, MK_YDZDX = 1, MK_YCBCR = 2, MK_Y100 = 3, MK_PRIME2 = 4 };

// transfer functions of convert.cpp:12-87 as used by the two steps of convert.cpp:1021-1109
enum TfKind : int { TF_NONE = 0, TF_PQ = 1, TF_RHO = 2, TF_GAMMA = 3 };

enum ScaleMode : int { SC_NONE = 0, SC_FULL = 1, SC_VIDEO = 2, SC_FLOAT_OUT = 3 };

struct PixK {
    // matrix stage
    int mat_kind;
    double wr, wg, wb;   // luma weights (convert.cpp:1177, 1182)
    int wri, wgi, wbi;   // the same weights as integers / 10000 (both families are four-decimal constants); 0 = not available
    double db, dr;       // chroma divisors 1.8814/1.4746 or 1.8556/1.5748
    double rdb, rdr;     // RN(1/db), RN(1/dr) for the reciprocal fast path
    float P, Q, RR, S;   // Y100/Y500 (convert.cpp:911-925)
    int half_m1;         // clip->Half - 1 of the tmp picture (convert.cpp:1200-1201)
    unsigned maxCV;      // clip->maxCV of the tmp picture (convert.cpp:1207-1213)
    // transfer stage (only when the transfer changes)
    int convert_transfer;
    int tf_linearise;    // TfKind applied when src transfer != LINEAR (convert.cpp:1024-1064)
    int tf_encode;       // TfKind applied when dst transfer != LINEAR (convert.cpp:1068-1109)
    int scale_mode;      // convert.cpp:1116-1144
    float mulY, addY, mulC, addC;
    // write_yuv stage (tiff.cpp:457-550) at the output depth
    int down_shift;
    unsigned loY, hiY, loC, hiC;
    // on-read clip of read_tiff (tiff.cpp:296-304), applied to 16-bit codes when enabled
    int clip_on_load;
    unsigned loadLo, loadHi;
};

// ---- x86-64 conversion semantics -------------------------------------------------------------

// (unsigned int)float on x86-64 is a 64-bit cvttss2si whose low half is kept; NaN and values outside the int64 range
// give the "integer indefinite" 0x8000000000000000, whose low half is 0 (CUDA's conversion would saturate instead).
__device__ __forceinline__ unsigned f2u_x86(float x)
{
    if (!(x > -9.2233720e18f && x < 9.2233720e18f)) return 0u;
    return (unsigned)__float2ll_rz(x);
}

// (int)double is a 32-bit cvttsd2si: NaN / out of range give INT_MIN.
__device__ __forceinline__ int d2i_x86(double x)
{
    int k = __double2int_rz(x);
    if (!(x > -2147483649.0 && x < 2147483648.0)) k = (int)0x80000000;
    return k;
}

__device__ __forceinline__ int f2i_x86(float x)
{
    int k = __float2int_rz(x);
    if (!(x > -2147483904.0f && x < 2147483648.0f)) k = (int)0x80000000;
    return k;
}

__device__ __forceinline__ float half_bits_to_float(unsigned h) { return __half2float(__ushort_as_half((unsigned short)h)); }

// ---- transfer functions in double, float in/out (convert.cpp:12-87) --------------------------

__device__ __forceinline__ float tf_pq_eotf(float V)
{
    double vp = pow((double)V, 1.0 / 78.84375);
    double num = fmax(__dadd_rn(vp, -0.8359375), 0.0);
    double den = __dadd_rn(18.8515625, -__dmul_rn(18.6875, vp));
    return __double2float_rn(pow(__ddiv_rn(num, den), 1.0 / 0.1593017578));
}

__device__ __forceinline__ float tf_pq_oetf(float L)
{
    double lp = pow((double)L, 0.1593017578);
    double num = __dadd_rn(0.8359375, __dmul_rn(18.8515625, lp));
    double den = __dadd_rn(1.0, __dmul_rn(18.6875, lp));
    return __double2float_rn(pow(__ddiv_rn(num, den), 78.84375));
}

// BT.1886 with the fixed call arguments gamma=2.4f, Lw=1, Lb=0: a = 1, b = 0 (convert.cpp:67-87, 1051-1057)
__device__ __forceinline__ float tf_gamma_eotf(float V)
{
    float vb = __fadd_rn(V, 0.0f);
    return __double2float_rn(__dmul_rn(1.0, pow(fmax((double)vb, 0.0), (double)2.4f)));
}

__device__ __forceinline__ float tf_gamma_oetf(float L)
{
    float la = __fdiv_rn(L, 1.0f);
    return __double2float_rn(__dadd_rn(pow(fmax((double)la, 0.0), __ddiv_rn(1.0, (double)2.4f)), -0.0));
}

// rho-gamma: pow(rho, V) and log(rho) resolve to the float overloads in the reference's C++
// (convert.cpp:24, 36); glibc's powf/logf are correctly rounded in practice, emulated by
// rounding the double result.
__device__ __forceinline__ float tf_rho_eotf(float V)
{
    float rv = __double2float_rn(pow(25.0, (double)V));
    double base = __ddiv_rn(__dadd_rn((double)rv, -1.0), 24.0);
    return __double2float_rn(pow(base, (double)2.4f));
}

__device__ __forceinline__ float tf_rho_oetf(float L)
{
    double lg = pow((double)L, __ddiv_rn(1.0, (double)2.4f));
    double num = log(__dadd_rn(1.0, __dmul_rn(24.0, lg)));
    float lrho = __double2float_rn(log(25.0));
    return __double2float_rn(__ddiv_rn(num, (double)lrho));
}

// normalised sample -> destination transfer domain (convert.cpp:1021-1109)
__device__ __forceinline__ float change_transfer(float x, int tf_linearise, int tf_encode)
{
    if (tf_linearise == TF_PQ) x = tf_pq_eotf(x);
    else if (tf_linearise == TF_RHO) x = tf_rho_eotf(x);
    else if (tf_linearise == TF_GAMMA) x = tf_gamma_eotf(x);
    if (tf_encode == TF_PQ) x = tf_pq_oetf(x);
    else if (tf_encode == TF_RHO) x = tf_rho_oetf(x);
    else if (tf_encode == TF_GAMMA) x = tf_gamma_oetf(x);
    return x;
}

// range scale after the transfer change (convert.cpp:1116-1144), U16 destination
__device__ __forceinline__ void scale_to_codes(float &G, float &B, float &R, const PixK &k)
{
    if (k.scale_mode == SC_FULL) {
        G = __fmul_rn(G, k.mulY);
        B = __fmul_rn(B, k.mulY);
        R = __fmul_rn(R, k.mulY);
    } else if (k.scale_mode == SC_VIDEO) {
        G = __fadd_rn(__fmul_rn(G, k.mulY), k.addY);
        B = __fadd_rn(__fmul_rn(B, k.mulC), k.addC);
        R = __fadd_rn(__fmul_rn(R, k.mulC), k.addC);
    }
}

// ---- colour-difference stage, exact form (convert.cpp:1146-1220) -----------------------------

template <int MK>
__device__ __forceinline__ void px_matrix_exact(float G, float B, float R, const PixK &k, unsigned &Yo,
                                                unsigned &Cbo, unsigned &Cro)
{
    unsigned y;
    long long cb, cr;
    if (MK == MK_PASS) {
        y = f2u_x86(G);
        cb = f2u_x86(B);
        cr = f2u_x86(R);
    } else {
        if (MK == MK_YDZDX) {
            y = f2u_x86(G);
            double hg = __dmul_rn((double)(-G), 0.5);   // -G/2.0 is exact
            cb = d2i_x86(__dadd_rn(__dadd_rn(hg, __dmul_rn((double)B, 0.5)), 0.5));
            cr = d2i_x86(__dadd_rn(__dadd_rn(hg, __dmul_rn((double)R, 0.5)), 0.5));
        } else if (MK == MK_YCBCR) {
            double s = __dadd_rn(__dadd_rn(__dmul_rn(k.wr, (double)R), __dmul_rn(k.wg, (double)G)),
                                 __dmul_rn(k.wb, (double)B));
            float tmpF = __double2float_rn(__dadd_rn(s, 0.5));
            y = f2u_x86(tmpF);
            cb = d2i_x86(__dadd_rn(__ddiv_rn((double)__fsub_rn(B, tmpF), k.db), 0.5));
            cr = d2i_x86(__dadd_rn(__ddiv_rn((double)__fsub_rn(R, tmpF), k.dr), 0.5));
        } else if (MK == MK_Y100) {
            y = f2u_x86(G);
            cb = d2i_x86(__dadd_rn((double)__fadd_rn(__fmul_rn(k.P, G), __fmul_rn(k.Q, B)), 0.5));
            cr = d2i_x86(__dadd_rn((double)__fadd_rn(__fmul_rn(k.RR, R), __fmul_rn(k.S, G)), 0.5));
        } else {   // MK_PRIME2 in the 4:4:4 stage is a copy (convert.cpp:1191-1194)
            y = f2u_x86(G);
            cb = f2u_x86(B);
            cr = f2u_x86(R);
        }
        cb += k.half_m1;
        cr += k.half_m1;
    }
    // Y is unsigned; chroma is compared through unsigned long, so negatives clamp to maxCV
    Yo = y > k.maxCV ? k.maxCV : y;
    Cbo = (unsigned long long)cb > (unsigned long long)k.maxCV ? k.maxCV : (unsigned)cb;
    Cro = (unsigned long long)cr > (unsigned long long)k.maxCV ? k.maxCV : (unsigned)cr;
}

constexpr int FRAC_BITS = 14;

// ---- colour-difference stage, fast form ------------------------------------------------------------
// The reference computes  k = (int)( d / c + 0.5 )  in double.  Here q = RZ(d * RN(1/c) + 0.5 + M)
// with M = 1.5*2^38 leaves floor(q * 2^14) in the low mantissa word; |q - exact| < 2^-36, so when the
// 14 fraction bits are neither all 0 nor all 1 the truncated integer is certain.  Otherwise (and
// for NaN, whose low word is 0) the pixel is redone by px_matrix_exact.  Returns true when safe.
__device__ __forceinline__ bool trunc_from_magic(double q, int &k)
{
    const int lo = __double2loint(q);
    k = (lo >> FRAC_BITS) + (int)((unsigned)lo >> 31);
    return ((unsigned)(lo + 1) & ((1u << FRAC_BITS) - 1u)) > 1u;
}

template <int MK>
__device__ __forceinline__ bool px_matrix_fast(float G, float B, float R, const PixK &k, unsigned &Y, unsigned &Cb,
                                               unsigned &Cr)
{
    const double MAGIC = 412316860416.0 + 0.5;   // 1.5 * 2^38 + the reference's +0.5
    bool ok = true;
    int cb, cr;
    if (MK == MK_PASS) {
        Y = min(__float2uint_rz(G), k.maxCV);
        Cb = min(__float2uint_rz(B), k.maxCV);
        Cr = min(__float2uint_rz(R), k.maxCV);
        // values of 2^32 and above wrap in the reference ((unsigned) keeps the low word): exact route
        return G >= 0.0f && B >= 0.0f && R >= 0.0f && G < 4294967296.0f && B < 4294967296.0f && R < 4294967296.0f;
    } else if (MK == MK_YCBCR) {
        const double s = __dadd_rn(__dadd_rn(__dmul_rn(k.wr, (double)R), __dmul_rn(k.wg, (double)G)),
                                   __dmul_rn(k.wb, (double)B));
        const float tmpF = __double2float_rn(__dadd_rn(s, 0.5));
        Y = min(__float2uint_rz(tmpF), k.maxCV);
        ok = tmpF >= 0.0f && tmpF < 4294967296.0f;
        ok &= trunc_from_magic(__fma_rz((double)__fsub_rn(B, tmpF), k.rdb, MAGIC), cb);
        ok &= trunc_from_magic(__fma_rz((double)__fsub_rn(R, tmpF), k.rdr, MAGIC), cr);
    } else if (MK == MK_YDZDX) {
        Y = min(__float2uint_rz(G), k.maxCV);
        ok = G >= 0.0f && G < 4294967296.0f;
        const double hg = __dmul_rn((double)G, -0.5);
        ok &= trunc_from_magic(__dadd_rz(__dadd_rn(hg, __dmul_rn((double)B, 0.5)), MAGIC), cb);
        ok &= trunc_from_magic(__dadd_rz(__dadd_rn(hg, __dmul_rn((double)R, 0.5)), MAGIC), cr);
    } else {   // MK_Y100
        Y = min(__float2uint_rz(G), k.maxCV);
        ok = G >= 0.0f && G < 4294967296.0f;
        ok &= trunc_from_magic(__dadd_rz((double)__fadd_rn(__fmul_rn(k.P, G), __fmul_rn(k.Q, B)), MAGIC), cb);
        ok &= trunc_from_magic(__dadd_rz((double)__fadd_rn(__fmul_rn(k.RR, R), __fmul_rn(k.S, G)), MAGIC), cr);
    }
    // negatives compare as huge unsigned and clamp to maxCV, like the reference's unsigned long compare
    Cb = min((unsigned)(cb + k.half_m1), k.maxCV);
    Cr = min((unsigned)(cr + k.half_m1), k.maxCV);
    return ok;
}

// Integer-source variant of px_matrix_fast (TIFF rows): identical arithmetic, but the integer codes become doubles /
// floats through magic-number adds instead of I2F + F2F conversions (16-32 lanes/clk/SM on B200).  u2d/u2f are exact.
__device__ __forceinline__ double u2d(unsigned v) { return __dadd_rn(__hiloint2double(0x43300000, (int)v), -4503599627370496.0); }
__device__ __forceinline__ float u2f(unsigned v) { return __fadd_rn(__uint_as_float(0x4B000000u | v), -8388608.0f); }   // v < 2^23

// FAM: 0 = constants from PixK at run time; 2020 / 709 = the family's constants as immediates (no constant-bank reloads
// per pixel; the host picks the instantiation whose constants equal PixK's).
template <int FAM> struct YccFam {
    static constexpr int WR = FAM == 2020 ? 2627 : 2126, WG = FAM == 2020 ? 6780 : 7152, WB = FAM == 2020 ? 593 : 722;
    static constexpr double DB = FAM == 2020 ? 1.8814 : 1.8556, DR = FAM == 2020 ? 1.4746 : 1.5748;
};

template <int MK, int FAM = 0>
__device__ __forceinline__ bool px_matrix_fast_u16(unsigned g, unsigned b, unsigned r, const PixK &k, unsigned &Y,
                                                   unsigned &Cb, unsigned &Cr)
{
    const double MAGICD = 412316860416.0 + 0.5;   // 1.5 * 2^38 + the reference's +0.5
    bool ok = true;
    int cb, cr;
    if (MK == MK_YCBCR) {
        // tmpF = (float)((wr*R + wg*G) + wb*B + 0.5), the sum in double (convert.cpp:1177).  For integer samples and the
        // four-decimal weights this is RN_float(N / 10000) with the INTEGER N = wri*R + wgi*G + wbi*B + 5000 < 2^31:
        // N / 10^4 is at least 1.6e-3 of a float half-ulp away from every float rounding boundary ((2k+1) * 2^(e-24) =
        // N / (2^4 * 625) has no solution, and |N * 2^(20-e) - 625 (2k+1)| >= 1), i.e. >= 9.5e-11 relative, while both the
        // reference's three rounded products and two rounded sums and this single rounded product stay within 5e-16.
        // One integer-to-double conversion and one DMUL replace three conversions, three DMUL and three DADD.
        float tmpF;
        if (FAM) {
            const int n = (int)r * YccFam<FAM>::WR + (int)g * YccFam<FAM>::WG + (int)b * YccFam<FAM>::WB + 5000;
            tmpF = __double2float_rn(__dmul_rn(__int2double_rn(n), 1e-4));
        } else if (k.wri) {
            const int n = (int)r * k.wri + (int)g * k.wgi + (int)b * k.wbi + 5000;
            tmpF = __double2float_rn(__dmul_rn(u2d((unsigned)n), 1e-4));
        } else {
            const double s = __dadd_rn(__dadd_rn(__dmul_rn(k.wr, u2d(r)), __dmul_rn(k.wg, u2d(g))), __dmul_rn(k.wb, u2d(b)));
            tmpF = __double2float_rn(__dadd_rn(s, 0.5));
        }
        // tmpF >= 0.5 here: truncation = floor, taken by a round-toward-zero add of 2^23
        Y = min(__float_as_uint(__fadd_rz(tmpF, 8388608.0f)) & 0x7FFFFFu, k.maxCV);
        const double rdb = FAM ? 1.0 / YccFam<FAM>::DB : k.rdb, rdr = FAM ? 1.0 / YccFam<FAM>::DR : k.rdr;
        const float bf = FAM ? __uint2float_rn(b) : u2f(b), rf = FAM ? __uint2float_rn(r) : u2f(r);     // exact either way
        ok &= trunc_from_magic(__fma_rz((double)__fsub_rn(bf, tmpF), rdb, MAGICD), cb);
        ok &= trunc_from_magic(__fma_rz((double)__fsub_rn(rf, tmpF), rdr, MAGICD), cr);
    } else {   // MK_YDZDX on integers: (int)(-G/2.0 + B/2.0 + 0.5) == (B - G + 1) / 2 with C's truncating division
        Y = min(g, k.maxCV);
        const int tb = (int)b - (int)g + 1, tr = (int)r - (int)g + 1;
        cb = (tb - (tb >> 31)) >> 1;
        cr = (tr - (tr >> 31)) >> 1;
    }
    Cb = min((unsigned)(cb + k.half_m1), k.maxCV);
    Cr = min((unsigned)(cr + k.half_m1), k.maxCV);
    return ok;
}

// ---- write_yuv's shift + range clamp (tiff.cpp:457-550) --------------------------------------
__device__ __forceinline__ unsigned out_clamp(unsigned v, int shift, unsigned lo, unsigned hi)
{
    v >>= shift;
    v = v < lo ? lo : v;
    return v > hi ? hi : v;
}

// ---- FIR taps in the reference's float order (convert.cpp:305-311, 365-370) ------------------
// 7-tap horizontal, co-sited: inputs are s[x-5],s[x-3],s[x-1],s[x],s[x+1],s[x+3],s[x+5]
__device__ __forceinline__ unsigned fir_h7(float m5, float m3, float m1, float c, float p1, float p3, float p5,
                                           float maxCV)
{
    const float k21 = 21.0f / 512.0f, k52 = 52.0f / 512.0f, k159 = 159.0f / 512.0f, k256 = 256.0f / 512.0f;
    float t = __fmul_rn(k21, __fadd_rn(m5, p5));
    t = __fsub_rn(t, __fmul_rn(k52, __fadd_rn(m3, p3)));
    t = __fadd_rn(t, __fmul_rn(k159, __fadd_rn(m1, p1)));
    t = __fadd_rn(t, __fmul_rn(k256, c));
    t = __fadd_rn(t, 0.5f);
    t = fminf(t, maxCV);
    t = fmaxf(t, 0.0f);
    return (unsigned)__float2int_rz(t);
}

// same filter, result kept as a float integer: the (unsigned short) truncation of the clamped value is a
// round-toward-zero add of 2^23 (t >= 0 after the clamp), so no F2I / I2F pair is needed
__device__ __forceinline__ float fir_h7_f(float m5, float m3, float m1, float c, float p1, float p3, float p5, float maxCV)
{
    const float k21 = 21.0f / 512.0f, k52 = 52.0f / 512.0f, k159 = 159.0f / 512.0f, k256 = 256.0f / 512.0f;
    float t = __fmul_rn(k21, __fadd_rn(m5, p5));
    t = __fsub_rn(t, __fmul_rn(k52, __fadd_rn(m3, p3)));
    t = __fadd_rn(t, __fmul_rn(k159, __fadd_rn(m1, p1)));
    t = __fadd_rn(t, __fmul_rn(k256, c));
    t = __fadd_rn(t, 0.5f);
    t = fminf(t, maxCV);
    t = fmaxf(t, 0.0f);
    return __fadd_rn(__fadd_rz(t, 8388608.0f), -8388608.0f);
}

// 12-tap vertical, half-phase: r[0..11] are rows y-5 .. y+6
__device__ __forceinline__ unsigned fir_v12(const float r[12], float maxCV)
{
    const float k228 = 228.0f / 512.0f, k70 = 70.0f / 512.0f, k37 = 37.0f / 512.0f, k21 = 21.0f / 512.0f,
                k11 = 11.0f / 512.0f, k5 = 5.0f / 512.0f;
    float t = __fmul_rn(k228, __fadd_rn(r[5], r[6]));
    t = __fadd_rn(t, __fmul_rn(k70, __fadd_rn(r[4], r[7])));
    t = __fsub_rn(t, __fmul_rn(k37, __fadd_rn(r[3], r[8])));
    t = __fsub_rn(t, __fmul_rn(k21, __fadd_rn(r[2], r[9])));
    t = __fadd_rn(t, __fmul_rn(k11, __fadd_rn(r[1], r[10])));
    t = __fadd_rn(t, __fmul_rn(k5, __fadd_rn(r[0], r[11])));
    t = __fadd_rn(t, 0.5f);
    t = fminf(t, maxCV);
    t = fmaxf(t, 0.0f);
    return (unsigned)__float2int_rz(t);
}

// same, result as the bit pattern of 1.5*2^23 + trunc(t) (0x4B400000 + value): the form the packed kernels' epilogue uses
__device__ __forceinline__ int fir_v12_magic(const float r[12], float maxCV)
{
    const float k228 = 228.0f / 512.0f, k70 = 70.0f / 512.0f, k37 = 37.0f / 512.0f, k21 = 21.0f / 512.0f,
                k11 = 11.0f / 512.0f, k5 = 5.0f / 512.0f;
    float t = __fmul_rn(k228, __fadd_rn(r[5], r[6]));
    t = __fadd_rn(t, __fmul_rn(k70, __fadd_rn(r[4], r[7])));
    t = __fsub_rn(t, __fmul_rn(k37, __fadd_rn(r[3], r[8])));
    t = __fsub_rn(t, __fmul_rn(k21, __fadd_rn(r[2], r[9])));
    t = __fadd_rn(t, __fmul_rn(k11, __fadd_rn(r[1], r[10])));
    t = __fadd_rn(t, __fmul_rn(k5, __fadd_rn(r[0], r[11])));
    t = __fadd_rn(t, 0.5f);
    t = fminf(t, maxCV);
    t = fmaxf(t, 0.0f);
    return __float_as_int(__fadd_rz(t, 12582912.0f));
}

}   // namespace h2y
