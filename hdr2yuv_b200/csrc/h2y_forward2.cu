// h2y_forward2.cu -- the fast forward kernels behind h2y_forward (the general kernel is h2y_forward.cu).
//
//   k_forward_exr420<.., SRC=0>   EXR route (half source through the per-frame LUT, 4:2:0 FIR, tmp depth <= 12), small
//                                 batches: CTA-shared 64-row ring, vertical filter register-blocked 4 outputs x 18 rows
//   k_forward_exr420_rows         EXR route, large batches: warp-autonomous, vertical filter as six running register
//                                 accumulators per chroma column, no ring, no CTA barrier inside a frame
//   k_forward_exr420<.., SRC=1>   integer rows (TIFF), 4:2:0 FIR at 16-bit tmp depth: same ring, reference-order arithmetic
//   k_forward_u16_444             integer rows, 4:4:4 output: a streaming map at HBM speed
//
// Common to the EXR kernels, and why they are still bit-exact (error bound in DESIGN.md 4):
//  * Per-pixel arithmetic runs on Blackwell's packed fp32x2 pipe (FFMA2 / FADD2: two lanes of fp32 per issue slot).
//    The colour-difference stage is evaluated in fp32 with a guard band: every truncation is taken twice, at x-G and
//    x+G, with a round-down add of 1.5*2^23 (the integer falls out of the mantissa, no F2I); when both agree the
//    truncated integer is certain, because the fp32 evaluation is within 7.2/8 G of the reference's double evaluation
//    (G = 2^(depth-22) = 8 half-ulps of the largest values; the terms of the bound are listed in DESIGN.md 4).
//    The rare pixel whose two truncations differ (0.15 % at 10 bits) is redone with the general kernel's FP64 routine,
//    inlined (an out-of-line call cost ~2900 cycles per 16-row step through ABI spills); four pixels share a branch.
//  * Frames whose LUT extremes prove that Cb/Cr stay clear of matrix_convert's clamp (two_lut_frame) take the TWO
//    instantiation: two pre-scaled LUT copies, chroma truncated with FRND.TRUNC, Half-1 added by the filter constant.
//  * The range scale (convert.cpp:1139-1144) keeps its two separately rounded fp32 operations, so the values entering
//    the matrix are the reference's own floats.
//  * Chroma stays in float from the truncation to the .yuv store; the u16 quantisation of the reference's `dst422`
//    intermediate and of tmp444 is reproduced by round-down adds + integer clamps on the magic-number bit pattern.
//    At tmp depth <= 12 every FIR term is an integer / 512 below 2^24, so the FMA order is free (SURVEY.md Appendix
//    A.7) and both filters are FFMA2 chains on {Cb,Cr} pairs.
//  * The LUT copy in shared memory is indexed by the raw half code; frames with negative, infinite or NaN samples
//    (never "clean") are left to the general kernel, which handles every case.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "h2y_f32x2.cuh"
#include "h2y_internal.h"

// Guard-band half width G = 2^(depth - H2Y_GUARD_SHIFT); see the error bound in DESIGN.md section 4.  A build switch
// only for the A/B timings in profiles/r01/variants.md (21 was the first, more conservative, choice).
#ifndef H2Y_GUARD_SHIFT
#define H2Y_GUARD_SHIFT 22
#endif
#ifndef H2Y_PAIRS_PER_BRANCH_THREE
#define H2Y_PAIRS_PER_BRANCH_THREE 1
#endif
// Pixel pairs per guard-band branch (1, 2 or 4): 2 measured fastest (profiles/r01/variants.md).
#ifndef H2Y_PAIRS_PER_BRANCH_THREE_NC
#define H2Y_PAIRS_PER_BRANCH_THREE_NC 1
#endif
#ifndef H2Y_PAIRS_PER_BRANCH
#define H2Y_PAIRS_PER_BRANCH 2
#endif

namespace h2y {

namespace {
constexpr int THREADS = 512;
constexpr int RING_ROWS = 64, RING_COLS = 120;            // float2 per column
constexpr int RING_PITCH = RING_COLS * 2;                 // floats per ring row
constexpr int LUT_MAX_CODES = 0x7C00;                     // clean frames: codes 0 .. 0x7BFF

// position of chroma column c (0..119) inside a ring row, in float2 units: the horizontal stage stores
// outputs {0,1} and {2,3} of a lane with two conflict-free 16-byte stores, so columns 4l, 4l+1 live in
// the first half of the row and 4l+2, 4l+3 in the second.
__device__ __forceinline__ int ring_pos(int c) { return ((c & 2) ? RING_COLS / 2 : 0) + ((c >> 2) << 1) + (c & 1); }
}   // namespace

struct Fwd2Args {
    const uint8_t *src;
    size_t src_stride;
    uint8_t *dst;
    size_t dst_stride;
    int w, h, nframes;
    int strip_w, nstrips, seg_rows, nsegs, nitems;
    float guard;             // G
    float wr, wg, wb, rdb, rdr;   // fp32 images of the double constants (their error is inside the guard budget)
    float lumc, cbc, crc;         // 0.5-Gy; chroma constants with the lowered luma folded in
    float twoGy, twoGb, twoGr;    // band widths 2G of the luma, Cb and Cr chains
    PixK k;
    const FrameK *framek;
    const float *luts;
    unsigned long long *fallback_count;
};

// one 8-pixel load: 3 (RGB) or 4 (RGBA) 16-byte vectors
template <int NCH> struct RawPx { uint4 v[NCH]; };

template <int NCH>
__device__ __forceinline__ void load_px8(RawPx<NCH> &raw, const uint8_t *frame, int w, int row, int x)
{
    const uint4 *p = reinterpret_cast<const uint4 *>(frame + ((size_t)row * w + x) * (2 * NCH));
#pragma unroll
    for (int i = 0; i < NCH; i++) raw.v[i] = __ldg(p + i);
}

template <int NCH>
__device__ __forceinline__ void split_codes(const RawPx<NCH> &raw, unsigned g[8], unsigned b[8], unsigned r[8])
{
    if (NCH == 3) {
        unsigned s[24];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const uint4 &v = raw.v[i];
            s[8 * i + 0] = v.x & 0xffffu; s[8 * i + 1] = v.x >> 16; s[8 * i + 2] = v.y & 0xffffu; s[8 * i + 3] = v.y >> 16;
            s[8 * i + 4] = v.z & 0xffffu; s[8 * i + 5] = v.z >> 16; s[8 * i + 6] = v.w & 0xffffu; s[8 * i + 7] = v.w >> 16;
        }
#pragma unroll
        for (int q = 0; q < 8; q++) { r[q] = s[3 * q]; g[q] = s[3 * q + 1]; b[q] = s[3 * q + 2]; }
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint4 &v = raw.v[i];
            r[2 * i] = v.x & 0xffffu; g[2 * i] = v.x >> 16; b[2 * i] = v.y & 0xffffu;
            r[2 * i + 1] = v.z & 0xffffu; g[2 * i + 1] = v.z >> 16; b[2 * i + 1] = v.w & 0xffffu;
        }
    }
}

// the reference-exact route for a pixel inside the guard band (rare: kept out of line)
template <int MK>
__device__ __forceinline__ void pixel_exact(float G, float B, float R, const PixK &k, unsigned &Y, unsigned &Cb, unsigned &Cr)
{
    if (!px_matrix_fast<MK>(G, B, R, k, Y, Cb, Cr)) px_matrix_exact<MK>(G, B, R, k, Y, Cb, Cr);
}

// Per-launch constants.  CFG 0 reads them from the launch arguments; CFG 10 / 12 are the headline configurations
// (BT.2020nc, 10- or 12-bit tmp and output depth, video range) with every constant an immediate: fewer live registers
// and no constant-bank reloads inside the pixel loop (measured: 1.49 -> 1.39 ms per 60 4K frames).  The values are
// what launch_forward_exr420 computes for those configurations; the host checks that before choosing them.
template <int CFG> struct KC {
    // CFG = 0: run time.  CFG = D (10 or 12): BT.2020nc at D-bit tmp/output depth, video range, no output shift
    static constexpr int D = CFG ? CFG : 10, S = 1 << (D - 8);
    // guard bands per chain, in units of u = 2^(D-25): 6u for luma (needs 5.125u), 7u for Cb (6.3u), 8u for Cr (7.2u);
    // DESIGN.md 4, "Guard-band truncation"
    static constexpr float G = 1.0f / (float)(1 << (H2Y_GUARD_SHIFT - D)), GY = 0.75f * G, GB = 0.875f * G, GR = G;
    static constexpr float RDB = (float)(1.0 / 1.8814), RDR = (float)(1.0 / 1.4746);
#define KCF(name, rt, ct) __device__ __forceinline__ static float name(const Fwd2Args &a) { return CFG ? (ct) : (rt); }
#define KCI(name, rt, ct) __device__ __forceinline__ static int name(const Fwd2Args &a) { return CFG ? (ct) : (int)(rt); }
    KCF(mulY, a.k.mulY, (float)(235 * S)) KCF(mulC, a.k.mulC, (float)(240 * S)) KCF(addY, a.k.addY, (float)(16 * S)) KCF(addC, a.k.addC, (float)(16 * S))
    KCF(wr, a.wr, (float)0.2627) KCF(wg, a.wg, (float)0.6780) KCF(wb, a.wb, (float)0.0593)
    KCF(rdb, a.rdb, RDB) KCF(rdr, a.rdr, RDR)
    KCF(lumc, a.lumc, 0.5f - GY) KCF(twoGy, a.twoGy, 2.0f * GY) KCF(twoGb, a.twoGb, 2.0f * GB) KCF(twoGr, a.twoGr, 2.0f * GR)
    KCF(cbc, a.cbc, 0.5f - GB - GY * RDB) KCF(crc, a.crc, 0.5f - GR - GY * RDR)
    KCI(maxCV, a.k.maxCV, (1 << D) - 1) KCI(half_m1, a.k.half_m1, (1 << (D - 1)) - 1) KCI(shift, a.k.down_shift, 0)
    KCI(loY, a.k.loY, 16 * S) KCI(hiY, a.k.hiY, 235 * S) KCI(loC, a.k.loC, 16 * S) KCI(hiC, a.k.hiC, 240 * S)
#undef KCF
#undef KCI
};

// Which frames take the TWO instantiation (two pre-scaled LUT copies in shared memory, chroma without clamp).  The
// frame's codes must fit the copies, and no Cb/Cr the frame can produce may come near matrix_convert's clamp
// (convert.cpp:1207-1213): the LUT is monotone over a clean frame's codes, so its values at code_lo / code_hi bound
// every scaled sample and with them the colour differences.  Evaluated identically by every CTA of both
// instantiations (plain global reads), so the two launches split the batch consistently.
template <int CFG>
__device__ __forceinline__ bool two_lut_frame(const Fwd2Args &a, const FrameK &fk)
{
    typedef KC<CFG> C;
    if (!fk.lut2_ok) return false;
    const float *gl = a.luts + (size_t)fk.lut_slot[0] * 65536;
    const float vlo = __ldg(gl + fk.code_lo), vhi = __ldg(gl + fk.code_hi);
    const float ylo = vlo * C::mulY(a) + C::addY(a), yhi = vhi * C::mulY(a) + C::addY(a);
    const float clo = vlo * C::mulC(a) + C::addC(a), chi = vhi * C::mulC(a) + C::addC(a);
    // extremes of (B - Y')/db and (R - Y')/dr over the box, with Y' = wr R + wg G + wb B + 0.5
    const float wr = C::wr(a), wg = C::wg(a), wb = C::wb(a);
    const float cb_hi = (chi - (wr * clo + wg * ylo + wb * chi)) * C::rdb(a), cb_lo = (clo - (wr * chi + wg * yhi + wb * clo) - 1.0f) * C::rdb(a);
    const float cr_hi = (chi - (wr * chi + wg * ylo + wb * clo)) * C::rdr(a), cr_lo = (clo - (wr * clo + wg * yhi + wb * chi) - 1.0f) * C::rdr(a);
    const float top = (float)(C::maxCV(a) - C::half_m1(a)) - 2.0f, bot = 2.0f - (float)C::half_m1(a);
    return vlo >= 0.0f && cb_hi < top && cr_hi < top && cb_lo > bot && cr_lo > bot && yhi < (float)C::maxCV(a);
}

// The same proof for a three-table frame (FrameK::clean3): each channel has its own table and code range, so the box
// of scaled samples has a side per channel.  The table entry that stands for code 0 (FrameK::zero_entry) is the
// channel's smallest value.  Evaluated identically by every CTA of both three-table instantiations.
template <int CFG>
__device__ __forceinline__ bool three_nc_frame(const Fwd2Args &a, const FrameK &fk)
{
    typedef KC<CFG> C;
    float lo[3], hi[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float *gl = a.luts + (size_t)fk.lut_slot[c] * 65536;
        const float vlo = __ldg(gl + (fk.zero_entry[c] ? 0u : fk.ch_lo[c])), vhi = __ldg(gl + fk.ch_hi[c]);
        if (!(vlo >= 0.0f)) return false;
        const float mul = c == 0 ? C::mulY(a) : C::mulC(a), add = c == 0 ? C::addY(a) : C::addC(a);
        lo[c] = vlo * mul + add; hi[c] = vhi * mul + add;
    }
    // planes 0/1/2 = G/B/R; extremes of (B - Y')/db and (R - Y')/dr over the box, with Y' = wr R + wg G + wb B + 0.5
    const float wr = C::wr(a), wg = C::wg(a), wb = C::wb(a);
    const float cb_hi = (hi[1] - (wr * lo[2] + wg * lo[0] + wb * hi[1])) * C::rdb(a), cb_lo = (lo[1] - (wr * hi[2] + wg * hi[0] + wb * lo[1]) - 1.0f) * C::rdb(a);
    const float cr_hi = (hi[2] - (wr * hi[2] + wg * lo[0] + wb * lo[1])) * C::rdr(a), cr_lo = (lo[2] - (wr * lo[2] + wg * hi[0] + wb * hi[1]) - 1.0f) * C::rdr(a);
    const float top = (float)(C::maxCV(a) - C::half_m1(a)) - 2.0f, bot = 2.0f - (float)C::half_m1(a);
    return cb_hi < top && cr_hi < top && cb_lo > bot && cr_lo > bot && hi[0] < (float)C::maxCV(a);
}

// LUT entry `code` of the table at shared-window address `lut` (+OFS bytes).  An explicit shared-space load: the
// generic form makes ptxas rebuild the window base (S2UR + UMOV + ULEA) in front of every group of gathers.
template <int OFS>
__device__ __forceinline__ float lds_lut(unsigned lut, unsigned code)
{
    float v;
    unsigned addr;
    asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(addr) : "r"(code), "r"(lut));       // one IMAD (FMA pipe), not shift + add
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFS));
    return v;
}

// ---- per-lane: 8 pixels -> luma (floor bits, not yet clamped) and chroma as floats ---------------------------
// TWO:   `lut` holds two pre-scaled copies (luma scale, then chroma scale at +LUT2_CODES); chroma comes without
//        Half-1 and without matrix_convert's clamp (two_lut_frame() has checked that it cannot bind).
// lutB / lutR: table bases of the B and R channels when each channel has its own table (three-table frames); the
// callers with one table pass `lut` three times and the compiler sees one value.
// PRESC (three-table frames): the per-channel tables already carry their range scale (G: luma, B/R: chroma), like TWO's
// two copies; the clamp and the integer chroma path stay unless TWO is set as well (three_nc_frame() has checked it).
template <int MK, int CFG = 0, bool TWO = false, bool PRESC = false>
__device__ __forceinline__ void pixels8(const Fwd2Args &a, unsigned lut, unsigned lutB, unsigned lutR, const unsigned g[8],
                                            const unsigned b[8], const unsigned r[8], unsigned ybits[8], u64 chroma[8])
{
    typedef KC<CFG> C;
    const PixK &k = a.k;
#ifdef H2Y_EXPERIMENT_NO_MATH
    // timing diagnostic only (wrong output): keeps the loads, filters and stores, drops the gathers and the matrix
#pragma unroll
    for (int i = 0; i < 8; i++) {
        ybits[i] = (unsigned)MAGIC_BITS + (g[i] & 1023u);
        chroma[i] = pk((float)(b[i] & 511u), (float)(r[i] & 511u));
    }
    return;
#endif
    const u64 addY2 = pk(C::addY(a), C::addY(a)), addC2 = pk(C::addC(a), C::addC(a));
    const u64 magic2 = pk(MAGIC, MAGIC);
    const u64 twoGy2 = pk(C::twoGy(a), C::twoGy(a)), twoGb2 = pk(C::twoGb(a), C::twoGb(a)), twoGr2 = pk(C::twoGr(a), C::twoGr(a));
    const float wr = C::wr(a), wg = C::wg(a), wb = C::wb(a), rdb = C::rdb(a), rdr = C::rdr(a);
    const u64 wr2 = pk(wr, wr), wg2 = pk(wg, wg), wb2 = pk(wb, wb);
    const u64 rdb2 = pk(rdb, rdb), rdr2 = pk(rdr, rdr);
    const u64 lumc2 = pk(C::lumc(a), C::lumc(a));
    const u64 cbc2 = pk(C::cbc(a), C::cbc(a)), crc2 = pk(C::crc(a), C::crc(a));
    const int cbias = C::half_m1(a) - MAGIC_BITS;
    const unsigned maxCV = (unsigned)C::maxCV(a);
    // PG pixel pairs share one guard-band branch: fewer, larger basic blocks for the scheduler at the price of keeping
    // the scaled samples of PG pairs alive until the branch
    constexpr int PG = PRESC ? (TWO ? H2Y_PAIRS_PER_BRANCH_THREE_NC : H2Y_PAIRS_PER_BRANCH_THREE) : H2Y_PAIRS_PER_BRANCH;
#pragma unroll
    for (int q0 = 0; q0 < 8; q0 += 2 * PG) {
        u64 G2[PG], B2[PG], R2[PG];
        unsigned cbi[2 * PG], cri[2 * PG];
        float tcb[2 * PG], tcr[2 * PG];
        bool flag[2 * PG];
#pragma unroll
        for (int p = 0; p < PG; p++) {
            const int q = q0 + 2 * p;
            // LUT gather + range scale (two rounded operations each, convert.cpp:1141-1143)
            if (PRESC) {
                G2[p] = pk(lds_lut<0>(lut, g[q]), lds_lut<0>(lut, g[q + 1]));
                B2[p] = pk(lds_lut<0>(lutB, b[q]), lds_lut<0>(lutB, b[q + 1]));
                R2[p] = pk(lds_lut<0>(lutR, r[q]), lds_lut<0>(lutR, r[q + 1]));
            } else if (TWO) {
                G2[p] = pk(lds_lut<0>(lut, g[q]), lds_lut<0>(lut, g[q + 1]));
                B2[p] = pk(lds_lut<LUT2_CODES * 4>(lut, b[q]), lds_lut<LUT2_CODES * 4>(lut, b[q + 1]));
                R2[p] = pk(lds_lut<LUT2_CODES * 4>(lut, r[q]), lds_lut<LUT2_CODES * 4>(lut, r[q + 1]));
            } else {
                G2[p] = fadd2(fmul2s(lds_lut<0>(lut, g[q]), lds_lut<0>(lut, g[q + 1]), C::mulY(a)), addY2);
                B2[p] = fadd2(fmul2s(lds_lut<0>(lutB, b[q]), lds_lut<0>(lutB, b[q + 1]), C::mulC(a)), addC2);
                R2[p] = fadd2(fmul2s(lds_lut<0>(lutR, r[q]), lds_lut<0>(lutR, r[q + 1]), C::mulC(a)), addC2);
            }
            u64 y1, y2, base;
            if (MK == MK_YCBCR) {
                const u64 slo = ffma2(wg2, G2[p], ffma2(wr2, R2[p], ffma2(wb2, B2[p], lumc2)));     // luma + 0.5 - G
                y1 = fadd2_rm(slo, magic2);
                y2 = fadd2_rm(fadd2(slo, twoGy2), magic2);
                base = slo;
            } else {                                     // Y'DzDx: Y = (unsigned)G', exact
                y1 = y2 = fadd2_rm(G2[p], magic2);
                base = G2[p];
            }
            const u64 cbl = ffma2(fsub2(B2[p], base), rdb2, cbc2), crl = ffma2(fsub2(R2[p], base), rdr2, crc2);
            const u64 cbh = fadd2(cbl, twoGb2), crh = fadd2(crl, twoGr2);
            const u64 cb1 = fadd2_rm(cbl, magic2), cb2 = fadd2_rm(cbh, magic2);
            const u64 cr1 = fadd2_rm(crl, magic2), cr2 = fadd2_rm(crh, magic2);
            int Y1[2], Y2[2], B1[2], Bq[2], R1[2], Rq[2], xb[2], xr[2];
            unpk(y1, Y1[0], Y1[1]); unpk(y2, Y2[0], Y2[1]);
            unpk(cb1, B1[0], B1[1]); unpk(cb2, Bq[0], Bq[1]);
            unpk(cr1, R1[0], R1[1]); unpk(cr2, Rq[0], Rq[1]);
            unpk(cbl, xb[0], xb[1]); unpk(crl, xr[0], xr[1]);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                ybits[q + e] = (unsigned)Y1[e];
                // An unflagged sample has no integer in (x, x+2G], where x is the lowered value: the reference's value
                // lies strictly inside that window, so it truncates like either end and has the same sign.
                if (TWO) {
                    // FRND.TRUNC of the upper end is the float the filter wants, written straight into the {Cb,Cr}
                    // pair: one XU instruction per sample instead of sign fix + rebias + I2F
                    tcb[2 * p + e] = truncf(e ? phi(cbh) : plo(cbh));
                    tcr[2 * p + e] = truncf(e ? phi(crh) : plo(crh));
                } else {
                    // trunc toward zero = floor + 1 for negative non-integers
                    cbi[2 * p + e] = (unsigned)(B1[e] + cbias + (int)((unsigned)xb[e] >> 31));
                    cri[2 * p + e] = (unsigned)(R1[e] + cbias + (int)((unsigned)xr[e] >> 31));
                }
                flag[2 * p + e] = Y1[e] != Y2[e] || B1[e] != Bq[e] || R1[e] != Rq[e];
            }
        }
        bool any = false;
#pragma unroll
        for (int i = 0; i < 2 * PG; i++) any = any || flag[i];
        if (any) {
            // within the guard band of an integer: take the reference-exact route for those pixels
#pragma unroll
            for (int i = 0; i < 2 * PG; i++)
                if (flag[i]) {
                    // gathered again rather than kept alive across the branch (registers are the scarce resource here)
                    float Gs, Bs, Rs;
                    if (PRESC) {
                        Gs = lds_lut<0>(lut, g[q0 + i]); Bs = lds_lut<0>(lutB, b[q0 + i]); Rs = lds_lut<0>(lutR, r[q0 + i]);
                    } else if (TWO) {
                        Gs = lds_lut<0>(lut, g[q0 + i]); Bs = lds_lut<LUT2_CODES * 4>(lut, b[q0 + i]); Rs = lds_lut<LUT2_CODES * 4>(lut, r[q0 + i]);
                    } else {
                        Gs = __fadd_rn(__fmul_rn(lds_lut<0>(lut, g[q0 + i]), C::mulY(a)), C::addY(a));
                        Bs = __fadd_rn(__fmul_rn(lds_lut<0>(lutB, b[q0 + i]), C::mulC(a)), C::addC(a));
                        Rs = __fadd_rn(__fmul_rn(lds_lut<0>(lutR, r[q0 + i]), C::mulC(a)), C::addC(a));
                    }
                    unsigned Ye, Cbe, Cre;
                    pixel_exact<MK>(Gs, Bs, Rs, k, Ye, Cbe, Cre);
                    ybits[q0 + i] = Ye + (unsigned)MAGIC_BITS;
                    cbi[i] = Cbe; cri[i] = Cre;
                    tcb[i] = (float)((int)Cbe - C::half_m1(a)); tcr[i] = (float)((int)Cre - C::half_m1(a));
                    if (a.fallback_count) atomicAdd(a.fallback_count, 1ull);     // diagnostics only (null in production launches)
                }
        }
#pragma unroll
        for (int i = 0; i < 2 * PG; i++)
            // matrix_convert's clamp through unsigned long: negatives land on maxCV (convert.cpp:1210-1213)
            chroma[q0 + i] = TWO ? pk(tcb[i], tcr[i]) : pk((float)(int)min(cbi[i], maxCV), (float)(int)min(cri[i], maxCV));
    }
}

// packed 16-bit min / max (ptxas fuses two of them into one VIMNMX3.U16x2)
__device__ __forceinline__ unsigned vmin2(unsigned a, unsigned b) { unsigned r; asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned vmax2(unsigned a, unsigned b) { unsigned r; asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// write_yuv on the luma floor bits: >> shift, range clamp, pack (the low 16 bits of the clamped value are the code).
// NOLOW: the caller's frames cannot produce a luma below the range floor, so only the upper clamp is applied.  That is
// every frame two_lut_frame() / three_nc_frame() accept: their scaled samples are >= minVR = minVRC (LUT values >= 0),
// the three weights sum to 1 and the rounding constant adds 0.5, so luma >= minVR + 0.5 - (a few fp32 ulps).
template <int CFG, bool NOLOW = false>
__device__ __forceinline__ uint4 pack_luma(const Fwd2Args &a, const unsigned ybits[8])
{
    typedef KC<CFG> C;
    if (CFG) {
        // shift is 0 and Y < 2^16: pack first, then clamp two codes per instruction
        const unsigned lo2 = (unsigned)C::loY(a) * 0x10001u, hi2 = (unsigned)C::hiY(a) * 0x10001u;
        if (NOLOW)
            return make_uint4(vmin2(__byte_perm(ybits[0], ybits[1], 0x5410), hi2), vmin2(__byte_perm(ybits[2], ybits[3], 0x5410), hi2),
                              vmin2(__byte_perm(ybits[4], ybits[5], 0x5410), hi2), vmin2(__byte_perm(ybits[6], ybits[7], 0x5410), hi2));
        return make_uint4(clamp_u16x2(__byte_perm(ybits[0], ybits[1], 0x5410), lo2, hi2), clamp_u16x2(__byte_perm(ybits[2], ybits[3], 0x5410), lo2, hi2),
                          clamp_u16x2(__byte_perm(ybits[4], ybits[5], 0x5410), lo2, hi2), clamp_u16x2(__byte_perm(ybits[6], ybits[7], 0x5410), lo2, hi2));
    }
    const int shift = C::shift(a);
    const int ylo = C::loY(a) + (MAGIC_BITS >> shift), yhi = C::hiY(a) + (MAGIC_BITS >> shift);
    unsigned yv[8];
#pragma unroll
    for (int i = 0; i < 8; i++) yv[i] = (unsigned)clamp3((int)ybits[i] >> shift, ylo, yhi);
    return make_uint4(__byte_perm(yv[0], yv[1], 0x5410), __byte_perm(yv[2], yv[3], 0x5410),
                      __byte_perm(yv[4], yv[5], 0x5410), __byte_perm(yv[6], yv[7], 0x5410));
}

// horizontal 7-tap at even x on a {Cb,Cr} pair (convert.cpp:290-321); exact integer arithmetic in fp32
// `c0` is the rounding 0.5, plus Half-1 when the caller's samples come without it (the taps sum to exactly 1)
__device__ __forceinline__ u64 fir_h7_pair(u64 m5, u64 m3, u64 m1, u64 c, u64 p1, u64 p3, u64 p5, int hi_bits, float c0 = 0.5f)
{
    const u64 k21 = pk(21.0f / 512.0f, 21.0f / 512.0f), k52n = pk(-52.0f / 512.0f, -52.0f / 512.0f),
              k159 = pk(159.0f / 512.0f, 159.0f / 512.0f), k256 = pk(0.5f, 0.5f), half2v = pk(c0, c0);
    u64 t = ffma2(k21, m5, half2v);
    t = ffma2(k21, p5, t);
    t = ffma2(k52n, m3, t);
    t = ffma2(k52n, p3, t);
    t = ffma2(k159, m1, t);
    t = ffma2(k159, p1, t);
    t = ffma2(k256, c, t);
    // clamp to [0, maxCV] and truncate into the u16 intermediate: floor first, clamp the integers
    const u64 fl = fadd2_rm(t, pk(MAGIC, MAGIC));
    const int a = clamp3(ilo(fl), MAGIC_BITS, hi_bits), b = clamp3(ihi(fl), MAGIC_BITS, hi_bits);
    return fadd2(pk(__int_as_float(a), __int_as_float(b)), pk(-MAGIC, -MAGIC));
}

// ---- integer-source route (TIFF rows, tmp depth 16): reference arithmetic, no guard band --------------------
// At 16-bit scale fp32 has too few fraction bits for the guard-band trick and the FIR's float rounding is
// observable, so this route keeps the general kernel's FP64 colour difference
// (reciprocal multiply with 14 guard bits, exact division when unsure) and the reference's float operation order
// in both filters; it shares the ring / register-blocked vertical stage of k_forward_exr420.
// write_yuv's shift mask and packed luma range for pixels8_u16 (a caller may keep them in registers across its loop)
struct LumaPack { unsigned keep, lo2, hi2; };
__device__ __forceinline__ LumaPack luma_pack(const PixK &k)
{
    LumaPack p;
    p.keep = (0xffffu >> k.down_shift) * 0x10001u; p.lo2 = k.loY * 0x10001u; p.hi2 = k.hiY * 0x10001u;
    return p;
}
template <int MK, int FAM = 0>
__device__ __forceinline__ void pixels8_u16(const PixK &k, const LumaPack &lp, const unsigned g[8], const unsigned b[8], const unsigned r[8],
                                            uint4 &ypack, u64 chroma[8], unsigned &fallbacks)
{
    unsigned yv[8];
    // four pixels share one fallback branch (a pixel needs the exact route about once in 4000)
#pragma unroll
    for (int q0 = 0; q0 < 8; q0 += 4) {
        unsigned Y[4], Cb[4], Cr[4];
        bool ok[4];
#pragma unroll
        for (int i = 0; i < 4; i++) ok[i] = px_matrix_fast_u16<MK, FAM>(g[q0 + i], b[q0 + i], r[q0 + i], k, Y[i], Cb[i], Cr[i]);   // codes already clipped on load
        if (!(ok[0] && ok[1] && ok[2] && ok[3])) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (!ok[i]) {
                    px_matrix_exact<MK>((float)g[q0 + i], (float)b[q0 + i], (float)r[q0 + i], k, Y[i], Cb[i], Cr[i]);
                    fallbacks++;
                }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            yv[q0 + i] = Y[i];
            // u2f on both planes at once: 2^23 + code as a bit pattern, minus 2^23 (exact)
            chroma[q0 + i] = fadd2(pk(__uint_as_float(0x4B000000u | Cb[i]), __uint_as_float(0x4B000000u | Cr[i])), pk(-8388608.0f, -8388608.0f));
        }
    }
    // write_yuv on luma: >> shift and the range clamp on two packed codes per instruction (Y <= maxCV < 2^16)
    const unsigned keep = lp.keep, lo2 = lp.lo2, hi2 = lp.hi2;
    unsigned wv[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        wv[i] = clamp_u16x2((__byte_perm(yv[2 * i], yv[2 * i + 1], 0x5410) >> k.down_shift) & keep, lo2, hi2);
    ypack = make_uint4(wv[0], wv[1], wv[2], wv[3]);
}

// horizontal 7-tap in the reference's operation order (convert.cpp:305-317), both planes of a {Cb,Cr} pair
__device__ __forceinline__ u64 fir_h7_pair_ref(u64 m5, u64 m3, u64 m1, u64 c, u64 p1, u64 p3, u64 p5, float maxCVf)
{
    return pk(fir_h7_f(plo(m5), plo(m3), plo(m1), plo(c), plo(p1), plo(p3), plo(p5), maxCVf),
              fir_h7_f(phi(m5), phi(m3), phi(m1), phi(c), phi(p1), phi(p3), phi(p5), maxCVf));
}

// The same two filters for the warp-autonomous integer kernel: reference operation order on both planes of a {Cb,Cr}
// pair at once.  The ORDER of the additions is what the reference fixes (every partial sum is rounded); a product may be
// fused into the addition that consumes it whenever the product itself is exact, because then fma(k, s, t) = RN(t + k*s)
// is the reference's RN(t + RN(k*s)).  The taps are n/512 and the samples are integers <= 65535 (sums of two <= 131070),
// so k*s = n*s/512 is exact when n*s < 2^24: true for n = 5, 11, 21, 37, 52, 70 (70 * 131070 = 9.2 M) and for 256 (one
// sample, a power of two), false for 159 and 228, whose products stay separately rounded multiplications.  (ptxas
// contracts a packed multiply feeding a packed add into FFMA2 by itself, see h2y_f32x2.cuh, so those two products are
// scalar FMULs; the 228 product is the first term and feeds an FFMA2 as its addend, which cannot be contracted.)
// The reference clamps the float to [0, maxCV] and truncates; flooring first (round-down add of 1.5*2^23) and
// clamping the integers gives the same code for every t, and leaves the bit pattern the callers want.
__device__ __forceinline__ u64 fmulk2(float kf, u64 v) { return pk(__fmul_rn(kf, plo(v)), __fmul_rn(kf, phi(v))); }
__device__ __forceinline__ u64 pk1(float v) { return pk(v, v); }

// convert.cpp:305-317; returns the clamped, truncated samples as floats
__device__ __forceinline__ u64 fir_h7_pair_ord(u64 m5, u64 m3, u64 m1, u64 c, u64 p1, u64 p3, u64 p5, int hi_bits)
{
    u64 t = fmulk2(21.0f / 512.0f, fadd2(m5, p5));                      // exact
    t = ffma2(pk1(-52.0f / 512.0f), fadd2(m3, p3), t);                  // RN(t - c52 * s), product exact
    t = fadd2(t, fmulk2(159.0f / 512.0f, fadd2(m1, p1)));               // product rounded on its own
    t = ffma2(pk1(256.0f / 512.0f), c, t);                              // product exact
    t = fadd2(t, pk(0.5f, 0.5f));
    const u64 fl = fadd2_rm(t, pk(MAGIC, MAGIC));
    const int a = clamp3(ilo(fl), MAGIC_BITS, hi_bits), b = clamp3(ihi(fl), MAGIC_BITS, hi_bits);
    return fadd2(pk(__int_as_float(a), __int_as_float(b)), pk(-MAGIC, -MAGIC));
}

// convert.cpp:365-374, r[0..11] = rows y-5 .. y+6; returns the floor as MAGIC_BITS + n in both halves (not yet clamped)
__device__ __forceinline__ u64 fir_v12_pair_ord(const u64 r[12])
{
    u64 t = fmulk2(228.0f / 512.0f, fadd2(r[5], r[6]));                 // rounded on its own (first term)
    t = ffma2(pk1(70.0f / 512.0f), fadd2(r[4], r[7]), t);               // products exact from here on
    t = ffma2(pk1(-37.0f / 512.0f), fadd2(r[3], r[8]), t);
    t = ffma2(pk1(-21.0f / 512.0f), fadd2(r[2], r[9]), t);
    t = ffma2(pk1(11.0f / 512.0f), fadd2(r[1], r[10]), t);
    t = ffma2(pk1(5.0f / 512.0f), fadd2(r[0], r[11]), t);
    t = fadd2(t, pk(0.5f, 0.5f));
    return fadd2_rm(t, pk(MAGIC, MAGIC));
}

template <int MK, int NCH, int SRC = 0>     // SRC: 0 = half source through the LUT (fp32 guard band), 1 = integer source (reference arithmetic)
__global__ void __launch_bounds__(THREADS, 1) k_forward_exr420(const Fwd2Args a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);                    // [RING_ROWS][RING_PITCH]
    float *lut_s = ring + RING_ROWS * RING_PITCH;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PixK &k = a.k;
    const int w = a.w, h = a.h, wh = w >> 1;
    const int hi_bits = MAGIC_BITS + (int)k.maxCV;
    const int shift = k.down_shift;
    const int clo = (int)k.loC + (MAGIC_BITS >> shift), chi = (int)k.hiC + (MAGIC_BITS >> shift);
    int cur_slot = -1;
    unsigned cur_lo = 1, cur_hi = 0;
    unsigned fallbacks = 0;

    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int strip = item % a.nstrips;
        const int seg = (item / a.nstrips) % a.nsegs;
        const int frame = item / (a.nstrips * a.nsegs);
        if (SRC == 0 && !a.framek[frame].clean) continue;      // v1 converts this frame (uniform per CTA)
        const uint8_t *fsrc = a.src + (size_t)frame * a.src_stride;
        uint16_t *fY = reinterpret_cast<uint16_t *>(a.dst + (size_t)frame * a.dst_stride);
        uint16_t *fCb = fY + (size_t)w * h;
        uint16_t *fCr = fCb + (size_t)wh * (h >> 1);

        // ---- shared-memory LUT for the frame's code range ----
        __syncthreads();                                // previous item's readers (ring and LUT) are done
        if (SRC == 0) {
            const FrameK &fk = a.framek[frame];
            const unsigned lo = fk.code_lo, hi = fk.code_hi;
            if (fk.lut_slot[0] != cur_slot || lo < cur_lo || hi > cur_hi) {
                const float *gl = a.luts + (size_t)fk.lut_slot[0] * 65536;
                for (unsigned c = lo + threadIdx.x; c <= hi; c += THREADS) lut_s[c] = __ldg(gl + c);
                cur_slot = fk.lut_slot[0]; cur_lo = lo; cur_hi = hi;
                __syncthreads();
            }
        }
        const unsigned lut_sa = (unsigned)__cvta_generic_to_shared(lut_s);   // indexed by the raw code; entries cur_lo..cur_hi are valid

        const int x0 = strip * a.strip_w;
        const int ys = seg * a.seg_rows, ye = min(ys + a.seg_rows, h);
        const int xl = x0 + 8 * (lane - 1);
        const bool lane_in_pic = xl >= 0 && xl < w;
        const bool lane_interior = lane >= 1 && lane < 31 && xl < min(x0 + a.strip_w, w);
        const int r0 = ys - 6;
        const int nsteps = (ye - ys + 12 + 15) / 16;

        RawPx<NCH> raw;
        {
            const int row = r0 + warp;
            if (row >= 0 && row < h && lane_in_pic) load_px8<NCH>(raw, fsrc, w, row, xl);
        }
        for (int s = 0; s < nsteps; s++) {
            const int row = r0 + 16 * s + warp;
            const bool row_ok = row >= 0 && row < h && row < ye + 6;
            u64 ch[8];
#pragma unroll
            for (int q = 0; q < 8; q++) ch[q] = 0ull;
            if (row_ok && lane_in_pic) {
                unsigned g[8], b[8], r[8];
                if (SRC == 1 && k.clip_on_load) {               // read_tiff's clip (tiff.cpp:296-304), two samples per instruction
                    const unsigned lo2 = k.loadLo * 0x10001u, hi2 = k.loadHi * 0x10001u;
#pragma unroll
                    for (int i = 0; i < NCH; i++) {
                        unsigned *wv = reinterpret_cast<unsigned *>(&raw.v[i]);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            asm("max.u16x2 %0, %0, %1;" : "+r"(wv[j]) : "r"(lo2));
                            asm("min.u16x2 %0, %0, %1;" : "+r"(wv[j]) : "r"(hi2));
                        }
                    }
                }
                split_codes<NCH>(raw, g, b, r);
                uint4 ypack;
                if (SRC == 0) { unsigned yb[8]; pixels8<MK>(a, lut_sa, lut_sa, lut_sa, g, b, r, yb, ch); ypack = pack_luma<0>(a, yb); }
                else pixels8_u16<MK>(a.k, luma_pack(a.k), g, b, r, ypack, ch, fallbacks);
                if (lane_interior && row >= ys && row < ye) *reinterpret_cast<uint4 *>(fY + (size_t)row * w + xl) = ypack;
            }
            {   // prefetch the next step's row
                const int nrow = row + 16;
                if (s + 1 < nsteps && nrow >= 0 && nrow < h && nrow < ye + 6 && lane_in_pic) load_px8<NCH>(raw, fsrc, w, nrow, xl);
            }
            if (row_ok) {   // warp-uniform: horizontal filter, neighbours through shuffles
                float l3x = __shfl_up_sync(0xffffffffu, plo(ch[3]), 1), l3y = __shfl_up_sync(0xffffffffu, phi(ch[3]), 1);
                float l5x = __shfl_up_sync(0xffffffffu, plo(ch[5]), 1), l5y = __shfl_up_sync(0xffffffffu, phi(ch[5]), 1);
                float l7x = __shfl_up_sync(0xffffffffu, plo(ch[7]), 1), l7y = __shfl_up_sync(0xffffffffu, phi(ch[7]), 1);
                float n1x = __shfl_down_sync(0xffffffffu, plo(ch[1]), 1), n1y = __shfl_down_sync(0xffffffffu, phi(ch[1]), 1);
                float n3x = __shfl_down_sync(0xffffffffu, plo(ch[3]), 1), n3y = __shfl_down_sync(0xffffffffu, phi(ch[3]), 1);
                u64 l3 = pk(l3x, l3y), l5 = pk(l5x, l5y), l7 = pk(l7x, l7y), n1 = pk(n1x, n1y), n3 = pk(n3x, n3y);
                if (xl == 0) l3 = l5 = l7 = ch[0];                  // replicate s[0]     (convert.cpp:295-300)
                if (xl + 8 >= w) n1 = n3 = ch[7];                   // replicate s[W-1]
                if (lane_interior) {
                    u64 o0, o1, o2, o3;
                    if (SRC == 0) {
                        o0 = fir_h7_pair(l3, l5, l7, ch[0], ch[1], ch[3], ch[5], hi_bits);
                        o1 = fir_h7_pair(l5, l7, ch[1], ch[2], ch[3], ch[5], ch[7], hi_bits);
                        o2 = fir_h7_pair(l7, ch[1], ch[3], ch[4], ch[5], ch[7], n1, hi_bits);
                        o3 = fir_h7_pair(ch[1], ch[3], ch[5], ch[6], ch[7], n1, n3, hi_bits);
                    } else {
                        const float mf = (float)k.maxCV;
                        o0 = fir_h7_pair_ref(l3, l5, l7, ch[0], ch[1], ch[3], ch[5], mf);
                        o1 = fir_h7_pair_ref(l5, l7, ch[1], ch[2], ch[3], ch[5], ch[7], mf);
                        o2 = fir_h7_pair_ref(l7, ch[1], ch[3], ch[4], ch[5], ch[7], n1, mf);
                        o3 = fir_h7_pair_ref(ch[1], ch[3], ch[5], ch[6], ch[7], n1, n3, mf);
                    }
                    float *rr = ring + (size_t)((row + RING_ROWS) & (RING_ROWS - 1)) * RING_PITCH + (lane - 1) * 4;
                    *reinterpret_cast<float4 *>(rr) = make_float4(plo(o0), phi(o0), plo(o1), phi(o1));
                    *reinterpret_cast<float4 *>(rr + RING_COLS) = make_float4(plo(o2), phi(o2), plo(o3), phi(o3));
                }
            }
            __syncthreads();
            // ---- vertical 12-tap (convert.cpp:333-377): half the CTA, alternating, 4 outputs per thread ----
            const int vt = (int)threadIdx.x - ((s & 1) ? THREADS / 2 : 0);
            if (vt >= 0 && vt < 2 * RING_COLS) {
                const int jg = vt / RING_COLS, c = vt - jg * RING_COLS;
                const int j0 = (ys >> 1) + 8 * s - 6 + 4 * jg;       // first of this thread's 4 output rows
                const int col = (x0 >> 1) + c;
                if (j0 + 3 >= (ys >> 1) && j0 < (ye >> 1) && col < min((x0 + a.strip_w) >> 1, wh)) {
                    const u64 kv[12] = {pk(5.0f / 512.0f, 5.0f / 512.0f), pk(11.0f / 512.0f, 11.0f / 512.0f),
                                        pk(-21.0f / 512.0f, -21.0f / 512.0f), pk(-37.0f / 512.0f, -37.0f / 512.0f),
                                        pk(70.0f / 512.0f, 70.0f / 512.0f), pk(228.0f / 512.0f, 228.0f / 512.0f),
                                        pk(228.0f / 512.0f, 228.0f / 512.0f), pk(70.0f / 512.0f, 70.0f / 512.0f),
                                        pk(-37.0f / 512.0f, -37.0f / 512.0f), pk(-21.0f / 512.0f, -21.0f / 512.0f),
                                        pk(11.0f / 512.0f, 11.0f / 512.0f), pk(5.0f / 512.0f, 5.0f / 512.0f)};
                    u64 acc[4];
#pragma unroll
                    for (int o = 0; o < 4; o++) acc[o] = pk(0.5f, 0.5f);
                    const float *rp = ring + 2 * ring_pos(c);
                    const int rfirst = 2 * j0 - 5;
                    const bool linear = rfirst >= 0 && rfirst + 17 <= h - 1 && ((rfirst & (RING_ROWS - 1)) + 17 < RING_ROWS);
                    if (SRC == 1) {
                        // reference operation order per output (convert.cpp:365-374): the 18 rows are loaded once
                        u64 rows[18];
#pragma unroll
                        for (int rr = 0; rr < 18; rr++) {
                            int rowi = rfirst + rr;
                            rowi = rowi < 0 ? 0 : (rowi > h - 1 ? h - 1 : rowi);
                            const float2 v = *reinterpret_cast<const float2 *>(rp + (size_t)(rowi & (RING_ROWS - 1)) * RING_PITCH);
                            rows[rr] = pk(v.x, v.y);
                        }
                        const float mf = (float)k.maxCV;
#pragma unroll
                        for (int o = 0; o < 4; o++) {
                            float rx[12], ry[12];
#pragma unroll
                            for (int t = 0; t < 12; t++) { rx[t] = plo(rows[2 * o + t]); ry[t] = phi(rows[2 * o + t]); }
                            // keep the results in the magic-number form the common epilogue expects
                            acc[o] = pk(__int_as_float(fir_v12_magic(rx, mf)), __int_as_float(fir_v12_magic(ry, mf)));
                        }
                    } else if (linear) {
                        // interior, no ring wrap: 18 loads at compile-time offsets
                        const float *base = rp + (size_t)(rfirst & (RING_ROWS - 1)) * RING_PITCH;
#pragma unroll
                        for (int rr = 0; rr < 18; rr++) {
                            const float2 v = *reinterpret_cast<const float2 *>(base + rr * RING_PITCH);
                            const u64 pv = pk(v.x, v.y);
#pragma unroll
                            for (int o = 0; o < 4; o++) {
                                const int tap = rr - 2 * o;
                                if (tap >= 0 && tap < 12) acc[o] = ffma2(kv[tap], pv, acc[o]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int rr = 0; rr < 18; rr++) {
                            int rowi = rfirst + rr;
                            rowi = rowi < 0 ? 0 : (rowi > h - 1 ? h - 1 : rowi);     // replicate (convert.cpp:337-347)
                            const float2 v = *reinterpret_cast<const float2 *>(rp + (size_t)(rowi & (RING_ROWS - 1)) * RING_PITCH);
                            const u64 pv = pk(v.x, v.y);
#pragma unroll
                            for (int o = 0; o < 4; o++) {
                                const int tap = rr - 2 * o;
                                if (tap >= 0 && tap < 12) acc[o] = ffma2(kv[tap], pv, acc[o]);
                            }
                        }
                    }
#pragma unroll
                    for (int o = 0; o < 4; o++) {
                        const int j = j0 + o;
                        if (j >= (ys >> 1) && j < (ye >> 1)) {
                            // clamp [0,maxCV] + truncation + write_yuv's shift and range clamp collapse to one
                            // integer clamp of the floored value (all bounds are integers, the map is monotone)
                            const u64 fl = SRC == 1 ? acc[o] : fadd2_rm(acc[o], pk(MAGIC, MAGIC));
                            const size_t off = (size_t)j * wh + col;
                            fCb[off] = (uint16_t)clamp3(ilo(fl) >> shift, clo, chi);
                            fCr[off] = (uint16_t)clamp3(ihi(fl) >> shift, clo, chi);
                        }
                    }
                }
            }
        }
    }
    if (a.fallback_count && fallbacks) atomicAdd(a.fallback_count, (unsigned long long)fallbacks);
}

// =================================================================================================
// v3: warp-autonomous variant for large batches.  A warp owns a column strip and a long run of rows and
// keeps the vertical filter in REGISTERS: the 12-tap filter at even rows is evaluated as six running
// {Cb,Cr} accumulators per chroma column (an output row starts every second input row and lives for
// twelve).  Each new horizontally filtered row is folded into the six accumulators with FFMA2s whose
// destination is the neighbouring accumulator on odd rows, so the rotation costs no moves.  There is no
// shared-memory ring, no CTA barrier inside a frame and no vertical-filter phase; shared memory holds
// only the LUT.  Rows are split statically: worker k gets rows [k*R/K, (k+1)*R/K) of the batch's R rows,
// so every warp does the same amount of work and the 11 halo rows are paid once per ~900 rows.
struct Fwd3Args {
    Fwd2Args b;
    int sub;                 // row workers per CTA (16 / strips when a picture has fewer than 16 strips)
    int wps;                 // warps per row worker = strips handled side by side
    int split_by_lut2;       // two instantiations share the frames: TWO takes those with lut2_ok, the other the rest
    int split_three_nc;      // likewise for three-table frames: TWO (clamp-free) takes those three_nc_frame() accepts
    long total_rows;         // nframes * h
    // plan reuse (h2y_internal.h, SpecSeed / SpecCtl), SPEC instantiations only: every frame is converted with the one
    // predicted FrameK at b.framek[0] while the warp gathers the frame's extrema into `slots`; `bail[frame]` is raised
    // when a code falls outside the predicted LUT window (before it is used as an index).  The frames the verify step
    // hands back are converted by the general kernel (h2y_forward.cu).
    const SpecCtl *ctl;
    int *bail;
    unsigned *slots;
};

constexpr int THREADS3 = 512, WARPS3 = THREADS3 / 32;
// Rows per row worker from which h2y_forward takes the rows kernels by itself.  A worker pays 11 halo rows per run, so
// small batches are where the CTA-ring kernel used to win; measured again at the end of round 2 (profiles/r02/variants.md):
// 4K EXR, two frames (29 rows per worker) 0.128 ms against 0.162 ms, one frame (14) 0.105 against 0.110; 1080p TIFF,
// 16 frames (58) 0.130 against 0.192, 4 frames (14) 0.076 against 0.070.
constexpr long ROWS_MIN_PER_WORKER = 24;
// SPEC instantiations: the warps' code bounds sit behind the LUT copy (24 bytes per warp)
constexpr unsigned SPEC_BOUNDS_BYTES = 24 * 16;
template <bool TWO> __device__ __host__ constexpr unsigned spec_bounds_offset() { return (TWO ? 2u * LUT2_CODES : (unsigned)LUT_MAX_CODES) * 4u; }
     // 16 warps (12 x 168 registers measured 6 % slower: latency hiding wins)

// THREE: the instantiation for frames whose channels need a table each (FrameK::clean3); with TWO as well: the
// clamp-free chroma path on those three tables (frames that three_nc_frame() accepts)

// maximum / minimum over the warp of both 16-bit halves (REDUX: the result is warp-uniform)
__device__ __forceinline__ unsigned warp_max_u16x2(unsigned v)
{
    return __reduce_max_sync(0xffffffffu, v & 0xffffu) | (__reduce_max_sync(0xffffffffu, v >> 16) << 16);
}
__device__ __forceinline__ unsigned warp_min_u16x2(unsigned v)
{
    return __reduce_min_sync(0xffffffffu, v & 0xffffu) | (__reduce_min_sync(0xffffffffu, v >> 16) << 16);
}

// one {max, min} pair of the warp's code bounds, re-read from shared memory every row (volatile: not kept in registers)
__device__ __forceinline__ void lds_bounds(unsigned addr, unsigned &mx, unsigned &mn)
{
    asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(mx), "=r"(mn) : "r"(addr));
}

// order-preserving key of a finite non-negative half code, as k_stats_vec / k_plan use it (h2y_stats.cu: fkey)
__device__ __forceinline__ unsigned half_code_key(unsigned code) { return __float_as_uint(half_bits_to_float(code)) | 0x80000000u; }

template <int MK, int NCH, int CFG, bool TWO, bool THREE, bool SPEC = false>
__device__ __forceinline__ void rows_body(const Fwd3Args &A)
{
    if (SPEC) { if (A.ctl->skip) return; }                       // no usable seed: every frame is handed back
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *lut_s = reinterpret_cast<float *>(smem_raw);
    // shared-window address of the LUT, held in an ordinary register (the asm hides that it is uniform: as a uniform
    // value ptxas rebuilds it from SR_CgaCtaId in front of every group of gathers)
    unsigned lut_sa;
    asm volatile("mov.u32 %0, %1;" : "=r"(lut_sa) : "r"((unsigned)__cvta_generic_to_shared(lut_s)));
    const Fwd2Args &a = A.b;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PixK &k = a.k;
    const int w = a.w, h = a.h, wh = w >> 1;
    typedef KC<CFG> C;
    const int hi_bits = MAGIC_BITS + C::maxCV(a);
    const int shift = C::shift(a);
    const int clo = C::loC(a) + (MAGIC_BITS >> shift), chi = C::hiC(a) + (MAGIC_BITS >> shift);
    const float hc0 = TWO ? (float)C::half_m1(a) + 0.5f : 0.5f;     // pixels8<TWO> leaves Half-1 to the filter

    // this warp's worker and strip set
    const int wk = warp / A.wps;                         // worker within the CTA
    const int sfirst = warp - wk * A.wps;                // first strip; further strips every wps
    const long K = (long)gridDim.x * A.sub;
    const long kid = (long)blockIdx.x * A.sub + (wk < A.sub ? wk : 0);
    const bool active = wk < A.sub;
    // row ranges (even boundaries); the CTA's union range decides which frames it walks
    const long g0 = ((kid * A.total_rows) / K) & ~1L, g1 = kid + 1 == K ? A.total_rows : (((kid + 1) * A.total_rows) / K) & ~1L;
    const long c0 = ((((long)blockIdx.x * A.sub) * A.total_rows) / K) & ~1L;
    const long c1 = (long)(blockIdx.x + 1) * A.sub == K ? A.total_rows : (((((long)blockIdx.x + 1) * A.sub) * A.total_rows) / K) & ~1L;
    if (c1 <= c0) return;
    const int f_first = (int)(c0 / h), f_last = (int)((c1 - 1) / h);
    int cur_slot = -1;
    unsigned cur_lo = 1, cur_hi = 0;
    unsigned lut3G = lut_sa, lut3B = lut_sa, lut3R = lut_sa;     // per-channel table bases (THREE); one table otherwise
    unsigned floor_rg = 0, floor_br = 0, floor_gb = 0;           // THREE: codes are raised to the table's first entry

    const u64 kv[12] = {pk(5.0f / 512.0f, 5.0f / 512.0f), pk(11.0f / 512.0f, 11.0f / 512.0f),
                        pk(-21.0f / 512.0f, -21.0f / 512.0f), pk(-37.0f / 512.0f, -37.0f / 512.0f),
                        pk(70.0f / 512.0f, 70.0f / 512.0f), pk(228.0f / 512.0f, 228.0f / 512.0f),
                        pk(228.0f / 512.0f, 228.0f / 512.0f), pk(70.0f / 512.0f, 70.0f / 512.0f),
                        pk(-37.0f / 512.0f, -37.0f / 512.0f), pk(-21.0f / 512.0f, -21.0f / 512.0f),
                        pk(11.0f / 512.0f, 11.0f / 512.0f), pk(5.0f / 512.0f, 5.0f / 512.0f)};

    for (int frame = f_first; frame <= f_last; frame++) {
        const FrameK &fk = a.framek[SPEC ? 0 : frame];
        if (THREE ? !fk.clean3 : !fk.clean) continue;    // uniform per CTA: another launch converts this frame
        if (!THREE && A.split_by_lut2 && TWO != two_lut_frame<CFG>(a, fk)) continue;   // the other instantiation converts this frame
        if (THREE && A.split_three_nc && TWO != three_nc_frame<CFG>(a, fk)) continue;
        // ---- LUT for this frame (CTA-wide) ----
        if (THREE) {
            // three tables back to back, channel c holding codes ch_lo[c] .. ch_hi[c]; the base handed to the gathers is
            // moved down by ch_lo[c] entries so that they index with the raw code
            if (frame != cur_slot) {
                __syncthreads();
                unsigned off = 0;
                for (int ch3 = 0; ch3 < 3; ch3++) {
                    const float *gl = a.luts + (size_t)fk.lut_slot[ch3] * 65536;
                    const unsigned lo = fk.ch_lo[ch3], hi = fk.ch_hi[ch3];
                    // with the range scale applied exactly as convert.cpp:1141-1143 rounds it (G: luma, B and R: chroma)
                    const float mul = ch3 == 0 ? C::mulY(a) : C::mulC(a), add = ch3 == 0 ? C::addY(a) : C::addC(a);
                    for (unsigned c = lo + threadIdx.x; c <= hi; c += THREADS3)
                        lut_s[off + c - lo] = __fadd_rn(__fmul_rn(__ldg(gl + (c == lo && fk.zero_entry[ch3] ? 0u : c)), mul), add);
                    const unsigned base = lut_sa + 4u * off - 4u * lo;
                    if (ch3 == 0) lut3G = base; else if (ch3 == 1) lut3B = base; else lut3R = base;
                    off += hi - lo + 1;
                }
                // packed floors for the three word kinds of an RGB row, (R,G) (B,R) (G,B); RGBA rows have (R,G) (B,A)
                floor_rg = fk.ch_lo[2] | (fk.ch_lo[0] << 16);
                floor_br = fk.ch_lo[1] | (NCH == 3 ? fk.ch_lo[2] << 16 : 0u);
                floor_gb = fk.ch_lo[0] | (fk.ch_lo[1] << 16);
                cur_slot = frame;
                __syncthreads();
            }
        } else {
            const unsigned lo = fk.code_lo, hi = fk.code_hi;
            if (fk.lut_slot[0] != cur_slot || lo < cur_lo || hi > cur_hi) {
                __syncthreads();                         // every warp is done with the previous LUT
                const float *gl = a.luts + (size_t)fk.lut_slot[0] * 65536;
                if (TWO) {
                    // two copies with the range scale already applied, exactly as convert.cpp:1141-1143 rounds it
                    for (unsigned c = lo + threadIdx.x; c <= hi; c += THREADS3) {
                        const float v = __ldg(gl + c);
                        lut_s[c] = __fadd_rn(__fmul_rn(v, C::mulY(a)), C::addY(a));
                        lut_s[LUT2_CODES + c] = __fadd_rn(__fmul_rn(v, C::mulC(a)), C::addC(a));
                    }
                } else
                for (unsigned c = lo + threadIdx.x; c <= hi; c += THREADS3) lut_s[c] = __ldg(gl + c);
                cur_slot = fk.lut_slot[0]; cur_lo = lo; cur_hi = hi;
                __syncthreads();
            }
        }
        if (!active) continue;
        const long fbase = (long)frame * h;
        const int ys = (int)(max(g0, fbase) - fbase), ye = (int)(min(g1, fbase + h) - fbase);
        if (ys >= ye) continue;
        const uint8_t *fsrc = a.src + (size_t)frame * a.src_stride;
        uint16_t *fY = reinterpret_cast<uint16_t *>(a.dst + (size_t)frame * a.dst_stride);
        uint16_t *fCb = fY + (size_t)w * h;
        uint16_t *fCr = fCb + (size_t)wh * (h >> 1);

        for (int strip = sfirst; strip < a.nstrips; strip += A.wps) {
            const int x0 = strip * a.strip_w;
            const int xl = x0 + 8 * (lane - 1);
            const bool lane_in_pic = xl >= 0 && xl < w;
            const bool lane_interior = lane >= 1 && lane < 31 && xl < min(x0 + a.strip_w, w);
            const int xload = lane_in_pic ? xl : (xl < 0 ? 0 : w - 8);          // halo lanes outside the picture read a valid
                                                                                 // address; their values are replaced below
            const bool left_edge = xl == 0, right_edge = xl + 8 >= w;
            const bool strip_edge = __any_sync(0xffffffffu, left_edge || right_edge);
            u64 acc[6][4];
#pragma unroll
            for (int i = 0; i < 6; i++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[i][c] = 0ull;
            // SPEC: extrema of the raw codes seen so far by the WARP, per word kind: (R,G) (B,R) (G,B) for RGB rows, (R,G)
            // (B,A) for RGBA, as {max, min} pairs in 24 bytes of shared memory behind the LUT (spec_bounds).  A lane-private
            // running min / max would need six more registers in a loop that has none to spare, and with 232 KB of the SM
            // carved out as shared memory the spills that follow go to L2 (measured: 1.10 -> 1.69 ms).  The bounds are read
            // back every row (three broadcast loads) and only written when a row moves them.
            volatile uint2 *sb = reinterpret_cast<volatile uint2 *>(smem_raw + spec_bounds_offset<TWO>()) + 3 * warp;
            const unsigned sb_sa = lut_sa + spec_bounds_offset<TWO>() + 24u * (unsigned)warp;
            if (SPEC) {
                if (lane < 3) { sb[lane].x = 0u; sb[lane].y = 0xFFFFFFFFu; }
                __syncwarp();
            }
            // no code may index past the LUT copy in shared memory; codes inside it but outside the predicted window read
            // stale entries, which the verify step catches through the gathered extrema (the frame is converted again)
            constexpr unsigned SPEC_CAP2 = ((TWO ? LUT2_CODES : (unsigned)LUT_MAX_CODES) - 1u) * 0x10001u;

            const int rfirst = ys - 6, rlast = ye + 4;                          // rows feeding outputs ys/2 .. ye/2-1 (both even)
            // running pointers: source row (clamped = edge replicate), Y row, chroma output row
            const unsigned spitch = (unsigned)w * (2 * NCH);
            const uint8_t *sp = fsrc + (size_t)min(max(rfirst, 0), h - 1) * spitch + (size_t)xload * (2 * NCH);
            uint16_t *yp = fY + (ptrdiff_t)rfirst * w + xl;
            uint16_t *cbp = fCb + ((ptrdiff_t)(rfirst >> 1) - 3) * wh + (xl >> 1);
            const int crd = (int)(fCr - fCb);                                    // elements; a plane is far below 2^31

            // one row: 8 pixels -> Y store, chroma, horizontal 7-tap -> o[4]
            // advance the source pointer from (clamped) row r to (clamped) row r+1
            auto next_src = [&](int r) { sp += ((unsigned)r < (unsigned)(h - 1)) ? spitch : 0; };
            auto row_split = [&](const RawPx<NCH> &raw, unsigned g[8], unsigned b[8], unsigned rr[8]) {
                if (THREE) {
                    // raise every code to its table's first entry (only exact zeros move: FrameK::zero_entry)
                    RawPx<NCH> cl = raw;
#pragma unroll
                    for (int i = 0; i < NCH; i++) {
                        unsigned *wv = reinterpret_cast<unsigned *>(&cl.v[i]);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int kind = NCH == 3 ? (4 * i + j) % 3 : (j & 1);
                            const unsigned fl = NCH == 3 ? (kind == 0 ? floor_rg : (kind == 1 ? floor_br : floor_gb)) : (kind == 0 ? floor_rg : floor_br);
                            asm("max.u16x2 %0, %0, %1;" : "+r"(wv[j]) : "r"(fl));
                        }
                    }
                    split_codes<NCH>(cl, g, b, rr);
                } else
                split_codes<NCH>(raw, g, b, rr);
            };
            auto row_front = [&](const unsigned g[8], const unsigned b[8], const unsigned rr[8], int r, u64 o[4]) {
                unsigned yb[8];
                u64 ch[8];
                if (THREE) pixels8<MK, CFG, TWO, true>(a, lut3G, lut3B, lut3R, g, b, rr, yb, ch);
                else pixels8<MK, CFG, TWO>(a, lut_sa, lut_sa, lut_sa, g, b, rr, yb, ch);
                const uint4 ypack = pack_luma<CFG, TWO>(a, yb);
                if (lane_interior && r >= ys && r < ye) *reinterpret_cast<uint4 *>(yp) = ypack;
                float l3x = __shfl_up_sync(0xffffffffu, plo(ch[3]), 1), l3y = __shfl_up_sync(0xffffffffu, phi(ch[3]), 1);
                float l5x = __shfl_up_sync(0xffffffffu, plo(ch[5]), 1), l5y = __shfl_up_sync(0xffffffffu, phi(ch[5]), 1);
                float l7x = __shfl_up_sync(0xffffffffu, plo(ch[7]), 1), l7y = __shfl_up_sync(0xffffffffu, phi(ch[7]), 1);
                float n1x = __shfl_down_sync(0xffffffffu, plo(ch[1]), 1), n1y = __shfl_down_sync(0xffffffffu, phi(ch[1]), 1);
                float n3x = __shfl_down_sync(0xffffffffu, plo(ch[3]), 1), n3y = __shfl_down_sync(0xffffffffu, phi(ch[3]), 1);
                u64 l3 = pk(l3x, l3y), l5 = pk(l5x, l5y), l7 = pk(l7x, l7y), n1 = pk(n1x, n1y), n3 = pk(n3x, n3y);
                if (strip_edge) {                                   // warp-uniform (from a vote): inner strips skip 10 selects
                    if (left_edge) l3 = l5 = l7 = ch[0];            // replicate s[0]     (convert.cpp:295-300)
                    if (right_edge) n1 = n3 = ch[7];                // replicate s[W-1]
                }
                o[0] = fir_h7_pair(l3, l5, l7, ch[0], ch[1], ch[3], ch[5], hi_bits, hc0);
                o[1] = fir_h7_pair(l5, l7, ch[1], ch[2], ch[3], ch[5], ch[7], hi_bits, hc0);
                o[2] = fir_h7_pair(l7, ch[1], ch[3], ch[4], ch[5], ch[7], n1, hi_bits, hc0);
                o[3] = fir_h7_pair(ch[1], ch[3], ch[5], ch[6], ch[7], n1, n3, hi_bits, hc0);
            };

            // even row r = 2m: acc[i] is output j = m-3+i and receives tap 11-2i; acc[0] completes
            auto row_even = [&](const u64 o[4], int r) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
#pragma unroll
                    for (int i = 0; i < 6; i++) acc[i][c] = ffma2(kv[11 - 2 * i], o[c], acc[i][c]);
                }
                const int j = (r >> 1) - 3;
                if (lane_interior && j >= (ys >> 1) && j < (ye >> 1)) {
                    unsigned cbv[4], crv[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        // clamp [0,maxCV] + truncation + write_yuv's shift and range clamp: one integer clamp of the floor
                        int lo_, hi_;
                        unpk(fadd2_rm(acc[0][c], pk(MAGIC, MAGIC)), lo_, hi_);
                        if (CFG) {
                            // the floor is in [-2^15, 2^15): its low half is the s16 value, clamped two at a time below
                            cbv[c] = (unsigned)lo_; crv[c] = (unsigned)hi_;
                        } else {
                            cbv[c] = (unsigned)clamp3(lo_ >> shift, clo, chi);
                            crv[c] = (unsigned)clamp3(hi_ >> shift, clo, chi);
                        }
                    }
                    uint2 cbo = make_uint2(__byte_perm(cbv[0], cbv[1], 0x5410), __byte_perm(cbv[2], cbv[3], 0x5410));
                    uint2 cro = make_uint2(__byte_perm(crv[0], crv[1], 0x5410), __byte_perm(crv[2], crv[3], 0x5410));
                    if (CFG) {
                        const unsigned lo2 = (unsigned)C::loC(a) * 0x10001u, hi2 = (unsigned)C::hiC(a) * 0x10001u;
                        cbo.x = clamp_s16x2(cbo.x, lo2, hi2); cbo.y = clamp_s16x2(cbo.y, lo2, hi2);
                        cro.x = clamp_s16x2(cro.x, lo2, hi2); cro.y = clamp_s16x2(cro.y, lo2, hi2);
                    }
                    *reinterpret_cast<uint2 *>(cbp) = cbo;
                    *reinterpret_cast<uint2 *>(cbp + crd) = cro;
                }
                cbp += wh;
            };
            // odd row r = 2m+1: the accumulators shift down by one output (the FFMA2 writes the neighbour);
            // old acc[i+1] receives tap 10-2i and a new output starts in acc[5] with tap 0
            auto row_odd = [&](const u64 o[4]) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
#pragma unroll
                    for (int i = 0; i < 5; i++) acc[i][c] = ffma2(kv[10 - 2 * i], o[c], acc[i + 1][c]);
                    acc[5][c] = ffma2(kv[0], o[c], pk(0.5f, 0.5f));
                }
            };

            // one row body for both parities (measured: a two-row trip with two sample buffers spills and is 20 % slower,
            // loading the next row only after the pixel stage exposes the load latency: profiles/r01/variants.md)
            RawPx<NCH> raw;
            load_px8<NCH>(raw, sp, 0, 0, 0);
            int r = rfirst;
#pragma unroll 1
            for (; r <= rlast; r++) {
                unsigned g[8], b[8], rr[8];
#ifdef H2Y_LOAD_BEFORE_SPLIT
                RawPx<NCH> cur = raw;
                next_src(r);
                if (r < rlast) load_px8<NCH>(raw, sp, 0, 0, 0);                 // prefetch the next row
                row_split(cur, g, b, rr);
#else
                if (SPEC) {
                    // A row costs two 3-input min / max per word kind (the warp's bound is one of the inputs) and one
                    // vote: does any lane hold a code outside the bounds?  Rarely (a running extremum of n rows moves
                    // O(log n) times); then the bounds are widened by warp reductions.  Halo rows and halo lanes are
                    // counted twice: harmless.  No code above the predicted window may reach the gathers: the whole frame
                    // is handed back instead.
                    unsigned smx0, smn0, smx1, smn1, smx2 = 0u, smn2 = 0xFFFFFFFFu;
                    lds_bounds(sb_sa, smx0, smn0); lds_bounds(sb_sa + 8, smx1, smn1);
                    if (NCH == 3) lds_bounds(sb_sa + 16, smx2, smn2);
                    unsigned rx0, rx1, rx2 = 0u, rn0, rn1, rn2 = 0xFFFFFFFFu;
                    if (NCH == 3) {
                        rn0 = vmin2(vmin2(smn0, raw.v[0].x), vmin2(raw.v[0].w, vmin2(raw.v[1].z, raw.v[2].y)));
                        rn1 = vmin2(vmin2(smn1, raw.v[0].y), vmin2(raw.v[1].x, vmin2(raw.v[1].w, raw.v[2].z)));
                        rn2 = vmin2(vmin2(smn2, raw.v[0].z), vmin2(raw.v[1].y, vmin2(raw.v[2].x, raw.v[2].w)));
                        rx0 = vmax2(vmax2(smx0, raw.v[0].x), vmax2(raw.v[0].w, vmax2(raw.v[1].z, raw.v[2].y)));
                        rx1 = vmax2(vmax2(smx1, raw.v[0].y), vmax2(raw.v[1].x, vmax2(raw.v[1].w, raw.v[2].z)));
                        rx2 = vmax2(vmax2(smx2, raw.v[0].z), vmax2(raw.v[1].y, vmax2(raw.v[2].x, raw.v[2].w)));
                    } else {
                        rn0 = smn0; rn1 = smn1; rx0 = smx0; rx1 = smx1;
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            rn0 = vmin2(rn0, vmin2(raw.v[i].x, raw.v[i].z)); rx0 = vmax2(rx0, vmax2(raw.v[i].x, raw.v[i].z));
                            rn1 = vmin2(rn1, vmin2(raw.v[i].y, raw.v[i].w)); rx1 = vmax2(rx1, vmax2(raw.v[i].y, raw.v[i].w));
                        }
                    }
                    unsigned moved = (rx0 ^ smx0) | (rx1 ^ smx1) | (rn0 ^ smn0) | (rn1 ^ smn1);
                    if (NCH == 3) moved |= (rx2 ^ smx2) | (rn2 ^ smn2);
                    if (__any_sync(0xffffffffu, moved != 0u)) {
                        rx0 = warp_max_u16x2(rx0); rx1 = warp_max_u16x2(rx1); rn0 = warp_min_u16x2(rn0); rn1 = warp_min_u16x2(rn1);
                        if (NCH == 3) { rx2 = warp_max_u16x2(rx2); rn2 = warp_min_u16x2(rn2); }
                        __syncwarp();                                           // every lane has read the old bounds
                        if (lane == 0) {
                            sb[0].x = rx0; sb[0].y = rn0; sb[1].x = rx1; sb[1].y = rn1;
                            if (NCH == 3) { sb[2].x = rx2; sb[2].y = rn2; }
                        }
                        __syncwarp();
                        const unsigned top = NCH == 3 ? vmax2(vmax2(rx0, rx1), rx2) : vmax2(rx0, rx1 & 0xffffu);   // alpha is not a colour sample
                        if (vmax2(top, SPEC_CAP2) != SPEC_CAP2) break;           // warp-uniform
                    }
                }
                // the codes leave the sample registers first, then the next row is loaded into the same registers: no copy
                row_split(raw, g, b, rr);
                next_src(r);
                if (r < rlast) load_px8<NCH>(raw, sp, 0, 0, 0);                 // prefetch the next row
#endif
#ifdef H2Y_EXPERIMENT_L2_PREFETCH_ROWS
                // timing experiment: pull the row H2Y_EXPERIMENT_L2_PREFETCH_ROWS further down into L2
                if (r + 1 + H2Y_EXPERIMENT_L2_PREFETCH_ROWS < h && r + 1 >= 0) {
                    const uint8_t *pf = sp + (size_t)H2Y_EXPERIMENT_L2_PREFETCH_ROWS * spitch;
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(pf));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(pf + 16 * (NCH - 1)));
                }
#endif
                u64 o[4];
                row_front(g, b, rr, r, o);
                yp += w;
                if ((r & 1) == 0) row_even(o, r); else row_odd(o);
            }
            if (SPEC) {
                if (r <= rlast) { if (lane == 0) A.bail[frame] = 1; continue; }     // left the row loop early
                // the run's extrema (already reduced over the warp) join the frame's statistics slots in the layout k_plan
                // reads (h2y_stats.cu)
                if (lane == 0) {
                    const unsigned smx0 = sb[0].x, smn0 = sb[0].y, smx1 = sb[1].x, smn1 = sb[1].y, smx2 = sb[2].x, smn2 = sb[2].y;
                    unsigned cmin[3], cmax[3];                              // G, B, R
                    if (NCH == 3) {
                        cmin[0] = min(smn0 >> 16, smn2 & 0xffffu); cmax[0] = max(smx0 >> 16, smx2 & 0xffffu);
                        cmin[1] = min(smn1 & 0xffffu, smn2 >> 16); cmax[1] = max(smx1 & 0xffffu, smx2 >> 16);
                        cmin[2] = min(smn0 & 0xffffu, smn1 >> 16); cmax[2] = max(smx0 & 0xffffu, smx1 >> 16);
                    } else {
                        cmin[0] = smn0 >> 16; cmax[0] = smx0 >> 16;
                        cmin[1] = smn1 & 0xffffu; cmax[1] = smx1 & 0xffffu;
                        cmin[2] = smn0 & 0xffffu; cmax[2] = smx0 & 0xffffu;
                    }
                    unsigned *sl = A.slots + (size_t)frame * 12;
#pragma unroll
                    for (int ch3 = 0; ch3 < 3; ch3++) {
                        atomicMin(sl + 2 * ch3, half_code_key(cmin[ch3]));
                        atomicMax(sl + 2 * ch3 + 1, half_code_key(cmax[ch3]));
                    }
                    atomicMin(sl + 6, min(cmin[0], min(cmin[1], cmin[2])));
                    atomicMax(sl + 7, max(cmax[0], max(cmax[1], cmax[2])));
                }
            }
        }
    }
}

template <int MK, int NCH, int CFG = 0, bool TWO = false, bool THREE = false, bool SPEC = false>
__global__ void __launch_bounds__(THREADS3, 1) k_forward_exr420_rows(const Fwd3Args A)
{
    rows_body<MK, NCH, CFG, TWO, THREE, SPEC>(A);
}

// The instantiations of one call convert disjoint frames and read only what the prologue wrote, so they need not wait
// for one another: the first runs on the caller's stream, the others on the context's auxiliary streams, forked behind
// an event recorded after the prologue and joined by events before anything else is queued (aux_fork / aux_stream /
// aux_join, h2y_api.cu).  Each moves onto the SMs as its predecessor's CTAs leave them.  (Round 1 overlapped them with
// programmatic stream serialization and no griddepcontrol.wait; the stream ordering that relied on is not documented.)
template <class K>
static cudaError_t launch_rows_on(K kernel, int grid, size_t smem, cudaStream_t st, const Fwd3Args &A)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, THREADS3, smem, st>>>(A);
    return cudaGetLastError();
}

// =================================================================================================
// Integer rows (TIFF), 4:2:0 FIR, large batches: the warp-autonomous layout of k_forward_exr420_rows with the
// reference's arithmetic.  At 16-bit scale the vertical filter's operation order is observable (SURVEY.md Appendix
// A.7), so its twelve input rows are kept, not folded into running sums: each warp owns a 12-slot ring of horizontally
// filtered rows in shared memory ([slot][half][lane] float4, only its own lanes touch it, no barrier), and every second
// row it reads the window back and evaluates convert.cpp:365-374 in order.
constexpr int VSLOTS = 12;
__device__ __forceinline__ void ring_sts(unsigned a, u64 x, u64 y)
{
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" :: "r"(a), "l"(x), "l"(y));
}
__device__ __forceinline__ void ring_lds(unsigned a, u64 &x, u64 &y)       // a = base + constant: ptxas folds the constant
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(a));
}

template <int MK, int NCH, int FAM = 0>
__global__ void __launch_bounds__(THREADS3, 1) k_forward_u16_420_rows(const Fwd3Args A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Fwd2Args &a = A.b;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // [slot][half][lane] float4, by shared-window address held in an ordinary register (the asm hides that part of it is
    // uniform: otherwise ptxas rebuilds the address from SR_TID / SR_CgaCtaId for every store and every window read)
    unsigned ring_sa;
    asm volatile("mov.u32 %0, %1;" : "=r"(ring_sa) : "r"((unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)warp * (VSLOTS * 1024u) + (unsigned)lane * 16u));
    const PixK &k = a.k;
    const int w = a.w, h = a.h, wh = w >> 1;
    const int hi_bits = MAGIC_BITS + (int)k.maxCV;
    const int shift = k.down_shift;
    const int clo = (int)k.loC + (MAGIC_BITS >> shift), chi = (int)k.hiC + (MAGIC_BITS >> shift);
    unsigned fallbacks = 0;
    // the packed clip limits stay in registers (otherwise they are rebuilt from the constant bank for every row)
    unsigned clip_lo2, clip_hi2;
    asm volatile("mov.u32 %0, %1;" : "=r"(clip_lo2) : "r"(k.loadLo * 0x10001u));
    asm volatile("mov.u32 %0, %1;" : "=r"(clip_hi2) : "r"(k.loadHi * 0x10001u));
    LumaPack lpack = luma_pack(k);
    asm volatile("" : "+r"(lpack.keep), "+r"(lpack.lo2), "+r"(lpack.hi2));

    const int wk = warp / A.wps, sfirst = warp - wk * A.wps;
    if (wk >= A.sub) return;
    const long K = (long)gridDim.x * A.sub, kid = (long)blockIdx.x * A.sub + wk;
    const long g0 = ((kid * A.total_rows) / K) & ~1L, g1 = kid + 1 == K ? A.total_rows : (((kid + 1) * A.total_rows) / K) & ~1L;
    if (g1 <= g0) return;

    for (int frame = (int)(g0 / h); frame <= (int)((g1 - 1) / h); frame++) {
        const long fbase = (long)frame * h;
        const int ys = (int)(max(g0, fbase) - fbase), ye = (int)(min(g1, fbase + h) - fbase);
        if (ys >= ye) continue;
        const uint8_t *fsrc = a.src + (size_t)frame * a.src_stride;
        uint16_t *fY = reinterpret_cast<uint16_t *>(a.dst + (size_t)frame * a.dst_stride);
        uint16_t *fCb = fY + (size_t)w * h;
        uint16_t *fCr = fCb + (size_t)wh * (h >> 1);

        for (int strip = sfirst; strip < a.nstrips; strip += A.wps) {
            const int x0 = strip * a.strip_w;
            const int xl = x0 + 8 * (lane - 1);
            const bool lane_in_pic = xl >= 0 && xl < w;
            const bool lane_interior = lane >= 1 && lane < 31 && xl < min(x0 + a.strip_w, w);
            const int xload = lane_in_pic ? xl : (xl < 0 ? 0 : w - 8);
            const bool left_edge = xl == 0, right_edge = xl + 8 >= w;
            const bool strip_edge = __any_sync(0xffffffffu, left_edge || right_edge);    // warp-uniform: inner strips skip the selects
            const int rfirst = ys - 6, rlast = ye + 4;              // rows feeding outputs ys/2 .. ye/2-1 (rfirst even)
            const unsigned spitch = (unsigned)w * (2 * NCH);
            const uint8_t *sp = fsrc + (size_t)min(max(rfirst, 0), h - 1) * spitch + (size_t)xload * (2 * NCH);
            uint16_t *yp = fY + (ptrdiff_t)rfirst * w + xl;
            uint16_t *cbp = fCb + ((ptrdiff_t)(rfirst >> 1) - 3) * wh + (xl >> 1);
            const int crd = (int)(fCr - fCb);
            unsigned so = 0;                                         // ring slot of row r, times 1024; row r - t sits t slots back

            RawPx<NCH> raw;
            load_px8<NCH>(raw, sp, 0, 0, 0);
#pragma unroll 1
            for (int r = rfirst; r <= rlast; r++) {
                if (k.clip_on_load) {                                               // read_tiff's clip (tiff.cpp:296-304)
                    const unsigned lo2 = clip_lo2, hi2 = clip_hi2;
#pragma unroll
                    for (int i = 0; i < NCH; i++) {
                        unsigned *wv = reinterpret_cast<unsigned *>(&raw.v[i]);
#pragma unroll
                        for (int j = 0; j < 4; j++) wv[j] = clamp_u16x2(wv[j], lo2, hi2);
                    }
                }
                // the codes leave the sample registers first, then the next row is loaded into the same registers: no copy
                unsigned g[8], b[8], rr[8];
                split_codes<NCH>(raw, g, b, rr);
                sp += ((unsigned)r < (unsigned)(h - 1)) ? spitch : 0;               // clamped rows = edge replicate
                if (r < rlast) load_px8<NCH>(raw, sp, 0, 0, 0);                     // prefetch the next row
                uint4 ypack;
                u64 ch[8];
                pixels8_u16<MK, FAM>(k, lpack, g, b, rr, ypack, ch, fallbacks);
                if (lane_interior && r >= ys && r < ye) *reinterpret_cast<uint4 *>(yp) = ypack;
                yp += w;
                float l3x = __shfl_up_sync(0xffffffffu, plo(ch[3]), 1), l3y = __shfl_up_sync(0xffffffffu, phi(ch[3]), 1);
                float l5x = __shfl_up_sync(0xffffffffu, plo(ch[5]), 1), l5y = __shfl_up_sync(0xffffffffu, phi(ch[5]), 1);
                float l7x = __shfl_up_sync(0xffffffffu, plo(ch[7]), 1), l7y = __shfl_up_sync(0xffffffffu, phi(ch[7]), 1);
                float n1x = __shfl_down_sync(0xffffffffu, plo(ch[1]), 1), n1y = __shfl_down_sync(0xffffffffu, phi(ch[1]), 1);
                float n3x = __shfl_down_sync(0xffffffffu, plo(ch[3]), 1), n3y = __shfl_down_sync(0xffffffffu, phi(ch[3]), 1);
                u64 l3 = pk(l3x, l3y), l5 = pk(l5x, l5y), l7 = pk(l7x, l7y), n1 = pk(n1x, n1y), n3 = pk(n3x, n3y);
                if (strip_edge) {
                    if (left_edge) l3 = l5 = l7 = ch[0];            // replicate s[0]     (convert.cpp:295-300)
                    if (right_edge) n1 = n3 = ch[7];                // replicate s[W-1]
                }
                const u64 o0 = fir_h7_pair_ord(l3, l5, l7, ch[0], ch[1], ch[3], ch[5], hi_bits);
                const u64 o1 = fir_h7_pair_ord(l5, l7, ch[1], ch[2], ch[3], ch[5], ch[7], hi_bits);
                const u64 o2 = fir_h7_pair_ord(l7, ch[1], ch[3], ch[4], ch[5], ch[7], n1, hi_bits);
                const u64 o3 = fir_h7_pair_ord(ch[1], ch[3], ch[5], ch[6], ch[7], n1, n3, hi_bits);
                ring_sts(ring_sa + so, o0, o1);
                ring_sts(ring_sa + so + 512u, o2, o3);
                __syncwarp();
                const int j = (r >> 1) - 3;                         // even row r = 2j+6 completes output row j
                if ((r & 1) == 0 && j >= (ys >> 1) && j < (ye >> 1)) {
                    unsigned cbv[4], crv[4];
                    // row r-11+t sits in slot (slot + 1 + t) mod 12: one of two bases per tap (a compare and a select),
                    // the tap's offset is an immediate of the load; the twelve addresses serve both halves
                    const unsigned a0 = ring_sa + so + 1024u, a1 = a0 - VSLOTS * 1024u;
                    unsigned wa[12];
#pragma unroll
                    for (int t = 0; t < 12; t++) wa[t] = so >= (unsigned)(VSLOTS - 1 - t) * 1024u ? a1 : a0;
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        u64 win[2][12];                             // rows 2j-5 .. 2j+6 of this lane's columns 2*half, 2*half+1
#pragma unroll
                        for (int t = 0; t < 12; t++) ring_lds(wa[t] + (unsigned)(1024 * t + 512 * half), win[0][t], win[1][t]);
#pragma unroll
                        for (int cc = 0; cc < 2; cc++) {
                            // clamp [0,maxCV] + truncation + write_yuv's shift and range clamp: one integer clamp of the floor
                            int lo_, hi_;
                            unpk(fir_v12_pair_ord(win[cc]), lo_, hi_);
                            cbv[2 * half + cc] = (unsigned)clamp3(lo_ >> shift, clo, chi);
                            crv[2 * half + cc] = (unsigned)clamp3(hi_ >> shift, clo, chi);
                        }
                    }
                    if (lane_interior) {
                        *reinterpret_cast<uint2 *>(cbp) = make_uint2(__byte_perm(cbv[0], cbv[1], 0x5410), __byte_perm(cbv[2], cbv[3], 0x5410));
                        *reinterpret_cast<uint2 *>(cbp + crd) = make_uint2(__byte_perm(crv[0], crv[1], 0x5410), __byte_perm(crv[2], crv[3], 0x5410));
                    }
                }
                if ((r & 1) == 0) cbp += wh;
                so = so + 1024u == VSLOTS * 1024u ? 0u : so + 1024u;
            }
            __syncwarp();
        }
    }
    if (a.fallback_count && fallbacks) atomicAdd(a.fallback_count, (unsigned long long)fallbacks);
}

// ---- integer source, 4:4:4 output: a pure streaming map (no filter), one 8-pixel group per thread step ---------
template <int MK, int NCH>
__global__ void __launch_bounds__(256) k_forward_u16_444(const Fwd2Args a, long groups_per_frame)
{
    const PixK &k = a.k;
    const long total = groups_per_frame * a.nframes;
    const size_t plane = (size_t)a.w * a.h;
    for (long gi = (long)blockIdx.x * blockDim.x + threadIdx.x; gi < total; gi += (long)gridDim.x * blockDim.x) {
        const int frame = (int)(gi / groups_per_frame);
        const long px = (gi - (long)frame * groups_per_frame) * 8;
        RawPx<NCH> raw;
        load_px8<NCH>(raw, a.src + (size_t)frame * a.src_stride + (size_t)px * (2 * NCH), 0, 0, 0);
        if (k.clip_on_load) {                                   // read_tiff's clip (tiff.cpp:296-304)
            const unsigned lo2 = k.loadLo * 0x10001u, hi2 = k.loadHi * 0x10001u;
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                unsigned *wv = reinterpret_cast<unsigned *>(&raw.v[i]);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    asm("max.u16x2 %0, %0, %1;" : "+r"(wv[j]) : "r"(lo2));
                    asm("min.u16x2 %0, %0, %1;" : "+r"(wv[j]) : "r"(hi2));
                }
            }
        }
        unsigned g[8], b[8], r[8], Y[8], Cb[8], Cr[8];
        split_codes<NCH>(raw, g, b, r);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (!px_matrix_fast_u16<MK>(g[q], b[q], r[q], k, Y[q], Cb[q], Cr[q]))
                px_matrix_exact<MK>((float)g[q], (float)b[q], (float)r[q], k, Y[q], Cb[q], Cr[q]);
            Y[q] = out_clamp(Y[q], k.down_shift, k.loY, k.hiY);
            Cb[q] = out_clamp(Cb[q], k.down_shift, k.loC, k.hiC);
            Cr[q] = out_clamp(Cr[q], k.down_shift, k.loC, k.hiC);
        }
        uint16_t *fo = reinterpret_cast<uint16_t *>(a.dst + (size_t)frame * a.dst_stride) + px;
        *reinterpret_cast<uint4 *>(fo) = make_uint4(Y[0] | (Y[1] << 16), Y[2] | (Y[3] << 16), Y[4] | (Y[5] << 16), Y[6] | (Y[7] << 16));
        *reinterpret_cast<uint4 *>(fo + plane) = make_uint4(Cb[0] | (Cb[1] << 16), Cb[2] | (Cb[3] << 16), Cb[4] | (Cb[5] << 16), Cb[6] | (Cb[7] << 16));
        *reinterpret_cast<uint4 *>(fo + 2 * plane) = make_uint4(Cr[0] | (Cr[1] << 16), Cr[2] | (Cr[3] << 16), Cr[4] | (Cr[5] << 16), Cr[6] | (Cr[7] << 16));
    }
}

// ---- host side -----------------------------------------------------------------------------------------

bool forward_u16_420_supported(const h2y_forward_params &p, const PixK &k)
{
    if (p.src.layout != H2Y_LAYOUT_RGB16 && p.src.layout != H2Y_LAYOUT_RGBA16) return false;
    if (k.convert_transfer) return false;
    if (p.dst.chroma_format_idc == H2Y_CHROMA_444) {                       // streaming map, any height
        if (k.mat_kind != MK_YCBCR && k.mat_kind != MK_YDZDX) return false;
        return ((long)p.src.width * p.src.height) % 8 == 0;
    }
    if (p.dst.chroma_format_idc != H2Y_CHROMA_420 || p.chroma_resampler_type == 0) return false;
    if (k.mat_kind != MK_YCBCR && k.mat_kind != MK_YDZDX) return false;
    const int w = p.src.width, h = p.src.height;
    return w >= 8 && (w & 7) == 0 && h >= 2 && (h & 1) == 0;
}

static void ring_items(Fwd2Args &a, int nframes, int grid_max)
{
    a.strip_w = 240;
    a.nstrips = (a.w + a.strip_w - 1) / a.strip_w;
    long want_items = 6L * grid_max;
    int nsegs = (int)((want_items + (long)nframes * a.nstrips - 1) / ((long)nframes * a.nstrips));
    int max_segs = a.h / 96 > 0 ? a.h / 96 : 1;
    if (nsegs > max_segs) nsegs = max_segs;
    if (nsegs < 1) nsegs = 1;
    int seg_rows = (a.h + nsegs - 1) / nsegs;
    seg_rows = (seg_rows + 15) / 16 * 16;
    a.seg_rows = seg_rows;
    a.nsegs = (a.h + seg_rows - 1) / seg_rows;
    a.nitems = nframes * a.nsegs * a.nstrips;
}

h2y_status launch_forward_u16_420(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, const void *d_src,
                                  size_t src_stride, void *d_dst, size_t dst_stride, int nframes, cudaStream_t st)
{
    Fwd2Args a;
    memset(&a, 0, sizeof(a));
    a.src = (const uint8_t *)d_src; a.src_stride = src_stride;
    a.dst = (uint8_t *)d_dst; a.dst_stride = dst_stride;
    a.w = p.src.width; a.h = p.src.height; a.nframes = nframes;
    a.k = k;
    if (p.dst.chroma_format_idc == H2Y_CHROMA_444) {
        const long groups = (long)a.w * a.h / 8;
        long want = (groups * nframes + 255) / 256;
        const int blocks = (int)(want < 16L * c->sm_count ? want : 16L * c->sm_count);
        const int nch4 = layout_channels(p.src.layout);
        if (k.mat_kind == MK_YCBCR) {
            if (nch4 == 3) k_forward_u16_444<MK_YCBCR, 3><<<blocks, 256, 0, st>>>(a, groups);
            else k_forward_u16_444<MK_YCBCR, 4><<<blocks, 256, 0, st>>>(a, groups);
        } else {
            if (nch4 == 3) k_forward_u16_444<MK_YDZDX, 3><<<blocks, 256, 0, st>>>(a, groups);
            else k_forward_u16_444<MK_YDZDX, 4><<<blocks, 256, 0, st>>>(a, groups);
        }
        c->launches++;
        H2Y_CUDA(c, cudaGetLastError());
        return H2Y_OK;
    }
    const int nch = layout_channels(p.src.layout);
    {
        // large batches: the warp-autonomous kernel (rows of the whole batch split evenly over all warps)
        a.strip_w = 240;
        a.nstrips = (a.w + a.strip_w - 1) / a.strip_w;
        Fwd3Args A3;
        memset(&A3, 0, sizeof(A3));
        A3.wps = 1;
        for (int d = 1; d <= WARPS3; d++) if (WARPS3 % d == 0 && a.nstrips % d == 0) A3.wps = d;
        A3.sub = WARPS3 / A3.wps;
        A3.total_rows = (long)nframes * a.h;
        const long rows_per_worker = A3.total_rows / ((long)c->sm_count * A3.sub);
        const bool want_rows = c->sw.fwd_kernel ? c->sw.fwd_kernel == 2 : rows_per_worker >= ROWS_MIN_PER_WORKER;    // forced by tests and experiments
        if (want_rows && A3.total_rows >= 2) {
            int g3 = c->sm_count;
            while (g3 > 1 && A3.total_rows / ((long)g3 * A3.sub) < 16) g3 >>= 1;   // forced on a tiny batch
            A3.b = a;
            const size_t smem3 = (size_t)WARPS3 * VSLOTS * 64 * sizeof(float4);
#define LR(MKV, NC, FAMV)                                                                                                  \
    do {                                                                                                                   \
        H2Y_CUDA(c, cudaFuncSetAttribute(k_forward_u16_420_rows<MKV, NC, FAMV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3)); \
        k_forward_u16_420_rows<MKV, NC, FAMV><<<g3, THREADS3, smem3, st>>>(A3);                                           \
    } while (0)
            // the two Y'CbCr families get their constants as immediates
            const int fam = k.mat_kind != MK_YCBCR || c->sw.no_specialised ? 0
                            : (k.wri == 2627 && k.wgi == 6780 && k.wbi == 593 && k.db == 1.8814 && k.dr == 1.4746 ? 2020
                               : (k.wri == 2126 && k.wgi == 7152 && k.wbi == 722 && k.db == 1.8556 && k.dr == 1.5748 ? 709 : 0));
            if (k.mat_kind == MK_YCBCR) {
                if (fam == 2020) { if (nch == 3) LR(MK_YCBCR, 3, 2020); else LR(MK_YCBCR, 4, 2020); }
                else if (fam == 709) { if (nch == 3) LR(MK_YCBCR, 3, 709); else LR(MK_YCBCR, 4, 709); }
                else { if (nch == 3) LR(MK_YCBCR, 3, 0); else LR(MK_YCBCR, 4, 0); }
            } else { if (nch == 3) LR(MK_YDZDX, 3, 0); else LR(MK_YDZDX, 4, 0); }
#undef LR
            c->launches++;
            H2Y_CUDA(c, cudaGetLastError());
            return H2Y_OK;
        }
    }
    ring_items(a, nframes, c->sm_count);
    const size_t smem = (size_t)RING_ROWS * RING_PITCH * sizeof(float);
    const int grid = a.nitems < c->sm_count ? a.nitems : c->sm_count;
#define LU(MKV, NC)                                                                                                        \
    do {                                                                                                                   \
        H2Y_CUDA(c, cudaFuncSetAttribute(k_forward_exr420<MKV, NC, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_forward_exr420<MKV, NC, 1><<<grid, THREADS, smem, st>>>(a);                                                     \
    } while (0)
    if (k.mat_kind == MK_YCBCR) { if (nch == 3) LU(MK_YCBCR, 3); else LU(MK_YCBCR, 4); }
    else { if (nch == 3) LU(MK_YDZDX, 3); else LU(MK_YDZDX, 4); }
#undef LU
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}


bool forward_exr420_supported(const h2y_forward_params &p, const PixK &k, int tmp_bit_depth)
{
    if (!layout_is_half(p.src.layout) || !k.convert_transfer) return false;
    if (p.dst.chroma_format_idc != H2Y_CHROMA_420 || p.chroma_resampler_type == 0) return false;
    if (k.mat_kind != MK_YCBCR && k.mat_kind != MK_YDZDX) return false;
    if (k.scale_mode != SC_VIDEO && k.scale_mode != SC_FULL) return false;
    if (tmp_bit_depth > 12 || tmp_bit_depth < 8) return false;
    const int w = p.src.width, h = p.src.height;
    return w >= 8 && (w & 7) == 0 && h >= 2 && (h & 1) == 0;
}

template <int MK, int NCH>
static h2y_status launch_v2(h2y_ctx_impl *c, const Fwd2Args &a, int grid, size_t smem, cudaStream_t st)
{
    H2Y_CUDA(c, cudaFuncSetAttribute(k_forward_exr420<MK, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_forward_exr420<MK, NCH><<<grid, THREADS, smem, st>>>(a);
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

// Geometry of a rows-kernel launch and whether the batch is large enough for it
static bool rows_plan(const h2y_ctx_impl *c, const Fwd2Args &a, int nframes, Fwd3Args *A3, int *g3)
{
    memset(A3, 0, sizeof(*A3));
    A3->b = a;
    // warps per row worker: the largest count that divides both the strips and the CTA's warps
    A3->wps = 1;
    for (int d = 1; d <= WARPS3; d++) if (WARPS3 % d == 0 && a.nstrips % d == 0) A3->wps = d;
    A3->sub = WARPS3 / A3->wps;
    A3->total_rows = (long)nframes * a.h;
    const long rows_per_worker = A3->total_rows / ((long)c->sm_count * A3->sub);
    const bool want_rows = c->sw.fwd_kernel ? c->sw.fwd_kernel == 2 : rows_per_worker >= ROWS_MIN_PER_WORKER;
    if (!want_rows || A3->total_rows < 2) return false;
    *g3 = c->sm_count;
    while (*g3 > 1 && A3->total_rows / ((long)*g3 * A3->sub) < 16) *g3 >>= 1;       // forced on a tiny batch
    return true;
}

// the headline configurations get their constants as immediates (KC<10>, KC<12>)
static bool exr_cfgd(const h2y_ctx_impl *c, const PixK &k, const Fwd2Args &a, int tmp_bit_depth)
{
    const int sc = 1 << (tmp_bit_depth - 8);
    return k.mat_kind == MK_YCBCR && k.wr == 0.2627 && k.wg == 0.6780 && k.wb == 0.0593 && k.db == 1.8814 && k.dr == 1.4746 &&
           (tmp_bit_depth == 10 || tmp_bit_depth == 12) && k.scale_mode == SC_VIDEO && k.down_shift == 0 &&
           a.k.mulY == (float)(235 * sc) && a.k.mulC == (float)(240 * sc) && a.k.addY == (float)(16 * sc) &&
           a.k.addC == (float)(16 * sc) && (int)k.loY == 16 * sc && (int)k.hiY == 235 * sc && (int)k.loC == 16 * sc &&
           (int)k.hiC == 240 * sc && (int)k.maxCV == (1 << tmp_bit_depth) - 1 && c->sw.guard_log2 < 0 && !c->sw.no_specialised;
}

static void exr_args(const h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, int tmp_bit_depth, const void *d_src,
                     size_t src_stride, void *d_dst, size_t dst_stride, int nframes, const FrameK *d_framek,
                     const float *d_luts, Fwd2Args *pa)
{
    Fwd2Args &a = *pa;
    a.src = (const uint8_t *)d_src; a.src_stride = src_stride;
    a.dst = (uint8_t *)d_dst; a.dst_stride = dst_stride;
    a.w = p.src.width; a.h = p.src.height; a.nframes = nframes;
    a.k = k;
    if (k.scale_mode == SC_FULL) { a.k.mulC = k.mulY; a.k.addY = 0.0f; a.k.addC = 0.0f; }   // x*maxCV (+0 is exact)
    a.framek = d_framek; a.luts = d_luts;
    a.fallback_count = nullptr;
    // the fp32 evaluation is within 7.2/8 G of the reference at this depth: G = 2^(depth-22)
    a.guard = 1.0f / (float)(1 << (H2Y_GUARD_SHIFT - tmp_bit_depth));
    if (c->sw.guard_log2 >= 0) a.guard = exp2f(-(float)c->sw.guard_log2);   // timing experiments only: breaks parity
    a.wr = (float)k.wr; a.wg = (float)k.wg; a.wb = (float)k.wb;
    a.rdb = k.mat_kind == MK_YCBCR ? (float)k.rdb : 0.5f;
    a.rdr = k.mat_kind == MK_YCBCR ? (float)k.rdr : 0.5f;
    // chroma is taken against the lowered luma (sf - G) on the Y'CbCr route: fold the +G*rd back into the constant
    // per chain: luma 6u, Cb 7u, Cr 8u = a.guard (the bound of each chain: DESIGN.md 4); one width for all when an
    // experiment sets it
    const float gy = c->sw.guard_log2 >= 0 ? a.guard : 0.75f * a.guard, gb = c->sw.guard_log2 >= 0 ? a.guard : 0.875f * a.guard, gr = a.guard;
    a.lumc = 0.5f - gy;
    a.twoGy = 2.0f * gy; a.twoGb = 2.0f * gb; a.twoGr = 2.0f * gr;
    a.cbc = k.mat_kind == MK_YCBCR ? 0.5f - gb - gy * a.rdb : 0.5f - gb;
    a.crc = k.mat_kind == MK_YCBCR ? 0.5f - gr - gy * a.rdr : 0.5f - gr;
    a.strip_w = 240;
    a.nstrips = (a.w + a.strip_w - 1) / a.strip_w;
    const int grid_max = c->sm_count;
    long want_items = 6L * grid_max;
    int nsegs = (int)((want_items + (long)nframes * a.nstrips - 1) / ((long)nframes * a.nstrips));
    int max_segs = a.h / 96 > 0 ? a.h / 96 : 1;
    if (nsegs > max_segs) nsegs = max_segs;
    if (nsegs < 1) nsegs = 1;
    int seg_rows = (a.h + nsegs - 1) / nsegs;
    seg_rows = (seg_rows + 15) / 16 * 16;
    a.seg_rows = seg_rows;
    a.nsegs = (a.h + seg_rows - 1) / seg_rows;
    a.nitems = nframes * a.nsegs * a.nstrips;
}

// Can this call take the single-pass route (SPEC instantiations exist for the compiled-in configurations only)?
bool forward_exr420_spec_supported(const h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, int tmp_bit_depth, int nframes)
{
    if (!forward_exr420_supported(p, k, tmp_bit_depth)) return false;
    Fwd2Args a;
    exr_args(c, p, k, tmp_bit_depth, nullptr, 0, nullptr, 0, nframes, nullptr, nullptr, &a);
    Fwd3Args A3;
    int g3;
    return exr_cfgd(c, k, a, tmp_bit_depth) && rows_plan(c, a, nframes, &A3, &g3);
}

template <int NC, int DD>
static h2y_status launch_cfgd(h2y_ctx_impl *c, Fwd3Args &A3, int g3, cudaStream_t st, bool spec)
{
    const size_t smem2 = (size_t)2 * LUT2_CODES * sizeof(float), smem3 = (size_t)LUT_MAX_CODES * sizeof(float),
                 smemT = (size_t)LUT3_FLOATS * sizeof(float);
    A3.split_by_lut2 = 1;
    A3.split_three_nc = 1;
    h2y_status s = aux_fork(c, st);
    if (s != H2Y_OK) return s;
    if (spec) {
        // single pass: the frames are converted with the predicted plan; two-copy or single-copy LUT by the seed's window
        H2Y_CUDA(c, launch_rows_on(k_forward_exr420_rows<MK_YCBCR, NC, DD, true, false, true>, g3, smem2 + SPEC_BOUNDS_BYTES, st, A3));
        H2Y_CUDA(c, launch_rows_on(k_forward_exr420_rows<MK_YCBCR, NC, DD, false, false, true>, g3, smem3 + SPEC_BOUNDS_BYTES, aux_stream(c, 0), A3));
        c->launches += 2;
        return H2Y_OK;
    }
    // frames whose code range allows two pre-scaled LUT copies, then the rest; three-table frames, with and without
    // matrix_convert's chroma clamp
    H2Y_CUDA(c, launch_rows_on(k_forward_exr420_rows<MK_YCBCR, NC, DD, true>, g3, smem2, st, A3));
    H2Y_CUDA(c, launch_rows_on(k_forward_exr420_rows<MK_YCBCR, NC, DD, false>, g3, smem3, aux_stream(c, 0), A3));
    // the two three-table instantiations share one auxiliary stream (clamp-free first: it takes most frames).  On two
    // streams they start together, each CTA holds a whole SM (232 KB of shared memory) until its ~900 rows are done, and
    // whichever kernel gets the SMs first decides the tail: natural content measured 1.14 or 1.27 ms from run to run
    H2Y_CUDA(c, launch_rows_on(k_forward_exr420_rows<MK_YCBCR, NC, DD, true, true>, g3, smemT, aux_stream(c, 1), A3));
    H2Y_CUDA(c, launch_rows_on(k_forward_exr420_rows<MK_YCBCR, NC, DD, false, true>, g3, smemT, aux_stream(c, 1), A3));
    c->launches += 4;
    return H2Y_OK;
}

template <int MKV, int NC>
static h2y_status launch_generic_rows(h2y_ctx_impl *c, Fwd3Args &A3, int g3, cudaStream_t st)
{
    const size_t smem3 = (size_t)LUT_MAX_CODES * sizeof(float), smemT = (size_t)LUT3_FLOATS * sizeof(float);
    h2y_status s = aux_fork(c, st);
    if (s != H2Y_OK) return s;
    H2Y_CUDA(c, launch_rows_on(k_forward_exr420_rows<MKV, NC>, g3, smem3, st, A3));
    H2Y_CUDA(c, launch_rows_on(k_forward_exr420_rows<MKV, NC, 0, false, true>, g3, smemT, aux_stream(c, 0), A3));
    c->launches += 2;
    return H2Y_OK;
}

// The caller joins the auxiliary streams (aux_join) after it has queued the general-kernel sweep.
h2y_status launch_forward_exr420(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, int tmp_bit_depth,
                                 const void *d_src, size_t src_stride, void *d_dst, size_t dst_stride, int nframes,
                                 const FrameK *d_framek, const float *d_luts, cudaStream_t st, int *took_three_table_frames,
                                 const SpecLaunch *sl)
{
    Fwd2Args a;
    exr_args(c, p, k, tmp_bit_depth, d_src, src_stride, d_dst, dst_stride, nframes, d_framek, d_luts, &a);
    const int nch = layout_channels(p.src.layout);
    // large batches: the warp-autonomous kernel (rows split evenly over all warps of the GPU)
    Fwd3Args A3;
    int g3 = 0;
    if (rows_plan(c, a, nframes, &A3, &g3)) {
        if (sl) { A3.ctl = sl->ctl; A3.bail = sl->bail; A3.slots = sl->slots; }
        // frames that need a table per channel (FrameK::clean3) are served by further instantiations of the same kernel
        if (took_three_table_frames) *took_three_table_frames = 1;
        const bool spec = sl && sl->spec;
        if (exr_cfgd(c, k, a, tmp_bit_depth)) {
            if (tmp_bit_depth == 10) return nch == 3 ? launch_cfgd<3, 10>(c, A3, g3, st, spec) : launch_cfgd<4, 10>(c, A3, g3, st, spec);
            return nch == 3 ? launch_cfgd<3, 12>(c, A3, g3, st, spec) : launch_cfgd<4, 12>(c, A3, g3, st, spec);
        }
        if (spec) return H2Y_ERR_UNSUPPORTED;                     // forward_exr420_spec_supported() said otherwise
        if (k.mat_kind == MK_YCBCR) return nch == 3 ? launch_generic_rows<MK_YCBCR, 3>(c, A3, g3, st) : launch_generic_rows<MK_YCBCR, 4>(c, A3, g3, st);
        return nch == 3 ? launch_generic_rows<MK_YDZDX, 3>(c, A3, g3, st) : launch_generic_rows<MK_YDZDX, 4>(c, A3, g3, st);
    }
    if (sl && sl->spec) return H2Y_ERR_UNSUPPORTED;
    const size_t smem = (size_t)RING_ROWS * RING_PITCH * sizeof(float) + (size_t)LUT_MAX_CODES * sizeof(float);
    const int grid = a.nitems < c->sm_count ? a.nitems : c->sm_count;
    if (k.mat_kind == MK_YCBCR)
        return nch == 3 ? launch_v2<MK_YCBCR, 3>(c, a, grid, smem, st) : launch_v2<MK_YCBCR, 4>(c, a, grid, smem, st);
    return nch == 3 ? launch_v2<MK_YDZDX, 3>(c, a, grid, smem, st) : launch_v2<MK_YDZDX, 4>(c, a, grid, smem, st);
}

}   // namespace h2y
