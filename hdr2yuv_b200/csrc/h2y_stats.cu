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
#include <cstdlib>
#include <cstring>

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

// stats slots: [frame][12] = {min key, max key} x 3 channels, then min and max RAW 16-bit code, then (half sources,
// vector kernel) per channel G,B,R the smallest NONZERO raw code minus one (0xffffffff: not collected)
constexpr int SLOTS = 12;
__global__ void k_stats_init(unsigned *slots, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int j = i % SLOTS;
        slots[i] = j >= 8 ? 0xffffffffu : ((j & 1) ? 0u : 0xffffffffu);
    }
}

// min(acc, w - 1) on both 16-bit halves, one VIADDMNMX: zero codes wrap to 0xffff and drop out of the minimum
__device__ __forceinline__ unsigned nzmin(unsigned acc, unsigned w)
{
    unsigned t;
    asm("add.u16x2 %0, %1, %2;" : "=r"(t) : "r"(w), "r"(0xffffffffu));
    asm("min.u16x2 %0, %0, %1;" : "+r"(acc) : "r"(t));
    return acc;
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
            atomicMin(&slots[blockIdx.y * SLOTS + c * 2 + 0], mn[c]);
            atomicMax(&slots[blockIdx.y * SLOTS + c * 2 + 1], mx[c]);
        }
    }
    // this scalar route does not track the raw code range: the frame is never "clean"
    if (threadIdx.x == 0 && blockIdx.x == 0) { slots[blockIdx.y * SLOTS + 6] = 0u; slots[blockIdx.y * SLOTS + 7] = 0xffffu; }
}

// ---- vectorised statistics for the fused path -----------------------------------------------------
// 16-byte loads, packed 16-bit min/max (HMNMX2 for half bit patterns, VIMNMX.U16x2 for integer codes):
// 3 instructions per pixel, so the pass runs at HBM speed.  Interleaved 3-channel data repeats every
// 48 bytes with three word kinds (R,G) (B,R) (G,B); 4-channel data has two, (R,G) and (B,A).
template <bool HALF> struct Pk;
template <> struct Pk<true> {
    static __device__ __forceinline__ unsigned mn(unsigned a, unsigned b)
    {
        __half2 r = __hmin2(*reinterpret_cast<__half2 *>(&a), *reinterpret_cast<__half2 *>(&b));
        return *reinterpret_cast<unsigned *>(&r);
    }
    static __device__ __forceinline__ unsigned mx(unsigned a, unsigned b)
    {
        __half2 r = __hmax2(*reinterpret_cast<__half2 *>(&a), *reinterpret_cast<__half2 *>(&b));
        return *reinterpret_cast<unsigned *>(&r);
    }
    static constexpr unsigned MIN_INIT = 0x7C007C00u, MAX_INIT = 0xFC00FC00u;   // +inf, -inf
    // NaN never wins a '<' / '>' in the reference loop (common.cpp:121-133); minNum/maxNum drop it too
    static __device__ __forceinline__ bool key(unsigned h, unsigned &k)
    {
        if ((h & 0x7FFFu) > 0x7C00u) return false;
        k = fkey(half_bits_to_float(h));
        return true;
    }
};
template <> struct Pk<false> {
    static __device__ __forceinline__ unsigned mn(unsigned a, unsigned b)
    {
        unsigned r;
        asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
        return r;
    }
    static __device__ __forceinline__ unsigned mx(unsigned a, unsigned b)
    {
        unsigned r;
        asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
        return r;
    }
    static constexpr unsigned MIN_INIT = 0xFFFFFFFFu, MAX_INIT = 0u;
    static __device__ __forceinline__ bool key(unsigned c, unsigned &k) { k = c; return true; }
};

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{
    uint4 v;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// grid = (blocks, nframes); npix % 8 == 0, frames 16-byte aligned.  NCH = 3 or 4 interleaved, 0 = planar.
template <bool HALF, int NCH>
__device__ __forceinline__ void stats_vec_frame(const uint8_t *__restrict__ src, size_t frame_stride, long npix, int clip_on, unsigned lo,
                                                unsigned hi, unsigned *slots, const int frame)
{
    typedef Pk<HALF> P;
    const uint8_t *f = src + (size_t)frame * frame_stride;
    unsigned mnv[3], mxv[3];        // per word kind (interleaved) or per plane (planar), two lanes each
    unsigned umn = 0xFFFFFFFFu, umx = 0u;   // half input: extrema of the RAW codes of the colour channels (packed)
    unsigned nzv[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};     // half input, 3 channels: smallest nonzero code - 1, per word kind
#pragma unroll
    for (int i = 0; i < 3; i++) { mnv[i] = P::MIN_INIT; mxv[i] = P::MAX_INIT; }
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long)gridDim.x * blockDim.x;
    if (NCH == 3) {
        // Warp-contiguous 16-byte loads (whole sectors per instruction).  Vector n holds words of kind
        // (n + w) mod 3 at position w, kinds 0 = (R,G), 1 = (B,R), 2 = (G,B); the grid stride is a multiple of
        // 3 vectors, so a thread's phase n mod 3 never changes and it can accumulate per word POSITION.
        const long nvec = npix * 6 / 16;
        unsigned pmn[4], pmx[4], pnz[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { pmn[i] = P::MIN_INIT; pmx[i] = P::MAX_INIT; pnz[i] = 0xFFFFFFFFu; }
        const uint4 *p = reinterpret_cast<const uint4 *>(f);
        long n = tid;
        for (; n + 3 * nthr < nvec; n += 4 * nthr) {
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = ldg_stream(p + n + j * nthr);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // stepping by nthr (a multiple of 3) keeps the phase
                pmn[0] = P::mn(pmn[0], v[j].x); pmx[0] = P::mx(pmx[0], v[j].x);
                pmn[1] = P::mn(pmn[1], v[j].y); pmx[1] = P::mx(pmx[1], v[j].y);
                pmn[2] = P::mn(pmn[2], v[j].z); pmx[2] = P::mx(pmx[2], v[j].z);
                pmn[3] = P::mn(pmn[3], v[j].w); pmx[3] = P::mx(pmx[3], v[j].w);
                if (HALF) {
                    typedef Pk<false> U;
                    umx = U::mx(umx, U::mx(U::mx(v[j].x, v[j].y), U::mx(v[j].z, v[j].w)));
                    umn = U::mn(umn, U::mn(U::mn(v[j].x, v[j].y), U::mn(v[j].z, v[j].w)));
                    pnz[0] = nzmin(pnz[0], v[j].x); pnz[1] = nzmin(pnz[1], v[j].y);
                    pnz[2] = nzmin(pnz[2], v[j].z); pnz[3] = nzmin(pnz[3], v[j].w);
                }
            }
        }
        for (; n < nvec; n += nthr) {
            const uint4 v = ldg_stream(p + n);
            pmn[0] = P::mn(pmn[0], v.x); pmx[0] = P::mx(pmx[0], v.x);
            pmn[1] = P::mn(pmn[1], v.y); pmx[1] = P::mx(pmx[1], v.y);
            pmn[2] = P::mn(pmn[2], v.z); pmx[2] = P::mx(pmx[2], v.z);
            pmn[3] = P::mn(pmn[3], v.w); pmx[3] = P::mx(pmx[3], v.w);
            if (HALF) {
                typedef Pk<false> U;
                umx = U::mx(umx, U::mx(U::mx(v.x, v.y), U::mx(v.z, v.w)));
                umn = U::mn(umn, U::mn(U::mn(v.x, v.y), U::mn(v.z, v.w)));
                pnz[0] = nzmin(pnz[0], v.x); pnz[1] = nzmin(pnz[1], v.y); pnz[2] = nzmin(pnz[2], v.z); pnz[3] = nzmin(pnz[3], v.w);
            }
        }
        const int ph = (int)(tid % 3);                        // kind of word w is (ph + w) % 3
#pragma unroll
        for (int wd = 0; wd < 4; wd++) {
            const int kind = (ph + wd) % 3;
#pragma unroll
            for (int kk = 0; kk < 3; kk++)
                if (kind == kk) {
                    mnv[kk] = P::mn(mnv[kk], pmn[wd]); mxv[kk] = P::mx(mxv[kk], pmx[wd]);
                    nzv[kk] = Pk<false>::mn(nzv[kk], pnz[wd]);
                }
        }
    } else if (NCH == 4) {
        const long nvec = npix / 2;                           // 2 pixels per uint4: (R,G)(B,A)(R,G)(B,A)
        for (long g = tid; g < nvec; g += nthr) {
            const uint4 a = ldg_stream(reinterpret_cast<const uint4 *>(f) + g);
            mnv[0] = P::mn(mnv[0], P::mn(a.x, a.z));
            mxv[0] = P::mx(mxv[0], P::mx(a.x, a.z));
            mnv[1] = P::mn(mnv[1], P::mn(a.y, a.w));
            mxv[1] = P::mx(mxv[1], P::mx(a.y, a.w));
            if (HALF) {                     // alpha (high half of .y/.w) is not a colour sample: mask it to the B code
                typedef Pk<false> U;
                const unsigned by = __byte_perm(a.y, 0, 0x1010), bw = __byte_perm(a.w, 0, 0x1010);
                umx = U::mx(U::mx(umx, U::mx(a.x, a.z)), U::mx(by, bw));
                umn = U::mn(U::mn(umn, U::mn(a.x, a.z)), U::mn(by, bw));
            }
        }
    } else {
        const long nvec = npix / 8;
#pragma unroll
        for (int pl = 0; pl < 3; pl++) {
            const uint4 *p = reinterpret_cast<const uint4 *>(f + (size_t)pl * npix * 2);
            for (long g = tid; g < nvec; g += nthr) {
                const uint4 a = ldg_stream(p + g);
                mnv[pl] = P::mn(P::mn(mnv[pl], a.x), P::mn(a.y, P::mn(a.z, a.w)));
                mxv[pl] = P::mx(P::mx(mxv[pl], a.x), P::mx(a.y, P::mx(a.z, a.w)));
            }
        }
    }
    // warp reduce, then one warp reduces the CTA's 8 warps through shared memory: 8 atomics per CTA
#pragma unroll
    for (int i = 0; i < 3; i++)
        for (int o = 16; o > 0; o >>= 1) {
            mnv[i] = P::mn(mnv[i], __shfl_xor_sync(0xffffffffu, mnv[i], o));
            mxv[i] = P::mx(mxv[i], __shfl_xor_sync(0xffffffffu, mxv[i], o));
        }
    if (HALF) {
        typedef Pk<false> U;
        for (int o = 16; o > 0; o >>= 1) {
            umn = U::mn(umn, __shfl_xor_sync(0xffffffffu, umn, o));
            umx = U::mx(umx, __shfl_xor_sync(0xffffffffu, umx, o));
#pragma unroll
            for (int i = 0; i < 3; i++) nzv[i] = U::mn(nzv[i], __shfl_xor_sync(0xffffffffu, nzv[i], o));
        }
    }
    __shared__ unsigned red[8][12];
    if ((threadIdx.x & 31) == 0) {
        unsigned *r = red[threadIdx.x >> 5];
        r[0] = mnv[0]; r[1] = mnv[1]; r[2] = mnv[2]; r[3] = mxv[0]; r[4] = mxv[1]; r[5] = mxv[2]; r[6] = umn; r[7] = umx;
        r[8] = nzv[0]; r[9] = nzv[1]; r[10] = nzv[2];
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;              // the masked kernel's frame loop synchronises before `red` is written again
    {
        typedef Pk<false> U;
        const int wsrc = threadIdx.x & 7;
        unsigned a0 = red[wsrc][0], a1 = red[wsrc][1], a2 = red[wsrc][2], b0 = red[wsrc][3], b1 = red[wsrc][4], b2 = red[wsrc][5],
                 c0 = red[wsrc][6], c1 = red[wsrc][7];
        for (int o = 4; o > 0; o >>= 1) {
            a0 = P::mn(a0, __shfl_xor_sync(0xffffffffu, a0, o)); a1 = P::mn(a1, __shfl_xor_sync(0xffffffffu, a1, o));
            a2 = P::mn(a2, __shfl_xor_sync(0xffffffffu, a2, o)); b0 = P::mx(b0, __shfl_xor_sync(0xffffffffu, b0, o));
            b1 = P::mx(b1, __shfl_xor_sync(0xffffffffu, b1, o)); b2 = P::mx(b2, __shfl_xor_sync(0xffffffffu, b2, o));
            c0 = U::mn(c0, __shfl_xor_sync(0xffffffffu, c0, o)); c1 = U::mx(c1, __shfl_xor_sync(0xffffffffu, c1, o));
        }
        mnv[0] = a0; mnv[1] = a1; mnv[2] = a2; mxv[0] = b0; mxv[1] = b1; mxv[2] = b2; umn = c0; umx = c1;
        unsigned z0 = red[wsrc][8], z1 = red[wsrc][9], z2 = red[wsrc][10];
        for (int o = 4; o > 0; o >>= 1) {
            z0 = U::mn(z0, __shfl_xor_sync(0xffffffffu, z0, o)); z1 = U::mn(z1, __shfl_xor_sync(0xffffffffu, z1, o));
            z2 = U::mn(z2, __shfl_xor_sync(0xffffffffu, z2, o));
        }
        nzv[0] = z0; nzv[1] = z1; nzv[2] = z2;
    }
    if (HALF) {
        if (threadIdx.x == 0) {
            atomicMin(&slots[frame * SLOTS + 6], min(umn & 0xffffu, umn >> 16));
            atomicMax(&slots[frame * SLOTS + 7], max(umx & 0xffffu, umx >> 16));
            if (NCH == 3) {     // the (word kind, half) pairs of G, B, R: as for the extrema below
                atomicMin(&slots[frame * SLOTS + 8], min(nzv[0] >> 16, nzv[2] & 0xffffu));
                atomicMin(&slots[frame * SLOTS + 9], min(nzv[1] & 0xffffu, nzv[2] >> 16));
                atomicMin(&slots[frame * SLOTS + 10], min(nzv[0] & 0xffffu, nzv[1] >> 16));
            }
        }
    } else if (threadIdx.x == 0 && blockIdx.x == 0) {
        slots[frame * SLOTS + 6] = 0u; slots[frame * SLOTS + 7] = 0xffffu;     // integer codes: not the v2 route
    }
    if (threadIdx.x == 0) {
        // (word kind, half) pairs that hold each of G, B, R (slot order 0/1/2 = G/B/R)
        unsigned cmn[3][2], cmx[3][2];
#define LO16(x) ((x) & 0xffffu)
#define HI16(x) ((x) >> 16)
        if (NCH == 3) {
            cmn[0][0] = HI16(mnv[0]); cmn[0][1] = LO16(mnv[2]);   cmx[0][0] = HI16(mxv[0]); cmx[0][1] = LO16(mxv[2]);   // G
            cmn[1][0] = LO16(mnv[1]); cmn[1][1] = HI16(mnv[2]);   cmx[1][0] = LO16(mxv[1]); cmx[1][1] = HI16(mxv[2]);   // B
            cmn[2][0] = LO16(mnv[0]); cmn[2][1] = HI16(mnv[1]);   cmx[2][0] = LO16(mxv[0]); cmx[2][1] = HI16(mxv[1]);   // R
        } else if (NCH == 4) {
            cmn[0][0] = cmn[0][1] = HI16(mnv[0]);  cmx[0][0] = cmx[0][1] = HI16(mxv[0]);
            cmn[1][0] = cmn[1][1] = LO16(mnv[1]);  cmx[1][0] = cmx[1][1] = LO16(mxv[1]);
            cmn[2][0] = cmn[2][1] = LO16(mnv[0]);  cmx[2][0] = cmx[2][1] = LO16(mxv[0]);
        } else {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                cmn[c][0] = LO16(mnv[c]); cmn[c][1] = HI16(mnv[c]);
                cmx[c][0] = LO16(mxv[c]); cmx[c][1] = HI16(mxv[c]);
            }
        }
#undef LO16
#undef HI16
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                unsigned a = cmn[c][j], b = cmx[c][j], k;
                if (!HALF && clip_on) {     // the clip is monotone: clip(min) = min(clip)
                    a = a < lo ? lo : (a > hi ? hi : a);
                    b = b < lo ? lo : (b > hi ? hi : b);
                }
                if (P::key(a, k)) atomicMin(&slots[frame * SLOTS + c * 2 + 0], k);
                if (P::key(b, k)) atomicMax(&slots[frame * SLOTS + c * 2 + 1], k);
            }
    }
}

// grid = (blocks, nframes); npix % 8 == 0, frames 16-byte aligned.  NCH = 3 or 4 interleaved, 0 = planar.
template <bool HALF, int NCH>
__global__ void __launch_bounds__(256, 4)
k_stats_vec(const uint8_t *__restrict__ src, size_t frame_stride, long npix, int clip_on, unsigned lo, unsigned hi,
            unsigned *slots)
{
    stats_vec_frame<HALF, NCH>(src, frame_stride, npix, clip_on, lo, hi, slots, blockIdx.y);
}

// Behind a plan-reuse pass: only the frames it handed back (mask[frame] != 0), usually none.  A lean grid (a few rows of
// blocks that walk the frames) instead of one row of blocks per frame: the launch costs what an empty kernel costs.
template <bool HALF, int NCH>
__global__ void __launch_bounds__(256, 4)
k_stats_vec_masked(const uint8_t *__restrict__ src, size_t frame_stride, long npix, int clip_on, unsigned lo, unsigned hi,
                   unsigned *slots, const int *mask, const SpecCtl *ctl, int nframes)
{
    if (ctl->nflag == 0) return;
    for (int frame = blockIdx.y; frame < nframes; frame += gridDim.y) {
        if (!mask[frame]) continue;             // uniform per block
        stats_vec_frame<HALF, NCH>(src, frame_stride, npix, clip_on, lo, hi, slots, frame);
        __syncthreads();
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
// `mask` (behind a plan-reuse pass): frames with mask[frame] == 0 keep their FrameK and take no part in the slot search.
__device__ __forceinline__ void plan_frames(const unsigned *slots, FrameK *fk, int nframes, int is_float, int bit_depth, int is_half, const int *mask)
{
    const int n = nframes * 3;
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        if (mask && !mask[p / 3]) continue;
        FrameK &f = fk[p / 3];
        const int c = p % 3;
        if (c == 0) f.spec_done = 0;
        unsigned kmin = slots[(p / 3) * SLOTS + c * 2], kmax = slots[(p / 3) * SLOTS + c * 2 + 1];
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
        if (mask && !mask[p / 3]) continue;
        FrameK &f = fk[p / 3];
        const int c = p % 3;
        int slot = p;
        for (int q = 0; q < p; q++) {
            if (mask && !mask[q / 3]) continue;
            const FrameK &g = fk[q / 3];
            if (g.est_floor[q % 3] == f.est_floor[c] && g.est_ceiling[q % 3] == f.est_ceiling[c]) { slot = q; break; }
        }
        f.lut_slot[c] = slot;
    }
    __syncthreads();
    for (int fi = threadIdx.x; fi < nframes; fi += blockDim.x) {
        if (mask && !mask[fi]) continue;
        FrameK &f = fk[fi];
        f.same_lut = f.lut_slot[0] == f.lut_slot[1] && f.lut_slot[0] == f.lut_slot[2];
        // "clean": half source whose raw codes are all below +inf's (finite, sign bit clear, so code order is
        // value order), every normalisation range positive, one LUT for the three channels
        const unsigned ulo = slots[fi * SLOTS + 6], uhi = slots[fi * SLOTS + 7];
        f.code_lo = ulo; f.code_hi = uhi;
        f.clean = is_float && f.same_lut && ulo <= uhi && uhi < 0x7C00u && f.range[0] > 0.0f && f.range[1] > 0.0f &&
                  f.range[2] > 0.0f;
        f.lut2_ok = f.clean && uhi < LUT2_CODES;
        // a table per channel: the extrema are half values here, so their bit patterns bound the channel's codes
        unsigned need = 0;
        for (int c = 0; c < 3; c++) {
            f.ch_lo[c] = __half_as_ushort(__float2half_rn(f.fmin[c]));
            f.ch_hi[c] = __half_as_ushort(__float2half_rn(f.fmax[c]));
            // exact zeros (black bars) would stretch the table down to code 0: when the pass collected the smallest
            // nonzero code, the table starts one entry below it and that entry holds the value of code 0; the kernel
            // raises every code to ch_lo first, which only moves the zeros
            const unsigned nzm1 = slots[fi * SLOTS + 8 + c];
            f.zero_entry[c] = 0;
            if (f.ch_lo[c] == 0 && nzm1 < 0xffffu && nzm1 + 1 <= f.ch_hi[c]) { f.ch_lo[c] = nzm1; f.zero_entry[c] = 1; }
            need += f.ch_hi[c] >= f.ch_lo[c] ? f.ch_hi[c] - f.ch_lo[c] + 1 : LUT3_FLOATS + 1;
        }
        f.clean3 = is_half && !f.same_lut && ulo <= uhi && uhi < 0x7C00u && f.range[0] > 0.0f && f.range[1] > 0.0f &&
                   f.range[2] > 0.0f && need <= LUT3_FLOATS;
    }
}

__global__ void k_plan(const unsigned *slots, FrameK *fk, int nframes, int is_float, int bit_depth, int is_half)
{
    plan_frames(slots, fk, nframes, is_float, bit_depth, is_half, nullptr);
}
// behind a plan-reuse pass: the frames it handed back (usually none)
__global__ void k_plan_masked(const unsigned *slots, FrameK *fk, int nframes, int is_float, int bit_depth, int is_half, const int *mask,
                              const SpecCtl *ctl)
{
    if (ctl->nflag == 0) return;
    plan_frames(slots, fk, nframes, is_float, bit_depth, is_half, mask);
}

// LUT p (= frame * 3 + channel) is built only by its owner (lut_slot == p), one 16-bit code per thread
__device__ __forceinline__ void build_lut(const FrameK *fk, float *luts, int is_half, int tf_lin, int tf_enc, int clip_on, unsigned lo,
                                          unsigned hi, const int p)
{
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

// grid = (256, nframes*3)
__global__ void __launch_bounds__(256)
k_build_lut(const FrameK *fk, float *luts, int is_half, int tf_lin, int tf_enc, int clip_on, unsigned lo, unsigned hi)
{
    build_lut(fk, luts, is_half, tf_lin, tf_enc, clip_on, lo, hi, blockIdx.y);
}
// behind a plan-reuse pass: grid = (256, a few), the blocks walk the LUTs of the frames that were handed back
__global__ void __launch_bounds__(256)
k_build_lut_masked(const FrameK *fk, float *luts, int is_half, int tf_lin, int tf_enc, int clip_on, unsigned lo, unsigned hi,
                   const int *mask, const SpecCtl *ctl, int nframes)
{
    if (ctl->nflag == 0) return;
    for (int p = blockIdx.y; p < nframes * 3; p += gridDim.y)
        if (mask[p / 3]) build_lut(fk, luts, is_half, tf_lin, tf_enc, clip_on, lo, hi, p);
}

// ---- launchers ---------------------------------------------------------------------------------

// `mask` != nullptr: behind a plan-reuse pass -- the slots of the masked frames were re-initialised by k_spec_verify,
// the other frames keep what the SPEC kernel and the first k_plan produced
static h2y_status stats_and_luts(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, const void *d_src,
                                 size_t src_stride, int nframes, FrameK **d_framek, float **d_luts, cudaStream_t st,
                                 const int *mask)
{
    void *slots, *fk, *luts;
    h2y_status s;
    if ((s = scratch_reserve(c, SCR_STATS, (size_t)nframes * SLOTS * sizeof(unsigned), &slots)) != H2Y_OK) return s;
    if ((s = scratch_reserve(c, SCR_FRAMEK, (size_t)nframes * sizeof(FrameK), &fk)) != H2Y_OK) return s;
    if ((s = scratch_reserve(c, SCR_LUT, (size_t)nframes * 3 * 65536 * sizeof(float), &luts)) != H2Y_OK) return s;
    const long npix = (long)p.src.width * p.src.height;
    const int nslots = nframes * SLOTS;
    if (!mask) k_stats_init<<<(nslots + 255) / 256, 256, 0, st>>>((unsigned *)slots, nslots);
    const int blocks = (int)(((npix + 2047) / 2048) < 4L * c->sm_count ? ((npix + 2047) / 2048) : 4L * c->sm_count);
    dim3 grid(blocks, nframes);
    const bool half = layout_is_half(p.src.layout);
    const int nch = layout_is_planar(p.src.layout) ? 0 : layout_channels(p.src.layout);
    const bool vec = (npix % 8) == 0 && (((uintptr_t)d_src | src_stride) & 15) == 0;
    const uint8_t *sb = (const uint8_t *)d_src;
    unsigned *sl = (unsigned *)slots;
    const SpecCtl *ctl = nullptr;
    if (mask) {
        SpecDev sd;
        if ((s = spec_dev(c, &sd)) != H2Y_OK) return s;
        ctl = sd.ctl;
    }
    if (vec) {
        // 2 CTAs of 256 threads per SM per frame row keeps ~16 KB in flight per SM
        // grid.x * 256 threads must be a multiple of 3 (see k_stats_vec, 3-channel case)
        const int gx = (c->sw.stats_gx ? c->sw.stats_gx : (nframes >= 8 ? 2 * c->sm_count : 6 * c->sm_count)) / 3 * 3;
        dim3 vgrid(gx, nframes);
        if (mask) {
            dim3 lean(gx, nframes < 4 ? nframes : 4);
            if (!half) return H2Y_ERR_UNSUPPORTED;                   // the plan-reuse route is the half-float route
            if (nch == 3) k_stats_vec_masked<true, 3><<<lean, 256, 0, st>>>(sb, src_stride, npix, 0, 0, 0, sl, mask, ctl, nframes);
            else k_stats_vec_masked<true, 4><<<lean, 256, 0, st>>>(sb, src_stride, npix, 0, 0, 0, sl, mask, ctl, nframes);
        } else if (half) {
            if (nch == 3) k_stats_vec<true, 3><<<vgrid, 256, 0, st>>>(sb, src_stride, npix, 0, 0, 0, sl);
            else k_stats_vec<true, 4><<<vgrid, 256, 0, st>>>(sb, src_stride, npix, 0, 0, 0, sl);
        } else {
            if (nch == 3) k_stats_vec<false, 3><<<vgrid, 256, 0, st>>>(sb, src_stride, npix, k.clip_on_load, k.loadLo, k.loadHi, sl);
            else if (nch == 4) k_stats_vec<false, 4><<<vgrid, 256, 0, st>>>(sb, src_stride, npix, k.clip_on_load, k.loadLo, k.loadHi, sl);
            else k_stats_vec<false, 0><<<vgrid, 256, 0, st>>>(sb, src_stride, npix, k.clip_on_load, k.loadLo, k.loadHi, sl);
        }
    } else if (mask)
        return H2Y_ERR_UNSUPPORTED;      // the plan-reuse route only exists for vector-aligned frames
    else if (half)
        k_stats_codes<true><<<grid, 256, 0, st>>>((const uint16_t *)d_src, src_stride / 2, p.src.layout, npix, 0, 0, 0,
                                                  (unsigned *)slots);
    else
        k_stats_codes<false><<<grid, 256, 0, st>>>((const uint16_t *)d_src, src_stride / 2, p.src.layout, npix,
                                                   k.clip_on_load, k.loadLo, k.loadHi, (unsigned *)slots);
    if (mask) {
        k_plan_masked<<<1, 256, 0, st>>>((const unsigned *)slots, (FrameK *)fk, nframes, 1, p.src.bit_depth, 1, mask, ctl);
        k_build_lut_masked<<<dim3(256, 6), 256, 0, st>>>((const FrameK *)fk, (float *)luts, 1, k.tf_linearise, k.tf_encode,
                                                          k.clip_on_load, k.loadLo, k.loadHi, mask, ctl, nframes);
    } else {
        k_plan<<<1, 256, 0, st>>>((const unsigned *)slots, (FrameK *)fk, nframes, half ? 1 : 0, p.src.bit_depth, half ? 1 : 0);
        k_build_lut<<<dim3(256, nframes * 3), 256, 0, st>>>((const FrameK *)fk, (float *)luts, half ? 1 : 0, k.tf_linearise,
                                                             k.tf_encode, k.clip_on_load, k.loadLo, k.loadHi);
    }
    c->launches += 4;
    H2Y_CUDA(c, cudaGetLastError());
    *d_framek = (FrameK *)fk;
    *d_luts = (float *)luts;
    return H2Y_OK;
}

h2y_status launch_stats_and_luts(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, const void *d_src,
                                 size_t src_stride, int nframes, FrameK **d_framek, float **d_luts, cudaStream_t st)
{
    return stats_and_luts(c, p, k, d_src, src_stride, nframes, d_framek, d_luts, st, nullptr);
}

// ---- plan reuse: the passes around the SPEC instantiations of the rows kernel (h2y_forward2.cu) -----------------------

// device state of the context: SpecSeed | SpecCtl | predicted FrameK | seed LUT (65536 floats)
h2y_status spec_dev(h2y_ctx_impl *c, SpecDev *out)
{
    const size_t bytes = 1024 + 65536 * sizeof(float);
    if (!c->spec_dev) {
        H2Y_CUDA(c, cudaMalloc(&c->spec_dev, bytes));
        H2Y_CUDA(c, cudaMemset(c->spec_dev, 0, bytes));             // seed.valid = 0
        H2Y_CUDA(c, cudaHostAlloc((void **)&c->spec_fb, sizeof(SpecCtl), cudaHostAllocDefault));
        memset(c->spec_fb, 0, sizeof(SpecCtl));
    }
    uint8_t *b = (uint8_t *)c->spec_dev;
    out->seed = (SpecSeed *)b;
    out->ctl = (SpecCtl *)(b + 128);
    out->pred = (FrameK *)(b + 256);
    out->seed_lut = (float *)(b + 1024);
    static_assert(sizeof(SpecSeed) <= 128 && sizeof(SpecCtl) <= 128 && sizeof(FrameK) <= 768, "layout of the plan-reuse block");
    return H2Y_OK;
}

// Everything the SPEC kernel needs before it starts: the predicted FrameK (one for all frames), fresh statistics slots,
// cleared per-frame flags, the control block.
__global__ void k_spec_prepare(const SpecSeed *seed, SpecCtl *ctl, FrameK *pred, unsigned *slots, int *bail, int *flag,
                               int nframes, int seq)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nframes * SLOTS) {
        const int j = i % SLOTS;
        slots[i] = j >= 8 ? 0xffffffffu : ((j & 1) ? 0u : 0xffffffffu);
    }
    if (i < nframes) { bail[i] = 0; flag[i] = 0; }
    if (i == 0) {
        ctl->skip = !seed->valid;
        ctl->nflag = 0; ctl->mode = 0; ctl->nframes = nframes; ctl->seq = seq; ctl->uniform = 0;
        FrameK f;
        memset(&f, 0, sizeof(f));
        for (int c = 0; c < 3; c++) {
            f.est_floor[c] = seed->est_floor; f.est_ceiling[c] = seed->est_ceiling;
            f.offset[c] = (float)seed->est_floor; f.range[c] = (float)(seed->est_ceiling - seed->est_floor);
            f.lut_slot[c] = 0;                              // the seed LUT
        }
        f.same_lut = 1;
        f.clean = seed->valid;
        f.code_lo = seed->code_lo; f.code_hi = seed->code_hi;
        f.lut2_ok = seed->valid && seed->lut2_ok;
        *pred = f;
    }
}

// Compare the plan k_plan derived from the gathered extrema with the predicted one.  A frame passes when the SPEC kernel
// did not hand it back, it is a clean single-table frame, its (int) floor / ceiling are the predicted ones (so the seed
// LUT is the function the reference applies, convert.cpp:1017-1019) and its codes lie inside the window the kernel had in
// shared memory.  Frames that fail get fresh statistics slots: the classic pass recomputes them from the samples.
__global__ void k_spec_verify(const FrameK *pred, FrameK *fk, const int *bail, int *flag, SpecCtl *ctl, unsigned *slots, int nframes,
                              int bit_depth)
{
    // the plan the reference would derive from the extrema the SPEC kernel gathered (every frame, one block)
    plan_frames(slots, fk, nframes, 1, bit_depth, 1, nullptr);
    __syncthreads();
    __shared__ int count;
    if (threadIdx.x == 0) count = 0;
    __syncthreads();
    const FrameK &q = *pred;
    for (int f = threadIdx.x; f < nframes; f += blockDim.x) {
        FrameK &a = fk[f];
        bool ok = !ctl->skip && !bail[f] && a.clean && a.same_lut && a.code_lo >= q.code_lo && a.code_hi <= q.code_hi;
        for (int c = 0; c < 3; c++) ok = ok && a.est_floor[c] == q.est_floor[c] && a.est_ceiling[c] == q.est_ceiling[c];
        flag[f] = !ok;
        a.spec_done = ok;
        if (!ok) {
            atomicAdd(&count, 1);
            for (int j = 0; j < SLOTS; j++) slots[f * SLOTS + j] = j >= 8 ? 0xffffffffu : ((j & 1) ? 0u : 0xffffffffu);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ctl->nflag = count;
        // the general kernel converts what was handed back (it spreads every frame over the whole GPU)
        ctl->mode = count == 0 ? 0 : 1;
    }
}

// After any call on the EXR rows route: the last frame's plan becomes the seed for the next call, with the window
// widened to every half code that truncates into [floor, ceiling] (so that later frames with slightly different extrema
// but the same (int) pair still pass), and the control block gets the numbers the host's policy reads (late).
__global__ void k_seed_update(const FrameK *fk, const float *luts, SpecSeed *seed, float *seed_lut, SpecCtl *ctl, int nframes,
                              int seq, int after_spec, int force)
{
    __shared__ int copy, uniform;
    const FrameK &l = fk[nframes - 1];
    if (threadIdx.x == 0) {
        copy = 0; uniform = 0;
        const bool single = l.clean && l.same_lut && l.est_floor[0] >= 0 && l.est_ceiling[0] > l.est_floor[0];
        if (after_spec && l.spec_done) {
            // confirmed against the seed: nothing changes
        } else if (single) {
            const unsigned lo = __half_as_ushort(__float2half_ru((float)l.est_floor[0]));
            const unsigned top = __half_as_ushort(__float2half_ru((float)l.est_ceiling[0] + 1.0f));   // smallest half >= ceiling + 1
            const unsigned hi = (top > 0x7C00u ? 0x7C00u : top) - 1u;
            if (hi >= lo && hi < 0x7C00u && l.code_lo >= lo && l.code_hi <= hi) {
                copy = force || !seed->valid || seed->est_floor != l.est_floor[0] || seed->est_ceiling != l.est_ceiling[0];
                seed->valid = 1;
                seed->est_floor = l.est_floor[0]; seed->est_ceiling = l.est_ceiling[0];
                seed->code_lo = lo; seed->code_hi = hi;
                seed->lut2_ok = hi < LUT2_CODES;
            } else seed->valid = 0;
        } else seed->valid = 0;
    }
    __syncthreads();
    if (copy) {
        const float *gl = luts + (size_t)l.lut_slot[0] * 65536;
        for (unsigned c = threadIdx.x; c < 0x7C00u; c += blockDim.x) seed_lut[c] = gl[c];
    }
    if (!after_spec) {
        // classic call: how many frames share the last frame's plan (what a plan-reuse pass would have confirmed)
        int n = 0;
        for (int f = threadIdx.x; f < nframes; f += blockDim.x) {
            const FrameK &a = fk[f];
            bool same = a.clean && a.same_lut;
            for (int c = 0; c < 3; c++) same = same && a.est_floor[c] == l.est_floor[0] && a.est_ceiling[c] == l.est_ceiling[0];
            n += same;
        }
        if (n) atomicAdd(&uniform, n);
        __syncthreads();
        if (threadIdx.x == 0) {
            ctl->skip = 0; ctl->mode = 0; ctl->nframes = nframes; ctl->seq = seq;
            ctl->uniform = uniform; ctl->nflag = nframes - uniform;
        }
    }
}

h2y_status launch_spec_prepare(h2y_ctx_impl *c, int nframes, int seq, FrameK **d_framek, unsigned **d_slots, int **d_bail,
                               int **d_flag, cudaStream_t st)
{
    SpecDev sd;
    h2y_status s = spec_dev(c, &sd);
    if (s != H2Y_OK) return s;
    void *slots, *fl;
    if ((s = scratch_reserve(c, SCR_STATS, (size_t)nframes * SLOTS * sizeof(unsigned), &slots)) != H2Y_OK) return s;
    if ((s = scratch_reserve(c, SCR_SPEC, (size_t)nframes * 2 * sizeof(int), &fl)) != H2Y_OK) return s;
    *d_slots = (unsigned *)slots;
    *d_bail = (int *)fl;
    *d_flag = (int *)fl + nframes;
    *d_framek = sd.pred;
    const int n = nframes * SLOTS;
    k_spec_prepare<<<(n + 255) / 256, 256, 0, st>>>(sd.seed, sd.ctl, sd.pred, *d_slots, *d_bail, *d_flag, nframes, seq);
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    return H2Y_OK;
}

// behind the SPEC kernel: plan from the gathered extrema, verify, then the classic statistics / plan / LUT pass for the
// frames that were handed back (each kernel returns at once for the others)
h2y_status launch_spec_verify_and_redo_prologue(h2y_ctx_impl *c, const h2y_forward_params &p, const PixK &k, const void *d_src,
                                                size_t src_stride, int nframes, float **d_luts, cudaStream_t st)
{
    SpecDev sd;
    h2y_status s = spec_dev(c, &sd);
    if (s != H2Y_OK) return s;
    void *fk;
    if ((s = scratch_reserve(c, SCR_FRAMEK, (size_t)nframes * sizeof(FrameK), &fk)) != H2Y_OK) return s;
    unsigned *slots = (unsigned *)c->scratch[SCR_STATS];
    int *bail = (int *)c->scratch[SCR_SPEC], *flag = bail + nframes;
    k_spec_verify<<<1, 256, 0, st>>>(sd.pred, (FrameK *)fk, bail, flag, sd.ctl, slots, nframes, p.src.bit_depth);
    c->launches += 1;
    H2Y_CUDA(c, cudaGetLastError());
    FrameK *dfk;
    return stats_and_luts(c, p, k, d_src, src_stride, nframes, &dfk, d_luts, st, flag);
}

h2y_status launch_seed_update(h2y_ctx_impl *c, const FrameK *d_framek, const float *d_luts, int nframes, int seq,
                              int after_spec, int force, cudaStream_t st)
{
    SpecDev sd;
    h2y_status s = spec_dev(c, &sd);
    if (s != H2Y_OK) return s;
    k_seed_update<<<1, 256, 0, st>>>(d_framek, d_luts, sd.seed, sd.seed_lut, sd.ctl, nframes, seq, after_spec, force);
    c->launches++;
    H2Y_CUDA(c, cudaGetLastError());
    H2Y_CUDA(c, cudaMemcpyAsync(c->spec_fb, sd.ctl, sizeof(SpecCtl), cudaMemcpyDeviceToHost, st));
    return H2Y_OK;
}

h2y_status launch_stats_planar(h2y_ctx_impl *c, const h2y_pic_desc &pic, const void *const d_planes[3],
                               FrameK **d_framek, cudaStream_t st)
{
    void *slots, *fk;
    h2y_status s;
    if ((s = scratch_reserve(c, SCR_STATS, SLOTS * sizeof(unsigned), &slots)) != H2Y_OK) return s;
    if ((s = scratch_reserve(c, SCR_FRAMEK, sizeof(FrameK), &fk)) != H2Y_OK) return s;
    int pw[3], ph[3];
    h2y_plane_dims(pic.width, pic.height, pic.chroma_format_idc, pw, ph);
    const long n0 = (long)pw[0] * ph[0], n12 = (long)pw[1] * ph[1];
    k_stats_init<<<1, 256, 0, st>>>((unsigned *)slots, SLOTS);
    const int blocks = (int)(((n0 + 2047) / 2048) < 4L * c->sm_count ? ((n0 + 2047) / 2048) : 4L * c->sm_count);
    const bool isf = pic.pic_buffer_type == H2Y_PIC_TYPE_F32;
    if (isf)
        k_stats_f32<<<blocks, 256, 0, st>>>((const float *)d_planes[0], (const float *)d_planes[1],
                                            (const float *)d_planes[2], n0, (unsigned *)slots);
    else
        k_stats_u16_planes<<<blocks, 256, 0, st>>>((const uint16_t *)d_planes[0], (const uint16_t *)d_planes[1],
                                                   (const uint16_t *)d_planes[2], n0, n12, (unsigned *)slots);
    k_plan<<<1, 256, 0, st>>>((const unsigned *)slots, (FrameK *)fk, 1, isf ? 1 : 0, pic.bit_depth, 0);
    c->launches += 3;
    H2Y_CUDA(c, cudaGetLastError());
    *d_framek = (FrameK *)fk;
    return H2Y_OK;
}

}   // namespace h2y
